"""`PosePredictor`: the per-iteration render-and-compare model, same call contract as the
reference's module (reference: cosypose/models/pose.py:18-132) but every stage runs inside
libcosyb200.so.  One iteration is two engine calls with the renderer in between:

    boxes_rend, boxes_crop, K_crop = engine.prepare_iter(K, TCO, label_ids)        # pose.py:45-67
    renders = renderer.render(obj_infos, TCO, K_crop, resolution)                  # pose.py:100-102
    pose9, TCO_out = engine.refine_iter(images, im_ids, boxes_crop, renders, ...)  # pose.py:104-108

A renderer that exposes `prerendered(n_iterations, batch_size)` (all views known up front, as in
the synthetic benchmark) lets the whole loop run in one engine call (`refine_n`); so does the
engine's own rasteriser (`rendering.CudaRasterizer`, attribute `in_engine`), which draws every
iteration's views on the device between the crop geometry and the network.
"""
from types import SimpleNamespace

import numpy as np
import torch

from ..engine import Engine


class PosePredictor:
    def __init__(self, engine, slot, renderer, mesh_db, render_size=(240, 320), pose_dim=9, cfg=None):
        if pose_dim != 9:
            raise ValueError(f'pose_dim={pose_dim} not supported')
        assert tuple(render_size) == (240, 320), 'the engine is built for 240x320 renders'
        assert isinstance(engine, Engine)
        self.engine = engine
        self.slot = slot
        self.renderer = renderer
        self.mesh_db = mesh_db
        self.render_size = tuple(render_size)
        self.pose_dim = pose_dim
        self.cfg = cfg if cfg is not None else SimpleNamespace(init_method='v0', backbone_str='efficientnet-b3',
                                                               n_pose_dims=9)

    # nn.Module-style no-ops so callers written against the reference keep working
    def eval(self):
        return self

    def cuda(self):
        return self

    def float(self):
        return self

    def load_state_dict(self, state_dict, strict=True):
        self.engine.load_pose_model(self.slot, state_dict)
        return self

    def _label_ids(self, labels):
        return torch.from_numpy(self.mesh_db.label_ids(labels)).to(self.engine.device)

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    @torch.no_grad()
    def forward(self, images, K, labels, TCO, n_iterations=1):
        """Reference signature: `images` [B,3,H,W] and `K` [B,3,3] are already gathered per
        hypothesis (pose_predictor.py:41-42)."""
        bsz = images.shape[0]
        assert K.shape == (bsz, 3, 3)
        assert TCO.shape == (bsz, 4, 4)
        assert len(labels) == bsz
        im_ids = torch.arange(bsz, dtype=torch.int32, device=self.engine.device)
        return self.forward_indexed(images, im_ids, K, labels, TCO, n_iterations)

    @torch.no_grad()
    def forward_indexed(self, images, im_ids, K, labels, TCO, n_iterations=1):
        """Same outputs, but `images` [Nim,3,H,W] are shared and hypothesis b reads
        images[im_ids[b]]; `K` [B,3,3] is per hypothesis."""
        eng = self.engine
        bsz = len(labels)
        assert K.shape == (bsz, 3, 3)
        assert TCO.shape == (bsz, 4, 4)
        images = images.contiguous()
        K = K.contiguous().float()
        label_ids = self._label_ids(labels)
        im_ids = im_ids.to(device=eng.device, dtype=torch.int32).contiguous()
        img_hw = images.shape[-2:]
        outputs = dict()
        TCO_input = TCO.detach().contiguous().float()

        stack = None
        if hasattr(self.renderer, 'prerendered'):
            stack = self.renderer.prerendered(n_iterations, bsz)
        in_engine = stack is None and getattr(self.renderer, 'in_engine', False)
        if stack is not None or in_engine:
            # views known up front, or rasterised by the engine itself: the whole loop is one engine call
            out = eng.refine_n(self.slot, images, im_ids, K, label_ids, stack, TCO_input, n_iter=n_iterations)
            for n in range(n_iterations):
                outputs[f'iteration={n + 1}'] = {
                    'TCO_input': TCO_input if n == 0 else out['TCO_output'][n - 1],
                    'TCO_output': out['TCO_output'][n],
                    'K_crop': out['K_crop'][n],
                    'model_outputs': {'pose': out['pose'][n]},
                    'boxes_rend': out['boxes_rend'][n],
                    'boxes_crop': out['boxes_crop'][n],
                }
            return outputs

        obj_infos = [dict(name=l) for l in labels]
        for n in range(n_iterations):
            boxes_rend, boxes_crop, K_crop = eng.prepare_iter(K, TCO_input, label_ids, img_hw)
            renders = self.renderer.render(obj_infos=obj_infos, TCO=TCO_input, K=K_crop,
                                           resolution=self.render_size)
            renders = renders.to(eng.device).contiguous()
            if renders.dtype != torch.uint8:
                renders = renders.float()
            pose9, TCO_output = eng.refine_iter(self.slot, images, im_ids, boxes_crop, renders, K_crop, TCO_input)
            outputs[f'iteration={n + 1}'] = {
                'TCO_input': TCO_input,
                'TCO_output': TCO_output,
                'K_crop': K_crop,
                'model_outputs': {'pose': pose9},
                'boxes_rend': boxes_rend,
                'boxes_crop': boxes_crop,
            }
            TCO_input = TCO_output
        return outputs
