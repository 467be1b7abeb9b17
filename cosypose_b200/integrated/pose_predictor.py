"""`CoarseRefinePosePredictor` with the reference's interface
(reference: cosypose/integrated/pose_predictor.py:14-107): detections or initial poses in,
per-iteration `PandasTensorCollection`s out, keyed `coarse/iteration=n`, `refiner/iteration=n`
or `external_coarse`.  Differences are internal: hypotheses index the shared image batch instead
of copying a full frame each (pose_predictor.py:41), and all arithmetic runs in libcosyb200.so.
"""
from collections import defaultdict

import numpy as np
import torch

from ..utils import tensor_collection as tc


class CoarseRefinePosePredictor:
    def __init__(self, coarse_model=None, refiner_model=None, bsz_objects=64):
        self.coarse_model = coarse_model
        self.refiner_model = refiner_model
        self.bsz_objects = bsz_objects

    def eval(self):
        return self

    @torch.no_grad()
    def batched_model_predictions(self, model, images, K, obj_data, n_iterations=1):
        preds = defaultdict(list)
        n = len(obj_data)
        device = model.engine.device
        for start in range(0, n, self.bsz_objects):
            ids = np.arange(start, min(n, start + self.bsz_objects))
            obj_inputs = obj_data[ids]
            labels = obj_inputs.infos['label'].values
            im_ids_np = obj_inputs.infos['batch_im_id'].values.astype(np.int64)
            im_ids = torch.as_tensor(im_ids_np, device=device)
            K_ = K[im_ids]
            outputs = model.forward_indexed(images=images, im_ids=im_ids, K=K_, labels=labels,
                                            TCO=obj_inputs.poses, n_iterations=n_iterations)
            for it in range(1, n_iterations + 1):
                o = outputs[f'iteration={it}']
                preds[f'iteration={it}'].append(tc.PandasTensorCollection(
                    obj_inputs.infos, poses=o['TCO_output'], poses_input=o['TCO_input'],
                    K_crop=o['K_crop'], boxes_rend=o['boxes_rend'], boxes_crop=o['boxes_crop']))
        return {k: tc.concatenate(v) for k, v in preds.items()}

    def make_TCO_init(self, detections, K):
        model = self.coarse_model
        eng = model.engine
        im_ids = torch.as_tensor(detections.infos['batch_im_id'].values.astype(np.int64), device=eng.device)
        K_ = K[im_ids].contiguous().float()
        boxes = detections.bboxes.to(eng.device, torch.float32).contiguous()
        label_ids = torch.from_numpy(model.mesh_db.label_ids(detections.infos['label'].values)).to(eng.device)
        zup = model.cfg.init_method == 'z-up+auto-depth'
        TCO_init = eng.tco_init(boxes, K_, label_ids, zup=zup)
        return tc.PandasTensorCollection(infos=detections.infos, poses=TCO_init)

    def get_predictions(self, images, K, detections=None, data_TCO_init=None,
                        n_coarse_iterations=1, n_refiner_iterations=1):
        preds = dict()
        if data_TCO_init is None:
            assert detections is not None
            assert self.coarse_model is not None
            assert n_coarse_iterations > 0
            data_TCO_init = self.make_TCO_init(detections, K)
            coarse_preds = self.batched_model_predictions(self.coarse_model, images, K, data_TCO_init,
                                                          n_iterations=n_coarse_iterations)
            for n in range(1, n_coarse_iterations + 1):
                preds[f'coarse/iteration={n}'] = coarse_preds[f'iteration={n}']
            data_TCO = coarse_preds[f'iteration={n_coarse_iterations}']
        else:
            assert n_coarse_iterations == 0
            data_TCO = data_TCO_init
            preds['external_coarse'] = data_TCO

        if n_refiner_iterations >= 1:
            assert self.refiner_model is not None
            refiner_preds = self.batched_model_predictions(self.refiner_model, images, K, data_TCO,
                                                           n_iterations=n_refiner_iterations)
            for n in range(1, n_refiner_iterations + 1):
                preds[f'refiner/iteration={n}'] = refiner_preds[f'iteration={n}']
            data_TCO = refiner_preds[f'iteration={n_refiner_iterations}']
        return data_TCO, preds
