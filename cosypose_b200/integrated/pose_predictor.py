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
                        n_coarse_iterations=1, n_refiner_iterations=1, shard=False):
        """shard=True (torch.distributed initialised, one process per GPU): every rank passes the SAME global
        detections / initial poses, refines the contiguous shard `sharding.shard_bounds` gives it and the per-iteration
        records of all ranks are exchanged with ONE all-gather, so every rank returns the full result (what the
        multiview matching stage needs).  Hypotheses are independent, so the result equals the 1-GPU result."""
        if shard:
            return self._get_predictions_sharded(images, K, detections, data_TCO_init, n_coarse_iterations,
                                                 n_refiner_iterations)
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

    def _get_predictions_sharded(self, images, K, detections, data_TCO_init, n_coarse_iterations, n_refiner_iterations):
        from .. import sharding
        rank, ws = sharding.world()
        full = detections if data_TCO_init is None else data_TCO_init
        assert full is not None
        n = len(full)
        start, stop = sharding.shard_bounds(n, rank, ws)
        ids = np.arange(start, stop)
        model = self.coarse_model or self.refiner_model
        engine = getattr(model, 'engine', None)
        if len(ids):
            local = full[ids]
            local.infos = local.infos.reset_index(drop=True)
            kw = dict(detections=local) if data_TCO_init is None else dict(data_TCO_init=local)
            _, preds_local = self.get_predictions(images, K, n_coarse_iterations=n_coarse_iterations,
                                                  n_refiner_iterations=n_refiner_iterations, **kw)
            keys = list(preds_local.keys())
            records = sharding.pack_records([preds_local[k] for k in keys])
        else:   # more ranks than hypotheses: this rank contributes padding only
            keys = ([f'coarse/iteration={i}' for i in range(1, n_coarse_iterations + 1)] if data_TCO_init is None
                    else ['external_coarse']) + [f'refiner/iteration={i}' for i in range(1, n_refiner_iterations + 1)]
            device = engine.device if engine is not None else K.device
            records = torch.zeros((0, len(keys) * sharding.RECORD_FLOATS), dtype=torch.float32, device=device)
        gathered = sharding.gather_records(records, n, engine=engine)
        preds = dict()
        for k, fields in zip(keys, sharding.unpack_records(gathered, len(keys))):
            if k == 'external_coarse':
                preds[k] = data_TCO_init
                continue
            preds[k] = tc.PandasTensorCollection(full.infos, **fields)
        return preds[keys[-1]], preds
