"""Runner-level batching.  The reference's runners feed the pose predictor one view group at a
time (`batch_size=1`: evaluation/pred_runner/multiview_predictions.py:24, bop_predictions.py:26), so a scene with 5
detections launches the trunk on 5 hypotheses.  `HypothesisQueue` collects the detections of many frames / scenes,
refines them in ONE `get_predictions` call (full chunks of `bsz_objects`, sharded over the ranks when asked), and hands
every scene its own rows back, in the scene's original order, ready for `MultiviewScenePredictor`."""
import numpy as np
import pandas as pd
import torch

from ..utils import tensor_collection as tc


class HypothesisQueue:
    def __init__(self, pose_predictor, n_coarse_iterations=1, n_refiner_iterations=4, shard=False):
        self.pose_predictor = pose_predictor
        self.n_coarse, self.n_refine, self.shard = n_coarse_iterations, n_refiner_iterations, shard
        self.reset()

    def reset(self):
        self._images, self._K, self._dets, self._keys, self._n_images = [], [], [], [], 0

    def __len__(self):
        return sum(len(d) for d in self._dets)

    def put(self, key, images, K, detections):
        """One view group: images [V,3,H,W], K [V,3,3], detections with infos[label, batch_im_id (0..V-1), ...] and
        bboxes [n,4].  `key` identifies the group in the results (e.g. (scene_id, group_id))."""
        assert images.shape[0] == K.shape[0]
        infos = detections.infos.copy()
        assert infos['batch_im_id'].between(0, images.shape[0] - 1).all()
        infos['queue_key'] = [key] * len(infos)
        infos['queue_row'] = np.arange(len(infos))
        infos['batch_im_id'] = infos['batch_im_id'].values + self._n_images       # index into the concatenated frames
        self._dets.append(tc.PandasTensorCollection(infos=infos, bboxes=detections.bboxes))
        self._images.append(images)
        self._K.append(K)
        self._keys.append(key)
        self._n_images += images.shape[0]

    def flush(self):
        """Refines everything queued; returns {key: (final collection, preds dict)} with each group's rows in the order
        they were put and `batch_im_id` local to the group again."""
        if not self._dets:
            return {}
        images, K = torch.cat(self._images, dim=0), torch.cat(self._K, dim=0)
        dets = tc.concatenate(self._dets)
        offsets = np.cumsum([0] + [im.shape[0] for im in self._images])
        final, preds = self.pose_predictor.get_predictions(images, K, detections=dets, n_coarse_iterations=self.n_coarse,
                                                           n_refiner_iterations=self.n_refine, shard=self.shard)
        out = {}
        keys_col = final.infos['queue_key'].values
        for g, key in enumerate(self._keys):
            ids = np.array([i for i, k in enumerate(keys_col) if k == key])

            def take(coll):
                c = coll[ids]
                c.infos = c.infos.drop(columns=['queue_key', 'queue_row']).reset_index(drop=True)
                c.infos['batch_im_id'] = c.infos['batch_im_id'].values - offsets[g]
                return c
            out[key] = (take(final), {k: take(v) for k, v in preds.items()})
        self.reset()
        return out
