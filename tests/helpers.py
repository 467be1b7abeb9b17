"""Shared builders for the parity tests: the same seeded workload for the engine and the oracle."""
import numpy as np
import pandas as pd
import torch

from cosypose_b200 import synthetic as syn

_SD_CACHE = {}


def state_dict(seed):
    if seed not in _SD_CACHE:
        _SD_CACHE[seed] = syn.make_pose_state_dict(seed)
    return _SD_CACHE[seed]


class Workload:
    """Inputs of one CoarseRefinePosePredictor.get_predictions call."""

    def __init__(self, n_images, dets, n_labels, n_coarse, n_refine, sym_counts=(1,)):
        self.labels = syn.make_labels(n_labels)
        self.points, self.sym, self.n_sym = syn.make_mesh_tables(n_labels, sym_counts=sym_counts)
        self.boxes, self.label_ids, self.im_ids = syn.make_detections(n_images, dets, n_labels)
        self.n = len(self.label_ids)
        self.images = syn.make_images(n_images)
        self.K = syn.make_camera_K(n_images)
        self.n_coarse, self.n_refine = n_coarse, n_refine
        self.views_c = syn.make_renders(max(n_coarse, 1), self.n, seed=11)
        self.views_r = syn.make_renders(max(n_refine, 1), self.n, seed=12)

    def infos(self):
        return pd.DataFrame(dict(label=[self.labels[i] for i in self.label_ids],
                                 batch_im_id=self.im_ids, score=np.ones(self.n)))

    def oracle_render_fn(self):
        def fn(stage, it, sl, TCO, K_crop):
            return (self.views_c if stage == 'coarse' else self.views_r)[it, sl]
        return fn


def build_predictor(w, device, bsz_objects=64, per_call_renderer=False, max_batch=None):
    """Engine-backed CoarseRefinePosePredictor for workload `w` (weights seed 0 / 1)."""
    from cosypose_b200.engine import Engine
    from cosypose_b200.integrated.pose_predictor import CoarseRefinePosePredictor
    from cosypose_b200.lib3d.rigid_mesh_database import BatchedMeshes
    from cosypose_b200.models.pose import PosePredictor
    from cosypose_b200.rendering import PerCallRenderer, PreRenderedViews
    eng = Engine(device, max_batch=max_batch or bsz_objects)
    mesh_db = BatchedMeshes.from_tables(w.labels, w.points, w.sym, w.n_sym)
    mesh_db.install(eng)
    stages = ([w.views_c] if w.n_coarse else []) + ([w.views_r] if w.n_refine else [])
    views = PreRenderedViews(stages, bsz_objects, device=eng.device)
    renderer = PerCallRenderer(views) if per_call_renderer else views
    coarse = PosePredictor(eng, 0, renderer, mesh_db).load_state_dict(state_dict(0))
    refiner = PosePredictor(eng, 1, renderer, mesh_db).load_state_dict(state_dict(1))
    return CoarseRefinePosePredictor(coarse, refiner, bsz_objects=bsz_objects), eng, views


class Scene:
    """Synthetic multiview scene (SURVEY.md section 8d, config 4 recipe) rebuilt from the meta of a
    golden file: candidates, cameras, AABB mesh tables."""

    def __init__(self, n_views, n_objects, n_labels, sym_counts=(1,), unique_labels=True, seed=0):
        from cosypose_b200.engine import aabb_corners
        self.labels = syn.make_labels(n_labels)
        pts, self.sym, self.n_sym = syn.make_mesh_tables(n_labels, n_points=64, sym_counts=tuple(int(s) for s in sym_counts))
        self.aabb = torch.as_tensor(aabb_corners(pts.numpy()))
        s = syn.make_multiview_scene(n_views, n_objects, n_labels, seed=seed, unique_labels=unique_labels)
        self.view_ids, self.label_ids, self.scores = s['view_ids'], s['label_ids'], s['scores']
        self.poses, self.K, self.TWC, self.TWO = s['poses'], s['K'], s['TWC'], s['TWO']
        self.n_views = n_views

    def candidates(self, device=None):
        from cosypose_b200.utils import tensor_collection as tc
        infos = pd.DataFrame(dict(view_id=self.view_ids, label=[self.labels[i] for i in self.label_ids],
                                  score=self.scores, scene_id=0, group_id=0, batch_im_id=self.view_ids))
        poses = self.poses if device is None else self.poses.to(device)
        return tc.PandasTensorCollection(infos=infos, poses=poses)

    def cameras(self, device=None):
        from cosypose_b200.utils import tensor_collection as tc
        infos = pd.DataFrame(dict(view_id=np.arange(self.n_views), scene_id=0, batch_im_id=np.arange(self.n_views)))
        K, TWC = (self.K, self.TWC) if device is None else (self.K.to(device), self.TWC.to(device))
        return tc.PandasTensorCollection(infos=infos, K=K, TWC=TWC)

    def mesh_db(self):
        from cosypose_b200.lib3d.rigid_mesh_database import BatchedMeshes
        return BatchedMeshes.from_tables(self.labels, self.aabb, self.sym, self.n_sym)
