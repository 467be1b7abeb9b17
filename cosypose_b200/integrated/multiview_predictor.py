"""`MultiviewScenePredictor` with the reference's interface
(cosypose/integrated/multiview_predictor.py:14-127): score filter -> candidate matching -> view
groups -> bundle adjustment per group -> reprojection of the scene into every view.
"""
import numpy as np
import pandas as pd
import torch

from ..engine import Engine
from ..multiview.bundle_adjustment import MultiviewRefinement, make_view_groups
from ..multiview.ransac import multiview_candidate_matching
from ..utils import tensor_collection as tc


class MultiviewScenePredictor:
    def __init__(self, mesh_db, n_sym=64, ba_aabb=True, ba_n_points=None, engine=None, device=None):
        """`mesh_db`: BatchedMeshes holding full vertex clouds and (already expanded) symmetry sets.
        `n_sym` is kept for signature compatibility: continuous symmetries are discretised when the
        tables are built (reference: lib3d/rigid_mesh_database.py:21-56)."""
        assert ba_n_points is None or not ba_aabb
        self.engine = engine if engine is not None else Engine(device, max_batch=1)
        self.mesh_db_ransac = mesh_db.batched(aabb=True)
        self.mesh_db_ransac.install(self.engine, with_points=False)
        self.mesh_db_ba = mesh_db.batched(aabb=ba_aabb, resample_n_points=ba_n_points)
        self.mesh_db_ba.engine = self.engine        # same symmetry tables; BA passes its own points

    def reproject_scene(self, objects, cameras):
        """TCO = inv(TWC) @ TWO for every (object, view), object-major (reference: :20-41)."""
        n_o, n_v = len(objects), len(cameras)
        dev = self.engine.device
        o, v = np.repeat(np.arange(n_o, dtype=np.int32), n_v), np.tile(np.arange(n_v, dtype=np.int32), n_o)
        d_ov = torch.from_numpy(np.stack((o, v))).to(dev)
        poses = self.engine.compose_inv(cameras.TWC.to(dev, torch.float32).contiguous(),
                                        objects.TWO.to(dev, torch.float32).contiguous(), d_ov[1], d_ov[0])
        oi, ci = objects.infos, cameras.infos
        infos = pd.DataFrame(dict(
            scene_id=ci['scene_id'].values[v], view_id=ci['view_id'].values[v],
            score=oi['score'].values[o] + 1.0, view_group=oi['view_group'].values[o],
            label=oi['label'].values[o], batch_im_id=ci['batch_im_id'].values[v],
            obj_id=oi['obj_id'].values[o], from_ba=True))
        return tc.PandasTensorCollection(infos=infos, poses=poses)

    def predict_scene_state(self, candidates, cameras, score_th=0.3, use_known_camera_poses=False,
                            ransac_n_iter=2000, ransac_dist_threshold=0.02, ba_n_iter=100):
        predictions = dict()
        cand_inputs = candidates
        scene_ids = np.unique(candidates.infos['scene_id'].to_numpy())
        assert len(scene_ids) == 1
        scene_id = scene_ids.item()
        group_id = np.unique(candidates.infos['group_id'].to_numpy()).item()
        keep = np.flatnonzero(candidates.infos['score'].to_numpy() >= score_th)
        candidates = candidates[keep]
        predictions['cand_inputs'] = candidates

        matching = multiview_candidate_matching(
            candidates=candidates, mesh_db=self.mesh_db_ransac, n_ransac_iter=ransac_n_iter,
            dist_threshold=ransac_dist_threshold, cameras=cameras if use_known_camera_poses else None)
        pairs_TC1C2 = matching['pairs_TC1C2']
        candidates = matching['filtered_candidates']
        predictions['cand_matched'] = candidates

        group_infos = make_view_groups(pairs_TC1C2)
        view_to_group = dict(zip(group_infos['view_id'].to_numpy().tolist(), group_infos['view_group'].to_numpy().tolist()))
        cand_views = candidates.infos['view_id'].to_numpy().tolist()
        if all(v in view_to_group for v in cand_views):     # the left merge of the reference (:72), without pandas
            infos = candidates.infos.assign(view_group=np.asarray([view_to_group[v] for v in cand_views],
                                                                  dtype=group_infos['view_group'].to_numpy().dtype))
            candidates = tc.PandasTensorCollection(infos=infos, **dict(candidates.tensors)).to(self.engine.device)
        else:
            candidates = candidates.merge_df(group_infos, on='view_id').to(self.engine.device)

        pred_objects, pred_cameras, pred_reproj, pred_reproj_init = [], [], [], []
        cand_groups = candidates.infos['view_group'].to_numpy()
        for view_group in np.unique(cand_groups).tolist():
            candidate_ids = np.flatnonzero(cand_groups == view_group)
            problem = MultiviewRefinement(candidates=candidates[candidate_ids], cameras=cameras,
                                          pairs_TC1C2=pairs_TC1C2, mesh_db=self.mesh_db_ba)
            # every scene collection the problem emits carries its group / scene columns (:88-93)
            tags = dict(view_group=view_group, group_id=group_id, scene_id=scene_id)
            problem.obj_infos = problem.obj_infos.assign(**tags)
            problem.cam_infos = problem.cam_infos.assign(**tags)
            ba = problem.solve(n_iterations=ba_n_iter, optimize_cameras=not use_known_camera_poses)
            pred_reproj.append(self.reproject_scene(ba['objects'], ba['cameras']))
            pred_reproj_init.append(self.reproject_scene(ba['objects_init'], ba['cameras_init']))
            pred_objects.append(ba['objects'])
            pred_cameras.append(ba['cameras'])

        predictions['scene/objects'] = tc.concatenate(pred_objects)
        predictions['scene/cameras'] = tc.concatenate(pred_cameras)
        predictions['ba_output'] = tc.concatenate(pred_reproj)
        predictions['ba_input'] = tc.concatenate(pred_reproj_init)
        cand_inputs = tc.PandasTensorCollection(infos=cand_inputs.infos,
                                                poses=cand_inputs.poses.to(self.engine.device))
        predictions['ba_output+all_cand'] = tc.concatenate([predictions['ba_output'], cand_inputs])
        return predictions
