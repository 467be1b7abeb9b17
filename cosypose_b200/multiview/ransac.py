"""Multiview candidate matching (stage 2), same interface and outputs as the reference
(cosypose/multiview/ransac.py:137-199), with the arithmetic in libcosyb200.so:

    seeds / tentative matches   engine.ransac_infos     host C++   (cosypose_cext.cpp:36-105)
    camera-pose hypotheses      Engine.ransac_models    one launch over ALL seeds (ransac.py:19-64)
    scoring                     Engine.ransac_score     one launch over ALL rows  (ransac.py:67-88)
    inlier voting               Engine.ransac_inliers_dev  three small launches (cosypose_cext.cpp:107-216)
    scene-level matching        scipy strongly-connected components, pandas (ransac.py:91-134)

The reference chunks the two device stages (`model_bsz`, `score_bsz`) and synchronises per chunk for
a host-side argmin (lib3d/symmetric_distances.py:13-16); here the distances stay on the device and only
the voted integer lists come back.  `model_bsz` / `score_bsz` are accepted and ignored.
"""
import numpy as np
import pandas as pd
import torch
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import connected_components

from .. import engine as E
from ..utils import tensor_collection as tc
from ..utils.timer import Timer


def scene_level_matching(candidates, inliers):
    """Objects = strongly connected components (size >= 2) of the inlier-match graph
    (reference: multiview/ransac.py:91-116)."""
    cand1, cand2 = inliers['inlier_matches_cand1'], inliers['inlier_matches_cand2']
    n_cand = len(candidates)
    graph = csr_matrix((np.ones(len(cand1), dtype=int), (cand1, cand2)), shape=(n_cand, n_cand))
    _, ids = connected_components(graph, directed=True, connection='strong')
    sizes = np.bincount(ids, minlength=1)[ids] if n_cand else np.zeros(0, dtype=int)

    # kept components renumbered 0..n-1 in ascending component id; `obj_id` is appended as the last column
    keep = np.flatnonzero(sizes >= 2)
    _, dense = np.unique(ids[keep], return_inverse=True)
    cand_infos = candidates.infos.iloc[keep].reset_index(drop=True)
    cand_infos['obj_id'] = dense
    poses = candidates.poses[torch.as_tensor(cand_infos['cand_id'].values.astype(np.int64),
                                             device=candidates.poses.device)]
    return tc.PandasTensorCollection(infos=cand_infos, poses=poses)


def make_obj_infos(matched_candidates):
    """One row per object: summed score, number of candidates (reference: ransac.py:119-125)."""
    infos = matched_candidates.infos
    obj_ids, first, inv = np.unique(infos['obj_id'].to_numpy(), return_index=True, return_inverse=True)
    score = infos['score'].to_numpy()
    # per object: sum of the scores in row order, number of rows, label of the first row
    total = np.zeros(len(obj_ids), dtype=np.result_type(score.dtype, np.float32))
    np.add.at(total, inv, score)
    return pd.DataFrame(dict(obj_id=obj_ids, score=total, label=infos['label'].to_numpy()[first],
                             n_cand=np.bincount(inv, minlength=len(obj_ids)).astype(int)))


def get_best_viewpair_pose_est(TC1C2, seeds, inliers):
    best = inliers['best_hypotheses']
    infos = pd.DataFrame(dict(view1=seeds[0][best], view2=seeds[1][best]))
    return tc.PandasTensorCollection(infos=infos, TC1C2=TC1C2[torch.as_tensor(best.astype(np.int64), device=TC1C2.device)])


def multiview_candidate_matching(candidates, mesh_db, model_bsz=1e3, score_bsz=1e5, dist_threshold=0.02,
                                 cameras=None, n_ransac_iter=20, n_min_inliers=3):
    """`mesh_db` is a BatchedMeshes installed into an engine (`mesh_db.engine`)."""
    eng = mesh_db.engine
    timer_models, timer_score, timer_misc = Timer(), Timer(), Timer()
    known_poses = cameras is not None
    if known_poses:
        n_ransac_iter = 1

    timer_misc.start()
    candidates.infos['cand_id'] = np.arange(len(candidates))
    view_ids = candidates.infos['view_id'].values.astype(np.int32)
    label_ids = mesh_db.label_ids(candidates.infos['label'].values)
    timer_misc.pause()

    timer_models.start()
    seeds, tmatches = E.ransac_infos(view_ids, label_ids, n_ransac_iter, 0)
    poses = candidates.poses.to(eng.device, torch.float32).contiguous()
    d_labels = torch.from_numpy(label_ids).to(eng.device)
    d_seeds = torch.from_numpy(seeds).to(eng.device)
    if not known_poses:
        TC1C2 = eng.ransac_models(poses, d_labels, d_seeds)
    else:
        cam_idx = pd.Series(np.arange(len(cameras)), index=cameras.infos['view_id'].values)
        i1 = torch.from_numpy(cam_idx.loc[seeds[0]].values.astype(np.int32)).to(eng.device)
        i2 = torch.from_numpy(cam_idx.loc[seeds[1]].values.astype(np.int32)).to(eng.device)
        TWC = cameras.TWC.to(eng.device, torch.float32).contiguous()
        TC1C2 = eng.compose_inv(TWC, TWC, i1, i2)
    timer_models.pause()

    timer_score.start()
    d_tm = torch.from_numpy(tmatches).to(eng.device)
    dists = eng.ransac_score(poses, d_labels, d_tm, TC1C2)
    # voting on the device (cosyb200_ransac_inliers_dev): the distances never leave the GPU; the host
    # implementation E.ransac_inliers (bit-identical, pinned to the reference's extension) stays as the checker
    inliers = eng.ransac_inliers_dev(seeds[0], seeds[1], d_tm, dists, dist_threshold, n_min_inliers)
    timer_score.pause()

    timer_misc.start()
    pairs_TC1C2 = get_best_viewpair_pose_est(TC1C2, seeds, inliers)
    filtered_candidates = scene_level_matching(candidates, inliers)
    scene_infos = make_obj_infos(filtered_candidates)
    timer_misc.pause()

    return dict(filtered_candidates=filtered_candidates, scene_infos=scene_infos, pairs_TC1C2=pairs_TC1C2,
                time_models=timer_models.stop(), time_score=timer_score.stop(), time_misc=timer_misc.stop(),
                seeds=seeds, tmatches=tmatches, dists=dists, inliers=inliers, TC1C2=TC1C2)
