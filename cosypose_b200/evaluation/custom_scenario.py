"""Multi-view reconstruction of a custom scenario from files (reference:
cosypose/scripts/run_custom_scenario.py:95-192).  A scenario directory holds

    candidates.csv       single-view 6D pose candidates, BOP19 result format (metres are written as mm)
    scene_camera.json    BOP camera file: {"<view id>": {"cam_K": [...]}}
    models/              models_info.json + obj_%06d.ply (millimetres)

and receives `results/subscene=<g>/predicted_scene.json` (objects + cameras in the scene frame) and
`scene_reprojected.csv` (every object in every camera frame, BOP19 format) per view group.

    python -m cosypose_b200.evaluation.custom_scenario --scenario <dir>
"""
import argparse
from pathlib import Path

import numpy as np
import torch

from . import bop_io


def nms3d(preds, th=0.04, poses_attr='poses'):
    """Greedy 3-D non-maximum suppression on the translations, best score first (reference:
    visualization/multiview.py:28-52).  The reference writes the kept poses to `.poses` whatever `poses_attr` is, which
    leaves e.g. `TWO` unfiltered and misaligned with the filtered infos; here the named tensor (and every other tensor
    of the collection) is filtered with the infos."""
    T = getattr(preds, poses_attr).detach().cpu().numpy()
    scores = preds.infos['score'].to_numpy()
    all_t = T[:, :3, -1]
    tested, keep = set(), []
    for idx in np.argsort(-scores, kind='stable').tolist():
        if idx in tested:
            continue
        dists = np.linalg.norm(all_t[idx] - all_t, axis=-1)
        dists[idx] = np.inf
        tested.update(np.where(dists <= th)[0].tolist())
        keep.append(idx)
    return preds[np.asarray(keep, dtype=np.int64)]


def run_custom_scenario(scenario_dir, sv_score_th=0.3, n_symmetries_rot=64, ransac_n_iter=2000,
                        ransac_dist_threshold=0.02, ba_n_iter=10, nms_th=0.04, device=None, log=print):
    from ..integrated.multiview_predictor import MultiviewScenePredictor
    scenario_dir = Path(scenario_dir)
    device = torch.device('cuda', 0) if device is None else torch.device(device)
    candidates = bop_io.read_csv_candidates(scenario_dir / 'candidates.csv').float().to(device)
    candidates.infos['group_id'] = 0
    scene_ids = np.unique(candidates.infos['scene_id'])
    assert len(scene_ids) == 1, 'candidates.csv must hold the candidates of ONE scene'
    scene_id = scene_ids.item()
    view_ids = np.unique(candidates.infos['view_id'])
    log(f'Loaded {len(candidates)} candidates in {len(view_ids)} views.')

    cameras = bop_io.read_cameras(scenario_dir / 'scene_camera.json', view_ids).float().to(device)
    cameras.infos['scene_id'] = scene_id
    cameras.infos['batch_im_id'] = np.arange(len(view_ids))

    mesh_db, _ = bop_io.mesh_db_from_bop_models(scenario_dir / 'models', n_sym=n_symmetries_rot)
    log(f'Loaded {len(mesh_db.labels)} 3D object models.')

    mv_predictor = MultiviewScenePredictor(mesh_db, device=device)
    predictions = mv_predictor.predict_scene_state(candidates, cameras, score_th=sv_score_th,
                                                   use_known_camera_poses=False, ransac_n_iter=ransac_n_iter,
                                                   ransac_dist_threshold=ransac_dist_threshold, ba_n_iter=ba_n_iter)
    objects, cams, reproj = predictions['scene/objects'], predictions['scene/cameras'], predictions['ba_output']
    written = []
    for view_group in np.unique(objects.infos['view_group']).tolist():
        objects_ = objects[np.where(objects.infos['view_group'] == view_group)[0]]
        cameras_ = cams[np.where(cams.infos['view_group'] == view_group)[0]]
        reproj_ = reproj[np.where(reproj.infos['view_group'] == view_group)[0]]
        objects_ = nms3d(objects_, th=nms_th, poses_attr='TWO')
        view_group_dir = scenario_dir / 'results' / f'subscene={view_group}'
        view_group_dir.mkdir(exist_ok=True, parents=True)
        log(f'Subscene {view_group} has {len(objects_)} objects and {len(cameras_)} cameras.')
        bop_io.save_scene_json(objects_, cameras_, view_group_dir / 'predicted_scene.json')
        bop_io.tc_to_csv(reproj_, view_group_dir / 'scene_reprojected.csv')
        written.append(view_group_dir)
    return dict(predictions=predictions, result_dirs=written)


def main():
    parser = argparse.ArgumentParser(description='Multi-view scene reconstruction of a scenario directory (same options as the '
                                                 "reference's scripts/run_custom_scenario.py)")
    parser.add_argument('--scenario', required=True, type=str, help='scenario directory')
    parser.add_argument('--sv_score_th', default=0.3, type=float, help='candidates scoring below this are dropped')
    parser.add_argument('--n_symmetries_rot', default=64, type=int,
                        help='rotations a continuous symmetry is discretised into')
    parser.add_argument('--ransac_n_iter', default=2000, type=int, help='RANSAC seeds per ordered view pair')
    parser.add_argument('--ransac_dist_threshold', default=0.02, type=float,
                        help='inlier bound on the symmetric distance of a tentative match, metres')
    parser.add_argument('--ba_n_iter', default=10, type=int, help='Levenberg-Marquardt iterations of the bundle adjustment')
    parser.add_argument('--nms_th', default=0.04, type=float, help='objects closer than this (metres) to a better one are suppressed')
    args = parser.parse_args()
    run_custom_scenario(args.scenario, args.sv_score_th, args.n_symmetries_rot, args.ransac_n_iter,
                        args.ransac_dist_threshold, args.ba_n_iter, args.nms_th)


if __name__ == '__main__':
    main()
