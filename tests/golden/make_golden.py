"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Only runnable in the build container (the GPU box has no /root/reference); the outputs are
committed.  Inputs are regenerated from seeds by cosypose_b200.synthetic in the tests, so the
fixtures hold reference OUTPUTS (plus input checksums to detect generator drift).

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1] / "baseline"))

import ref_harness  # noqa: E402
from cosypose_b200 import synthetic as syn  # noqa: E402


def checksum(t):
    return float(np.asarray(t, dtype=np.float64).sum())


class _RefRenderer:
    """renderer.render(obj_infos, TCO, K, resolution) replaying pre-generated views in the
    reference's call order (per stage, per chunk, per iteration)."""

    def __init__(self, stages, bsz):
        self.calls = []
        for views in stages:
            for s in range(0, views.shape[1], bsz):
                for it in range(views.shape[0]):
                    self.calls.append(views[it, s:s + bsz])
        self.i = 0

    def render(self, obj_infos, TCO, K, resolution=(240, 320), **kw):
        out = self.calls[self.i % len(self.calls)]
        self.i += 1
        assert out.shape[0] == len(obj_infos)
        return out


def ref_mesh_db(cosypose, labels, points, sym, n_sym):
    from cosypose.lib3d.rigid_mesh_database import BatchedMeshes
    infos = {l: dict(label=l, n_points=points.shape[1], n_sym=int(n)) for l, n in zip(labels, n_sym)}
    return BatchedMeshes(infos, labels, points.clone(), sym.clone()).float()


def ref_pose_model(cosypose, sd, renderer, mesh_db):
    from types import SimpleNamespace
    from cosypose.training.pose_models_cfg import create_model_pose
    cfg = SimpleNamespace(backbone_str='efficientnet-b3', n_pose_dims=9, init_method='v0')
    model = create_model_pose(cfg, renderer, mesh_db)
    missing = model.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if 'num_batches_tracked' not in k], missing
    assert not missing.unexpected_keys, missing
    model.cfg = cfg
    return model.eval()


def golden_single_view(cosypose, name, n_images, dets, n_labels, n_coarse, n_refine, bsz, zup=False):
    import cosypose.utils.tensor_collection as tc
    from cosypose.integrated.pose_predictor import CoarseRefinePosePredictor
    labels = syn.make_labels(n_labels)
    points, sym, n_sym = syn.make_mesh_tables(n_labels)
    mesh_db = ref_mesh_db(cosypose, labels, points, sym, n_sym)
    sd_c, sd_r = syn.make_pose_state_dict(0), syn.make_pose_state_dict(1)
    boxes, label_ids, im_ids = syn.make_detections(n_images, dets, n_labels)
    n = len(label_ids)
    images = syn.make_images(n_images)
    K = syn.make_camera_K(n_images)
    views_c = syn.make_renders(n_coarse, n, seed=11)
    views_r = syn.make_renders(n_refine, n, seed=12)
    renderer = _RefRenderer([views_c, views_r], bsz)
    coarse = ref_pose_model(cosypose, sd_c, renderer, mesh_db)
    refiner = ref_pose_model(cosypose, sd_r, renderer, mesh_db)
    if zup:
        coarse.cfg.init_method = 'z-up+auto-depth'
    pred = CoarseRefinePosePredictor(coarse, refiner, bsz_objects=bsz)
    infos = pd.DataFrame(dict(label=[labels[i] for i in label_ids], batch_im_id=im_ids,
                              score=np.ones(n)))
    detections = tc.PandasTensorCollection(infos=infos, bboxes=boxes)
    with torch.no_grad():
        final, preds = pred.get_predictions(images, K, detections=detections,
                                            n_coarse_iterations=n_coarse, n_refiner_iterations=n_refine)
    out = dict(meta=np.array([n_images, dets, n_labels, n_coarse, n_refine, bsz, int(zup)]),
               chk_images=checksum(images), chk_boxes=checksum(boxes), chk_views=checksum(views_r),
               chk_sd=checksum(sd_r['backbone._conv_head.weight']), final_poses=final.poses.numpy())
    for k, v in preds.items():
        for t in ('poses', 'poses_input', 'K_crop', 'boxes_rend', 'boxes_crop'):
            out[f'{k}/{t}'] = getattr(v, t).numpy()
    np.savez_compressed(HERE / f'{name}.npz', **out)
    print(name, 'final pose[0]:\n', final.poses[0].numpy())


def golden_backbone(cosypose, name='backbone_b2'):
    """Block-boundary taps of the reference EfficientNet on one seeded input (B=2)."""
    labels = syn.make_labels(3)
    points, sym, n_sym = syn.make_mesh_tables(3)
    mesh_db = ref_mesh_db(cosypose, labels, points, sym, n_sym)
    sd = syn.make_pose_state_dict(0)
    model = ref_pose_model(cosypose, sd, None, mesh_db)
    x = syn.make_net_input(2, seed=21)
    taps = {}
    hooks = [model.backbone._bn0.register_forward_hook(lambda m, i, o: taps.__setitem__('stem_bn', o))]
    for i, blk in enumerate(model.backbone._blocks):
        hooks.append(blk.register_forward_hook(lambda m, inp, o, i=i: taps.__setitem__(f'block{i}', o)))
    with torch.no_grad():
        feat = model.backbone(x)
        pose = model.net_forward(x)['pose']
    out = dict(chk_x=checksum(x), pose=pose.numpy(), pooled=feat.flatten(2).mean(-1).numpy())
    # full tensors are large; keep block outputs for a few blocks + per-block statistics for all
    for i in (0, 1, 2, 5, 8, 13, 18, 25):
        out[f'block{i}'] = taps[f'block{i}'][0].permute(1, 2, 0).numpy().astype(np.float32)[::4, ::4]
    out['block_mean_abs'] = np.array([taps[f'block{i}'].abs().mean().item() for i in range(26)])
    out['block_sum'] = np.array([taps[f'block{i}'].double().sum().item() for i in range(26)])
    np.savez_compressed(HERE / f'{name}.npz', **out)
    print(name, 'pose[0]:', pose[0].numpy(), 'mean|act|:', out['block_mean_abs'].round(3))


def golden_multiview(cosypose, name, n_views, n_objects, n_labels, sym_counts, unique_labels, n_iter, seed=0):
    import cosypose.utils.tensor_collection as tc
    from cosypose.multiview.ransac import multiview_candidate_matching
    from cosypose.lib3d.symmetric_distances import symmetric_distance_batched_fast
    labels = syn.make_labels(n_labels)
    points, sym, n_sym = syn.make_mesh_tables(n_labels, n_points=64, sym_counts=sym_counts)
    from cosypose.lib3d.mesh_ops import get_meshes_bounding_boxes
    aabb = get_meshes_bounding_boxes(points)
    mesh_db = ref_mesh_db(cosypose, labels, aabb, sym, n_sym)
    scene = syn.make_multiview_scene(n_views, n_objects, n_labels, seed=seed, unique_labels=unique_labels)
    infos = pd.DataFrame(dict(view_id=scene['view_ids'], label=[labels[i] for i in scene['label_ids']],
                              score=scene['scores'], scene_id=0, group_id=0,
                              batch_im_id=scene['view_ids']))
    cands = tc.PandasTensorCollection(infos=infos, poses=scene['poses'])
    out_m = multiview_candidate_matching(cands, mesh_db, n_ransac_iter=n_iter, dist_threshold=0.02)
    fc = out_m['filtered_candidates']
    # symmetric distance probe
    rs = np.random.RandomState(5)
    ids1, ids2 = rs.randint(0, len(cands), 64), rs.randint(0, len(cands), 64)
    lab = infos['label'].values[ids1]
    d, _ = symmetric_distance_batched_fast(scene['poses'][ids1], scene['poses'][ids2], lab, mesh_db)
    out = dict(meta=np.array([n_views, n_objects, n_labels, int(unique_labels), n_iter, seed]),
               sym_counts=np.array(sym_counts), chk_poses=checksum(scene['poses']),
               filtered_cand_id=fc.infos['cand_id'].values.astype(np.int64),
               filtered_obj_id=fc.infos['obj_id'].values.astype(np.int64),
               filtered_poses=fc.poses.numpy(),
               scene_obj_id=out_m['scene_infos']['obj_id'].values.astype(np.int64),
               scene_n_cand=out_m['scene_infos']['n_cand'].values.astype(np.int64),
               scene_score=out_m['scene_infos']['score'].values.astype(np.float64),
               pairs_view1=out_m['pairs_TC1C2'].infos['view1'].values.astype(np.int64),
               pairs_view2=out_m['pairs_TC1C2'].infos['view2'].values.astype(np.int64),
               pairs_TC1C2=out_m['pairs_TC1C2'].TC1C2.numpy(),
               symdist_ids1=ids1, symdist_ids2=ids2, symdist=d.numpy())
    np.savez_compressed(HERE / f'{name}.npz', **out)
    print(name, 'matched', len(fc), 'pairs', len(out['pairs_view1']))


def golden_scene_state(cosypose, name, n_views, n_objects, n_labels, sym_counts, n_ransac, ba_n_iter, seed=0):
    """MultiviewScenePredictor.predict_scene_state of the reference (matching + BA + reprojection)."""
    import cosypose.utils.tensor_collection as tc
    from cosypose.integrated.multiview_predictor import MultiviewScenePredictor
    from cosypose.lib3d.mesh_ops import get_meshes_bounding_boxes
    labels = syn.make_labels(n_labels)
    points, sym, n_sym = syn.make_mesh_tables(n_labels, n_points=64, sym_counts=sym_counts)
    mesh_db = ref_mesh_db(cosypose, labels, get_meshes_bounding_boxes(points), sym, n_sym)
    scene = syn.make_multiview_scene(n_views, n_objects, n_labels, seed=seed, unique_labels=True)
    infos = pd.DataFrame(dict(view_id=scene['view_ids'], label=[labels[i] for i in scene['label_ids']],
                              score=scene['scores'], scene_id=0, group_id=0, batch_im_id=scene['view_ids']))
    cands = tc.PandasTensorCollection(infos=infos, poses=scene['poses'])
    cam_infos = pd.DataFrame(dict(view_id=np.arange(n_views), scene_id=0, batch_im_id=np.arange(n_views)))
    cameras = tc.PandasTensorCollection(infos=cam_infos, K=scene['K'], TWC=scene['TWC'])
    pred = MultiviewScenePredictor.__new__(MultiviewScenePredictor)
    pred.mesh_db_ransac = mesh_db
    pred.mesh_db_ba = mesh_db
    out = pred.predict_scene_state(cands, cameras, ransac_n_iter=n_ransac, ba_n_iter=ba_n_iter)
    res = dict(meta=np.array([n_views, n_objects, n_labels, n_ransac, ba_n_iter, seed]),
               sym_counts=np.array(sym_counts), chk_poses=checksum(scene['poses']),
               objects_TWO=out['scene/objects'].TWO.numpy(),
               objects_obj_id=out['scene/objects'].infos['obj_id'].values.astype(np.int64),
               objects_n_cand=out['scene/objects'].infos['n_cand'].values.astype(np.int64),
               cameras_TWC=out['scene/cameras'].TWC.numpy(),
               cameras_view_id=out['scene/cameras'].infos['view_id'].values.astype(np.int64),
               ba_output_poses=out['ba_output'].poses.numpy(),
               ba_output_view_id=out['ba_output'].infos['view_id'].values.astype(np.int64),
               ba_output_obj_id=out['ba_output'].infos['obj_id'].values.astype(np.int64),
               ba_input_poses=out['ba_input'].poses.numpy(),
               n_all=np.array([len(out['ba_output+all_cand'])]))
    np.savez_compressed(HERE / f'{name}.npz', **res)
    print(name, 'objects', len(res['objects_obj_id']), 'cameras', len(res['cameras_view_id']))


def main():
    torch.manual_seed(0)
    cosypose = ref_harness.import_reference()
    torch.set_num_threads(8)
    golden_backbone(cosypose)
    # BASELINE.json config 1: single crop, 1 object, 1 coarse + 1 refine
    golden_single_view(cosypose, 'single_view_cfg1', 1, 1, 3, 1, 1, 64)
    # small multi-chunk case: 2 images x 3 detections, 1 coarse + 2 refine, chunks of 4
    golden_single_view(cosypose, 'single_view_small', 2, 3, 5, 1, 2, 4)
    golden_single_view(cosypose, 'single_view_zup', 1, 2, 3, 1, 1, 64, zup=True)
    golden_multiview(cosypose, 'multiview_small', 4, 6, 8, (1,), True, 50)
    golden_multiview(cosypose, 'multiview_sym', 3, 5, 4, (1, 2, 4, 8), False, 30, seed=3)
    golden_multiview(cosypose, 'multiview_cfg4', 8, 16, 21, (1,), True, 2000)
    golden_scene_state(cosypose, 'scene_state_small', 4, 6, 8, (1,), 50, 3)
    golden_scene_state(cosypose, 'scene_state_sym', 4, 5, 6, (1, 2, 4), 50, 10, seed=2)


if __name__ == '__main__':
    main()
