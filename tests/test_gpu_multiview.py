"""GPU parity of multiview matching and bundle adjustment: engine kernels (through the C-ABI) vs the
CPU oracle and vs golden vectors from the unmodified reference.  Integer / index outputs must be
bit-exact, distances and poses fp32-close (tolerances in each test)."""
import numpy as np
import pytest
import torch

from helpers import Scene

pytestmark = pytest.mark.gpu


def _scene(g):
    n_views, n_objects, n_labels, unique, n_iter, seed = [int(x) for x in g['meta']]
    return Scene(n_views, n_objects, n_labels, g['sym_counts'], bool(unique), seed), n_iter


def _engine(sc):
    from cosypose_b200.engine import Engine
    eng = Engine(0, max_batch=1)
    mesh_db = sc.mesh_db()
    mesh_db.install(eng, with_points=False)
    return eng, mesh_db


@pytest.mark.parametrize('sym_counts', [(1,), (1, 2, 4, 8), (3, 64, 5)])
def test_symmetric_distance_and_ransac_kernels(sym_counts):
    """All group sizes (1..32 lanes per row), ragged symmetry counts incl. the maximum (64)."""
    from cosypose_b200 import engine as E
    from oracle import cext_oracle, multiview_oracle as mo
    sc = Scene(4, 7, 9, sym_counts, False, 4)
    eng, _ = _engine(sc)
    dev = eng.device
    rs = np.random.RandomState(1)
    n = 301
    i1, i2 = rs.randint(0, len(sc.poses), n), rs.randint(0, len(sc.poses), n)
    lab = sc.label_ids[i1]
    d_ref, b_ref = mo.symmetric_distance(sc.poses[i1], sc.poses[i2], lab, sc.aabb, sc.sym)
    d, b = eng.symmetric_distance(sc.poses[i1].to(dev).contiguous(), sc.poses[i2].to(dev).contiguous(),
                                  torch.as_tensor(lab, dtype=torch.int32, device=dev))
    assert (d.cpu() - d_ref).abs().max() < 2e-6
    # the selected symmetry may differ only between (near-)equivalent ones
    same = b.cpu().long() == b_ref
    assert same.float().mean() > 0.9
    seeds, tm = E.ransac_infos(sc.view_ids, sc.label_ids, 15, 0)
    sd = dict(zip(['view1', 'view2', 'match1_cand1', 'match1_cand2', 'match2_cand1', 'match2_cand2'], seeds))
    T_ref = mo.estimate_camera_poses(sc.poses, sc.label_ids, sd, sc.aabb, sc.sym, sc.n_sym)
    poses = sc.poses.to(dev).contiguous()
    labels = torch.as_tensor(sc.label_ids, dtype=torch.int32, device=dev)
    T = eng.ransac_models(poses, labels, torch.from_numpy(seeds).to(dev))
    # compare through the scores (a different but equivalent symmetry gives a different TC1C2 with the
    # same distance) and directly where the argmin is unambiguous
    tmd = dict(zip(['hypothesis_id', 'cand1', 'cand2'], tm))
    s_ref = mo.score_tmatches(sc.poses, sc.label_ids, tmd, T_ref, sc.aabb, sc.sym)
    s = eng.ransac_score(poses, labels, torch.from_numpy(tm).to(dev), T)
    assert (s.cpu() - s_ref).abs().max() < 5e-6
    if sym_counts == (1,):
        assert (T.cpu() - T_ref).abs().max() < 2e-6


@pytest.mark.parametrize('name', ['multiview_small', 'multiview_sym', 'multiview_cfg4'])
def test_candidate_matching_vs_reference(golden_dir, name):
    from cosypose_b200.multiview.ransac import multiview_candidate_matching
    g = np.load(golden_dir / f'{name}.npz')
    sc, n_iter = _scene(g)
    if name == 'multiview_cfg4':
        n_iter = 2000
    eng, mesh_db = _engine(sc)
    out = multiview_candidate_matching(sc.candidates(eng.device), mesh_db, n_ransac_iter=n_iter, dist_threshold=0.02)
    fc = out['filtered_candidates']
    assert np.array_equal(fc.infos['cand_id'].values, g['filtered_cand_id'])          # bit exact
    assert np.array_equal(fc.infos['obj_id'].values, g['filtered_obj_id'])
    assert np.array_equal(out['pairs_TC1C2'].infos['view1'].values, g['pairs_view1'])
    assert np.array_equal(out['pairs_TC1C2'].infos['view2'].values, g['pairs_view2'])
    assert np.array_equal(out['scene_infos']['n_cand'].values, g['scene_n_cand'])
    assert np.allclose(out['scene_infos']['score'].values, g['scene_score'])
    assert np.abs(fc.poses.cpu().numpy() - g['filtered_poses']).max() == 0
    assert np.abs(out['pairs_TC1C2'].TC1C2.cpu().numpy() - g['pairs_TC1C2']).max() < 1e-5
    if name == 'multiview_cfg4':
        assert out['seeds'].shape == (6, 13440) and out['tmatches'].shape == (3, 215040)
        assert len(fc) == 128 and len(out['pairs_TC1C2']) == 56


def test_known_camera_poses_branch():
    from cosypose_b200.multiview.ransac import multiview_candidate_matching
    sc = Scene(4, 6, 8, (1,), True, 0)
    eng, mesh_db = _engine(sc)
    out = multiview_candidate_matching(sc.candidates(eng.device), mesh_db, cameras=sc.cameras(eng.device),
                                       n_ransac_iter=50)
    # one hypothesis per ordered view pair, TC1C2 = inv(TWC1) @ TWC2
    seeds = out['seeds']
    assert seeds.shape[1] == 12
    ref = torch.linalg.inv(sc.TWC[seeds[0]]) @ sc.TWC[seeds[1]]
    assert (out['TC1C2'].cpu() - ref).abs().max() < 1e-5
    assert len(out['filtered_candidates']) == 24


def test_ba_linearize_vs_autograd_oracle():
    from oracle import multiview_oracle as mo
    sc = Scene(3, 4, 5, (1, 2), True, 1)
    eng, mesh_db = _engine(sc)
    dev = eng.device
    uniq = np.unique(sc.label_ids)
    cand_obj = np.searchsorted(uniq, sc.label_ids).astype(np.int32)
    gen = torch.Generator().manual_seed(0)
    TWO_9d = mo.extract_pose9d(sc.TWO[:len(uniq)]) + 0.01 * torch.randn((len(uniq), 9), generator=gen)
    TCW_9d = mo.extract_pose9d(mo.invert_T(sc.TWC)) + 0.01 * torch.randn((sc.n_views, 9), generator=gen)
    e_ref, loss_ref, J_ref, d_ref = mo.ba_linearize(sc.poses, cand_obj, sc.view_ids, sc.label_ids, TWO_9d, TCW_9d,
                                                    sc.K, sc.aabb, sc.sym, sc.n_sym)
    i32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.int32, device=dev)
    out = eng.ba_linearize(sc.poses.to(dev).contiguous(), i32(cand_obj), i32(sc.view_ids), i32(sc.label_ids),
                           TWO_9d.to(dev), TCW_9d.to(dev), sc.K.to(dev), sc.aabb.to(dev).contiguous())
    scale = e_ref.abs().max().clamp_min(1.0)
    assert (out['errors'].cpu() - e_ref).abs().max() < 2e-3 * scale      # pixels, fp32 projections
    assert abs(out['loss'].item() - loss_ref.item()) < 1e-3 * max(1.0, loss_ref.item())
    assert (out['align_dists'].cpu() - d_ref).abs().max() < 2e-3 * scale
    JtJ_ref, Jte_ref = J_ref.t() @ J_ref, J_ref.t() @ e_ref
    assert (out['JtJ'].cpu() - JtJ_ref).abs().max() < 1e-4 * JtJ_ref.abs().max()
    assert (out['Jte'].cpu() - Jte_ref).abs().max() < 1e-3 * Jte_ref.abs().max().clamp_min(1.0)
    # compact rows against the dense Jacobian
    n_obj = len(uniq)
    rows = out['Jc'].cpu()
    cid = np.repeat(np.arange(len(sc.label_ids)), 16)
    for r in (0, 17, 101, len(cid) - 1):
        o, v = cand_obj[cid[r]], sc.view_ids[cid[r]]
        dense = torch.cat((J_ref[r, 9 * o:9 * o + 9], J_ref[r, 9 * (n_obj + v):9 * (n_obj + v) + 9]))
        assert (rows[r] - dense).abs().max() < 1e-4 * dense.abs().max().clamp_min(1.0)


@pytest.mark.parametrize('case', ['cfg4', 'ties', 'repeated_labels'])
def test_device_voting_matches_host_bit_for_bit(case):
    """cosyb200_ransac_inliers_dev against the host implementation (itself bit-exact vs the reference's compiled
    extension, tests/test_abi.py): same inlier lists, same best hypotheses, same order - including exact distance
    ties (stable order), distances on the threshold, NaNs, and a best hypothesis with id 0 (dropped by both)."""
    from cosypose_b200 import engine as E
    rs = np.random.RandomState(11)
    if case == 'cfg4':
        sc = Scene(8, 16, 21, (1,), True, 0)
        view_ids, label_ids = np.asarray(sc.view_ids, dtype=np.int32), np.asarray(sc.label_ids, dtype=np.int32)
        n_iter = 2000
    elif case == 'ties':
        view_ids = np.repeat(np.arange(4), 6).astype(np.int32)
        label_ids = np.tile(np.arange(6), 4).astype(np.int32)
        n_iter = 30
    else:
        view_ids = np.repeat(np.arange(3), 12).astype(np.int32)
        label_ids = rs.randint(0, 3, size=36).astype(np.int32)       # many tentative matches per view pair
        n_iter = 200
    seeds, tmatches = E.ransac_infos(view_ids, label_ids, n_iter, 0)
    n_mtc = tmatches.shape[1]
    d = rs.uniform(0.0, 0.04, size=n_mtc).astype(np.float32)
    if case != 'cfg4':
        d = (np.round(d * 250) / 250).astype(np.float32)             # heavy exact ties, many values == 0.02
        d[rs.randint(0, n_mtc, size=max(1, n_mtc // 50))] = np.nan
    d[tmatches[0] == 0] = 0.001                                      # hypothesis 0 is the best of its pair
    eng = E.Engine(0, max_batch=1)
    ref = E.ransac_inliers(seeds[0], seeds[1], tmatches[0], tmatches[1], tmatches[2], d, 0.02, 3)
    out = eng.ransac_inliers_dev(seeds[0], seeds[1], torch.from_numpy(tmatches).to(eng.device),
                                 torch.from_numpy(d).to(eng.device), 0.02, 3)
    assert len(ref['best_hypotheses']) > 0
    for k in ('inlier_matches_cand1', 'inlier_matches_cand2', 'best_hypotheses'):
        assert np.array_equal(ref[k], out[k]), k
    eng.close()


def test_ba_linearize_f64_and_lm_solve():
    """float64 linearisation on the device against the autograd oracle evaluated in float64 (1e-9 relative), and the
    device Cholesky solve against numpy's float64 solve of the same damped system."""
    from oracle import multiview_oracle as mo
    sc = Scene(3, 4, 5, (1, 2), True, 1)
    eng, mesh_db = _engine(sc)
    dev = eng.device
    uniq = np.unique(sc.label_ids)
    cand_obj = np.searchsorted(uniq, sc.label_ids).astype(np.int32)
    gen = torch.Generator().manual_seed(0)
    TWO_9d = mo.extract_pose9d(sc.TWO[:len(uniq)]) + 0.01 * torch.randn((len(uniq), 9), generator=gen)
    TCW_9d = mo.extract_pose9d(mo.invert_T(sc.TWC)) + 0.01 * torch.randn((sc.n_views, 9), generator=gen)
    d = lambda t: t.double()
    e_ref, loss_ref, J_ref, _ = mo.ba_linearize(d(sc.poses), cand_obj, sc.view_ids, sc.label_ids, d(TWO_9d), d(TCW_9d),
                                                d(sc.K), d(sc.aabb), d(sc.sym), sc.n_sym)
    assert J_ref.dtype == torch.float64
    i32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.int32, device=dev)
    out = eng.ba_linearize_f64(sc.poses.to(dev).contiguous(), i32(cand_obj), i32(sc.view_ids), i32(sc.label_ids),
                               TWO_9d.to(dev), TCW_9d.to(dev), sc.K.to(dev), sc.aabb.to(dev).contiguous())
    JtJ_ref, Jte_ref = J_ref.t() @ J_ref, J_ref.t() @ e_ref
    assert out['JtJ'].dtype == torch.float64
    # the aligned candidate poses are fp32 data (cand_TCO @ S), everything after them is float64
    assert (out['JtJ'].cpu() - JtJ_ref).abs().max() < 1e-9 * JtJ_ref.abs().max()
    assert (out['Jte'].cpu() - Jte_ref).abs().max() < 1e-5 * Jte_ref.abs().max().clamp_min(1.0)
    assert abs(out['loss'].item() - loss_ref.item()) < 1e-5 * max(1.0, loss_ref.item())
    for lambd in (1e-3, 1e-7):
        h = eng.lm_solve(out['JtJ'], out['Jte'], lambd).cpu().double().numpy()
        A = out['JtJ'].cpu().numpy() + lambd * np.eye(JtJ_ref.shape[0])
        b = out['Jte'].cpu().numpy()
        # backward error of the returned (fp32-rounded) step; the system is nearly singular along the gauge directions
        # for small lambda, so two float64 solvers only agree in the residual there, not in h itself
        assert np.linalg.norm(A @ h - b) < 1e-6 * (np.linalg.norm(A, 2) * np.linalg.norm(h) + np.linalg.norm(b)), lambd
        if lambd == 1e-3:
            h_ref = np.linalg.solve(A, b)
            assert np.abs(h - h_ref).max() < 1e-5 * max(1.0, np.abs(h_ref).max()), np.abs(h - h_ref).max()


@pytest.mark.parametrize('name', ['scene_state_small', 'scene_state_sym'])
def test_predict_scene_state_vs_reference(golden_dir, name):
    """MultiviewScenePredictor.predict_scene_state end to end (matching + view groups + LM bundle
    adjustment + reprojection): integer outputs bit-exact, gauge-invariant poses within 1e-4 of the reference."""
    from cosypose_b200.integrated.multiview_predictor import MultiviewScenePredictor
    g = np.load(golden_dir / f'{name}.npz')
    n_views, n_objects, n_labels, n_ransac, ba_n_iter, seed = [int(x) for x in g['meta']]
    sc = Scene(n_views, n_objects, n_labels, g['sym_counts'], True, seed)
    pred = MultiviewScenePredictor(sc.mesh_db(), device=0)
    out = pred.predict_scene_state(sc.candidates(pred.engine.device), sc.cameras(pred.engine.device),
                                   ransac_n_iter=n_ransac, ba_n_iter=ba_n_iter)
    assert set(out) == {'cand_inputs', 'cand_matched', 'scene/objects', 'scene/cameras', 'ba_output', 'ba_input',
                        'ba_output+all_cand'}
    assert np.array_equal(out['scene/objects'].infos['obj_id'].values, g['objects_obj_id'])
    assert np.array_equal(out['scene/objects'].infos['n_cand'].values, g['objects_n_cand'])
    assert np.array_equal(out['scene/cameras'].infos['view_id'].values, g['cameras_view_id'])
    assert np.array_equal(out['ba_output'].infos['view_id'].values, g['ba_output_view_id'])
    assert np.array_equal(out['ba_output'].infos['obj_id'].values, g['ba_output_obj_id'])
    assert len(out['ba_output+all_cand']) == int(g['n_all'][0])
    assert np.abs(out['ba_input'].poses.cpu().numpy() - g['ba_input_poses']).max() < 1e-5
    # The world frame is a gauge freedom of the problem: the damped step moves it along a (numerically) null
    # direction, so TWO / TWC themselves are not reproducible even by the reference across thread counts (measured:
    # 9e-3 between 1 and 8 CPU threads).  Every gauge-invariant quantity is compared: object-in-camera poses
    # (ba_output), camera-to-camera transforms, objects in the first camera's frame.
    # The LM normal equations are ill conditioned in fp32: the reference run in float64 (scene_state_*_fp64.npz,
    # the same reference code on float64 tensors) is `ref32_minus_fp64` (1.1e-4 / 1.1e-5) away from the fp32
    # reference, so the engine (float64 linearisation + solve on the device) is held to 1e-4 against the float64
    # reference and to 1e-4 + that distance against the fp32 golden.
    g64 = np.load(golden_dir / f'{name}_fp64.npz')
    ours = out['ba_output'].poses.cpu().numpy().astype(np.float64)
    assert np.array_equal(out['ba_output'].infos['view_id'].values, g64['ba_output_view_id'])
    d64 = np.abs(ours - g64['ba_output_poses']).max()
    d32 = np.abs(ours - g['ba_output_poses']).max()
    print(name, 'ba_output vs reference fp64', d64, 'vs reference fp32', d32, 'ref32 vs ref64', float(g64['ref32_minus_fp64'][0]))
    assert d64 < 1e-4
    assert d32 < 1e-4 + float(g64['ref32_minus_fp64'][0])
    tol = 1e-4 + float(g64['ref32_minus_fp64'][0]) * 3
    TWC, TWC_g = out['scene/cameras'].TWC.cpu().numpy(), g['cameras_TWC']
    rel = np.linalg.inv(TWC[:1]) @ TWC
    rel_g = np.linalg.inv(TWC_g[:1]) @ TWC_g
    assert np.abs(rel - rel_g).max() < tol
    TWO, TWO_g = out['scene/objects'].TWO.cpu().numpy(), g['objects_TWO']
    assert np.abs(np.linalg.inv(TWC[:1]) @ TWO - np.linalg.inv(TWC_g[:1]) @ TWO_g).max() < tol
    # BA must not move the scene away from the ground truth it was generated from
    err_in = np.abs(g['ba_input_poses'][:, :3, 3] - out['ba_output'].poses.cpu().numpy()[:, :3, 3]).max()
    assert err_in < 0.1
