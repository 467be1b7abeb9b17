"""Pins oracle/multiview_oracle.py to the reference's multiview_candidate_matching outputs
(tests/golden/multiview_*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from helpers import Scene


def _scene(g):
    n_views, n_objects, n_labels, unique, n_iter, seed = [int(x) for x in g['meta']]
    return Scene(n_views, n_objects, n_labels, g['sym_counts'], bool(unique), seed), n_iter


@pytest.mark.parametrize('name', ['multiview_small', 'multiview_sym', 'multiview_cfg4'])
def test_matching_oracle_vs_reference(golden_dir, name):
    from oracle import multiview_oracle as mo
    g = np.load(golden_dir / f'{name}.npz')
    sc, n_iter = _scene(g)
    assert float(sc.poses.double().sum()) == pytest.approx(float(g['chk_poses']), rel=1e-9)
    if name == 'multiview_cfg4':
        n_iter = 2000
    out = mo.candidate_matching(sc.view_ids, sc.label_ids, sc.scores, sc.poses, sc.aabb, sc.sym, sc.n_sym,
                                n_ransac_iter=n_iter)
    # integer / index outputs: bit exact
    assert np.array_equal(out['filtered_cand_id'], g['filtered_cand_id'])
    assert np.array_equal(out['filtered_obj_id'], g['filtered_obj_id'])
    assert np.array_equal(out['pairs_view1'], g['pairs_view1'])
    assert np.array_equal(out['pairs_view2'], g['pairs_view2'])
    assert np.array_equal(out['scene_n_cand'], g['scene_n_cand'])
    assert np.allclose(out['scene_score'], g['scene_score'])
    assert np.abs(out['pairs_TC1C2'].numpy() - g['pairs_TC1C2']).max() < 1e-5
    d, _ = mo.symmetric_distance(sc.poses[g['symdist_ids1']], sc.poses[g['symdist_ids2']],
                                 sc.label_ids[g['symdist_ids1']], sc.aabb, sc.sym)
    assert np.abs(d.numpy() - g['symdist']).max() < 1e-6


def test_ba_oracle_jacobian_is_consistent():
    """The autograd Jacobian of the oracle agrees with central finite differences (float64)."""
    from oracle import multiview_oracle as mo
    sc = Scene(3, 4, 5, (1, 2), True, 1)
    n = len(sc.view_ids)
    cand_obj = np.searchsorted(np.unique(sc.label_ids), sc.label_ids)
    TWO_9d = mo.extract_pose9d(sc.TWO[:len(np.unique(sc.label_ids))]).double()
    TCW_9d = mo.extract_pose9d(mo.invert_T(sc.TWC)).double()
    args = (sc.poses.double(), cand_obj, sc.view_ids, sc.label_ids)
    kw = (sc.K.double(), sc.aabb.double(), sc.sym.double(), sc.n_sym)
    e0, loss, J, _ = mo.ba_linearize(*args, TWO_9d, TCW_9d, *kw)
    eps = 1e-6
    for col in (0, 4, 8, TWO_9d.numel() + 2, TWO_9d.numel() + 7):
        dp = torch.zeros(TWO_9d.numel() + TCW_9d.numel(), dtype=torch.float64)
        dp[col] = eps
        ep, *_ = mo.ba_linearize(*args, TWO_9d + dp[:TWO_9d.numel()].view(-1, 9), TCW_9d + dp[TWO_9d.numel():].view(-1, 9), *kw)
        em, *_ = mo.ba_linearize(*args, TWO_9d - dp[:TWO_9d.numel()].view(-1, 9), TCW_9d - dp[TWO_9d.numel():].view(-1, 9), *kw)
        fd = -(ep - em) / (2 * eps)          # errors = y - yhat, J = d yhat
        assert (fd - J[:, col]).abs().max() < 1e-4 * max(1.0, J[:, col].abs().max().item())


def test_ba_oracle_matches_reference_forward_jacobian(golden_dir):
    """oracle/multiview_oracle.ba_linearize against errors / loss / Jacobian of the reference's own
    MultiviewRefinement.forward_jacobian (tests/golden/ba_jacobian_ref.npz, make_golden_r2.golden_ba_jacobian):
    the BA oracle is pinned to the reference, not only to finite differences of itself."""
    import numpy as np
    import torch
    from helpers import Scene
    from oracle import multiview_oracle as mo
    g = np.load(golden_dir / 'ba_jacobian_ref.npz')
    n_views, n_objects, n_labels, n_ransac, seed = [int(x) for x in g['meta']]
    sc = Scene(n_views, n_objects, n_labels, g['sym_counts'], True, seed)
    ci = g['cand_index']
    cand_label = g['obj_label_id'][g['cand_obj']]
    assert np.array_equal(cand_label, np.asarray(sc.label_ids)[ci])
    e, loss, J, _ = mo.ba_linearize(sc.poses[ci], g['cand_obj'], g['cand_view'], cand_label,
                                    torch.from_numpy(g['TWO_9d']), torch.from_numpy(g['TCW_9d']), sc.K, sc.aabb,
                                    sc.sym, sc.n_sym)
    assert e.shape == g['errors'].shape and J.shape == g['J'].shape
    assert np.abs(e.detach().numpy() - g['errors']).max() < 1e-4 * max(1.0, np.abs(g['errors']).max())
    assert abs(loss.item() - float(g['loss'][0])) < 1e-5 * float(g['loss'][0])
    assert np.abs(J.numpy() - g['J']).max() < 1e-5 * np.abs(g['J']).max()
