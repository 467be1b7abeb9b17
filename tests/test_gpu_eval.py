"""ADD / ADD-S kernel against the reference's formulas (restated with torch on the CPU in
float64), BOP CSV round trip, and the runner-level hypothesis queue against per-group calls."""
import numpy as np
import pandas as pd
import pytest
import torch

from helpers import Workload, build_predictor

pytestmark = pytest.mark.gpu


def _transform(T, p):
    return (T[:, None, :3, :3] @ p[..., None]).squeeze(-1) + T[:, None, :3, 3]


def _ref_add(Tp, Tg, pts):          # lib3d/distances.py:5-9
    return _transform(Tg, pts) - _transform(Tp, pts)


def _ref_adds(Tp, Tg, pts):         # lib3d/distances.py:12-21, line by line
    a, b = _transform(Tp, pts), _transform(Tg, pts)
    d = b.unsqueeze(1) - a.unsqueeze(2)                  # [n, pred j, gt i, 3] = gt_i - pred_j
    assign = (d ** 2).sum(-1).argmin(dim=1)              # over the predicted points, for every ground-truth point
    ids_row = torch.arange(d.shape[0]).unsqueeze(1).repeat(1, d.shape[1])
    ids_col = torch.arange(d.shape[1]).unsqueeze(0).repeat(d.shape[0], 1)
    return d[ids_row, assign, ids_col]


def test_add_and_adds_errors():
    from cosypose_b200.engine import Engine
    from cosypose_b200.lib3d import distances
    eng = Engine(0, max_batch=1)
    dev = eng.device
    gen = torch.Generator().manual_seed(0)
    n, P = 5, 700

    def rand_T():
        q, _ = torch.linalg.qr(torch.randn((n, 3, 3), generator=gen))
        T = torch.eye(4).repeat(n, 1, 1)
        T[:, :3, :3] = q * torch.sign(torch.det(q))[:, None, None]
        T[:, :3, 3] = torch.randn((n, 3), generator=gen) * 0.1 + torch.tensor([0., 0., 1.])
        return T
    Tg = rand_T()
    Tp = Tg.clone()
    Tp[:, :3, 3] += 0.01 * torch.randn((n, 3), generator=gen)
    Tp[:, :3, :3] = Tp[:, :3, :3] @ rand_T()[:, :3, :3]          # a large rotation: ADD-S differs from ADD
    pts = (torch.rand((n, P, 3), generator=gen) - 0.5) * 0.1
    sym = torch.tensor([0, 1, 0, 1, 1], dtype=torch.int32)
    out = distances.pose_errors(Tp.to(dev), Tg.to(dev), pts.to(dev), is_symmetric=sym, engine=eng)
    ref = torch.where(sym[:, None, None].bool(), _ref_adds(Tp.double(), Tg.double(), pts.double()),
                      _ref_add(Tp.double(), Tg.double(), pts.double()))
    assert (out['norm_avg'].cpu().double() - ref.norm(dim=-1).mean(-1)).abs().max() < 1e-6
    assert (out['xyz_avg'].cpu().double() - ref.abs().mean(-2)).abs().max() < 1e-6
    assert (out['TCO_norm'].cpu().double() - (Tp - Tg)[:, :3, 3].double().norm(dim=-1)).abs().max() < 1e-7
    d_add = distances.dists_add(Tp.to(dev), Tg.to(dev), pts.to(dev), engine=eng).cpu().double()
    assert (d_add - _ref_add(Tp.double(), Tg.double(), pts.double())).abs().max() < 1e-6
    d_s = distances.dists_add_symmetric(Tp.to(dev), Tg.to(dev), pts.to(dev), engine=eng).cpu().double()
    r_s = _ref_adds(Tp.double(), Tg.double(), pts.double())
    # an fp32 near-tie may pick another (equally close) ground-truth point: compare the norms, and the vectors where
    # the assignment agrees
    assert (d_s.norm(dim=-1) - r_s.norm(dim=-1)).abs().max() < 1e-6
    assert ((d_s - r_s).abs().amax(-1) < 1e-6).float().mean() > 0.999
    eng.close()


def test_add_distances_vs_reference_golden(golden_dir):
    """dists_add / dists_add_symmetric against the outputs of the reference's own functions (add_distances.npz)."""
    from cosypose_b200.engine import Engine
    from cosypose_b200.lib3d import distances
    g = np.load(golden_dir / 'add_distances.npz')
    eng = Engine(0, max_batch=1)
    dev = eng.device
    Tp, Tg, pts = (torch.from_numpy(g[k]).to(dev) for k in ('T_pred', 'T_gt', 'points'))
    d = distances.dists_add(Tp, Tg, pts, engine=eng).cpu().numpy()
    assert np.abs(d - g['dists_add']).max() < 1e-6
    ds = distances.dists_add_symmetric(Tp, Tg, pts, engine=eng).cpu().numpy()
    assert np.abs(np.linalg.norm(ds, axis=-1) - np.linalg.norm(g['dists_add_symmetric'], axis=-1)).max() < 1e-6
    assert (np.abs(ds - g['dists_add_symmetric']).max(-1) < 1e-6).mean() > 0.999     # fp32 near-ties of the argmin
    eng.close()


def test_hypothesis_queue_equals_per_group_calls():
    """Three view groups refined in one batched call give each group exactly what a call of its own gives."""
    from cosypose_b200.evaluation.hypothesis_queue import HypothesisQueue
    from cosypose_b200.rendering import PreRenderedViews
    from cosypose_b200.utils import tensor_collection as tc
    dev = torch.device('cuda', 0)
    w = Workload(6, 3, 7, 1, 2)                     # 6 frames x 3 detections
    pred, eng, views = build_predictor(w, 0, bsz_objects=8)
    infos = w.infos()
    groups = {('s0', 0): [0, 1], ('s1', 0): [2, 3, 4], ('s2', 1): [5]}   # frames per group

    def group_inputs(frames):
        rows = np.where(np.isin(w.im_ids, frames))[0]
        gi = infos.iloc[rows].reset_index(drop=True).copy()
        gi['batch_im_id'] = gi['batch_im_id'].values - frames[0]
        det = tc.PandasTensorCollection(infos=gi, bboxes=w.boxes[rows].to(dev))
        return rows, w.images[frames].to(dev), w.K[frames].to(dev), det

    def views_for(rows):
        return PreRenderedViews([w.views_c[:, rows].contiguous(), w.views_r[:, rows].contiguous()], 8, device=dev)

    singles = {}
    for key, frames in groups.items():
        rows, im, K, det = group_inputs(frames)
        pred.coarse_model.renderer = pred.refiner_model.renderer = views_for(rows)
        singles[key], _ = pred.get_predictions(im, K, detections=det, n_coarse_iterations=1, n_refiner_iterations=2)
    q = HypothesisQueue(pred, 1, 2)
    all_rows = []
    for key, frames in groups.items():
        rows, im, K, det = group_inputs(frames)
        q.put(key, im, K, det)
        all_rows.append(rows)
    assert len(q) == 18
    pred.coarse_model.renderer = pred.refiner_model.renderer = views_for(np.concatenate(all_rows))
    out = q.flush()
    assert set(out) == set(groups) and len(q) == 0
    for key in groups:
        final, preds = out[key]
        assert list(final.infos['label']) == list(singles[key].infos['label'])
        assert list(final.infos['batch_im_id']) == list(singles[key].infos['batch_im_id'])
        assert (final.poses - singles[key].poses).abs().max() < 1e-5      # chunking changes GEMM tiling only
        assert set(preds) == {'coarse/iteration=1', 'refiner/iteration=1', 'refiner/iteration=2'}
    eng.close()
