"""ADD / ADD-S distances with the reference's function signatures (reference: cosypose/lib3d/distances.py:5-21),
computed by `cosyb200_pose_errors` (one CTA per pose pair, the predicted points in shared memory) instead of the
[n, P, P, 3] difference tensor, plus the error statistics of evaluation/meters/pose_meters.py:84-89 in the same pass."""
import torch


def _engine_for(t, engine):
    if engine is not None:
        return engine
    from ..engine import Engine
    key = t.device.index or 0
    if key not in _engine_for.cache:
        _engine_for.cache[key] = Engine(key, max_batch=1)
    return _engine_for.cache[key]


_engine_for.cache = {}


def dists_add(TXO_pred, TXO_gt, points, engine=None):
    """[n, P, 3]: T_gt p - T_pred p."""
    eng = _engine_for(TXO_pred, engine)
    return eng.pose_errors(TXO_pred.float().contiguous(), TXO_gt.float().contiguous(), points.float().contiguous(),
                           return_dists=True)['dists']


def dists_add_symmetric(TXO_pred, TXO_gt, points, engine=None):
    """[n, P, 3]: for every ground-truth point the difference to its closest predicted point (first minimum)."""
    eng = _engine_for(TXO_pred, engine)
    sym = torch.ones(TXO_pred.shape[0], dtype=torch.int32, device=TXO_pred.device)
    return eng.pose_errors(TXO_pred.float().contiguous(), TXO_gt.float().contiguous(), points.float().contiguous(),
                           symmetric=sym, return_dists=True)['dists']


def pose_errors(TXO_pred, TXO_gt, points, is_symmetric=None, engine=None):
    """`PoseErrorMeter.compute_errors` for error types ADD / ADD-S / ADD(-S) (pose_meters.py:53-92): is_symmetric
    None -> ADD for every pair, a bool / int vector -> ADD-S where set.  Returns norm_avg [n], xyz_avg [n,3],
    TCO_xyz [n,3], TCO_norm [n]."""
    eng = _engine_for(TXO_pred, engine)
    sym = None
    if is_symmetric is not None:
        sym = torch.as_tensor(is_symmetric, device=TXO_pred.device).to(torch.int32).contiguous()
    return eng.pose_errors(TXO_pred.float().contiguous(), TXO_gt.float().contiguous(), points.float().contiguous(),
                           symmetric=sym)
