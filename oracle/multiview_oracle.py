"""Torch-CPU fp32 restatement of multiview candidate matching and object-level bundle adjustment.
TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned against golden vectors produced by the
unmodified reference (tests/golden/multiview_*.npz, scene_state_*.npz; tests/test_oracle_multiview.py).
Paths cited are relative to /root/reference/cosypose/.
"""
from collections import defaultdict

import numpy as np
import torch
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import connected_components

from . import cext_oracle
from .pose_oracle import rotation_from_ortho6d


def invert_T(T):
    """lib3d/transform_ops.py:24-32."""
    out = T.clone()
    Rt = T[..., :3, :3].transpose(-2, -1)
    out[..., :3, :3] = Rt
    out[..., :3, [3]] = -Rt @ T[..., :3, [3]]
    return out


def transform_pts(T, pts):
    """lib3d/transform_ops.py:7-21; T [n,4,4] or [n,S,4,4], pts [n,P,3]."""
    if T.dim() == 4:
        pts = pts.unsqueeze(1)
    return (T.unsqueeze(-3)[..., :3, :3] @ pts.unsqueeze(-1)).squeeze(-1) + T.unsqueeze(-3)[..., :3, 3]


def symmetric_distance(T1, T2, label_ids, points, sym):
    """symmetric_distance_batched_fast, lib3d/symmetric_distances.py:38-57: all (identity padded)
    symmetries of the label; selection by mean squared distance, value = mean of norms."""
    lab = torch.as_tensor(np.asarray(label_ids), dtype=torch.long)
    n = T1.shape[0]
    if n == 0:
        return torch.empty(0), torch.empty(0, dtype=torch.long)
    pts, S = points[lab], sym[lab]
    p1 = transform_pts(T1.unsqueeze(1) @ S, pts)
    p2 = transform_pts(T2, pts).unsqueeze(1)
    d2 = ((p1 - p2) ** 2).sum(dim=-1)
    best = d2.mean(dim=-1).argmin(dim=1)
    return torch.sqrt(d2[torch.arange(n), best]).mean(dim=-1), best


def estimate_camera_poses(poses, label_ids, seeds, points, sym, n_sym):
    """multiview/ransac.py:19-47 over all seeds (dict of int arrays as cext returns)."""
    label_ids = np.asarray(label_ids)
    a, b, g, d = (np.asarray(seeds[k]) for k in ('match1_cand1', 'match1_cand2', 'match2_cand1', 'match2_cand2'))
    if len(a) == 0:
        return torch.zeros((0, 4, 4))
    TC1Oa, TObC2, TC1Og, TC2Od = poses[a], invert_T(poses[b]), poses[g], poses[d]
    lab_ab, lab_gd = label_ids[a], label_ids[g]
    ids_expand = np.repeat(np.arange(len(a)), np.asarray(n_sym)[lab_ab])
    sym_ids = np.concatenate([np.arange(k) for k in np.asarray(n_sym)[lab_ab]])
    S = sym[torch.as_tensor(lab_ab[ids_expand]), torch.as_tensor(sym_ids)]
    dists, _ = symmetric_distance(TC1Og[ids_expand], (TC1Oa[ids_expand] @ S @ TObC2[ids_expand]) @ TC2Od[ids_expand],
                                  lab_gd[ids_expand], points, sym)
    min_ids = cext_oracle.scatter_argmin(dists.numpy(), ids_expand)
    S_star = sym[torch.as_tensor(lab_ab), torch.as_tensor(sym_ids[min_ids])]
    return TC1Oa @ S_star @ TObC2


def score_tmatches(poses, label_ids, tmatches, TC1C2, points, sym):
    """multiview/ransac.py:67-88."""
    c1, c2, h = (np.asarray(tmatches[k]) for k in ('cand1', 'cand2', 'hypothesis_id'))
    d, _ = symmetric_distance(poses[c1], TC1C2[h] @ poses[c2], np.asarray(label_ids)[c1], points, sym)
    return d


def candidate_matching(view_ids, label_ids, scores, poses, points, sym, n_sym, n_ransac_iter=20,
                       dist_threshold=0.02, n_min_inliers=3):
    """multiview_candidate_matching, multiview/ransac.py:137-199 (cameras unknown)."""
    labels = [f'l{int(l)}' for l in label_ids]
    seeds, tm = cext_oracle.make_ransac_infos(list(view_ids), labels, n_ransac_iter, 0)
    TC1C2 = estimate_camera_poses(poses, label_ids, seeds, points, sym, n_sym)
    dists = score_tmatches(poses, label_ids, tm, TC1C2, points, sym)
    inl = cext_oracle.find_ransac_inliers(seeds['view1'], seeds['view2'], tm['hypothesis_id'], tm['cand1'],
                                          tm['cand2'], dists.numpy(), dist_threshold, n_min_inliers)
    n = len(view_ids)
    graph = csr_matrix((np.ones(len(inl['inlier_matches_cand1'])),
                        (inl['inlier_matches_cand1'], inl['inlier_matches_cand2'])), shape=(n, n))
    _, comp = connected_components(graph, directed=True, connection='strong')
    size = np.bincount(comp)[comp]
    keep = np.where(size >= 2)[0]
    _, obj_id = np.unique(comp[keep], return_inverse=True)
    best = inl['best_hypotheses']
    n_obj = int(obj_id.max()) + 1 if len(obj_id) else 0
    scene_score = np.array([np.asarray(scores)[keep][obj_id == o].sum() for o in range(n_obj)])
    scene_n_cand = np.array([(obj_id == o).sum() for o in range(n_obj)])
    return dict(filtered_cand_id=keep, filtered_obj_id=obj_id, filtered_poses=poses[keep],
                pairs_view1=seeds['view1'][best], pairs_view2=seeds['view2'][best], pairs_TC1C2=TC1C2[best],
                scene_score=scene_score, scene_n_cand=scene_n_cand, dists=dists, seeds=seeds, tmatches=tm,
                TC1C2=TC1C2, inliers=inl)


# ------------------------------------------------------------------------------ bundle adjustment

def transform_from_pose9d(p):
    """lib3d/transform_ops.py:53-64."""
    T = torch.zeros(p.shape[:-1] + (4, 4), dtype=p.dtype)
    T[..., :3, :3] = rotation_from_ortho6d(p[..., :6])
    T[..., :3, 3] = p[..., 6:]
    T[..., 3, 3] = 1
    return T


def extract_pose9d(T):
    """multiview/bundle_adjustment.py:159-162."""
    return torch.cat((T[..., :3, :2].transpose(-1, -2).flatten(-2, -1), T[..., :3, -1]), dim=-1)


def project_points(pts, K, T):
    """lib3d/camera_geometry.py:4-15 (no z clamp)."""
    P = K @ T[:, :3]
    ph = torch.cat((pts, torch.ones(pts.shape[:-1] + (1,))), dim=-1)
    suv = (P.unsqueeze(1) @ ph.unsqueeze(-1)).squeeze(-1)
    return (suv / suv[..., [-1]])[..., :2]


def ba_align(cand_TCO, cand_obj, cand_view, cand_label, TWO_9d, TCW_9d, K, points, sym, n_sym):
    """align_TCO_cand, bundle_adjustment.py:164-173 -> symmetric_distance_reprojected
    (lib3d/symmetric_distances.py:105-121): real symmetries only, first minimum."""
    TCO = transform_from_pose9d(TCW_9d)[cand_view] @ transform_from_pose9d(TWO_9d)[cand_obj]
    ns = np.asarray(n_sym)[cand_label]
    ids_expand = np.repeat(np.arange(len(cand_label)), ns)
    sym_ids = np.concatenate([np.arange(k) for k in ns])
    S = sym[torch.as_tensor(cand_label[ids_expand]), torch.as_tensor(sym_ids)]
    pts = points[torch.as_tensor(cand_label[ids_expand])]
    Kx = K[torch.as_tensor(cand_view[ids_expand])]
    a = project_points(pts, Kx, cand_TCO[ids_expand] @ S)
    b = project_points(pts, Kx, TCO[ids_expand])
    d = torch.norm(a - b, dim=-1, p=2).mean(dim=-1)
    min_ids = cext_oracle.scatter_argmin(d.numpy(), ids_expand)
    S_star = sym[torch.as_tensor(cand_label), torch.as_tensor(sym_ids[min_ids])]
    return d[torch.as_tensor(min_ids.astype(np.int64))], cand_TCO @ S_star


def ba_linearize(cand_TCO, cand_obj, cand_view, cand_label, TWO_9d, TCW_9d, K, points, sym, n_sym,
                 residuals_threshold=25.0):
    """forward_jacobian, bundle_adjustment.py:175-214, with autograd over per-residual replicas as the
    reference does.  Returns errors [n_res], loss, dense J [n_res, 9*(n_obj+n_view)], align dists."""
    dists, aligned = ba_align(cand_TCO, cand_obj, cand_view, cand_label, TWO_9d, TCW_9d, K, points, sym, n_sym)
    n_cand, n_pts = len(cand_label), points.shape[1]
    cid = np.repeat(np.arange(n_cand), n_pts * 2)
    pid = np.tile(np.repeat(np.arange(n_pts), 2), n_cand)
    xy = np.tile(np.arange(2), n_cand * n_pts)
    oid, vid = cand_obj[cid], cand_view[cid]
    n_res = len(cid)
    ar = torch.arange(n_res)
    TCW_r = TCW_9d.unsqueeze(0).repeat(n_res, 1, 1).requires_grad_()
    TWO_r = TWO_9d.unsqueeze(0).repeat(n_res, 1, 1).requires_grad_()
    TCO_n = transform_from_pose9d(TCW_r)[ar, torch.as_tensor(vid)] @ transform_from_pose9d(TWO_r)[ar, torch.as_tensor(oid)]
    K_n = K[torch.as_tensor(vid)]
    pts_n = points[torch.as_tensor(cand_label[cid]), torch.as_tensor(pid)].unsqueeze(1)
    yhat = project_points(pts_n, K_n, TCO_n).squeeze(1)[ar, torch.as_tensor(xy)]
    y = project_points(pts_n, K_n, aligned[torch.as_tensor(cid)]).squeeze(1)[ar, torch.as_tensor(xy)]
    errors = (y - yhat).detach()
    loss = torch.min(errors ** 2, torch.ones_like(errors) * residuals_threshold).mean()
    yhat.sum().backward()
    J = torch.cat((TWO_r.grad.flatten(-2, -1), TCW_r.grad.flatten(-2, -1)), dim=-1)
    return errors, loss, J, dists


def ba_optimize(cand_TCO, cand_obj, cand_view, cand_label, TWO_9d, TCW_9d, K, points, sym, n_sym,
                n_iterations=50, optimize_cameras=True, residuals_threshold=25, lambd0=1e-3, L_down=9,
                L_up=11, eps=1e-5):
    """optimize_lm, bundle_adjustment.py:224-278."""
    args = (cand_TCO, cand_obj, cand_view, cand_label)
    n_two = TWO_9d.numel()
    n_params = n_two + TCW_9d.numel()
    prev_update, lambd, done = False, lambd0, False
    history = defaultdict(list)
    for n in range(n_iterations):
        if not prev_update:
            errors, loss, J, _ = ba_linearize(*args, TWO_9d, TCW_9d, K, points, sym, n_sym, residuals_threshold)
        history['loss'].append(float(loss))
        history['lambda'].append(lambd)
        if done:
            break
        A = J.t() @ J + lambd * torch.eye(n_params)
        h = (torch.pinverse(A) @ (J.t() @ errors.view(-1, 1))).flatten()
        TWO_new = TWO_9d + h[:n_two].view(-1, 9)
        TCW_new = TCW_9d + h[n_two:].view(-1, 9) if optimize_cameras else TCW_9d
        errors, next_loss, J, _ = ba_linearize(*args, TWO_new, TCW_new, K, points, sym, n_sym, residuals_threshold)
        rho = loss - next_loss
        if rho.abs() < eps:
            done = True
        elif rho > eps:
            TWO_9d, TCW_9d, loss = TWO_new, TCW_new, next_loss
            lambd = max(lambd / L_down, 1e-7)
            prev_update = True
        else:
            lambd = min(lambd * L_up, 1e7)
            prev_update = False
    return TWO_9d, TCW_9d, history
