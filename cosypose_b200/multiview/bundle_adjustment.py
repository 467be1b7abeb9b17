"""Object-level bundle adjustment (stage 3) with the reference's interface
(cosypose/multiview/bundle_adjustment.py:22-351): `make_view_groups`, `SamplerError`,
`MultiviewRefinement(candidates, cameras, pairs_TC1C2, mesh_db).solve(...)`.

What runs where:
  * every linearisation (symmetry alignment of the candidates, reprojection residuals, analytic
    Jacobian, J^T J and J^T e) is one engine call, `Engine.ba_linearize` (cosyb200_ba_linearize);
    the reference differentiates through [n_residuals, n_objects + n_views, 9] replicated
    parameters with autograd (:175-214);
  * the Levenberg-Marquardt state machine (:224-278) and the chained initialisation (:112-157) are
    host control flow, as in the reference; the damped normal equations are solved on the host
    with a pseudo-inverse, exactly where the reference solves them
    (`torch.pinverse(A.cpu())`, :216-222).
"""
from collections import defaultdict

import numpy as np
import pandas as pd
import torch
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import connected_components

from ..utils import tensor_collection as tc
from ..utils.timer import Timer
from .ransac import make_obj_infos


def make_view_groups(pairs_TC1C2):
    """Views linked by an estimated relative pose form a group (reference: :22-35)."""
    v1, v2 = pairs_TC1C2.infos['view1'].values, pairs_TC1C2.infos['view2'].values
    views, inv = np.unique(np.concatenate((v1, v2)), return_inverse=True)
    n = len(views)
    graph = csr_matrix((np.ones(len(v1)), (inv[:len(v1)], inv[len(v1):])), shape=(n, n))
    _, ids = connected_components(graph, directed=True, connection='strong')
    return pd.DataFrame(dict(view_id=views, view_group=ids))


class SamplerError(Exception):
    pass


def _rot_from_6d(p):
    """numpy float32 version of compute_rotation_matrix_from_ortho6d (lib3d/rotations.py:6-21)."""
    a, b = p[..., 0:3], p[..., 3:6]
    x = a / np.linalg.norm(a, axis=-1, keepdims=True)
    z = np.cross(x, b)
    z = z / np.linalg.norm(z, axis=-1, keepdims=True)
    y = np.cross(z, x)
    return np.stack((x, y, z), axis=-1)


def transform_from_pose9d(p):
    """compute_transform_from_pose9d (lib3d/transform_ops.py:53-64) on a float32 numpy array."""
    p = np.asarray(p, dtype=np.float32)
    T = np.zeros(p.shape[:-1] + (4, 4), dtype=np.float32)
    T[..., :3, :3] = _rot_from_6d(p)
    T[..., :3, 3] = p[..., 6:]
    T[..., 3, 3] = 1
    return T


def extract_pose9d(T):
    """[R[:,0], R[:,1], t] (reference: :159-162)."""
    T = np.asarray(T, dtype=np.float32)
    return np.concatenate((T[..., :3, 0], T[..., :3, 1], T[..., :3, 3]), axis=-1)


def invert_T(T):
    """lib3d/transform_ops.py:24-32 on float32 numpy arrays."""
    T = np.asarray(T, dtype=np.float32)
    out = T.copy()
    Rt = np.swapaxes(T[..., :3, :3], -1, -2)
    out[..., :3, :3] = Rt
    out[..., :3, 3] = -(Rt @ T[..., :3, 3:4])[..., 0]
    return out


class LazyHistory(dict):
    """The optimisation history of `MultiviewRefinement.solve`; the per-iteration scene collections are materialised
    when 'objects' or 'cameras' is first read."""

    def __init__(self, history, problem):
        super().__init__(history)
        self._problem = problem

    def _materialise(self):
        if not dict.__contains__(self, 'objects'):
            pairs = self._problem._scene_infos_many(list(zip(dict.__getitem__(self, 'TWO_9d'),
                                                             dict.__getitem__(self, 'TCW_9d'))))
            dict.__setitem__(self, 'objects', [o for o, _ in pairs])
            dict.__setitem__(self, 'cameras', [c for _, c in pairs])

    def __getitem__(self, key):
        if key in ('objects', 'cameras'):
            self._materialise()
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        return key in ('objects', 'cameras') or dict.__contains__(self, key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def __reduce__(self):       # pickles as the plain, fully materialised dict
        self._materialise()
        return (dict, (dict(dict.items(self)),))

    def keys(self):
        self._materialise()
        return dict.keys(self)

    def items(self):
        self._materialise()
        return dict.items(self)


class MultiviewRefinement:
    def __init__(self, candidates, cameras, pairs_TC1C2, mesh_db):
        self.mesh_db = mesh_db
        self.engine = eng = mesh_db.engine
        self.device = eng.device

        cand_views = candidates.infos['view_id'].to_numpy()
        cand_objs = candidates.infos['obj_id'].to_numpy()
        view_ids = np.unique(cand_views)
        pv1, pv2 = pairs_TC1C2.infos['view1'].to_numpy(), pairs_TC1C2.infos['view2'].to_numpy()
        keep = np.flatnonzero(np.isin(pv1, view_ids) & np.isin(pv2, view_ids))
        cam_keep = np.flatnonzero(np.isin(cameras.infos['view_id'].to_numpy(), view_ids))
        if len(cam_keep) != len(cameras):
            cameras = cameras[cam_keep]

        self.cam_infos = cameras.infos
        self.view_to_id = {v: n for n, v in enumerate(self.cam_infos['view_id'].to_numpy().tolist())}
        self.K = cameras.K.to(self.device, torch.float32).contiguous()
        self.n_views = len(self.cam_infos)

        self.obj_infos = make_obj_infos(candidates)
        self.obj_to_id = {o: n for n, o in enumerate(self.obj_infos['obj_id'].to_numpy().tolist())}
        self.n_objects = len(self.obj_infos)
        self.points = mesh_db.points.to(self.device, torch.float32).contiguous()   # [L, P, 3]
        self.n_points = self.points.shape[1]

        self.cand = candidates
        self.cand_TCO = candidates.poses.to(self.device, torch.float32).contiguous()
        self.cand_view_ids = np.array([self.view_to_id[v] for v in cand_views.tolist()], dtype=np.int32)
        self.cand_obj_ids = np.array([self.obj_to_id[o] for o in cand_objs.tolist()], dtype=np.int32)
        self.n_candidates = len(self.cand_TCO)
        # one upload for the three per-candidate index columns
        idx = np.stack((self.cand_view_ids, self.cand_obj_ids,
                        mesh_db.label_ids(candidates.infos['label'].to_numpy()))).astype(np.int32)
        d_idx = torch.from_numpy(idx).to(self.device)
        self._d_view, self._d_obj, self._d_label = d_idx[0], d_idx[1], d_idx[2]

        self.visibility = np.zeros((self.n_objects, self.n_views), dtype=bool)
        self.visibility[self.cand_obj_ids, self.cand_view_ids] = True

        # one download for the relative camera poses and the candidate poses
        n_pairs = len(keep)
        if n_pairs:
            sel = torch.as_tensor(keep, device=pairs_TC1C2.TC1C2.device)
            both = torch.cat((pairs_TC1C2.TC1C2[sel].to(self.device, torch.float32), self.cand_TCO), dim=0).cpu().numpy()
        else:
            both = self.cand_TCO.cpu().numpy()
        TC1C2, cand_np = both[:n_pairs], both[n_pairs:]
        TC2C1 = invert_T(TC1C2) if n_pairs else TC1C2
        self.v2v1_TC2C1 = {(self.view_to_id[v2], self.view_to_id[v1]): T
                           for v1, v2, T in zip(pv1[keep].tolist(), pv2[keep].tolist(), TC2C1)}
        self.ov_TCO = {(o, v): T for o, v, T in zip(self.cand_obj_ids.tolist(), self.cand_view_ids.tolist(), cand_np)}

    # -- initialisation (reference: :112-157) --------------------------------------------------
    def sample_initial_TWO_TWC(self, seed):
        TWO = np.full((self.n_objects, 4, 4), np.nan, dtype=np.float32)
        TWC = np.full((self.n_views, 4, 4), np.nan, dtype=np.float32)
        rs = np.random.RandomState(seed)
        views_ordered = rs.permutation(np.arange(self.n_views))
        objects_ordered = rs.permutation(np.arange(self.n_objects))

        TWC[views_ordered[0]] = np.eye(4, dtype=np.float32)
        done = {int(views_ordered[0])}
        todo = set(range(self.n_views)) - done
        n_pass = 0
        while todo:
            for v1 in views_ordered:
                if v1 not in todo:
                    continue
                for v2 in views_ordered:
                    if v2 in done and (v2, v1) in self.v2v1_TC2C1:
                        TWC[v1] = TWC[v2] @ self.v2v1_TC2C1[(v2, v1)]
                        todo.remove(v1)
                        done.add(int(v1))
                        break
            n_pass += 1
            if n_pass >= 20:
                raise SamplerError('Cannot find an initialization')
        for o in objects_ordered:
            for v in views_ordered:
                if self.visibility[o, v]:
                    TWO[o] = TWC[v] @ self.ov_TCO[(int(o), int(v))]
                    break
        return TWO, TWC

    def _linearize(self, TWO_9d, TCW_9d, residuals_threshold=25.0, normal_equations=True):
        """Residuals, analytic Jacobian, J^T J, J^T e and the loss in float64 on the device (the reference
        differentiates fp32 residuals with autograd, :175-214, and is itself ~1e-4 from its float64 evaluation)."""
        return self.engine.ba_linearize_f64(self.cand_TCO, self._d_obj, self._d_view, self._d_label,
                                            TWO_9d, TCW_9d, self.K, self.points,
                                            residuals_threshold=residuals_threshold,
                                            normal_equations=normal_equations)

    def align_TCO_cand(self, TWO_9d, TCW_9d):
        out = self._linearize(TWO_9d, TCW_9d, normal_equations=False)
        return out['align_dists'], out['aligned']

    def robust_initialization_TWO_TCW(self, n_init=1):
        best = None
        for n in range(n_init):
            TWO, TWC = self.sample_initial_TWO_TWC(n)
            TWO_9d = torch.from_numpy(extract_pose9d(TWO)).to(self.device)
            TCW_9d = torch.from_numpy(extract_pose9d(invert_T(TWC))).to(self.device)
            dists, _ = self.align_TCO_cand(TWO_9d, TCW_9d)
            score = dists.mean().item()
            if best is None or score < best[0]:
                best = (score, TWO_9d, TCW_9d)
        return best[1], best[2]

    # -- Levenberg-Marquardt (reference: :216-278) -----------------------------------------------
    def compute_lm_step(self, JtJ, Jte, lambd):
        """h = (JtJ + lambda I)^-1 Jte: the reference's `pinverse` on the CPU (:216-222) of a positive definite matrix,
        here a float64 Cholesky solve on the device (cosyb200_lm_solve)."""
        return self.engine.lm_solve(JtJ, Jte, lambd)

    def optimize_lm(self, TWO_9d, TCW_9d, optimize_cameras=True, n_iterations=50, residuals_threshold=25,
                    lambd0=1e-3, L_down=9, L_up=11, eps=1e-5):
        n_two = TWO_9d.numel()
        prev_iter_is_update = False
        lambd, done = lambd0, False
        history = defaultdict(list)
        lin = None
        for n in range(n_iterations):
            if not prev_iter_is_update:
                lin = self._linearize(TWO_9d, TCW_9d, residuals_threshold)
                loss = lin['loss'].item()
            history['TWO_9d'].append(TWO_9d)
            history['TCW_9d'].append(TCW_9d)
            history['loss'].append(loss)
            history['lambda'].append(lambd)
            history['iteration'].append(n)
            if done:
                break
            h = self.compute_lm_step(lin['JtJ'], lin['Jte'], lambd)
            TWO_new = TWO_9d + h[:n_two].view(self.n_objects, 9)
            TCW_new = TCW_9d + h[n_two:].view(self.n_views, 9) if optimize_cameras else TCW_9d
            lin_new = self._linearize(TWO_new, TCW_new, residuals_threshold)
            next_loss = lin_new['loss'].item()
            rho = loss - next_loss
            if abs(rho) < eps:
                done = True
            elif rho > eps:
                TWO_9d, TCW_9d, loss, lin = TWO_new, TCW_new, next_loss, lin_new
                lambd = max(lambd / L_down, 1e-7)
                prev_iter_is_update = True
            else:
                lambd = min(lambd * L_up, 1e7)
                prev_iter_is_update = False   # re-linearised at the kept point on the next pass
        return TWO_9d, TCW_9d, history

    # -- outputs ---------------------------------------------------------------------------------
    def make_scene_infos(self, TWO_9d, TCW_9d):
        return self._scene_infos_many([(TWO_9d, TCW_9d)])[0]

    def _scene_infos_many(self, states):
        """(objects, cameras) collections for a list of (TWO_9d, TCW_9d) states: one download of all parameter
        vectors, the 9-D -> 4x4 conversion on the host (as the reference, :280-293), one upload."""
        n_o, n_v = self.n_objects, self.n_views
        flat = torch.cat([x.reshape(-1, 9) for st in states for x in st], dim=0).cpu().numpy()
        T = transform_from_pose9d(flat)
        per = n_o + n_v
        for k in range(len(states)):
            T[k * per + n_o:(k + 1) * per] = invert_T(T[k * per + n_o:(k + 1) * per])
        d_T = torch.from_numpy(T).to(self.device)
        out = []
        for k in range(len(states)):
            objects = tc.PandasTensorCollection(infos=self.obj_infos, TWO=d_T[k * per:k * per + n_o])
            cameras = tc.PandasTensorCollection(infos=self.cam_infos, TWC=d_T[k * per + n_o:(k + 1) * per], K=self.K)
            out.append((objects, cameras))
        return out

    def convert_history(self, history):
        """`history['objects']` / `history['cameras']` (one collection per LM iteration, :295-303) are only read by
        the reference's visualisation: they are built on first access."""
        return LazyHistory(history, self)

    def solve(self, sample_n_init=1, **lm_kwargs):
        timer_init, timer_opt, timer_misc = Timer(), Timer(), Timer()
        timer_init.start()
        TWO_9d_init, TCW_9d_init = self.robust_initialization_TWO_TCW(n_init=sample_n_init)
        timer_init.pause()
        timer_opt.start()
        TWO_9d_opt, TCW_9d_opt, history = self.optimize_lm(TWO_9d_init, TCW_9d_init, **lm_kwargs)
        timer_opt.pause()
        timer_misc.start()
        (objects, cameras), (objects_init, cameras_init) = self._scene_infos_many(
            [(TWO_9d_opt, TCW_9d_opt), (TWO_9d_init, TCW_9d_init)])
        history = self.convert_history(history)
        timer_misc.pause()
        return dict(objects_init=objects_init, cameras_init=cameras_init, objects=objects, cameras=cameras,
                    history=history, time_init=timer_init.stop(), time_opt=timer_opt.stop(),
                    time_misc=timer_misc.stop())
