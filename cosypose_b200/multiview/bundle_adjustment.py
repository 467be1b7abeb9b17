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


class MultiviewRefinement:
    def __init__(self, candidates, cameras, pairs_TC1C2, mesh_db):
        self.mesh_db = mesh_db
        self.engine = eng = mesh_db.engine
        self.device = eng.device

        view_ids = np.unique(candidates.infos['view_id'])
        keep = np.isin(pairs_TC1C2.infos['view1'], view_ids) & np.isin(pairs_TC1C2.infos['view2'], view_ids)
        pairs_TC1C2 = pairs_TC1C2[np.where(keep)[0]]
        cameras = cameras[np.where(np.isin(cameras.infos['view_id'], view_ids))[0]]

        self.cam_infos = cameras.infos
        self.view_to_id = {v: n for n, v in enumerate(self.cam_infos['view_id'])}
        self.K = cameras.K.to(self.device, torch.float32).contiguous()
        self.n_views = len(self.cam_infos)

        self.obj_infos = make_obj_infos(candidates)
        self.obj_to_id = {o: n for n, o in enumerate(self.obj_infos['obj_id'])}
        self.n_objects = len(self.obj_infos)
        self.points = mesh_db.points.to(self.device, torch.float32).contiguous()   # [L, P, 3]
        self.n_points = self.points.shape[1]

        self.cand = candidates
        self.cand_TCO = candidates.poses.to(self.device, torch.float32).contiguous()
        self.cand_view_ids = np.array([self.view_to_id[v] for v in candidates.infos['view_id']], dtype=np.int32)
        self.cand_obj_ids = np.array([self.obj_to_id[o] for o in candidates.infos['obj_id']], dtype=np.int32)
        self.n_candidates = len(self.cand_TCO)
        self._d_view = torch.from_numpy(self.cand_view_ids).to(self.device)
        self._d_obj = torch.from_numpy(self.cand_obj_ids).to(self.device)
        self._d_label = torch.from_numpy(mesh_db.label_ids(candidates.infos['label'].values)).to(self.device)

        self.visibility = np.zeros((self.n_objects, self.n_views), dtype=bool)
        self.visibility[self.cand_obj_ids, self.cand_view_ids] = True

        TC1C2 = pairs_TC1C2.TC1C2.detach().cpu().numpy() if len(pairs_TC1C2) else np.zeros((0, 4, 4), np.float32)
        self.v2v1_TC2C1 = {(self.view_to_id[v2], self.view_to_id[v1]): invert_T(T)
                           for v1, v2, T in zip(pairs_TC1C2.infos['view1'], pairs_TC1C2.infos['view2'], TC1C2)}
        cand_np = self.cand_TCO.cpu().numpy()
        self.ov_TCO = {(int(o), int(v)): T for o, v, T in zip(self.cand_obj_ids, self.cand_view_ids, cand_np)}

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
        TWO = transform_from_pose9d(TWO_9d.cpu().numpy())
        TWC = invert_T(transform_from_pose9d(TCW_9d.cpu().numpy()))
        objects = tc.PandasTensorCollection(infos=self.obj_infos, TWO=torch.from_numpy(TWO).to(self.device))
        cameras = tc.PandasTensorCollection(infos=self.cam_infos, TWC=torch.from_numpy(TWC).to(self.device), K=self.K)
        return objects, cameras

    def convert_history(self, history):
        history['objects'], history['cameras'] = [], []
        for TWO_9d, TCW_9d in zip(history['TWO_9d'], history['TCW_9d']):
            objects, cameras = self.make_scene_infos(TWO_9d, TCW_9d)
            history['objects'].append(objects)
            history['cameras'].append(cameras)
        return history

    def solve(self, sample_n_init=1, **lm_kwargs):
        timer_init, timer_opt, timer_misc = Timer(), Timer(), Timer()
        timer_init.start()
        TWO_9d_init, TCW_9d_init = self.robust_initialization_TWO_TCW(n_init=sample_n_init)
        timer_init.pause()
        timer_opt.start()
        TWO_9d_opt, TCW_9d_opt, history = self.optimize_lm(TWO_9d_init, TCW_9d_init, **lm_kwargs)
        timer_opt.pause()
        timer_misc.start()
        objects, cameras = self.make_scene_infos(TWO_9d_opt, TCW_9d_opt)
        objects_init, cameras_init = self.make_scene_infos(TWO_9d_init, TCW_9d_init)
        history = self.convert_history(history)
        timer_misc.pause()
        return dict(objects_init=objects_init, cameras_init=cameras_init, objects=objects, cameras=cameras,
                    history=history, time_init=timer_init.stop(), time_opt=timer_opt.stop(),
                    time_misc=timer_misc.stop())
