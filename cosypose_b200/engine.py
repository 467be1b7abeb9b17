"""Thin Python wrapper over one cosyb200 handle: torch tensors in, raw pointers out.

PyTorch is used for device memory and streams only; every computation below is a call into
libcosyb200.so.  Tensors must be fp32 / int32, contiguous and on the engine's CUDA device.
"""
import ctypes
import os
from ctypes import c_char_p, c_int32, c_int64, c_void_p

import numpy as np
import torch

from . import _lib
from . import effnet_spec as spec

RENDER_H, RENDER_W = spec.RENDER_H, spec.RENDER_W


def _ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(None)


def _np_ptr(a):
    return c_void_p(a.ctypes.data) if a is not None else c_void_p(None)


class Engine:
    def __init__(self, device=None, max_batch=64):
        L = _lib.lib()
        if not torch.cuda.is_available():
            raise _lib.EngineError('cosypose_b200 needs a CUDA device (sm_100a); there is no CPU path')
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device('cuda', device if isinstance(device, int) else torch.device(device).index or 0)
        self.max_batch = int(max_batch)
        h = c_void_p()
        _lib.check(L.cosyb200_create(ctypes.byref(h), self.device.index, self.max_batch), 'create')
        self._h = h
        self._L = L
        self.n_labels = 0
        self.s_max = 0
        self.loaded = [False, False]
        if os.environ.get('COSYB200_GRAPH') == '0':      # profiling runs: plain launches instead of graph replay
            self.set_option('graph', 0)

    def close(self):
        if getattr(self, '_h', None):
            self._L.cosyb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ------------------------------------------------------------------------------
    def _stream(self):
        return c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, t, dtype, shape=None, name='tensor'):
        assert isinstance(t, torch.Tensor), f'{name} must be a tensor'
        assert t.device == self.device, f'{name} is on {t.device}, engine on {self.device}'
        assert t.dtype == dtype, f'{name} must be {dtype}, got {t.dtype}'
        assert t.is_contiguous(), f'{name} must be contiguous'
        if shape is not None:
            assert tuple(t.shape) == tuple(shape), f'{name} has shape {tuple(t.shape)}, expected {tuple(shape)}'
        return t

    def _chk_renders(self, renders, lead):
        """fp32 NCHW [..,3,240,320] -> 0, uint8 NHWC [..,240,320,3] -> 1 (render_u8 flag)."""
        if renders.dtype == torch.uint8:
            self._chk(renders, torch.uint8, tuple(lead) + (RENDER_H, RENDER_W, 3), 'renders (uint8 NHWC)')
            return 1
        self._chk(renders, torch.float32, tuple(lead) + (3, RENDER_H, RENDER_W), 'renders')
        return 0

    def _new(self, *shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # -- setup --------------------------------------------------------------------------------
    def load_pose_model(self, slot, state_dict):
        """state_dict in the reference's naming; int64 counters ignored."""
        items = [(k, v.detach().to('cpu', torch.float32).contiguous())
                 for k, v in state_dict.items() if torch.is_floating_point(v)]
        n = len(items)
        names = (c_char_p * n)(*[k.encode() for k, _ in items])
        ptrs = (c_void_p * n)(*[v.data_ptr() for _, v in items])
        numels = (c_int64 * n)(*[v.numel() for _, v in items])
        _lib.check(self._L.cosyb200_load_pose_model(self._h, slot, n, names, ptrs, numels), 'load_pose_model')
        self.loaded[slot] = True

    def set_meshes(self, points, symmetries, n_sym, aabb=None, point_ids=None):
        """points [L,P,3] (or None for a matching-only engine), symmetries [L,S,4,4] identity
        padded, n_sym [L]; aabb [L,8,3] defaults to the corners of `points`."""
        sym = np.ascontiguousarray(symmetries.detach().cpu().numpy() if isinstance(symmetries, torch.Tensor) else symmetries, dtype=np.float32)
        n_sym = np.ascontiguousarray(n_sym, dtype=np.int32)
        L, S = sym.shape[:2]
        pts = None
        if points is not None:
            pts = np.ascontiguousarray(points.detach().cpu().numpy() if isinstance(points, torch.Tensor) else points, dtype=np.float32)
            assert pts.shape[0] == L and pts.shape[2] == 3
            if point_ids is None:
                point_ids = sample_point_ids(pts.shape[1])
            point_ids = np.ascontiguousarray(point_ids, dtype=np.int64)
            if aabb is None:
                aabb = aabb_corners(pts)
        assert aabb is not None, 'aabb is required when points is None'
        aabb = np.ascontiguousarray(aabb.detach().cpu().numpy() if isinstance(aabb, torch.Tensor) else aabb, dtype=np.float32)
        assert aabb.shape == (L, 8, 3)
        _lib.check(self._L.cosyb200_set_meshes(
            self._h, L, pts.shape[1] if pts is not None else 0, _np_ptr(pts),
            len(point_ids) if pts is not None else 0, _np_ptr(point_ids) if pts is not None else c_void_p(None),
            S, _np_ptr(sym), _np_ptr(n_sym), _np_ptr(aabb)), 'set_meshes')
        self.n_labels, self.s_max = L, S

    # -- device rasteriser ------------------------------------------------------------------
    def set_render_meshes(self, vertices, colors, faces, face_offsets):
        """Triangle meshes of all labels in one table: vertices / colors [Nv,3] (colours in [0,1]), faces [Nf,3]
        indexing the vertex table, face_offsets [L+1] (label l owns faces face_offsets[l]:face_offsets[l+1])."""
        v = np.ascontiguousarray(vertices, dtype=np.float32)
        c = np.ascontiguousarray(colors, dtype=np.float32)
        f = np.ascontiguousarray(faces, dtype=np.int32)
        off = np.ascontiguousarray(face_offsets, dtype=np.int32)
        assert v.ndim == 2 and v.shape[1] == 3 and c.shape == v.shape and f.ndim == 2 and f.shape[1] == 3
        _lib.check(self._L.cosyb200_set_render_meshes(self._h, len(off) - 1, v.shape[0], _np_ptr(v), _np_ptr(c),
                                                      f.shape[0], _np_ptr(f), _np_ptr(off)), 'set_render_meshes')
        self.n_render_labels = len(off) - 1

    def render(self, label_ids, TCO, K, uint8=True, out=None, depth=False):
        """Views of B hypotheses at poses TCO [B,4,4] through intrinsics K [B,3,3]: uint8 [B,240,320,3] or
        float32 [B,3,240,320] in [0,1]; depth=True also returns camera z [B,240,320] (0 = background)."""
        B = TCO.shape[0]
        self._chk(label_ids, torch.int32, (B,), 'label_ids')
        self._chk(TCO, torch.float32, (B, 4, 4), 'TCO')
        self._chk(K, torch.float32, (B, 3, 3), 'K')
        if out is None:
            out = (torch.empty((B, 240, 320, 3), dtype=torch.uint8, device=self.device) if uint8
                   else self._new(B, 3, 240, 320))
        d = self._new(B, 240, 320) if depth else None
        _lib.check(self._L.cosyb200_render(self._h, B, _ptr(label_ids), _ptr(TCO), _ptr(K), _ptr(out), int(uint8),
                                           _ptr(d), self._stream()), 'render')
        return (out, d) if depth else out

    # -- single-view path ---------------------------------------------------------------------
    def tco_init(self, boxes, K, label_ids, zup=False):
        B = boxes.shape[0]
        self._chk(boxes, torch.float32, (B, 4), 'boxes')
        self._chk(K, torch.float32, (B, 3, 3), 'K')
        self._chk(label_ids, torch.int32, (B,), 'label_ids')
        TCO = self._new(B, 4, 4)
        _lib.check(self._L.cosyb200_tco_init(self._h, B, int(zup), _ptr(boxes), _ptr(K), _ptr(label_ids),
                                             _ptr(TCO), self._stream()), 'tco_init')
        return TCO

    def prepare_iter(self, K, TCO, label_ids, img_hw):
        B = K.shape[0]
        self._chk(K, torch.float32, (B, 3, 3), 'K')
        self._chk(TCO, torch.float32, (B, 4, 4), 'TCO')
        self._chk(label_ids, torch.int32, (B,), 'label_ids')
        boxes_rend, boxes_crop, K_crop = self._new(B, 4), self._new(B, 4), self._new(B, 3, 3)
        _lib.check(self._L.cosyb200_prepare_iter(
            self._h, B, int(img_hw[0]), int(img_hw[1]), _ptr(K), _ptr(TCO), _ptr(label_ids),
            _ptr(boxes_rend), _ptr(boxes_crop), _ptr(K_crop), self._stream()), 'prepare_iter')
        return boxes_rend, boxes_crop, K_crop

    def roi_crop(self, images, im_ids, boxes_crop):
        B = boxes_crop.shape[0]
        n_im, c, H, W = images.shape
        assert c == 3
        self._chk(images, torch.float32, name='images')
        self._chk(im_ids, torch.int32, (B,), 'im_ids')
        self._chk(boxes_crop, torch.float32, (B, 4), 'boxes_crop')
        crops = self._new(B, 3, RENDER_H, RENDER_W)
        _lib.check(self._L.cosyb200_roi_crop(self._h, B, _ptr(images), n_im, H, W, _ptr(im_ids),
                                             _ptr(boxes_crop), _ptr(crops), self._stream()), 'roi_crop')
        return crops

    def net_forward(self, slot, crops, renders, taps=False):
        B = crops.shape[0]
        self._chk(crops, torch.float32, (B, 3, RENDER_H, RENDER_W), 'crops')
        self._chk(renders, torch.float32, (B, 3, RENDER_H, RENDER_W), 'renders')
        pose9 = self._new(B, 9)
        tap_arr, tap_tensors = None, None
        if taps:
            shapes = spec.activation_shapes()[1:]   # stem, block0..25, head
            tap_tensors = {name: self._new(B, h, w, c) for name, h, w, c in shapes}
            tap_arr = (c_void_p * len(shapes))(*[tap_tensors[name].data_ptr() for name, *_ in shapes])
        _lib.check(self._L.cosyb200_net_forward(self._h, slot, B, _ptr(crops), _ptr(renders), _ptr(pose9),
                                                tap_arr, self._stream()), 'net_forward')
        return (pose9, tap_tensors) if taps else pose9

    def update_pose(self, TCO, K_crop, pose9):
        B = TCO.shape[0]
        self._chk(TCO, torch.float32, (B, 4, 4), 'TCO')
        self._chk(K_crop, torch.float32, (B, 3, 3), 'K_crop')
        self._chk(pose9, torch.float32, (B, 9), 'pose9')
        out = self._new(B, 4, 4)
        _lib.check(self._L.cosyb200_update_pose(self._h, B, _ptr(TCO), _ptr(K_crop), _ptr(pose9), _ptr(out),
                                                self._stream()), 'update_pose')
        return out

    def refine_iter(self, slot, images, im_ids, boxes_crop, renders, K_crop, TCO):
        B = TCO.shape[0]
        n_im, c, H, W = images.shape
        assert c == 3
        self._chk(images, torch.float32, name='images')
        self._chk(im_ids, torch.int32, (B,), 'im_ids')
        self._chk(boxes_crop, torch.float32, (B, 4), 'boxes_crop')
        u8 = self._chk_renders(renders, (B,))
        self._chk(K_crop, torch.float32, (B, 3, 3), 'K_crop')
        self._chk(TCO, torch.float32, (B, 4, 4), 'TCO')
        pose9, TCO_out = self._new(B, 9), self._new(B, 4, 4)
        _lib.check(self._L.cosyb200_refine_iter(
            self._h, slot, B, _ptr(images), n_im, H, W, _ptr(im_ids), _ptr(boxes_crop), _ptr(renders), u8,
            _ptr(K_crop), _ptr(TCO), _ptr(pose9), _ptr(TCO_out), self._stream()), 'refine_iter')
        return pose9, TCO_out

    def refine_n(self, slot, images, im_ids, K, label_ids, renders, TCO, out=None, n_iter=None):
        """n_iter = renders.shape[0] iterations on pre-rendered views; returns a dict of
        per-iteration tensors (TCO_output, K_crop, boxes_rend, boxes_crop, pose).
        renders=None with n_iter given: every iteration rasterises its own views on the device
        (set_render_meshes first)."""
        if renders is None:
            assert n_iter is not None and n_iter >= 1, 'refine_n without views needs n_iter'
            B = TCO.shape[0]
        else:
            n_iter, B = renders.shape[:2]
        n_im, c, H, W = images.shape
        self._chk(images, torch.float32, name='images')
        self._chk(im_ids, torch.int32, (B,), 'im_ids')
        self._chk(K, torch.float32, (B, 3, 3), 'K')
        self._chk(label_ids, torch.int32, (B,), 'label_ids')
        u8 = self._chk_renders(renders, (n_iter, B)) if renders is not None else 1
        self._chk(TCO, torch.float32, (B, 4, 4), 'TCO')
        if out is None:
            out = dict(TCO_output=self._new(n_iter, B, 4, 4), K_crop=self._new(n_iter, B, 3, 3),
                       boxes_rend=self._new(n_iter, B, 4), boxes_crop=self._new(n_iter, B, 4),
                       pose=self._new(n_iter, B, 9))
        _lib.check(self._L.cosyb200_refine_n(
            self._h, slot, B, n_iter, _ptr(images), n_im, H, W, _ptr(im_ids), _ptr(K), _ptr(label_ids),
            _ptr(renders) if renders is not None else c_void_p(None), u8, _ptr(TCO), _ptr(out['TCO_output']), _ptr(out['K_crop']), _ptr(out['boxes_rend']),
            _ptr(out['boxes_crop']), _ptr(out['pose']), self._stream()), 'refine_n')
        return out

    # -- multi-GPU exchange --------------------------------------------------------------------
    def nccl_init(self):
        """Creates the engine's own NCCL communicator over the ranks of the initialised torch.distributed
        group (the 128-byte unique id travels through that group's object broadcast)."""
        import ctypes
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _lib.check(self._L.cosyb200_nccl_unique_id(buf), 'nccl_unique_id')
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0)
        idbuf = ctypes.create_string_buffer(box[0], 128)
        _lib.check(self._L.cosyb200_nccl_comm_init(self._h, world, rank, idbuf), 'nccl_comm_init')
        self.nccl_ready = True

    def allgather_candidates(self, all_records, local=None):
        """all_records [world * per, rec] float32: every rank's `per` rows are exchanged with one ncclAllGather on
        the current stream; local=None means this rank's rows are already in place (rows rank*per ...)."""
        assert getattr(self, 'nccl_ready', False), 'call nccl_init() first'
        self._chk(all_records, torch.float32, name='all_records')
        import torch.distributed as dist
        world = dist.get_world_size()
        assert all_records.shape[0] % world == 0
        count = all_records.numel() // world
        if local is not None:
            self._chk(local, torch.float32, name='local')
            assert local.numel() == count
        _lib.check(self._L.cosyb200_allgather_candidates(self._h, _ptr(local), _ptr(all_records), count,
                                                         self._stream()), 'allgather_candidates')
        return all_records

    def set_option(self, name, value):
        _lib.check(self._L.cosyb200_set_option(self._h, name.encode(), int(value)), 'set_option')

    def debug_pointwise(self, impl, A, W_nk, bias, gate=None, rows_per_img=1, resid=None, swish=False):
        """One 1x1 convolution C = act((A*gate) @ W^T + bias) (+ resid); impl 0 = CUDA cores, 1 = tcgen05."""
        M, K = A.shape
        N = W_nk.shape[0]
        self._chk(A, torch.float32, (M, K), 'A')
        W = np.ascontiguousarray(W_nk.detach().cpu().numpy() if isinstance(W_nk, torch.Tensor) else W_nk, dtype=np.float32)
        b = np.ascontiguousarray(bias.detach().cpu().numpy() if isinstance(bias, torch.Tensor) else bias, dtype=np.float32)
        assert W.shape == (N, K) and b.shape == (N,)
        if gate is not None:
            self._chk(gate, torch.float32, (-(-M // rows_per_img), K), 'gate')
        if resid is not None:
            self._chk(resid, torch.float32, (M, N), 'resid')
        C = self._new(M, N)
        _lib.check(self._L.cosyb200_debug_pointwise(self._h, int(impl), M, N, K, _ptr(A), _np_ptr(W), _np_ptr(b),
                                                    _ptr(gate), int(rows_per_img), _ptr(resid), int(swish), _ptr(C),
                                                    self._stream()), 'debug_pointwise')
        return C

    # -- launch accounting --------------------------------------------------------------------
    CATEGORIES = ('geometry', 'roi_crop', 'stem', 'expand_1x1', 'depthwise', 'squeeze_excite',
                  'project_1x1', 'head_1x1', 'pool_fc_update', 'ransac', 'render')

    def profile_enable(self, on=True):
        _lib.check(self._L.cosyb200_profile_enable(self._h, int(on)), 'profile_enable')

    def profile_read(self, reset=True):
        """{category: (launches, device_ms)} since the last reset (ms only while profiling)."""
        n = np.zeros(len(self.CATEGORIES), dtype=np.int64)
        ms = np.zeros(len(self.CATEGORIES), dtype=np.float64)
        _lib.check(self._L.cosyb200_profile_read(self._h, int(reset), _np_ptr(n), _np_ptr(ms)), 'profile_read')
        return {c: (int(n[i]), float(ms[i])) for i, c in enumerate(self.CATEGORIES)}

    def profile_read_blocks(self, reset=True):
        """{category: [device_ms per MBConv block (index 31: outside the blocks)]} while profiling."""
        ms = np.zeros((len(self.CATEGORIES), 32), dtype=np.float64)
        _lib.check(self._L.cosyb200_profile_read_blocks(self._h, int(reset), _np_ptr(ms)), 'profile_read_blocks')
        return {c: ms[i].tolist() for i, c in enumerate(self.CATEGORIES)}

    # -- multiview ----------------------------------------------------------------------------
    def ransac_models(self, poses, cand_label_ids, seeds):
        n_seeds = seeds.shape[1]
        self._chk(poses, torch.float32, name='poses')
        self._chk(cand_label_ids, torch.int32, (poses.shape[0],), 'cand_label_ids')
        self._chk(seeds, torch.int32, (6, n_seeds), 'seeds')
        TC1C2 = self._new(n_seeds, 4, 4)
        _lib.check(self._L.cosyb200_ransac_models(self._h, n_seeds, _ptr(poses), _ptr(cand_label_ids),
                                                  _ptr(seeds), _ptr(TC1C2), self._stream()), 'ransac_models')
        return TC1C2

    def ransac_score(self, poses, cand_label_ids, tmatches, TC1C2):
        n = tmatches.shape[1]
        self._chk(poses, torch.float32, name='poses')
        self._chk(cand_label_ids, torch.int32, (poses.shape[0],), 'cand_label_ids')
        self._chk(tmatches, torch.int32, (3, n), 'tmatches')
        self._chk(TC1C2, torch.float32, name='TC1C2')
        dists = self._new(n)
        _lib.check(self._L.cosyb200_ransac_score(self._h, n, _ptr(poses), _ptr(cand_label_ids), _ptr(tmatches),
                                                 _ptr(TC1C2), _ptr(dists), self._stream()), 'ransac_score')
        return dists

    def symmetric_distance(self, T1, T2, label_ids):
        n = T1.shape[0]
        self._chk(T1, torch.float32, (n, 4, 4), 'T1')
        self._chk(T2, torch.float32, (n, 4, 4), 'T2')
        self._chk(label_ids, torch.int32, (n,), 'label_ids')
        dists, best = self._new(n), self._new(n, dtype=torch.int32)
        _lib.check(self._L.cosyb200_symmetric_distance(self._h, n, _ptr(T1), _ptr(T2), _ptr(label_ids),
                                                       _ptr(dists), _ptr(best), self._stream()), 'symmetric_distance')
        return dists, best


    def compose_inv(self, A, B, ia=None, ib=None):
        """inv(A[ia]) @ B[ib] for rigid transforms."""
        self._chk(A, torch.float32, name='A')
        self._chk(B, torch.float32, name='B')
        n = len(ia) if ia is not None else (len(ib) if ib is not None else A.shape[0])
        for idx in (ia, ib):
            if idx is not None:
                self._chk(idx, torch.int32, (n,), 'index')
        out = self._new(n, 4, 4)
        _lib.check(self._L.cosyb200_compose_inv(self._h, n, _ptr(A), _ptr(ia), _ptr(B), _ptr(ib), _ptr(out),
                                                self._stream()), 'compose_inv')
        return out

    def ba_linearize(self, cand_TCO, cand_obj, cand_view, cand_label, TWO_9d, TCW_9d, K, points,
                     residuals_threshold=25.0, normal_equations=True):
        n_cand, n_obj, n_view, n_pts = cand_TCO.shape[0], TWO_9d.shape[0], TCW_9d.shape[0], points.shape[1]
        self._chk(cand_TCO, torch.float32, (n_cand, 4, 4), 'cand_TCO')
        for t, nm in ((cand_obj, 'cand_obj'), (cand_view, 'cand_view'), (cand_label, 'cand_label')):
            self._chk(t, torch.int32, (n_cand,), nm)
        self._chk(TWO_9d, torch.float32, (n_obj, 9), 'TWO_9d')
        self._chk(TCW_9d, torch.float32, (n_view, 9), 'TCW_9d')
        self._chk(K, torch.float32, (n_view, 3, 3), 'K')
        self._chk(points, torch.float32, (points.shape[0], n_pts, 3), 'points')
        n_res, n_params = n_cand * n_pts * 2, 9 * (n_obj + n_view)
        out = dict(align_dists=self._new(n_cand), aligned=self._new(n_cand, 4, 4), errors=self._new(n_res),
                   Jc=self._new(n_res, 18), loss=self._new(1),
                   JtJ=self._new(n_params, n_params) if normal_equations else None,
                   Jte=self._new(n_params) if normal_equations else None)
        _lib.check(self._L.cosyb200_ba_linearize(
            self._h, n_cand, n_obj, n_view, n_pts, _ptr(cand_TCO), _ptr(cand_obj), _ptr(cand_view),
            _ptr(cand_label), _ptr(TWO_9d), _ptr(TCW_9d), _ptr(K), _ptr(points), float(residuals_threshold),
            _ptr(out['align_dists']), _ptr(out['aligned']), _ptr(out['errors']), _ptr(out['Jc']),
            _ptr(out['JtJ']), _ptr(out['Jte']), _ptr(out['loss']), self._stream()), 'ba_linearize')
        return out

    def ransac_inliers_dev(self, seeds_view1, seeds_view2, d_tmatches, dists, dist_threshold, n_min_inliers):
        """find_ransac_inliers on the device: seeds_view1/2 host int arrays (grouped by ordered view pair, as
        ransac_infos emits them), d_tmatches [3, n_mtc] int32 and dists [n_mtc] float32 on the device.  Only the two
        counters and the short ordered result lists are copied back."""
        v1 = np.ascontiguousarray(seeds_view1, dtype=np.int64)
        v2 = np.ascontiguousarray(seeds_view2, dtype=np.int64)
        n_seeds, n_mtc = len(v1), d_tmatches.shape[1]
        if n_seeds == 0 or n_mtc == 0:
            e = np.zeros(0, dtype=np.int32)
            return dict(inlier_matches_cand1=e, inlier_matches_cand2=e.copy(), best_hypotheses=e.copy())
        key = v1 * (int(v2.max()) + 1 if n_seeds else 1) + v2
        assert n_seeds >= 1 and np.all(np.diff(key) >= 0), 'seeds must be grouped by ascending (view1, view2)'
        starts = np.flatnonzero(np.concatenate(([True], np.diff(key) != 0)))
        pair_start = torch.from_numpy(np.concatenate((starts, [n_seeds])).astype(np.int32)).to(self.device)
        n_pairs = len(starts)
        self._chk(d_tmatches, torch.int32, (3, n_mtc), 'tmatches')
        self._chk(dists, torch.float32, (n_mtc,), 'dists')
        o1, o2 = self._new(n_mtc, dtype=torch.int32), self._new(n_mtc, dtype=torch.int32)
        ob = self._new(n_pairs, dtype=torch.int32)
        counts = self._new(2, dtype=torch.int64)
        _lib.check(self._L.cosyb200_ransac_inliers_dev(
            self._h, n_seeds, n_pairs, _ptr(pair_start), n_mtc, _ptr(d_tmatches[0]), _ptr(d_tmatches[1]),
            _ptr(d_tmatches[2]), _ptr(dists), float(dist_threshold), int(n_min_inliers), _ptr(o1), _ptr(o2), _ptr(ob),
            _ptr(counts), self._stream()), 'ransac_inliers_dev')
        no, nb = (int(c) for c in counts.cpu())
        return dict(inlier_matches_cand1=o1[:no].cpu().numpy(), inlier_matches_cand2=o2[:no].cpu().numpy(),
                    best_hypotheses=ob[:nb].cpu().numpy())

    def pose_errors(self, TXO_pred, TXO_gt, points, symmetric=None, return_dists=False):
        """ADD (symmetric[b] == 0) / ADD-S (!= 0) errors of n pose pairs; points [n,P,3] are the model points of each
        pair's label (reference: lib3d/distances.py:5-21, evaluation/meters/pose_meters.py:84-89)."""
        n, P = points.shape[0], points.shape[1]
        self._chk(TXO_pred, torch.float32, (n, 4, 4), 'TXO_pred')
        self._chk(TXO_gt, torch.float32, (n, 4, 4), 'TXO_gt')
        self._chk(points, torch.float32, (n, P, 3), 'points')
        if symmetric is not None:
            self._chk(symmetric, torch.int32, (n,), 'symmetric')
        dists = self._new(n, P, 3) if return_dists else None
        out = dict(norm_avg=self._new(n), xyz_avg=self._new(n, 3), TCO_xyz=self._new(n, 3), TCO_norm=self._new(n))
        _lib.check(self._L.cosyb200_pose_errors(self._h, n, P, _ptr(TXO_pred), _ptr(TXO_gt), _ptr(points), _ptr(symmetric),
                                                _ptr(dists), _ptr(out['norm_avg']), _ptr(out['xyz_avg']),
                                                _ptr(out['TCO_xyz']), _ptr(out['TCO_norm']), self._stream()), 'pose_errors')
        if return_dists:
            out['dists'] = dists
        return out

    def ba_linearize_f64(self, cand_TCO, cand_obj, cand_view, cand_label, TWO_9d, TCW_9d, K, points,
                         residuals_threshold=25.0, normal_equations=True):
        """ba_linearize evaluated in float64 on the device: JtJ / Jte / loss are float64 device tensors."""
        n_cand, n_obj, n_view, n_pts = cand_TCO.shape[0], TWO_9d.shape[0], TCW_9d.shape[0], points.shape[1]
        self._chk(cand_TCO, torch.float32, (n_cand, 4, 4), 'cand_TCO')
        for t, nm in ((cand_obj, 'cand_obj'), (cand_view, 'cand_view'), (cand_label, 'cand_label')):
            self._chk(t, torch.int32, (n_cand,), nm)
        self._chk(TWO_9d, torch.float32, (n_obj, 9), 'TWO_9d')
        self._chk(TCW_9d, torch.float32, (n_view, 9), 'TCW_9d')
        self._chk(K, torch.float32, (n_view, 3, 3), 'K')
        self._chk(points, torch.float32, (points.shape[0], n_pts, 3), 'points')
        n_params = 9 * (n_obj + n_view)
        out = dict(align_dists=self._new(n_cand), aligned=self._new(n_cand, 4, 4),
                   loss=self._new(1, dtype=torch.float64),
                   JtJ=self._new(n_params, n_params, dtype=torch.float64) if normal_equations else None,
                   Jte=self._new(n_params, dtype=torch.float64) if normal_equations else None)
        _lib.check(self._L.cosyb200_ba_linearize_f64(
            self._h, n_cand, n_obj, n_view, n_pts, _ptr(cand_TCO), _ptr(cand_obj), _ptr(cand_view),
            _ptr(cand_label), _ptr(TWO_9d), _ptr(TCW_9d), _ptr(K), _ptr(points), float(residuals_threshold),
            _ptr(out['align_dists']), _ptr(out['aligned']), _ptr(out['JtJ']), _ptr(out['Jte']), _ptr(out['loss']),
            self._stream()), 'ba_linearize_f64')
        return out

    def lm_solve(self, JtJ64, Jte64, lambd):
        """(JtJ + lambda I)^-1 Jte on the device (float64 Cholesky); returns the fp32 step [n]."""
        n = Jte64.shape[0]
        self._chk(JtJ64, torch.float64, (n, n), 'JtJ')
        self._chk(Jte64, torch.float64, (n,), 'Jte')
        step = self._new(n)
        _lib.check(self._L.cosyb200_lm_solve(self._h, n, _ptr(JtJ64), _ptr(Jte64), float(lambd), _ptr(step),
                                             c_void_p(None), self._stream()), 'lm_solve')
        return step


# -- host-side tables ----------------------------------------------------------------------------

def sample_point_ids(n_points_max, n_points=2000):
    """The fixed subset Meshes.sample_points(2000, deterministic=True) selects
    (reference: lib3d/mesh_ops.py:31-41): RandomState(0).choice without replacement."""
    return np.random.RandomState(0).choice(n_points_max, size=n_points, replace=False)


def aabb_corners(points):
    """8 axis-aligned box corners per label in the reference's vertex order
    (reference: lib3d/mesh_ops.py:15-28).  points [L,P,3] -> [L,8,3]."""
    pts = np.asarray(points)
    mn, mx = pts.min(axis=1), pts.max(axis=1)
    sel = [(0, 1, 1), (1, 1, 1), (1, 0, 1), (0, 0, 1), (0, 1, 0), (1, 1, 0), (1, 0, 0), (0, 0, 0)]
    out = np.empty((pts.shape[0], 8, 3), dtype=pts.dtype)
    for i, s in enumerate(sel):
        for a in range(3):
            out[:, i, a] = mx[:, a] if s[a] else mn[:, a]
    return out


# -- host-only integer stages (no GPU needed) ------------------------------------------------------

def ransac_infos(view_ids, label_ids, n_ransac_iter, seed=0):
    L = _lib.lib()
    v = np.ascontiguousarray(view_ids, dtype=np.int32)
    l = np.ascontiguousarray(label_ids, dtype=np.int32)
    assert v.shape == l.shape and v.ndim == 1
    ns, nm = c_int64(0), c_int64(0)
    _lib.check(L.cosyb200_ransac_infos(len(v), _np_ptr(v), _np_ptr(l), int(n_ransac_iter), int(seed),
                                       ctypes.byref(ns), ctypes.byref(nm), None, None), 'ransac_infos')
    seeds = np.empty((6, ns.value), dtype=np.int32)
    tm = np.empty((3, nm.value), dtype=np.int32)
    _lib.check(L.cosyb200_ransac_infos(len(v), _np_ptr(v), _np_ptr(l), int(n_ransac_iter), int(seed),
                                       ctypes.byref(ns), ctypes.byref(nm), _np_ptr(seeds), _np_ptr(tm)), 'ransac_infos')
    return seeds, tm


def ransac_inliers(seeds_view1, seeds_view2, mtc_hyp, mtc_c1, mtc_c2, dists, dist_threshold, n_min_inliers):
    L = _lib.lib()
    a = [np.ascontiguousarray(x, dtype=np.int32) for x in (seeds_view1, seeds_view2, mtc_hyp, mtc_c1, mtc_c2)]
    d = np.ascontiguousarray(dists, dtype=np.float32)
    n_seeds, n_mtc = len(a[0]), len(a[2])
    assert len(d) == n_mtc
    o1 = np.empty(max(n_mtc, 1), dtype=np.int32)
    o2 = np.empty(max(n_mtc, 1), dtype=np.int32)
    ob = np.empty(max(n_seeds, 1), dtype=np.int32)
    no, nb = c_int64(0), c_int64(0)
    _lib.check(L.cosyb200_ransac_inliers(
        n_seeds, _np_ptr(a[0]), _np_ptr(a[1]), n_mtc, _np_ptr(a[2]), _np_ptr(a[3]), _np_ptr(a[4]), _np_ptr(d),
        float(dist_threshold), int(n_min_inliers), _np_ptr(o1), _np_ptr(o2), ctypes.byref(no), _np_ptr(ob),
        ctypes.byref(nb)), 'ransac_inliers')
    return dict(inlier_matches_cand1=o1[:no.value].copy(), inlier_matches_cand2=o2[:no.value].copy(),
                best_hypotheses=ob[:nb.value].copy())


def scatter_argmin(values, group_ids, n_groups=None):
    L = _lib.lib()
    v = np.ascontiguousarray(values, dtype=np.float32)
    g = np.ascontiguousarray(group_ids, dtype=np.int32)
    assert v.shape == g.shape
    if n_groups is None:
        n_groups = int(len(np.unique(g)))
    out = np.empty(max(n_groups, 1), dtype=np.int32)
    _lib.check(L.cosyb200_scatter_argmin(len(v), _np_ptr(v), _np_ptr(g), n_groups, _np_ptr(out)), 'scatter_argmin')
    return out[:n_groups]


def expand_ids_for_symmetry(label_ids, n_sym_per_label):
    L = _lib.lib()
    l = np.ascontiguousarray(label_ids, dtype=np.int32)
    ns = np.ascontiguousarray(n_sym_per_label, dtype=np.int32)
    n = c_int64(0)
    _lib.check(L.cosyb200_expand_ids_for_symmetry(len(l), _np_ptr(l), _np_ptr(ns), ctypes.byref(n), None, None),
               'expand_ids_for_symmetry')
    ids = np.empty(max(n.value, 1), dtype=np.int32)
    sym = np.empty(max(n.value, 1), dtype=np.int32)
    _lib.check(L.cosyb200_expand_ids_for_symmetry(len(l), _np_ptr(l), _np_ptr(ns), ctypes.byref(n),
                                                  _np_ptr(ids), _np_ptr(sym)), 'expand_ids_for_symmetry')
    return ids[:n.value], sym[:n.value]
