"""Torch-CPU fp32 restatement of the single-view refinement path.  TEST INFRASTRUCTURE
(see oracle/__init__.py): checker for the CUDA engine and the timed CPU baseline.

Every function cites the reference lines it follows (paths relative to
/root/reference/).  Arithmetic is float32 and uses the same torch operators as the
reference, so on one machine the two agree to rounding; the RoI crop restates the
compiled torchvision operator's rule and is cross-checked against it in
tests/test_oracle_roi_align.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

from cosypose_b200 import effnet_spec as spec

RENDER_SIZE = (spec.RENDER_H, spec.RENDER_W)
N_SAMPLE_POINTS = 2000


# --------------------------------------------------------------------------- meshes

def sample_point_ids(n_points_max, n_points=N_SAMPLE_POINTS):
    """cosypose/lib3d/mesh_ops.py:31-41 with deterministic=True: a fixed subset that
    depends only on the padded point count."""
    return np.random.RandomState(0).choice(n_points_max, size=n_points, replace=False)


def select_points(points_table, label_ids, n_points=N_SAMPLE_POINTS):
    """cosypose/models/pose.py:50-51 (mesh_db.select + sample_points)."""
    ids = torch.as_tensor(sample_point_ids(points_table.shape[1], n_points))
    pts = points_table[torch.as_tensor(np.asarray(label_ids), dtype=torch.long)]
    return torch.index_select(pts, 1, ids)


# --------------------------------------------------------------------------- geometry

def project_points_robust(points_3d, K, TCO, z_min=0.1):
    """cosypose/lib3d/camera_geometry.py:18-31."""
    bsz, n_points = points_3d.shape[:2]
    pts = torch.cat((points_3d, torch.ones(bsz, n_points, 1)), dim=-1)
    P = K @ TCO[:, :3]
    suv = (P.unsqueeze(1) @ pts.unsqueeze(-1)).squeeze(-1)
    z = suv[..., -1]
    suv[..., -1] = torch.max(torch.ones_like(z) * z_min, z)
    suv = suv / suv[..., [-1]]
    return suv[..., :2]


def boxes_from_uv(uv):
    """cosypose/lib3d/camera_geometry.py:34-42."""
    x1 = uv[..., [0]].min(dim=1)[0]
    y1 = uv[..., [1]].min(dim=1)[0]
    x2 = uv[..., [0]].max(dim=1)[0]
    y2 = uv[..., [1]].max(dim=1)[0]
    return torch.cat((x1, y1, x2, y2), dim=1)


def deepim_boxes(rend_center_uv, obs_boxes, rend_boxes, im_size, lamb=1.4):
    """cosypose/lib3d/cropping.py:7-47 (clamp=False)."""
    lobs, robs, uobs, dobs = obs_boxes[:, [0, 2, 1, 3]].t()
    lrend, rrend, urend, drend = rend_boxes[:, [0, 2, 1, 3]].t()
    xc = rend_center_uv[..., 0, 0]
    yc = rend_center_uv[..., 0, 1]
    w, h = max(im_size), min(im_size)
    r = w / h
    xdist = torch.stack(((lobs - xc).abs(), (lrend - xc).abs(),
                         (robs - xc).abs(), (rrend - xc).abs()), dim=1).max(dim=1)[0]
    ydist = torch.stack(((uobs - yc).abs(), (urend - yc).abs(),
                         (dobs - yc).abs(), (drend - yc).abs()), dim=1).max(dim=1)[0]
    width = torch.max(xdist, ydist * r) * 2 * lamb
    height = torch.max(xdist / r, ydist) * 2 * lamb
    return torch.stack((xc - width / 2, yc - height / 2, xc + width / 2, yc + height / 2), dim=1)


def get_K_crop_resize(K, boxes, crop_resize=RENDER_SIZE):
    """cosypose/lib3d/camera_geometry.py:45-87 (orig_size is unused there)."""
    new_K = K.clone()
    final_width, final_height = float(max(crop_resize)), float(min(crop_resize))
    crop_width = boxes[:, 2] - boxes[:, 0]
    crop_height = boxes[:, 3] - boxes[:, 1]
    crop_cj = (boxes[:, 0] + boxes[:, 2]) / 2
    crop_ci = (boxes[:, 1] + boxes[:, 3]) / 2
    cx = K[:, 0, 2] + (crop_width - 1) / 2 - crop_cj
    cy = K[:, 1, 2] + (crop_height - 1) / 2 - crop_ci
    center_x = (crop_width - 1) / 2
    center_y = (crop_height - 1) / 2
    orig_cx_diff = cx - center_x
    orig_cy_diff = cy - center_y
    scale_x = final_width / crop_width
    scale_y = final_height / crop_height
    new_K[:, 0, 0] = scale_x * K[:, 0, 0]
    new_K[:, 1, 1] = scale_y * K[:, 1, 1]
    new_K[:, 0, 2] = (final_width - 1) / 2 + scale_x * orig_cx_diff
    new_K[:, 1, 2] = (final_height - 1) / 2 + scale_y * orig_cy_diff
    return new_K


def roi_align_crop(images, im_ids, boxes, output_size=RENDER_SIZE, sampling_ratio=4):
    """The compiled `torchvision.ops.roi_align(images[im_ids], rois, output_size,
    spatial_scale=1, sampling_ratio=4, aligned=False)` the reference calls at
    cosypose/lib3d/cropping.py:74 (torchvision 0.4.2 pinned in environment.yaml:10;
    algorithm: torchvision csrc/ops/cpu/roi_align_common.h `pre_calc_for_bilinear_interpolate`
    + roi_align_kernel.cpp forward).  Rule per sample point (y, x):
      contributes 0 if y < -1 or y > H or x < -1 or x > W; else clamp to >= 0;
      lo = int(coord); if lo >= size-1: lo = hi = size-1, coord = lo; else hi = lo+1;
      bilinear weights; the bin value is the sum over the 4x4 samples divided by 16.
    `images` is [Nim,C,H,W]; roi b reads image im_ids[b]."""
    n_im, c, H, W = images.shape
    ph_n, pw_n = output_size
    g = sampling_ratio
    out = torch.empty((len(boxes), c, ph_n, pw_n), dtype=torch.float32)
    f32 = torch.float32
    for b in range(len(boxes)):
        x1, y1, x2, y2 = [boxes[b, i].to(f32) for i in range(4)]
        roi_w = torch.max(x2 - x1, torch.tensor(1.0))
        roi_h = torch.max(y2 - y1, torch.tensor(1.0))
        bin_h = roi_h / ph_n
        bin_w = roi_w / pw_n

        def axis(start, bin_size, n_bins, size):
            p = torch.arange(n_bins, dtype=f32)[:, None]
            i = torch.arange(g, dtype=f32)[None, :]
            coord = start + p * bin_size + (i + 0.5) * bin_size / g      # [n_bins, g]
            valid = ~((coord < -1.0) | (coord > size))
            cc = torch.where(coord <= 0, torch.zeros_like(coord), coord)
            lo = cc.to(torch.int64)
            edge = lo >= size - 1
            lo = torch.where(edge, torch.full_like(lo, size - 1), lo)
            hi = torch.where(edge, lo, lo + 1)
            cc = torch.where(edge, lo.to(f32), cc)
            l = cc - lo.to(f32)
            h = 1.0 - l
            lo = lo.clamp(0, size - 1)     # invalid samples (masked below) may index anywhere
            hi = hi.clamp(0, size - 1)
            return valid, lo, hi, l, h

        vy, ylo, yhi, ly, hy = axis(y1, bin_h, ph_n, H)       # [ph, g]
        vx, xlo, xhi, lx, hx = axis(x1, bin_w, pw_n, W)       # [pw, g]
        img = images[int(im_ids[b])]                           # [C,H,W]
        # gather rows then columns: [C, ph, g, pw, g]
        r_lo = img[:, ylo.reshape(-1), :].reshape(c, ph_n, g, W)
        r_hi = img[:, yhi.reshape(-1), :].reshape(c, ph_n, g, W)
        xl, xh = xlo.reshape(-1), xhi.reshape(-1)
        v1 = r_lo[..., xl].reshape(c, ph_n, g, pw_n, g)
        v2 = r_lo[..., xh].reshape(c, ph_n, g, pw_n, g)
        v3 = r_hi[..., xl].reshape(c, ph_n, g, pw_n, g)
        v4 = r_hi[..., xh].reshape(c, ph_n, g, pw_n, g)
        hy_, ly_ = hy[None, :, :, None, None], ly[None, :, :, None, None]
        hx_, lx_ = hx[None, None, None, :, :], lx[None, None, None, :, :]
        val = (hy_ * hx_) * v1 + (hy_ * lx_) * v2 + (ly_ * hx_) * v3 + (ly_ * lx_) * v4
        mask = (vy[None, :, :, None, None] & vx[None, None, None, :, :]).to(f32)
        val = val * mask
        out[b] = val.sum(dim=(2, 4)) / float(g * g)
    return out


# --------------------------------------------------------------------------- trunk

def swish(x):
    """cosypose/models/efficientnet_utils.py:37-57 (forward: x * sigmoid(x))."""
    return x * torch.sigmoid(x)


def _bn(x, sd, prefix):
    """Eval-mode BatchNorm2d, eps=1e-3 (cosypose/models/efficientnet_utils.py:269)."""
    return F.batch_norm(x, sd[f'{prefix}.running_mean'], sd[f'{prefix}.running_var'],
                        sd[f'{prefix}.weight'], sd[f'{prefix}.bias'], False, 0.0, spec.BN_EPS)


def _same_conv(x, w, bias, stride, lo, hi, groups=1):
    """Conv2dStaticSamePadding (cosypose/models/efficientnet_utils.py:123-146): zero pad
    (lo, hi) on both axes - computed for a 300x300 image - then a pad-free conv."""
    if lo or hi:
        x = F.pad(x, (lo, hi, lo, hi))
    return F.conv2d(x, w, bias, stride, 0, 1, groups)


def mbconv(x, sd, b, taps=None):
    """MBConvBlock.forward, eval mode (cosypose/models/efficientnet.py:71-98)."""
    p = f'backbone._blocks.{b.idx}'
    inputs = x
    if b.e != 1:
        x = swish(_bn(F.conv2d(x, sd[f'{p}._expand_conv.weight']), sd, f'{p}._bn0'))
        if taps is not None:
            taps[f'block{b.idx}.expand'] = x
    x = swish(_bn(_same_conv(x, sd[f'{p}._depthwise_conv.weight'], None, b.s, b.pad_lo, b.pad_hi,
                             groups=b.cexp), sd, f'{p}._bn1'))
    if taps is not None:
        taps[f'block{b.idx}.dw'] = x
    sq = F.adaptive_avg_pool2d(x, 1)
    sq = F.conv2d(swish(F.conv2d(sq, sd[f'{p}._se_reduce.weight'], sd[f'{p}._se_reduce.bias'])),
                  sd[f'{p}._se_expand.weight'], sd[f'{p}._se_expand.bias'])
    if taps is not None:
        taps[f'block{b.idx}.gate'] = torch.sigmoid(sq).flatten(1)
    x = torch.sigmoid(sq) * x
    x = _bn(F.conv2d(x, sd[f'{p}._project_conv.weight']), sd, f'{p}._bn2')
    if b.skip:
        x = x + inputs
    return x


def extract_features(x, sd, taps=None):
    """EfficientNet.extract_features (cosypose/models/efficientnet.py:174-190).
    `taps` (optional dict) receives every block-boundary activation, NCHW."""
    x = swish(_bn(_same_conv(x, sd['backbone._conv_stem.weight'], None, 2, *spec.STEM_PAD),
                  sd, 'backbone._bn0'))
    if taps is not None:
        taps['stem'] = x
    for b in spec.BLOCKS:
        x = mbconv(x, sd, b, taps)
        if taps is not None:
            taps[f'block{b.idx}'] = x
    x = swish(_bn(F.conv2d(x, sd['backbone._conv_head.weight']), sd, 'backbone._bn1'))
    if taps is not None:
        taps['head'] = x
    return x


def net_forward(x, sd, taps=None):
    """PosePredictor.net_forward (cosypose/models/pose.py:81-87): mean pool + Linear(1536, 9)."""
    feat = extract_features(x, sd, taps).flatten(2).mean(dim=-1)
    if taps is not None:
        taps['pooled'] = feat
    return F.linear(feat, sd['pose_fc.weight'], sd['pose_fc.bias'])


# --------------------------------------------------------------------------- pose update

def rotation_from_ortho6d(poses):
    """cosypose/lib3d/rotations.py:6-21 - columns (x, y, z)."""
    x_raw, y_raw = poses[..., 0:3], poses[..., 3:6]
    x = x_raw / torch.norm(x_raw, p=2, dim=-1, keepdim=True)
    z = torch.cross(x, y_raw, dim=-1)
    z = z / torch.norm(z, p=2, dim=-1, keepdim=True)
    y = torch.cross(z, x, dim=-1)
    return torch.stack((x, y, z), -1)


def apply_imagespace_predictions(TCO, K, vxvyvz, dRCO):
    """cosypose/lib3d/cosypose_ops.py:10-31."""
    TCO_out = TCO.clone()
    zsrc = TCO[:, 2, [3]]
    vz = vxvyvz[:, [2]]
    ztgt = vz * zsrc
    vxvy = vxvyvz[:, :2]
    fxfy = K[:, [0, 1], [0, 1]]
    xsrcysrc = TCO[:, :2, 3]
    TCO_out[:, 2, 3] = ztgt.flatten()
    TCO_out[:, :2, 3] = ((vxvy / fxfy) + (xsrcysrc / zsrc.repeat(1, 2))) * ztgt.repeat(1, 2)
    TCO_out[:, :3, :3] = dRCO @ TCO[:, :3, :3]
    return TCO_out


def update_pose(TCO, K_crop, pose9):
    """PosePredictor.update_pose, pose_dim == 9 (cosypose/models/pose.py:69-79)."""
    dR = rotation_from_ortho6d(pose9[:, 0:6])
    return apply_imagespace_predictions(TCO, K_crop, pose9[:, 6:9], dR)


def TCO_init_from_boxes(boxes, K, z=1.0):
    """cosypose/lib3d/cosypose_ops.py:121-135 with z_range=(1.0, 1.0)."""
    bsz = boxes.shape[0]
    uv_centers = (boxes[:, [0, 1]] + boxes[:, [2, 3]]) / 2
    zt = torch.full((bsz, 1), z, dtype=boxes.dtype)
    fxfy = K[:, [0, 1], [0, 1]]
    cxcy = K[:, [0, 1], [2, 2]]
    xy_init = ((uv_centers - cxcy) * zt) / fxfy
    TCO = torch.eye(4, dtype=torch.float32).unsqueeze(0).repeat(bsz, 1, 1)
    TCO[:, :2, 3] = xy_init
    TCO[:, 2, 3] = zt.flatten()
    return TCO


def transform_pts(T, pts):
    """cosypose/lib3d/transform_ops.py:7-21 for T [B,4,4]."""
    return (T[:, None, :3, :3] @ pts.unsqueeze(-1)).squeeze(-1) + T[:, None, :3, 3]


def TCO_init_from_boxes_zup_autodepth(boxes_2d, model_points_3d, K):
    """cosypose/lib3d/cosypose_ops.py:138-173."""
    bsz = boxes_2d.shape[0]
    z_guess = 1.0
    fxfy = K[:, [0, 1], [0, 1]]
    cxcy = K[:, [0, 1], [2, 2]]
    TCO = torch.tensor([[0, 1, 0, 0], [0, 0, -1, 0], [-1, 0, 0, z_guess], [0, 0, 0, 1]],
                       dtype=torch.float32).repeat(bsz, 1, 1)
    bb_xy_centers = (boxes_2d[:, [0, 1]] + boxes_2d[:, [2, 3]]) / 2
    TCO[:, :2, 3] = ((bb_xy_centers - cxcy) * z_guess) / fxfy
    C_pts_3d = transform_pts(TCO, model_points_3d)
    deltax_3d = C_pts_3d[:, :, 0].max(dim=1).values - C_pts_3d[:, :, 0].min(dim=1).values
    deltay_3d = C_pts_3d[:, :, 1].max(dim=1).values - C_pts_3d[:, :, 1].min(dim=1).values
    bb_deltax = (boxes_2d[:, 2] - boxes_2d[:, 0]) + 1
    bb_deltay = (boxes_2d[:, 3] - boxes_2d[:, 1]) + 1
    z_from_dx = fxfy[:, 0] * deltax_3d / bb_deltax
    z_from_dy = fxfy[:, 1] * deltay_3d / bb_deltay
    z = (z_from_dy.unsqueeze(1) + z_from_dx.unsqueeze(1)) / 2
    TCO[:, :2, 3] = ((bb_xy_centers - cxcy) * z) / fxfy
    TCO[:, 2, 3] = z.flatten()
    return TCO


# --------------------------------------------------------------------------- iteration loop

def crop_inputs(images, im_ids, K, TCO, points):
    """PosePredictor.crop_inputs (cosypose/models/pose.py:45-67) + deepim_crops_robust
    (cosypose/lib3d/cropping.py:64-75).  K is already gathered per hypothesis."""
    uv = project_points_robust(points, K, TCO)
    boxes_rend = boxes_from_uv(uv)
    center_uv = project_points_robust(torch.zeros(len(K), 1, 3), K, TCO)
    boxes_crop = deepim_boxes(center_uv, boxes_rend, boxes_rend, im_size=images.shape[-2:])
    images_crop = roi_align_crop(images, im_ids, boxes_crop)
    K_crop = get_K_crop_resize(K, boxes_crop)
    return images_crop, K_crop, boxes_rend, boxes_crop


def pose_forward(images, im_ids, K, label_ids, TCO, sd, points_table, render_fn, n_iterations=1,
                 taps=None):
    """PosePredictor.forward (cosypose/models/pose.py:89-132).
    `K` is per image ([Nim,3,3]); `render_fn(iteration_index, TCO_input, K_crop)` returns the
    rendered views [B,3,240,320].  Returns the per-iteration list of output dicts."""
    im_ids_t = torch.as_tensor(np.asarray(im_ids), dtype=torch.long)
    K_ = K[im_ids_t]
    points = select_points(points_table, label_ids)
    outputs = []
    TCO_input = TCO
    with torch.no_grad():
        for n in range(n_iterations):
            images_crop, K_crop, boxes_rend, boxes_crop = crop_inputs(images, im_ids, K_, TCO_input, points)
            renders = render_fn(n, TCO_input, K_crop)
            x = torch.cat((images_crop, renders), dim=1)
            t = None
            if taps is not None:
                t = {}
                taps.append(t)
                t['images_crop'] = images_crop
            pose9 = net_forward(x, sd, t)
            TCO_output = update_pose(TCO_input, K_crop, pose9)
            outputs.append(dict(TCO_input=TCO_input, TCO_output=TCO_output, K_crop=K_crop,
                                pose=pose9, boxes_rend=boxes_rend, boxes_crop=boxes_crop))
            TCO_input = TCO_output
    return outputs


def coarse_refine_predictions(images, K, bboxes, label_ids, im_ids, sd_coarse, sd_refiner,
                              points_table, render_fn, n_coarse_iterations=1,
                              n_refiner_iterations=1, bsz_objects=64, TCO_init=None,
                              init_method='v0'):
    """CoarseRefinePosePredictor.get_predictions (cosypose/integrated/pose_predictor.py:76-107)
    over flat arrays.  `render_fn(stage, iteration_index, chunk_slice, TCO_input, K_crop)`.
    Returns {'coarse/iteration=n' | 'refiner/iteration=n': dict of concatenated tensors}."""
    n = len(label_ids)
    im_ids_t = torch.as_tensor(np.asarray(im_ids), dtype=torch.long)
    preds = {}

    def run(stage, sd, TCO, n_iter):
        chunks = []
        for s in range(0, n, bsz_objects):
            sl = slice(s, min(n, s + bsz_objects))
            outs = pose_forward(images, im_ids[sl], K, label_ids[sl], TCO[sl], sd, points_table,
                                lambda it, T, Kc: render_fn(stage, it, sl, T, Kc), n_iter)
            chunks.append(outs)
        for it in range(n_iter):
            preds[f'{stage}/iteration={it + 1}'] = {
                k: torch.cat([c[it][k] for c in chunks], dim=0) for k in chunks[0][it]}
        return preds[f'{stage}/iteration={n_iter}']['TCO_output']

    if TCO_init is None:
        assert n_coarse_iterations > 0
        if init_method == 'z-up+auto-depth':
            TCO = TCO_init_from_boxes_zup_autodepth(bboxes, select_points(points_table, label_ids), K[im_ids_t])
        else:
            TCO = TCO_init_from_boxes(bboxes, K[im_ids_t])
        TCO = run('coarse', sd_coarse, TCO, n_coarse_iterations)
    else:
        assert n_coarse_iterations == 0
        TCO = TCO_init
    if n_refiner_iterations >= 1:
        TCO = run('refiner', sd_refiner, TCO, n_refiner_iterations)
    return TCO, preds
