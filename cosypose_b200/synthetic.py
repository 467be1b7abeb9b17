"""Seeded synthetic workloads for tests and benchmarks (no network, no datasets).

Nothing here is on the product path: the predictors never import this module.
It builds (a) a pose-model `state_dict` in the reference's naming whose
BatchNorm statistics are *calibrated* so activations stay O(1) through all 26
blocks (PyTorch default init decays to ~1e-13 and makes
parity vacuous), (b) mesh / symmetry tables and (c) the inputs of the
BASELINE.json configurations.

Calibration runs one float64 pass of plain torch convolutions over a U[0,1)
batch and stores each layer's batch statistics as its running statistics
(the recipe of SURVEY.md section 8d).  float64 makes the resulting float32
weights independent of the host's thread count / BLAS summation order, so the
fixtures committed under tests/golden/ stay valid on any box.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import effnet_spec as spec


def _uniform(gen, shape, lo, hi):
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * (hi - lo) + lo)


def _raw_state_dict(seed):
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in spec.state_dict_layout().items():
        if name.endswith('running_mean'):
            t = torch.zeros(shape, dtype=torch.float64)
        elif name.endswith('running_var'):
            t = torch.ones(shape, dtype=torch.float64)
        elif '._bn2.weight' in name and spec.BLOCKS[int(name.split('.')[2])].skip:
            # small residual branches (as in trained nets): keeps the 19-block identity chain
            # from amplifying input-statistics shifts geometrically
            t = _uniform(gen, shape, 0.15, 0.25)
        elif '._bn' in name and name.endswith('.weight'):
            t = _uniform(gen, shape, 0.8, 1.2)
        elif '._bn' in name and name.endswith('.bias'):
            t = _uniform(gen, shape, -0.3, 0.3)
        elif '_se_' in name and name.endswith('.bias'):
            t = _uniform(gen, shape, -0.5, 0.5)
        elif name == 'pose_fc.bias':
            t = torch.tensor([1., 0, 0, 0, 1, 0, 0, 0, 1], dtype=torch.float64)
        elif name == 'pose_fc.weight':
            a = 0.3 / np.sqrt(shape[1])
            t = _uniform(gen, shape, -a, a)
        else:  # convolution weights: unit-gain uniform
            fan_in = int(np.prod(shape[1:]))
            a = np.sqrt(3.0 / fan_in)
            t = _uniform(gen, shape, -a, a)
        sd[name] = t
    return sd


def _swish(x):
    return x * torch.sigmoid(x)


def _pad_conv(x, w, s, lo, hi, groups=1):
    x = F.pad(x, (lo, hi, lo, hi))
    return F.conv2d(x, w, None, s, 0, 1, groups)


def _calibrate(sd, x):
    """Writes batch statistics into running_mean / running_var, layer by layer."""
    def bn(prefix, y):
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        sd[f'{prefix}.running_mean'] = mean
        sd[f'{prefix}.running_var'] = var
        g, b = sd[f'{prefix}.weight'], sd[f'{prefix}.bias']
        scale = g / torch.sqrt(var + spec.BN_EPS)
        return (y - mean[None, :, None, None]) * scale[None, :, None, None] + b[None, :, None, None]

    x = _swish(bn('backbone._bn0', _pad_conv(x, sd['backbone._conv_stem.weight'], 2, *spec.STEM_PAD)))
    for b in spec.BLOCKS:
        p = f'backbone._blocks.{b.idx}'
        inp = x
        if b.e != 1:
            x = _swish(bn(f'{p}._bn0', F.conv2d(x, sd[f'{p}._expand_conv.weight'])))
        x = _swish(bn(f'{p}._bn1', _pad_conv(x, sd[f'{p}._depthwise_conv.weight'],
                                           b.s, b.pad_lo, b.pad_hi, groups=b.cexp)))
        sq = x.mean(dim=(2, 3), keepdim=True)
        sq = _swish(F.conv2d(sq, sd[f'{p}._se_reduce.weight'], sd[f'{p}._se_reduce.bias']))
        sq = F.conv2d(sq, sd[f'{p}._se_expand.weight'], sd[f'{p}._se_expand.bias'])
        x = torch.sigmoid(sq) * x
        x = bn(f'{p}._bn2', F.conv2d(x, sd[f'{p}._project_conv.weight']))
        if b.skip:
            x = x + inp
    x = _swish(bn('backbone._bn1', F.conv2d(x, sd['backbone._conv_head.weight'])))
    return x


def _smooth_field(gen, shape, coarse=8, dtype=torch.float64):
    """U[0,1) noise on a grid `coarse` times smaller, bilinearly upsampled: a low-frequency
    texture, so that crops at any zoom keep similar statistics."""
    *lead, h, w = shape
    small = torch.rand((int(np.prod(lead)), 1, max(2, h // coarse), max(2, w // coarse)),
                       generator=gen, dtype=dtype)
    up = F.interpolate(small, size=(h, w), mode='bilinear', align_corners=False)
    return up.reshape(*lead, h, w)


def _render_like(gen, shape, dtype=torch.float64):
    """A textured object on a zero background, like a rendered view."""
    r = _smooth_field(gen, shape, coarse=4, dtype=dtype)
    *_, h, w = shape
    r[..., :h // 6, :] = 0
    r[..., -(h // 6):, :] = 0
    r[..., :, :w // 5] = 0
    r[..., :, -(w // 5):] = 0
    return r


def make_pose_state_dict(seed=0, calib_batch=4, head_std=0.02):
    """Reference-named float32 `state_dict` with calibrated BatchNorm statistics.

    Calibration data look like the path's real inputs (a smooth observed crop + a view with a
    zero background).  The pose head is then rescaled so that on the calibration batch its
    output is the identity update [1,0,0, 0,1,0, 0,0,1] +- head_std: iterated updates stay near
    the initial pose instead of diverging.  Different seeds stand for coarse / refiner weights."""
    sd = _raw_state_dict(seed)
    gen = torch.Generator().manual_seed(seed + 7919)
    shape = (1, 3, spec.RENDER_H, spec.RENDER_W)
    obs = []
    for i in range(calib_batch):   # crops at several zoom levels, some keeping pixel noise
        o = _smooth_field(gen, shape, coarse=(8, 32, 16, 64)[i % 4])
        if i % 2 == 0:
            o = o + 0.05 * (torch.rand(shape, generator=gen, dtype=torch.float64) - 0.5)
        obs.append(o)
    rend = _render_like(gen, (calib_batch, 3, spec.RENDER_H, spec.RENDER_W))
    x = torch.cat((torch.cat(obs, dim=0).clamp_(0, 1), rend), dim=1)
    with torch.no_grad():
        feat = _calibrate(sd, x).flatten(2).mean(dim=-1)           # [calib, 1536]
        y = feat @ sd['pose_fc.weight'].t()
        scale = head_std / y.std(dim=0, unbiased=False).clamp_min(1e-12)
        sd['pose_fc.weight'] = sd['pose_fc.weight'] * scale[:, None]
        sd['pose_fc.bias'] = sd['pose_fc.bias'] - (feat @ sd['pose_fc.weight'].t()).mean(dim=0)
    return {k: v.to(torch.float32).contiguous() for k, v in sd.items()}


# ---------------------------------------------------------------------------
# meshes and symmetries
# ---------------------------------------------------------------------------

def make_labels(n_labels):
    """Labels follow the reference's BOP naming `obj_%06d`
    (reference: cosypose/datasets/bop_object_datasets.py:13)."""
    return [f'obj_{i + 1:06d}' for i in range(n_labels)]


def _rotz(theta):
    c, s = np.cos(theta), np.sin(theta)
    T = np.eye(4)
    T[:2, :2] = [[c, -s], [s, c]]
    return T


def make_mesh_tables(n_labels, n_points=2500, seed=1, sym_counts=(1,), extent=0.1):
    """points [L, n_points, 3] (metres), symmetries [L, Smax, 4, 4] (identity padded,
    like reference rigid_mesh_database.py:55), n_sym [L].

    Symmetries are rotations about z by 2*pi*k/n, the structure
    `make_bop_symmetries` produces for a continuous axis
    (reference: cosypose/lib3d/symmetries.py:7-35)."""
    rs = np.random.RandomState(seed)
    ext = extent * (0.5 + rs.rand(n_labels, 1, 3))
    points = ((rs.rand(n_labels, n_points, 3) - 0.5) * ext).astype(np.float32)
    n_sym = np.array([sym_counts[i % len(sym_counts)] for i in range(n_labels)], dtype=np.int32)
    smax = int(n_sym.max())
    sym = np.tile(np.eye(4, dtype=np.float32), (n_labels, smax, 1, 1))
    for l in range(n_labels):
        for k in range(n_sym[l]):
            sym[l, k] = _rotz(2 * np.pi * k / n_sym[l]).astype(np.float32)
    return torch.from_numpy(points), torch.from_numpy(sym), n_sym


def make_camera_K(n, fx=600.0, fy=600.0, cx=320.0, cy=240.0):
    K = torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=torch.float32)
    return K[None].repeat(n, 1, 1)


def make_detections(n_images, dets_per_image, n_labels, seed=3, h=480, w=640):
    """bboxes [N,4] (x1,y1,x2,y2 inside the frame, side 60..250 px), label ids, image ids."""
    rs = np.random.RandomState(seed)
    n = n_images * dets_per_image
    side_w = rs.uniform(60, 250, n)
    side_h = rs.uniform(60, 250, n)
    x1 = rs.uniform(0, w - side_w)
    y1 = rs.uniform(0, h - side_h)
    boxes = np.stack([x1, y1, x1 + side_w, y1 + side_h], 1).astype(np.float32)
    label_ids = rs.randint(0, n_labels, n).astype(np.int64)
    im_ids = np.repeat(np.arange(n_images), dets_per_image).astype(np.int64)
    return torch.from_numpy(boxes), label_ids, im_ids


def make_images(n_images, seed=5, h=480, w=640):
    """Observed frames [Nim,3,H,W] in [0,1]: low-frequency texture + a little pixel noise."""
    gen = torch.Generator().manual_seed(seed)
    img = _smooth_field(gen, (n_images, 3, h, w), coarse=8, dtype=torch.float32)
    img = img + 0.05 * (torch.rand((n_images, 3, h, w), generator=gen, dtype=torch.float32) - 0.5)
    return img.clamp_(0, 1).contiguous()


def make_renders(n_iter, n, seed=11):
    """Pre-rendered views [n_iter, N, 3, 240, 320] in [0,1] with a zero background, standing in
    for the renderer the path treats as a side input
    (reference: cosypose/rendering/bullet_batch_renderer.py:46-90)."""
    gen = torch.Generator().manual_seed(seed)
    return _render_like(gen, (n_iter, n, 3, spec.RENDER_H, spec.RENDER_W), dtype=torch.float32).contiguous()


# ---------------------------------------------------------------------------
# multiview scene (BASELINE.json config 4)
# ---------------------------------------------------------------------------

def _rand_rotation(rs):
    q = rs.randn(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def make_multiview_scene(n_views=8, n_objects=16, n_labels=21, seed=0,
                         noise_t=0.002, unique_labels=True):
    """Candidates `inv(TWC) @ TWO` + translation noise for every (view, object).

    Returns dict(view_ids[N], label_ids[N], scores[N], poses[N,4,4], TWO, TWC, K[V,3,3])."""
    rs = np.random.RandomState(seed)
    if unique_labels:
        label_ids = rs.permutation(n_labels)[:n_objects]
    else:
        label_ids = rs.randint(0, n_labels, n_objects)
    TWO = np.tile(np.eye(4), (n_objects, 1, 1))
    for o in range(n_objects):
        TWO[o, :3, :3] = _rand_rotation(rs)
        TWO[o, :3, 3] = rs.randn(3) * 0.3
    TWC = np.tile(np.eye(4), (n_views, 1, 1))
    for v in range(n_views):
        # small random rotation about a random axis, camera ~1 m behind the scene
        R = _rand_rotation(rs)
        ang = np.eye(3) * 0.8 + R * 0.2
        u, _, vt = np.linalg.svd(ang)
        Rv = u @ vt
        if np.linalg.det(Rv) < 0:
            Rv = -Rv
        TWC[v, :3, :3] = Rv
        TWC[v, :3, 3] = rs.randn(3) * 0.1 + np.array([0, 0, -1.0])
    poses, view_ids, labels = [], [], []
    for v in range(n_views):
        TCW = np.linalg.inv(TWC[v])
        for o in range(n_objects):
            T = TCW @ TWO[o]
            T[:3, 3] += rs.randn(3) * noise_t
            poses.append(T)
            view_ids.append(v)
            labels.append(label_ids[o])
    poses = torch.from_numpy(np.stack(poses).astype(np.float32))
    return dict(view_ids=np.array(view_ids, dtype=np.int64),
                label_ids=np.array(labels, dtype=np.int64),
                scores=np.ones(len(view_ids), dtype=np.float32),
                poses=poses,
                TWO=torch.from_numpy(TWO.astype(np.float32)),
                TWC=torch.from_numpy(TWC.astype(np.float32)),
                K=make_camera_K(n_views))


def make_net_input(n, seed=21):
    """A 6-channel network input [n,6,240,320] with the statistics of the real path
    (smooth observed crop + rendered view on a zero background)."""
    gen = torch.Generator().manual_seed(seed)
    shape = (n, 3, spec.RENDER_H, spec.RENDER_W)
    obs = _smooth_field(gen, shape, coarse=16, dtype=torch.float32)
    return torch.cat((obs, _render_like(gen, shape, dtype=torch.float32)), dim=1).contiguous()


def make_icosphere(subdiv=2, radius=0.05):
    """Unit icosahedron subdivided `subdiv` times (20 * 4^subdiv faces), vertices on a sphere of `radius` metres."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    v = [np.asarray(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        mid, nf = {}, []

        def midpoint(a, b):
            key = (min(a, b), max(a, b))
            if key not in mid:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                mid[key] = len(v) - 1
            return mid[key]
        for a, b, c in f:
            ab, bc, ca = midpoint(a, b), midpoint(b, c), midpoint(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return (np.asarray(v) * radius).astype(np.float32), np.asarray(f, dtype=np.int32)


def make_render_meshes(n_labels, seed=2, subdiv=2, extent=0.1):
    """Per label a closed triangle mesh with smooth per-vertex colours: even labels a deformed icosphere
    (many pixel-sized triangles), odd labels a box of 12 triangles (triangles spanning the view), so both
    paths of the rasteriser are exercised.  Returns (vertices, faces, colors) lists."""
    rs = np.random.RandomState(seed)
    verts, faces, cols = [], [], []
    for l in range(n_labels):
        if l % 2 == 0:
            v, f = make_icosphere(subdiv, radius=0.5 * extent)
            scale = rs.uniform(0.6, 1.0, size=3).astype(np.float32)
            bump = 1.0 + 0.15 * np.sin(v @ rs.uniform(-60, 60, size=3).astype(np.float32))
            v = (v * bump[:, None] * scale).astype(np.float32)
        else:
            h = (0.5 * extent * rs.uniform(0.4, 1.0, size=3)).astype(np.float32)
            v = np.asarray([[sx * h[0], sy * h[1], sz * h[2]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)],
                           dtype=np.float32)
            f = np.asarray([(0, 1, 3), (0, 3, 2), (4, 6, 7), (4, 7, 5), (0, 4, 5), (0, 5, 1), (2, 3, 7), (2, 7, 6),
                            (0, 2, 6), (0, 6, 4), (1, 5, 7), (1, 7, 3)], dtype=np.int32)
        c = 0.5 + 0.5 * np.sin(v / extent * rs.uniform(4, 12, size=3) + rs.uniform(0, 6.28, size=3))
        verts.append(v)
        faces.append(f)
        cols.append(c.astype(np.float32))
    return verts, faces, cols
