"""CPU restatement of the device rasteriser (TEST INFRASTRUCTURE ONLY: imported by tests/, smoke() and nothing else).

PARITY UNPINNED against the reference renderer: the reference draws its views with pybullet's OpenGL renderer
(rendering/bullet_batch_renderer.py:46-90 -> bullet_scene_renderer.py:38-61 -> simulator/camera.py:168-180), and
pybullet is not installable in this image, so no frame of the reference exists to compare with.  What this file
pins is the camera model the reference builds (simulator/camera.py:10-34: pinhole from K with skew, window [0,w]x[0,h]
so pixel (i, j) is sampled at (j + 0.5, i + 0.5); near plane 0.01, camera.py:45; background 0,
bullet_scene_renderer.py:49) and the arithmetic of cosypose_b200/csrc/kernels_raster.cuh, operation for
operation in float32 (every numpy op below is one IEEE single-precision rounding, as every __f*_rn there), so
that the CUDA frames can be checked bit for bit.
"""
import numpy as np

F = np.float32
NEAR_Z = F(0.01)
H, W = 240, 320


def _edge(ax, ay, bx, by, px, py):
    """(b - a) x (p - a) with the end points in lexicographic order (kernels_raster.cuh: edge_fn)."""
    swap = (bx < ax) or (bx == ax and by < ay)
    sx, sy, ex, ey = (bx, by, ax, ay) if swap else (ax, ay, bx, by)
    e = (ex - sx) * (py - sy) - (ey - sy) * (px - sx)
    return -e if swap else e


def _setup(verts, face, T, K):
    fx, sk, cx, fy, cy = K[0, 0], K[0, 1], K[0, 2], K[1, 1], K[1, 2]
    u, v, iz, ok = [], [], [], True
    for i in range(3):
        X, Y, Z = verts[face[i]]
        xc = ((T[0, 0] * X + T[0, 1] * Y) + T[0, 2] * Z) + T[0, 3]
        yc = ((T[1, 0] * X + T[1, 1] * Y) + T[1, 2] * Z) + T[1, 3]
        zc = ((T[2, 0] * X + T[2, 1] * Y) + T[2, 2] * Z) + T[2, 3]
        if not zc >= NEAR_Z:
            ok = False
        with np.errstate(all='ignore'):
            u.append((fx * xc + sk * yc) / zc + cx)
            v.append((fy * yc) / zc + cy)
            iz.append(F(1.0) / zc)
    vid = [int(face[0]), int(face[1]), int(face[2])]
    with np.errstate(all='ignore'):
        area = _edge(u[0], v[0], u[1], v[1], u[2], v[2])
    if area < 0:
        for a in (u, v, iz, vid):
            a[1], a[2] = a[2], a[1]
        area = -area
    if not area > 0:
        ok = False
    return ok, u, v, iz, area, vid


def render(vertices, colors, faces, face_offsets, label_ids, TCO, K):
    """vertices / colors [Nv,3] float32, faces [Nf,3] int32, face_offsets [L+1]; label_ids [B], TCO [B,4,4],
    K [B,3,3] -> uint8 [B,240,320,3] and the depth / triangle-id buffers (float32 z, int64 id, -1 = background)."""
    vertices = np.asarray(vertices, dtype=F)
    colors = np.asarray(colors, dtype=F)
    B = len(label_ids)
    out = np.zeros((B, H, W, 3), dtype=np.uint8)
    zbuf = np.full((B, H, W), np.inf, dtype=F)
    ids = np.full((B, H, W), -1, dtype=np.int64)
    half = F(0.5)
    for b in range(B):
        T = np.asarray(TCO[b], dtype=F)
        Kb = np.asarray(K[b], dtype=F)
        lab = int(label_ids[b])
        for t in range(int(face_offsets[lab]), int(face_offsets[lab + 1])):
            ok, u, v, iz, area, vid = _setup(vertices, faces[t], T, Kb)
            if not ok:
                continue
            umin, umax, vmin, vmax = min(u), max(u), min(v), max(v)
            if not all(np.isfinite(x) for x in (umin, umax, vmin, vmax)):
                # the kernel saturates the conversions; boxes that far out never intersect the view unless they span it
                x0, x1 = (0 if umin < 0 else W), (W - 1 if umax > 0 else -1)
                y0, y1 = (0 if vmin < 0 else H), (H - 1 if vmax > 0 else -1)
            else:
                x0 = max(0, int(np.ceil(np.clip(umin - half, -1e9, 1e9))))
                x1 = min(W - 1, int(np.floor(np.clip(umax - half, -1e9, 1e9))))
                y0 = max(0, int(np.ceil(np.clip(vmin - half, -1e9, 1e9))))
                y1 = min(H - 1, int(np.floor(np.clip(vmax - half, -1e9, 1e9))))
            if x1 < x0 or y1 < y0:
                continue
            jj, ii = np.meshgrid(np.arange(x0, x1 + 1), np.arange(y0, y1 + 1))
            px, py = jj.astype(F) + half, ii.astype(F) + half

            def edge_v(a, c):
                ax, ay, bx, by = u[a], v[a], u[c], v[c]
                swap = (bx < ax) or (bx == ax and by < ay)
                sx, sy, ex, ey = (bx, by, ax, ay) if swap else (ax, ay, bx, by)
                e = (ex - sx) * (py - sy) - (ey - sy) * (px - sx)
                return -e if swap else e
            w0, w1, w2 = edge_v(1, 2), edge_v(2, 0), edge_v(0, 1)
            inside = (w0 >= 0) & (w1 >= 0) & (w2 >= 0)
            if not inside.any():
                continue
            with np.errstate(all='ignore'):
                b0, b1, b2 = w0 / area, w1 / area, w2 / area
                q = (b0 * iz[0] + b1 * iz[1]) + b2 * iz[2]
                z = F(1.0) / q
            inside &= q > 0
            zs, idv = zbuf[b, y0:y1 + 1, x0:x1 + 1], ids[b, y0:y1 + 1, x0:x1 + 1]
            # nearest wins; equal depth: the smaller triangle id (the 64-bit key order of the kernel)
            win = inside & ((z < zs) | ((z == zs) & (t < idv)))
            if not win.any():
                continue
            zs[win] = z[win]
            idv[win] = t
            o = out[b, y0:y1 + 1, x0:x1 + 1]
            wq = (b0 * iz[0], b1 * iz[1], b2 * iz[2])
            for c in range(3):
                s = (wq[0] * colors[vid[0], c] + wq[1] * colors[vid[1], c]) + wq[2] * colors[vid[2], c]
                val = np.minimum(np.maximum(s * z, F(0)), F(1))
                o[..., c][win] = np.rint(val * F(255))[win].astype(np.uint8)
    return out, zbuf, ids
