"""Pre-rendered views standing in for the renderer side input.

The path treats the rendered view as an input (reference renderer contract:
cosypose/models/pose.py:100-102 calling cosypose/rendering/bullet_batch_renderer.py:46-90, which
returns float [B,3,240,320] in [0,1]).  `PreRenderedViews` replays stacks generated up front, in
the order `CoarseRefinePosePredictor` asks for them: per stage, per chunk of `bsz_objects`
hypotheses, per iteration.
"""
import torch


class PreRenderedViews:
    def __init__(self, stages, bsz_objects=64, device=None):
        """stages: list of tensors [n_iter_s, N, 3, 240, 320], one per stage in call order
        (e.g. [coarse_views, refiner_views])."""
        self.bsz = bsz_objects
        self.chunks = []      # flat list of [n_iter, Bc, 3, H, W] contiguous tensors, call order
        for views in stages:
            n = views.shape[1]
            for s in range(0, n, bsz_objects):
                c = views[:, s:s + bsz_objects].contiguous()
                self.chunks.append(c.to(device) if device is not None else c)
        self.reset()

    @classmethod
    def from_uint8(cls, stages, bsz_objects=64):
        """stages: device tensors [n_iter_s, N, 240, 320, 3] uint8 (NHWC, as the reference's renderer
        delivers them); consumed by the engine without a conversion pass."""
        self = cls.__new__(cls)
        self.bsz = bsz_objects
        self.chunks = []
        for views in stages:
            for s in range(0, views.shape[1], bsz_objects):
                self.chunks.append(views[:, s:s + bsz_objects].contiguous())
        self.reset()
        return self

    @classmethod
    def from_chunks(cls, chunks, bsz_objects=64):
        """chunks: per (stage, chunk) stacks [n_iter, Bc, ...] in call order, used in place (no copy): a caller
        that refills the same device buffers every step keeps the engine's captured graphs valid."""
        self = cls.__new__(cls)
        self.bsz = bsz_objects
        self.chunks = [c for c in chunks]
        assert all(c.is_contiguous() for c in self.chunks)
        self.reset()
        return self

    def reset(self):
        self._chunk = 0
        self._it = 0

    def _advance(self, n_iter_taken):
        self._it += n_iter_taken
        if self._it >= self.chunks[self._chunk].shape[0]:
            self._it = 0
            self._chunk = (self._chunk + 1) % len(self.chunks)

    def prerendered(self, n_iterations, batch_size):
        """All views of the next `n_iterations` calls as one [n_iterations, B, 3, H, W] stack."""
        c = self.chunks[self._chunk]
        assert self._it == 0 and c.shape[0] == n_iterations and c.shape[1] == batch_size, \
            'pre-rendered stack does not match the call sequence'
        self._advance(n_iterations)
        return c

    def render(self, obj_infos, TCO, K, resolution=(240, 320), **kwargs):
        c = self.chunks[self._chunk]
        assert c.shape[1] == len(obj_infos), 'pre-rendered stack does not match the call sequence'
        out = c[self._it]
        self._advance(1)
        return out


class PerCallRenderer:
    """Wraps PreRenderedViews but hides `prerendered`, forcing the two-phase per-iteration path
    (prepare_iter -> render -> refine_iter) a real renderer needs."""

    def __init__(self, views):
        self.views = views

    def reset(self):
        self.views.reset()

    def render(self, *args, **kwargs):
        return self.views.render(*args, **kwargs)


class RenderMeshTable:
    """Triangle meshes of a label set packed for the engine's rasteriser: one vertex / colour table, one face
    table with global vertex ids, per-label face ranges.  Label order must match the `BatchedMeshes`
    the pose models use (its `label_to_id`)."""

    def __init__(self, labels, vertices, faces, colors=None):
        """labels: list of str; vertices[l] [Nv_l,3] float (metres, object frame); faces[l] [Nf_l,3] int
        (ids into vertices[l]); colors[l] [Nv_l,3] in [0,1] (default: mid grey)."""
        import numpy as np
        assert len(labels) == len(vertices) == len(faces)
        self.labels = list(labels)
        self.label_to_id = {l: i for i, l in enumerate(self.labels)}
        vs, cs, fs, off, base = [], [], [], [0], 0
        for l in range(len(labels)):
            v = np.asarray(vertices[l], dtype=np.float32).reshape(-1, 3)
            f = np.asarray(faces[l], dtype=np.int64).reshape(-1, 3)
            assert f.size == 0 or (f.min() >= 0 and f.max() < len(v)), f'faces of {labels[l]} index outside its vertices'
            c = (np.full_like(v, 0.5) if colors is None or colors[l] is None
                 else np.asarray(colors[l], dtype=np.float32).reshape(-1, 3))
            assert c.shape == v.shape
            vs.append(v)
            cs.append(c)
            fs.append((f + base).astype(np.int32))
            base += len(v)
            off.append(off[-1] + len(f))
        self.vertices = np.concatenate(vs, axis=0)
        self.colors = np.concatenate(cs, axis=0)
        self.faces = np.concatenate(fs, axis=0)
        self.face_offsets = np.asarray(off, dtype=np.int32)

    @classmethod
    def from_ply(cls, labels, paths, scale=0.001):
        """One PLY per label (BOP `models/obj_%06d.ply`: millimetres, per-vertex colours); `scale` converts to the
        metres of the poses (0.001 as the reference loads its BOP URDFs, cosypose/datasets/urdf_dataset.py:31)."""
        from .lib3d.ply import read_ply
        meshes = [read_ply(p) for p in paths]
        return cls(labels, [m['vertices'] * scale for m in meshes], [m['faces'] for m in meshes],
                   [m['colors'] for m in meshes])

    def label_ids(self, labels):
        import numpy as np
        return np.asarray([self.label_to_id[l] for l in labels], dtype=np.int32)


class CudaRasterizer:
    """Renderer with the reference's call contract (reference: rendering/bullet_batch_renderer.py:46-90:
    `render(obj_infos, TCO, K, resolution)` -> float [B,3,240,320] in [0,1], background 0) drawn by
    libcosyb200.so on the hypotheses' own GPU: no worker processes, queue, pickling or pinned copy.

    `in_engine = True` tells `PosePredictor` that the views need not leave the engine at all: all
    iterations of a batch then run as one `refine_n` call (and one CUDA graph) that rasterises between
    the crop geometry and the network."""
    in_engine = True

    def __init__(self, engine, mesh_table, as_uint8=True):
        self.engine = engine
        self.table = mesh_table
        self.as_uint8 = as_uint8
        engine.set_render_meshes(mesh_table.vertices, mesh_table.colors, mesh_table.faces, mesh_table.face_offsets)

    def render(self, obj_infos, TCO, K, resolution=(240, 320), render_depth=False):
        assert tuple(resolution) == (240, 320), 'the engine renders 240x320 views'
        dev = self.engine.device
        label_ids = torch.from_numpy(self.table.label_ids([o['name'] for o in obj_infos])).to(dev)
        TCO = torch.as_tensor(TCO).detach().to(dev, torch.float32).contiguous()
        K = torch.as_tensor(K).detach().to(dev, torch.float32).contiguous()
        bsz = len(TCO)
        assert TCO.shape == (bsz, 4, 4)
        assert K.shape == (bsz, 3, 3)
        return self.engine.render(label_ids, TCO, K, uint8=self.as_uint8, depth=render_depth)

    def reset(self):
        pass
