"""Label-indexed mesh tables used by the path: padded point clouds and symmetry sets.

Mirrors the interface the predictors use on the reference's `BatchedMeshes` / `Meshes`
(reference: cosypose/lib3d/rigid_mesh_database.py:59-94): `label_to_id`, `labels`, `points`
[L,Nmax,3], `symmetries` [L,Smax,4,4] (identity padded, :55), `n_sym_mapping`, `select(labels)`
and `sample_points(n, deterministic=True)`.  Mesh *loading* (trimesh) is outside the path; tables
are built from vertex arrays with the reference's padding rule (`pad_stack_tensors`, :97-120:
short clouds are padded by re-drawing their own points with RandomState(0), symmetry sets with
identity), because the fixed 2000-point subset depends on the padded length Nmax.
"""
import numpy as np
import torch

from ..engine import aabb_corners, sample_point_ids
from ..utils.tensor_collection import TensorCollection


def pad_stack(arrays, fill='select_random'):
    """Stack variable-length [n_i, ...] arrays to [L, n_max, ...] with the reference's padding."""
    n_max = max(len(a) for a in arrays)
    rs = np.random.RandomState(0)
    out = []
    for a in arrays:
        a = np.asarray(a)
        n_pad = n_max - len(a)
        if n_pad > 0:
            if isinstance(fill, str):
                assert fill == 'select_random'
                pad = a[rs.choice(np.arange(len(a)), size=n_pad)]
            else:
                pad = np.broadcast_to(np.asarray(fill, dtype=a.dtype), (n_pad,) + a.shape[1:])
            a = np.concatenate((a, pad), axis=0)
        out.append(a)
    return np.stack(out)


class Meshes(TensorCollection):
    def __init__(self, infos, labels, points, symmetries):
        super().__init__()
        self.infos = infos
        self.labels = np.asarray(labels)
        self.register_tensor('points', points)
        self.register_tensor('symmetries', symmetries)

    def sample_points(self, n_points, deterministic=False):
        assert n_points <= self.points.shape[1]
        rs = np.random.RandomState(0) if deterministic else np.random
        ids = rs.choice(self.points.shape[1], size=n_points, replace=False)
        return torch.index_select(self.points, 1, torch.as_tensor(ids).to(self.points.device))


class BatchedMeshes(TensorCollection):
    def __init__(self, infos, labels, points, symmetries):
        super().__init__()
        self.infos = infos
        self.label_to_id = {label: n for n, label in enumerate(labels)}
        self.labels = np.asarray(labels)
        self.register_tensor('points', points)
        self.register_tensor('symmetries', symmetries)

    @classmethod
    def from_vertex_lists(cls, labels, vertices, symmetries=None, aabb=False):
        """vertices: list of [n_i,3] arrays (metres); symmetries: list of [s_i,4,4] (or None)."""
        if aabb:
            vertices = [aabb_corners(np.asarray(v)[None])[0] for v in vertices]
        if symmetries is None:
            symmetries = [np.eye(4)[None] for _ in labels]
        infos = {l: dict(label=l, n_points=len(v), n_sym=len(s)) for l, v, s in zip(labels, vertices, symmetries)}
        points = torch.as_tensor(pad_stack(vertices), dtype=torch.float32)
        syms = torch.as_tensor(pad_stack(symmetries, fill=np.eye(4)), dtype=torch.float32)
        return cls(infos, labels, points, syms)

    @classmethod
    def from_tables(cls, labels, points, symmetries, n_sym):
        infos = {l: dict(label=l, n_points=int(points.shape[1]), n_sym=int(n)) for l, n in zip(labels, n_sym)}
        return cls(infos, labels, torch.as_tensor(points, dtype=torch.float32),
                   torch.as_tensor(symmetries, dtype=torch.float32))

    @property
    def n_sym_mapping(self):
        return {label: obj['n_sym'] for label, obj in self.infos.items()}

    def n_sym_array(self):
        return np.array([self.infos[l]['n_sym'] for l in self.labels], dtype=np.int32)

    def label_ids(self, labels):
        """Dense int32 ids of an iterable of labels (KeyError on an unknown label, like select)."""
        l2i = self.label_to_id
        return np.fromiter((l2i[l] for l in labels), dtype=np.int32, count=len(labels))

    def select(self, labels):
        ids = self.label_ids(labels).astype(np.int64)
        return Meshes(infos=[self.infos[l] for l in labels], labels=self.labels[ids],
                      points=self.points[ids], symmetries=self.symmetries[ids])

    def aabb(self):
        return torch.as_tensor(aabb_corners(self.points.cpu().numpy()))

    def batched(self, aabb=False, resample_n_points=None, n_sym=None):
        """Same role as the reference's MeshDataBase.batched (rigid_mesh_database.py:21-56) on tables
        that are already stacked: `aabb=True` replaces each cloud by its 8 box corners
        (lib3d/mesh_ops.py:15-28); `resample_n_points` keeps a seeded subset of the vertices (the
        reference samples the mesh surface with trimesh, which needs the faces and is outside the
        path).  Symmetry sets are shared."""
        if aabb:
            assert resample_n_points is None
            points = self.aabb().to(self.points.device, self.points.dtype)
        elif resample_n_points:
            ids = np.random.RandomState(0).choice(self.points.shape[1], size=resample_n_points,
                                                  replace=self.points.shape[1] < resample_n_points)
            points = self.points[:, torch.as_tensor(ids)]
        else:
            points = self.points
        out = BatchedMeshes(self.infos, list(self.labels), points.clone(), self.symmetries)
        out.engine = getattr(self, 'engine', None)
        return out

    def install(self, engine, with_points=True):
        """Uploads the tables into an engine handle (cosyb200_set_meshes) and remembers the
        engine (`self.engine`) for the multiview stages."""
        pts = self.points.cpu().numpy()
        ids = sample_point_ids(pts.shape[1]) if with_points and pts.shape[1] >= 2000 else None
        engine.set_meshes(pts if ids is not None else None, self.symmetries, self.n_sym_array(),
                          aabb=aabb_corners(pts), point_ids=ids)
        self.engine = engine
        return engine
