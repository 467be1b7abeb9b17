"""Properties of the rasteriser restatement (oracle/raster_oracle.py) that hold without a GPU, and the mesh table
packing the engine consumes."""
import numpy as np
from scipy.ndimage import binary_fill_holes
from scipy.spatial.transform import Rotation as R

from cosypose_b200 import synthetic
from cosypose_b200.rendering import RenderMeshTable
from oracle import raster_oracle as ro


def _scene(B=4, subdiv=2):
    v, f, c = synthetic.make_render_meshes(4, subdiv=subdiv)
    tab = RenderMeshTable(synthetic.make_labels(4), v, f, c)
    rs = np.random.RandomState(0)
    T = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
    T[:, :3, :3] = R.random(B, random_state=1).as_matrix().astype(np.float32)
    T[:, 2, 3] = rs.uniform(0.25, 0.6, B)
    K = np.tile(np.array([[900, 0, 160], [0, 900, 120], [0, 0, 1]], dtype=np.float32), (B, 1, 1))
    return tab, T, K


def test_mesh_table_packing():
    v, f, c = synthetic.make_render_meshes(3)
    tab = RenderMeshTable(['a', 'b', 'c'], v, f, c)
    assert tab.face_offsets.tolist() == [0, len(f[0]), len(f[0]) + len(f[1]), len(f[0]) + len(f[1]) + len(f[2])]
    assert tab.vertices.shape[0] == sum(len(x) for x in v) and tab.faces.dtype == np.int32
    # faces of label 1 index its own vertices inside the global table
    f1 = tab.faces[tab.face_offsets[1]:tab.face_offsets[2]]
    assert f1.min() == len(v[0]) and f1.max() == len(v[0]) + len(v[1]) - 1
    assert np.array_equal(tab.vertices[f1[0]], v[1][f[1][0]])
    assert tab.label_ids(['c', 'a']).tolist() == [2, 0]


def test_closed_meshes_render_without_holes_and_background_is_zero():
    tab, T, K = _scene()
    out, z, ids = ro.render(tab.vertices, tab.colors, tab.faces, tab.face_offsets, [0, 1, 2, 3], T, K)
    for b in range(4):
        m = ids[b] >= 0
        assert m.sum() > 2000
        assert not (binary_fill_holes(m) ^ m).any()          # shared edges never drop a pixel centre
        assert not out[b][~m].any()
        # every winning triangle belongs to the hypothesis's label
        assert ids[b][m].min() >= tab.face_offsets[b] and ids[b][m].max() < tab.face_offsets[b + 1]
        # the depth buffer holds camera-frame z of the surface: inside the object's z range
        zc = (tab.vertices @ T[b, :3, :3].T + T[b, :3, 3])[:, 2]
        assert z[b][m].min() >= zc.min() - 1e-5 and z[b][m].max() <= zc.max() + 1e-5


def test_projection_matches_pinhole_of_the_path():
    """A tiny triangle around a mesh point lands on the pixel whose centre is nearest to K @ TCO @ p
    (the projection of lib3d/camera_geometry.py:18-31), with the OpenGL half-pixel convention."""
    p = np.array([0.01, -0.02, 0.0], dtype=np.float32)
    e = 1.5e-3
    verts = np.stack([p + [e, 0, 0], p + [-e, e, 0], p + [-e, -e, 0]]).astype(np.float32)
    T = np.eye(4, dtype=np.float32)[None].copy()
    T[0, 2, 3] = 0.5
    K = np.array([[[600, 0, 160], [0, 600, 120], [0, 0, 1]]], dtype=np.float32)
    out, z, ids = ro.render(verts, np.ones_like(verts), np.array([[0, 1, 2]], dtype=np.int32), [0, 1], [0], T, K)
    uv = (K[0] @ (p + T[0, :3, 3]))
    u, v = uv[0] / uv[2], uv[1] / uv[2]
    ii, jj = np.nonzero(ids[0] >= 0)
    assert len(ii) >= 1
    assert abs(jj.mean() + 0.5 - u) < 1.0 and abs(ii.mean() + 0.5 - v) < 1.0
    assert (out[0][ii, jj] == 255).all()


def test_near_plane_and_invalid_poses_give_black_frames():
    tab, T, K = _scene(2)
    T[0, 2, 3] = -0.3
    T[1, 0, 0] = np.nan
    out, _, ids = ro.render(tab.vertices, tab.colors, tab.faces, tab.face_offsets, [0, 1], T, K)
    assert not out.any() and (ids < 0).all()


def _write_ply(path, v, f, c, fmt):
    import struct
    head = ['ply', f'format {fmt} 1.0', 'comment made by the test', f'element vertex {len(v)}',
            'property float x', 'property float y', 'property float z', 'property float nx', 'property float ny',
            'property float nz', 'property uchar red', 'property uchar green', 'property uchar blue',
            f'element face {len(f)}', 'property list uchar int vertex_indices', 'end_header']
    c8 = np.rint(c * 255).astype(np.uint8)
    with open(path, 'wb') as fh:
        fh.write(('\n'.join(head) + '\n').encode())
        if fmt == 'ascii':
            for p, col in zip(v, c8):
                fh.write(('%r %r %r 0 0 1 %d %d %d\n' % (float(p[0]), float(p[1]), float(p[2]), *col)).encode())
            for tri in f:
                fh.write(('3 %d %d %d\n' % tuple(tri)).encode())
        else:
            e = '<' if fmt == 'binary_little_endian' else '>'
            for p, col in zip(v, c8):
                fh.write(struct.pack(e + '6f3B', p[0], p[1], p[2], 0, 0, 1, *col))
            for tri in f:
                fh.write(struct.pack(e + 'B3i', 3, *tri))


def test_ply_reader_round_trip(tmp_path):
    from cosypose_b200.lib3d.ply import read_ply
    v, f, c = synthetic.make_render_meshes(2)
    for fmt in ('ascii', 'binary_little_endian', 'binary_big_endian'):
        p = tmp_path / f'{fmt}.ply'
        _write_ply(p, v[0] * 1000, f[0], c[0], fmt)
        m = read_ply(p)
        assert np.array_equal(m['faces'], f[0]) and m['faces'].dtype == np.int32
        assert np.allclose(m['vertices'], v[0] * 1000, rtol=1e-6)
        assert np.abs(m['colors'] - c[0]).max() <= 0.5 / 255 + 1e-6
    # quads are fan-triangulated; a file without colours gives colors=None
    q = tmp_path / 'quad.ply'
    q.write_text('ply\nformat ascii 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\n'
                 'element face 1\nproperty list uchar uint vertex_index\nend_header\n0 0 0\n1 0 0\n1 1 0\n0 1 0\n4 0 1 2 3\n')
    m = read_ply(q)
    assert m['colors'] is None and m['faces'].tolist() == [[0, 1, 2], [0, 2, 3]]
    tab = RenderMeshTable.from_ply(['a', 'b'], [tmp_path / 'ascii.ply', tmp_path / 'binary_little_endian.ply'])
    assert np.allclose(tab.vertices[:len(v[0])], v[0], atol=1e-7) and tab.face_offsets.tolist() == [0, len(f[0]), 2 * len(f[0])]
