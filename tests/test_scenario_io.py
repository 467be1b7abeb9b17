"""File formats either side of the multiview path: BOP models directory, symmetry sets, camera file,
scene JSON, 3-D NMS; and (GPU) the custom-scenario runner end to end on a synthetic scenario written to disk."""
import json

import numpy as np
import pandas as pd
import pytest
import torch

from cosypose_b200 import synthetic
from cosypose_b200.evaluation import bop_io
from cosypose_b200.evaluation.custom_scenario import nms3d, run_custom_scenario
from cosypose_b200.lib3d.symmetries import make_bop_symmetries
from cosypose_b200.utils import tensor_collection as tc
from test_oracle_raster import _write_ply


def test_make_bop_symmetries():
    assert np.array_equal(make_bop_symmetries({}), np.eye(4)[None])
    flip = np.diag([-1., -1., 1., 1.])
    flip[:3, 3] = (10., 0., 0.)                       # millimetres in models_info.json
    d = make_bop_symmetries(dict(symmetries_discrete=[flip.flatten().tolist()]))
    assert d.shape == (2, 4, 4) and np.array_equal(d[0], np.eye(4)) and np.allclose(d[1][:3, 3], (0.01, 0, 0))
    c = make_bop_symmetries(dict(symmetries_continuous=[dict(axis=[0, 0, 1], offset=[0, 0, 0])]), n_symmetries_continuous=4)
    assert c.shape == (4, 4, 4) and np.allclose(c[0], np.eye(4))
    assert np.allclose(c[1][:3, :3], [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-12)     # +90 degrees about z
    both = make_bop_symmetries(dict(symmetries_discrete=[flip.flatten().tolist()],
                                    symmetries_continuous=[dict(axis=[1, 0, 0], offset=[0, 0, 0])]), 8)
    assert both.shape == (16, 4, 4)
    assert np.allclose(both[8 + 2], c_rot('x', np.pi / 2) @ d[1])                        # continuous * discrete, discrete-major
    for M in both:
        assert np.allclose(M[:3, :3] @ M[:3, :3].T, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(M[:3, :3]), 1)


def c_rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    M = np.eye(4)
    i, j = dict(x=(1, 2), y=(2, 0), z=(0, 1))[axis]
    M[i, i], M[i, j], M[j, i], M[j, j] = c, -s, s, c
    return M


def _write_models(models_dir, n_labels, sym_every=3):
    models_dir.mkdir(parents=True)
    v, f, c = synthetic.make_render_meshes(n_labels, subdiv=1)
    infos = {}
    for l in range(n_labels):
        _write_ply(models_dir / f'obj_{l + 1:06d}.ply', v[l] * 1000, f[l], c[l], 'binary_little_endian' if l % 2 else 'ascii')
        infos[str(l + 1)] = dict(diameter=float(np.linalg.norm(np.ptp(v[l] * 1000, axis=0))))
        if l % sym_every == 1:
            infos[str(l + 1)]['symmetries_discrete'] = [np.diag([-1., -1., 1., 1.]).flatten().tolist()]
    (models_dir / 'models_info.json').write_text(json.dumps(infos))
    return v, f, c


def test_models_directory_to_tables(tmp_path):
    v, f, c = _write_models(tmp_path / 'models', 5)
    objs = bop_io.load_bop_object_models(tmp_path / 'models')
    assert [o['label'] for o in objs] == [f'obj_{i:06d}' for i in range(1, 6)]
    assert objs[1]['is_symmetric'] and not objs[0]['is_symmetric'] and objs[0]['mesh_units'] == 'mm'
    assert np.isclose(objs[2]['diameter_m'], objs[2]['diameter'] * 1e-3)
    mesh_db, table = bop_io.mesh_db_from_bop_models(tmp_path / 'models', n_sym=4)
    assert list(mesh_db.labels) == [o['label'] for o in objs]
    assert mesh_db.n_sym_array().tolist() == [1, 2, 1, 1, 2]
    assert mesh_db.symmetries.shape == (5, 2, 4, 4) and torch.equal(mesh_db.symmetries[0, 1], torch.eye(4))   # identity padded
    n0 = len(v[0])
    assert np.allclose(mesh_db.points[0, :n0].numpy(), v[0], atol=1e-7)                                       # metres
    assert table.face_offsets.tolist() == np.cumsum([0] + [len(x) for x in f]).tolist()
    assert mesh_db.infos['obj_000002']['symmetries_discrete'] and mesh_db.infos['obj_000002']['n_sym'] == 2


def test_read_cameras_and_scene_json(tmp_path):
    cams = {str(v): dict(cam_K=(np.array([[600 + v, 0, 320], [0, 610, 240], [0, 0, 1]], dtype=float)).flatten().tolist(),
                         depth_scale=1.0) for v in (3, 7, 11)}
    (tmp_path / 'scene_camera.json').write_text(json.dumps(cams))
    cameras = bop_io.read_cameras(tmp_path / 'scene_camera.json', np.array([7, 3]))
    assert cameras.K.shape == (2, 3, 3) and cameras.K[0, 0, 0] == 607 and list(cameras.infos['view_id']) == [7, 3]
    objects = tc.PandasTensorCollection(
        infos=pd.DataFrame(dict(score=[2.5, 1.0], label=['obj_000004', 'obj_000001'], n_cand=[3, 2], obj_id=[0, 1])),
        TWO=torch.eye(4).repeat(2, 1, 1))
    cameras = tc.PandasTensorCollection(infos=cameras.infos, TWC=torch.eye(4).repeat(2, 1, 1) * 2, K=cameras.K.float())
    bop_io.save_scene_json(objects, cameras, tmp_path / 'scene.json')
    scene = json.loads((tmp_path / 'scene.json').read_text())
    assert [o['label'] for o in scene['objects']] == ['obj_000004', 'obj_000001'] and scene['objects'][0]['n_cand'] == 3
    assert scene['objects'][0]['score'] == 2.5 and np.array(scene['objects'][1]['TWO']).shape == (4, 4)
    assert [c['view_id'] for c in scene['cameras']] == [7, 3] and scene['cameras'][0]['K'][0][0] == 607.0
    assert scene['cameras'][1]['TWC'][0][0] == 2.0


def test_nms3d_keeps_best_of_each_cluster():
    t = np.array([[0, 0, 0], [0.01, 0, 0], [0.5, 0, 0], [0.51, 0.01, 0], [1.0, 0, 0]], dtype=np.float32)
    TWO = torch.eye(4).repeat(5, 1, 1)
    TWO[:, :3, 3] = torch.from_numpy(t)
    objs = tc.PandasTensorCollection(infos=pd.DataFrame(dict(score=[1.0, 3.0, 2.0, 0.5, 0.1], label=list('abcde'))), TWO=TWO)
    out = nms3d(objs, th=0.04, poses_attr='TWO')
    assert list(out.infos['label']) == ['b', 'c', 'e']                      # best first; a and d suppressed
    assert torch.equal(out.TWO[:, :3, 3], torch.from_numpy(t[[1, 2, 4]]))   # poses stay aligned with their rows


@pytest.mark.gpu
def test_custom_scenario_end_to_end(tmp_path):
    """candidates.csv + scene_camera.json + models/ -> results/subscene=0/{predicted_scene.json, scene_reprojected.csv}:
    all 12 objects are recovered and their reprojections agree with the candidates they came from."""
    n_views, n_obj, n_labels = 4, 12, 12
    _write_models(tmp_path / 'models', n_labels, sym_every=100)
    s = synthetic.make_multiview_scene(n_views, n_obj, n_labels, seed=4)
    labels = [f'obj_{i + 1:06d}' for i in s['label_ids']]
    infos = pd.DataFrame(dict(scene_id=5, view_id=s['view_ids'] * 10, label=labels, score=s['scores']))
    bop_io.tc_to_csv(tc.PandasTensorCollection(infos=infos, poses=s['poses']), tmp_path / 'candidates.csv')
    cams = {str(v * 10): dict(cam_K=s['K'][v].flatten().tolist()) for v in range(n_views)}
    (tmp_path / 'scene_camera.json').write_text(json.dumps(cams))
    msgs = []
    out = run_custom_scenario(tmp_path, ransac_n_iter=200, ba_n_iter=3, log=msgs.append)
    assert len(out['result_dirs']) == 1 and out['result_dirs'][0].name == 'subscene=0'
    scene = json.loads((out['result_dirs'][0] / 'predicted_scene.json').read_text())
    assert len(scene['objects']) == n_obj and len(scene['cameras']) == n_views
    assert sorted(o['label'] for o in scene['objects']) == sorted(set(labels))
    assert all(o['n_cand'] == n_views for o in scene['objects'])
    assert sorted(c['view_id'] for c in scene['cameras']) == [0, 10, 20, 30]
    reproj = bop_io.read_csv_candidates(out['result_dirs'][0] / 'scene_reprojected.csv')
    assert len(reproj) == n_obj * n_views
    # every reprojected (view, label) pose agrees with the candidate it was estimated from in what bundle adjustment
    # minimises (image position; the candidates carry 2 mm of noise = up to ~2 px here): projected origin within 4 px, depth within 8 %
    cand = {(int(v), l): T for v, l, T in zip(infos['view_id'], labels, s['poses'])}
    K = s['K'][0]
    for i in range(len(reproj)):
        T0 = cand[(int(reproj.infos['view_id'][i]), reproj.infos['label'][i])]
        T1 = reproj.poses[i]
        uv0, uv1 = (K @ T0[:3, 3]) / T0[2, 3], (K @ T1[:3, 3]) / T1[2, 3]
        assert (uv0 - uv1).abs().max() < 4.0, (i, uv0, uv1)
        assert abs(T1[2, 3] / T0[2, 3] - 1) < 0.08
        assert (T1[:3, :3] - T0[:3, :3]).abs().max() < 0.05
    assert any('12 objects and 4 cameras' in m for m in msgs)
