"""BOP-format candidate / result files: the wire formats on either side of the path.
Reference: cosypose/scripts/run_custom_scenario.py:26-92 (`tc_to_csv`, `read_csv_candidates`, `read_cameras`,
`save_scene_json`), cosypose/datasets/bop_object_datasets.py:5-34 (`models_info.json` + `obj_%06d.ply`) and
bop_toolkit_lib/inout.py:265-294 (`save_bop_results`, version bop19: `scene_id,im_id,obj_id,score,R,t,time`, R row-major
with 9 and t with 3 space-separated numbers, t in millimetres)."""
import json
from pathlib import Path

import numpy as np
import pandas as pd
import torch

from ..utils import tensor_collection as tc

HEADER = 'scene_id,im_id,obj_id,score,R,t,time'


def tc_to_csv(predictions, csv_path):
    """predictions: PandasTensorCollection with infos[scene_id, view_id, label, score] and poses [n,4,4] (metres)."""
    poses = predictions.poses.detach().cpu().numpy()
    lines = [HEADER]
    for n in range(len(predictions)):
        row = predictions.infos.iloc[n]
        t = poses[n, :3, -1] * 1e3                                  # m -> mm
        R = poses[n, :3, :3]
        lines.append('{},{},{},{},{},{},{}'.format(
            row.scene_id, row.view_id, int(str(row.label).split('_')[-1]), row.score,
            ' '.join(map(str, R.flatten().tolist())), ' '.join(map(str, t.flatten().tolist())), -1.0))
    with open(csv_path, 'w') as f:
        f.write('\n'.join(lines))


def read_csv_candidates(csv_path):
    """-> PandasTensorCollection(infos[view_id, scene_id, score, label], poses [n,4,4] in metres)."""
    df = pd.read_csv(csv_path)
    infos = df.loc[:, ['im_id', 'scene_id', 'score', 'obj_id']].copy()
    infos['obj_id'] = infos['obj_id'].apply(lambda x: f'obj_{x:06d}')
    infos = infos.rename(dict(im_id='view_id', obj_id='label'), axis=1)
    R = np.stack(df['R'].apply(lambda x: list(map(float, x.split(' '))))).reshape(-1, 3, 3)
    t = np.stack(df['t'].apply(lambda x: list(map(float, x.split(' '))))).reshape(-1, 3) * 1e-3
    TCO = torch.eye(4, dtype=torch.float32).unsqueeze(0).repeat(len(R), 1, 1)
    TCO[:, :3, :3] = torch.tensor(R, dtype=torch.float32)
    TCO[:, :3, -1] = torch.tensor(t, dtype=torch.float32)
    return tc.PandasTensorCollection(poses=TCO, infos=infos)


def read_cameras(json_path, view_ids):
    """BOP `scene_camera.json` ({"<view id>": {"cam_K": [9 numbers], ...}}) -> PandasTensorCollection(infos[view_id],
    K [V,3,3]) in the order of `view_ids` (reference: run_custom_scenario.py:61-70)."""
    cameras = json.loads(Path(json_path).read_text())
    K = np.stack([np.array(cameras[str(v)]['cam_K'], dtype=np.float64).reshape(3, 3) for v in view_ids])
    return tc.PandasTensorCollection(K=torch.as_tensor(K), infos=pd.DataFrame(dict(view_id=view_ids)))


def save_scene_json(objects, cameras, results_scene_path):
    """{"objects": [{score, label, n_cand, TWO}], "cameras": [{view_id, TWC, K}]} (reference: :73-92)."""
    TWO, TWC, K = (x.detach().cpu().numpy() for x in (objects.TWO, cameras.TWC, cameras.K))
    list_objects, list_cameras = [], []
    for n in range(len(objects)):
        obj = {k: np.asarray(objects.infos.loc[n, k]).item() for k in ('score', 'label', 'n_cand')}
        obj['TWO'] = TWO[n].tolist()
        list_objects.append(obj)
    for n in range(len(cameras)):
        list_cameras.append(dict(view_id=np.asarray(cameras.infos.loc[n, 'view_id']).item(), TWC=TWC[n].tolist(),
                                 K=K[n].tolist()))
    Path(results_scene_path).write_text(json.dumps(dict(objects=list_objects, cameras=list_cameras)))


def load_bop_object_models(models_dir, mesh_units='mm'):
    """The object list of a BOP `models/` directory (reference: datasets/bop_object_datasets.py:5-34): one dict per
    entry of models_info.json with label `obj_%06d`, mesh_path, symmetries, diameter."""
    models_dir = Path(models_dir)
    infos = json.loads((models_dir / 'models_info.json').read_text())
    scale = 0.001 if mesh_units == 'mm' else 1.0
    objects = []
    for obj_id, bop_info in infos.items():
        label = f'obj_{int(obj_id):06d}'
        obj = dict(label=label, category=None, mesh_path=(models_dir / label).with_suffix('.ply').as_posix(),
                   mesh_units=mesh_units)
        for k in ('symmetries_discrete', 'symmetries_continuous'):
            obj[k] = bop_info.get(k, [])
        obj['is_symmetric'] = any(len(obj[k]) > 0 for k in ('symmetries_discrete', 'symmetries_continuous'))
        obj['diameter'] = bop_info['diameter']
        obj['diameter_m'] = bop_info['diameter'] * scale
        objects.append(obj)
    return objects


def mesh_db_from_bop_models(models_dir, n_sym=64, mesh_units='mm'):
    """`MeshDataBase.from_object_ds(BOPObjectDataset(dir))` of the reference (lib3d/rigid_mesh_database.py:11-56) as
    stacked tables: vertices of every `obj_%06d.ply` in metres, symmetry sets with continuous symmetries discretised
    into `n_sym` rotations.  Returns (BatchedMeshes, RenderMeshTable): the second feeds the device rasteriser."""
    from ..lib3d.ply import read_ply
    from ..lib3d.rigid_mesh_database import BatchedMeshes
    from ..lib3d.symmetries import make_bop_symmetries
    from ..rendering import RenderMeshTable
    objects = load_bop_object_models(models_dir, mesh_units)
    if mesh_units not in ('mm', 'm'):
        raise ValueError('Unit not supported', mesh_units)
    scale = 0.001 if mesh_units == 'mm' else 1.0
    meshes = [read_ply(o['mesh_path']) for o in objects]
    labels = [o['label'] for o in objects]
    verts = [m['vertices'] * np.float32(scale) for m in meshes]
    syms = [make_bop_symmetries(o, n_symmetries_continuous=n_sym, scale=scale) for o in objects]
    mesh_db = BatchedMeshes.from_vertex_lists(labels, verts, syms)
    for o in objects:
        mesh_db.infos[o['label']].update({k: v for k, v in o.items() if k != 'label'})
    table = RenderMeshTable(labels, verts, [m['faces'] for m in meshes], [m['colors'] for m in meshes])
    return mesh_db, table
