"""BOP-format candidate / result files: the wire formats on either side of the path.
Reference: cosypose/scripts/run_custom_scenario.py:26-58 (`tc_to_csv`, `read_csv_candidates`) and
bop_toolkit_lib/inout.py:265-294 (`save_bop_results`, version bop19: `scene_id,im_id,obj_id,score,R,t,time`, R row-major
with 9 and t with 3 space-separated numbers, t in millimetres)."""
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
