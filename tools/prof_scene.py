"""Host-side profile (cProfile) of MultiviewScenePredictor.predict_scene_state on the configs[3] scene.

    python tools/prof_scene.py [--n 20]
"""
import argparse
import cProfile
import pstats
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))

import torch  # noqa: E402
from helpers import Scene  # noqa: E402
from cosypose_b200.integrated.multiview_predictor import MultiviewScenePredictor  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=20)
ap.add_argument('--top', type=int, default=45)
args = ap.parse_args()
dev = torch.device('cuda', 0)
sc = Scene(8, 16, 21, (1,), True, seed=0)
mv = MultiviewScenePredictor(sc.mesh_db(), device=dev)
cands, cams = sc.candidates(dev), sc.cameras(dev)


def step():
    out = mv.predict_scene_state(cands, cams, ransac_n_iter=2000, ransac_dist_threshold=0.02, ba_n_iter=2)
    out['scene/objects'].TWO.cpu()


for _ in range(3):
    step()
torch.cuda.synchronize()
import time  # noqa: E402
t0 = time.perf_counter()
for _ in range(args.n):
    step()
torch.cuda.synchronize()
print(f'wall: {(time.perf_counter() - t0) / args.n * 1e3:.2f} ms per scene (no profiler)')
pr = cProfile.Profile()
pr.enable()
for _ in range(args.n):
    step()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(args.top)
