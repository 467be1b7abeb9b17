"""Per-MBConv-block device time of one refinement iteration at the benchmark batch (CUDA events around every
launch, engine profiling mode; warm caches, unlike an ncu launch list).

    python tools/block_times.py [--iters 3] [--opt name=value ...]
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))

import torch  # noqa: E402
from helpers import Workload, build_predictor  # noqa: E402
from cosypose_b200.utils import tensor_collection as tc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--opt', action='append', default=[])
args = ap.parse_args()
dev = torch.device('cuda', 0)
w = Workload(8, 8, 21, 1, 4)
pred, eng, views = build_predictor(w, 0, bsz_objects=64)
for o in args.opt:
    k, v = o.split('=')
    eng.set_option(k, int(v))
det = tc.PandasTensorCollection(infos=w.infos(), bboxes=w.boxes.to(dev))
images, K = w.images.to(dev), w.K.to(dev)


def run():
    views.reset()
    return pred.get_predictions(images, K, detections=det, n_coarse_iterations=1, n_refiner_iterations=4)


run()
torch.cuda.synchronize()
eng.profile_read(reset=True)
eng.profile_read_blocks(reset=True)
eng.profile_enable(True)
for _ in range(args.iters):
    run()
torch.cuda.synchronize()
tot = eng.profile_read(reset=True)
blk = eng.profile_read_blocks(reset=True)
eng.profile_enable(False)
n_fwd = args.iters * 5
spec = eng.block_specs() if hasattr(eng, 'block_specs') else None
print(f'per forward batch of 64 (us), mean over {n_fwd} forwards; options {args.opt}')
print('blk   expand      dw      se    proj   total')
sums = [0, 0, 0, 0]
for b in range(26):
    v = [blk[c][b] * 1e3 / n_fwd for c in ('expand_1x1', 'depthwise', 'squeeze_excite', 'project_1x1')]
    sums = [a + x for a, x in zip(sums, v)]
    print(f'{b:3d} {v[0]:8.1f} {v[1]:7.1f} {v[2]:7.1f} {v[3]:7.1f} {sum(v):7.1f}')
print(f'sum {sums[0]:8.1f} {sums[1]:7.1f} {sums[2]:7.1f} {sums[3]:7.1f} {sum(sums):7.1f}')
for c in ('geometry', 'roi_crop', 'stem', 'head_1x1', 'pool_fc_update'):
    print(f'{c:16s} {tot[c][1] * 1e3 / n_fwd:8.1f}')
print(f'all categories   {sum(ms for _, ms in tot.values()) * 1e3 / n_fwd:8.1f}')
