"""Minimal workload for ncu: a few refinement iterations at the benchmark batch (64 hypotheses).

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c <n> \
        -o gpurun_out/prof python tools/prof_iter.py [--iters 2] [--gemm-impl 1]
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
ap.add_argument('--iters', type=int, default=2)
ap.add_argument('--gemm-impl', type=int, default=1)
args = ap.parse_args()
dev = torch.device('cuda', 0)
w = Workload(8, 8, 21, 1, max(1, args.iters - 1))
pred, eng, views = build_predictor(w, 0, bsz_objects=64)
eng.set_option('gemm_impl', args.gemm_impl)
det = tc.PandasTensorCollection(infos=w.infos(), bboxes=w.boxes.to(dev))
final, _ = pred.get_predictions(w.images.to(dev), w.K.to(dev), detections=det, n_coarse_iterations=1,
                                n_refiner_iterations=max(1, args.iters - 1))
torch.cuda.synchronize()
print('done', final.poses[0, :3, 3].tolist())
