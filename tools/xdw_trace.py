"""Cycle trace of CTA 0 of the fused expand + depthwise kernel (kernels_xdw.cuh) for one block at batch 64."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import torch
from helpers import Workload, build_predictor
from cosypose_b200.engine import _ptr
from cosypose_b200 import _lib
from cosypose_b200.utils import tensor_collection as tc
dev = torch.device('cuda', 0)
w = Workload(8, 8, 21, 1, 1)
pred, eng, views = build_predictor(w, 0, bsz_objects=64)
trace = torch.zeros(1024, dtype=torch.int64, device=dev)
_lib.check(eng._L.cosyb200_debug_trace(eng._h, _ptr(trace)))
det = tc.PandasTensorCollection(infos=w.infos(), bboxes=w.boxes.to(dev))
images, K = w.images.to(dev), w.K.to(dev)
names = ['top', 'acc_full', 'bar0', 'drained', 'bar1', 'dw done', 'bar2']
for blk in [int(a) for a in sys.argv[1:]] or [2, 3, 5, 6, 8]:
    eng.set_option('trace_block', blk)
    for rep in range(2):
        trace.zero_()
        views.reset()
        pred.get_predictions(images, K, detections=det, n_coarse_iterations=1, n_refiner_iterations=0)
        torch.cuda.synchronize()
    t = trace.cpu().tolist()
    print(f'== block {blk}: chunk: ' + ' | '.join(names) + ' (cycles since the first stamp; deltas)')
    t0 = t[512]
    for g in range(12):
        v = [t[512 + 8 * g + j] - t0 for j in range(7)]
        print(f'  {g:2d} {v[0]:7d} ' + ' '.join(f'{v[j] - v[j - 1]:6d}' for j in range(1, 7)))
