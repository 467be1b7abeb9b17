"""Cycle trace of CTA 0 of the 3xFP16 1x1 kernel (kernels_pw2.cuh) for a few layer shapes: which role waits for what."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from cosypose_b200.engine import Engine, _ptr
from cosypose_b200 import _lib
eng = Engine(0, max_batch=1)
dev = eng.device
trace = torch.zeros(512, dtype=torch.int64, device=dev)
_lib.check(eng._L.cosyb200_debug_trace(eng._h, _ptr(trace)))
gen = torch.Generator().manual_seed(0)
CASES = ((307200, 192, 32, True, False, True, 4800), (19200, 816, 136, True, False, True, 300),
         (19200, 136, 816, False, True, False, 1), (4480, 1392, 232, True, False, True, 70),
         (4480, 232, 1392, False, True, False, 1))
for M, K, N, gate, sw, res, rows in CASES:
    A = torch.randn((M, K), generator=gen).to(dev)
    W = torch.randn((N, K), generator=gen) / K ** 0.5
    g = torch.rand((-(-M // rows), K), generator=gen).to(dev) if gate else None
    r = torch.randn((M, N), generator=gen).to(dev) if res else None
    for rep in range(2):
        trace.zero_()
        eng.debug_pointwise(2, A, W, torch.zeros(N), gate=g, rows_per_img=rows, resid=r, swish=sw)
        torch.cuda.synchronize()
    t = trace.cpu().tolist()
    rel = lambda i: (t[i] - t[0]) if t[i] else None
    print(f'== M={M} K={K} N={N} gate={gate} swish={sw} resid={res}')
    print('producer item: top | raw landed+bar | split done | slot free, sttm, arrived')
    for i in range(16):
        print(f'  {i:2d}', [rel(64 + 4 * i + j) for j in range(4)], '| lds+gate done:', [rel(256 + 4 * i)])
    print('mma item: before fullA wait | after')
    print('  ', [(rel(128 + 2 * g_), rel(129 + 2 * g_)) for g_ in range(24)])
    print('drain group: top | acc full | loaded, arrived | epilogue end')
    for i in range(12):
        print(f'  {i:2d}', [rel(192 + 4 * i + j) for j in range(4)])
