"""Prints the cycle trace of CTA 0 of the tensor-core 1x1 kernel for a few layer shapes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from cosypose_b200.engine import Engine, _ptr
from cosypose_b200 import _lib
eng = Engine(0, max_batch=1)
dev = eng.device
trace = torch.zeros(32, dtype=torch.int64, device=dev)
_lib.check(eng._L.cosyb200_debug_trace(eng._h, _ptr(trace)))
gen = torch.Generator().manual_seed(0)
CASES = ((128 * 148 * 8, 192, 32, True, False), (19200, 816, 136, True, False), (19200, 136, 816, False, True),
         (4480, 1392, 232, True, False), (4480, 232, 1392, False, True))
for M, K, N, gate, sw in CASES:
    A = torch.randn((M, K), generator=gen).to(dev)
    W = torch.randn((N, K), generator=gen) / K ** 0.5
    g = torch.rand((-(-M // 70), K), generator=gen).to(dev) if gate else None
    for rep in range(2):
        trace.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        eng.debug_pointwise(1, A, W, torch.zeros(N), gate=g, rows_per_img=70, swish=sw)
        e1.record()
        torch.cuda.synchronize()
    t = trace.cpu().tolist()
    rel = lambda i: (t[i] - t[0]) if t[i] else None
    print(f'M={M} K={K} N={N} gate={gate} swish={sw}: setup_done {rel(1)} drain_done {rel(2)} exit {rel(3)} cycles')
    print('   producer g=6: top, cp.async landed, bar, issued next, split done, slot free, sttm done, arrived :', [rel(8 + i) for i in range(8)])
    print('   mma commit      :', [rel(16 + i) for i in range(8)])
    print('   drain g=6: top, acc full, loaded+added, arrived :', [rel(24 + i) for i in range(4)])
    print('   epilogue tile 1: start', rel(28), 'staged', rel(29), 'stored', rel(30))
    print('   mma g=6: top', rel(4), 'waits done', rel(5), 'mmas issued', rel(6), 'commits issued', rel(22))
