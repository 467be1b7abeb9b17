"""Bring-up / tuning aid for the 3xFP16 1x1 kernel (kernels_pw2.cuh), GPU only: error against float64 and device
time per trunk shape for the three implementations.

    python tools/pw2_check.py [--batch 64] [--quick]
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from cosypose_b200.engine import Engine
from cosypose_b200 import effnet_spec as spec

HW = {}
for i, (name, hh, ww, c) in enumerate(spec.activation_shapes()[1:-1]):
    HW[i] = (hh, ww)   # index 0 = stem output = input of block 0

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--quick', action='store_true')
ap.add_argument('--impls', default='0,1,2')
args = ap.parse_args()
impls = [int(x) for x in args.impls.split(',')]
eng = Engine(0, max_batch=1)
dev = eng.device
gen = torch.Generator().manual_seed(0)


def ref64(A, W, bias, gate, rows, resid, swish):
    A64 = A.double()
    if gate is not None:
        A64 = A64 * gate.double().repeat_interleave(rows, dim=0)[:A.shape[0]]
    y = A64 @ W.double().t() + bias.double()
    if swish:
        y = y * torch.sigmoid(y)
    if resid is not None:
        y = y + resid.double()
    return y


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return min(a.elapsed_time(b) for a, b in evs) * 1e3


# small correctness cases first (ragged M, every kind)
print('== small cases (M = 227) ==')
for K, N, kind in ((24, 144, 'expand'), (40, 24, 'project'), (192, 32, 'project'), (96, 576, 'expand'),
                   (816, 136, 'project'), (1392, 232, 'project'), (384, 1536, 'head'), (2304, 384, 'project')):
    M, rows = 227, 70
    A = torch.randn((M, K), generator=gen)
    W = torch.randn((N, K), generator=gen) / np.sqrt(K)
    bias = torch.randn(N, generator=gen)
    gate = torch.rand((4, K), generator=gen) if kind == 'project' else None
    resid = torch.randn((M, N), generator=gen) if kind == 'project' else None
    swish = kind != 'project'
    ref = ref64(A, W, bias, gate, rows, resid, swish)
    line = f'K={K:5d} N={N:5d} {kind:8s}'
    for impl in impls:
        out = eng.debug_pointwise(impl, A.to(dev), W, bias, gate.to(dev) if gate is not None else None, rows,
                                  resid.to(dev) if resid is not None else None, swish)
        err = (out.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
        line += f' | impl{impl} err {err:.2e}'
    print(line, flush=True)

print('== accumulation bias (all-positive operands, mean signed relative error) ==')
for K, N in ((32, 192), (192, 32), (1392, 232), (2304, 384)):
    M = 1024
    A = torch.randn((M, K), generator=gen).abs() + 0.1
    W = (torch.randn((N, K), generator=gen).abs() + 0.1) / np.sqrt(K)
    bias = torch.zeros(N)
    ref = A.double() @ W.double().t()
    line = f'K={K:5d} N={N:5d}'
    for impl in impls:
        out = eng.debug_pointwise(impl, A.to(dev), W, bias, swish=False).cpu().double()
        rel = (out - ref) / ref
        line += f' | impl{impl} mean {rel.mean():+.2e} max {rel.abs().max():.2e}'
    print(line, flush=True)

if args.quick:
    sys.exit(0)
print(f'== trunk layers at batch {args.batch}: device us per launch ==')
tot = {i: 0.0 for i in impls}
layers = []
for i, b in enumerate(spec.BLOCKS):
    if b.e != 1:
        layers.append((f'b{i} expand', args.batch * HW[i][0] * HW[i][1], b.cin, b.cexp, 'expand', 1))
    layers.append((f'b{i} project', args.batch * HW[i + 1][0] * HW[i + 1][1], b.cexp, b.cout,
                   'project_res' if b.skip else 'project', HW[i + 1][0] * HW[i + 1][1]))
layers.append(('head', args.batch * 70, 384, 1536, 'expand', 1))
for name, M, K, N, kind, rows in layers:
    A = torch.randn((M, K), generator=gen).to(dev)
    W = torch.randn((N, K), generator=gen) / np.sqrt(K)
    bias = torch.randn(N, generator=gen)
    gate = torch.rand((-(-M // rows), K), generator=gen).to(dev) if kind.startswith('project') else None
    resid = torch.randn((M, N), generator=gen).to(dev) if kind == 'project_res' else None
    swish = kind == 'expand'
    line = f'{name:12s} M={M:8d} K={K:5d} N={N:5d}'
    outs = {}
    for impl in impls:
        # debug_pointwise packs + uploads the weights on every call and synchronises: time the kernel with the
        # engine's own per-launch events instead
        eng.profile_read(reset=True)
        eng.profile_enable(True)
        for _ in range(3):
            outs[impl] = eng.debug_pointwise(impl, A, W, bias, gate, rows, resid, swish)
        r = eng.profile_read(reset=True)
        eng.profile_enable(False)
        us = r['expand_1x1'][1] / 3 * 1e3
        tot[impl] += us
        line += f' | impl{impl} {us:7.1f}'
    if 0 in outs and 2 in outs:
        d = (outs[2] - outs[0]).abs().max().item() / outs[0].abs().max().item()
        line += f' | 2 vs 0: {d:.1e}'
    print(line, flush=True)
print('total us per forward batch: ' + ', '.join(f'impl{i} {tot[i]:.0f}' for i in impls))
