"""Measures the error of the 1x1 kernels against float64 (GPU only): all-positive operands expose any
accumulation bias, normal operands give the typical error."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from cosypose_b200.engine import Engine
eng = Engine(0, max_batch=1)
dev = eng.device
gen = torch.Generator().manual_seed(0)
for K, N in ((24, 144), (192, 32), (384, 1536), (1392, 232), (2304, 384)):
    M = 4096
    for kind in ('positive', 'normal'):
        A = torch.randn((M, K), generator=gen)
        W = torch.randn((N, K), generator=gen) / K ** 0.5
        if kind == 'positive':
            A, W = A.abs() + 0.1, (W.abs() + 0.1 / K ** 0.5)
        bias = torch.zeros(N)
        ref = A.double() @ W.double().t()
        ref = ref * torch.sigmoid(ref)
        row = []
        for name, impl in (('cuda-core fp32', 0), ('tcgen05 3xTF32', 1)):
            out = eng.debug_pointwise(impl, A.to(dev), W, bias, swish=True).cpu().double()
            err = (out - ref) / ref.abs().max()
            row.append(f'{name}: max {err.abs().max():.2e} mean {err.mean():+.2e}')
        print(f'K={K:5d} N={N:5d} {kind:9s}| ' + ' | '.join(row))
