"""Debug aid: repeatability of the tensor-core 1x1 kernel on the block-3 shapes at the benchmark batch."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from cosypose_b200.engine import Engine
eng = Engine(0, max_batch=1)
dev = eng.device
gen = torch.Generator().manual_seed(0)
CASES = [(307200, 32, 192, 'expand'), (307200, 192, 32, 'project'), (1228800, 24, 144, 'expand'), (307200, 144, 32, 'p_nores')]
for dbg in (0, 1, 2):
    eng.set_option('tc_dbg', dbg)
    eng.set_option('tc_groups', 1)
    for M, K, N, kind in CASES:
        rows = 4800
        A = torch.randn((M, K), generator=gen).to(dev)
        W = torch.randn((N, K), generator=gen) / np.sqrt(K)
        bias = torch.randn(N, generator=gen)
        gate = torch.rand((M // rows, K), generator=gen).to(dev) if kind != 'expand' else None
        resid = torch.randn((M, N), generator=gen).to(dev) if kind == 'project' else None
        first = eng.debug_pointwise(1, A, W, bias, gate, rows, resid, kind == 'expand').clone()
        nbad = []
        for _ in range(10):
            again = eng.debug_pointwise(1, A, W, bias, gate, rows, resid, kind == 'expand')
            d = (first != again)
            if d.any():
                rows_bad = d.any(dim=1).nonzero().flatten()
                nbad.append((int(d.sum()), int(rows_bad[0]), int(rows_bad[-1]), int(d.any(dim=0).nonzero().flatten()[0])))
        print('dbg', dbg, (M, K, N, kind), 'mismatching repeats:', nbad if nbad else 'none')
