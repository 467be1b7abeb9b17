"""Device time of the distinct 1x1-convolution shapes of blocks 9-25 + head at the benchmark batch for every n-tile
column count of k_pw2 (engine option pw2_nt), next to the cost model's own choice (pw2_nt = 0).

    python tools/pw2_tune.py [--batch 64]
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from cosypose_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--only', default='')
args = ap.parse_args()
eng = Engine(0, max_batch=1)
dev = eng.device
gen = torch.Generator().manual_seed(0)
B = args.batch
shapes = [('b9 expand', B * 300, 96, 576, 'expand', 1), ('b14 expand', B * 300, 136, 816, 'expand', 1),
          ('b19 expand', B * 70, 232, 1392, 'expand', 1), ('b25 expand', B * 70, 384, 2304, 'expand', 1),
          ('head', B * 70, 384, 1536, 'expand', 1),
          ('b8 project', B * 300, 288, 96, 'project', 300), ('b9 project', B * 300, 576, 96, 'project_res', 300),
          ('b13 project', B * 300, 576, 136, 'project', 300), ('b14 project', B * 300, 816, 136, 'project_res', 300),
          ('b18 project', B * 70, 816, 232, 'project', 70), ('b19 project', B * 70, 1392, 232, 'project_res', 70),
          ('b24 project', B * 70, 1392, 384, 'project', 70), ('b25 project', B * 70, 2304, 384, 'project_res', 70)]
for name, M, K, N, kind, rows in shapes:
    if args.only and args.only not in name:
        continue
    A = torch.randn((M, K), generator=gen).to(dev)
    W = torch.randn((N, K), generator=gen) / np.sqrt(K)
    bias = torch.randn(N, generator=gen)
    gate = torch.rand((-(-M // rows), K), generator=gen).to(dev) if kind.startswith('project') else None
    resid = torch.randn((M, N), generator=gen).to(dev) if kind == 'project_res' else None
    swish = kind == 'expand'
    n16 = -(-N // 16)
    nt_min = -(-n16 * 16 // 192)
    line = f'{name:12s} M={M:6d} K={K:5d} N={N:5d} |'
    best = None
    for nt in [0, 0] + list(range(nt_min, min(n16, nt_min + 6) + 1)):
        eng.set_option('pw2_nt', nt)
        eng.profile_read(reset=True)
        eng.profile_enable(True)
        try:
            for _ in range(4):
                eng.debug_pointwise(2, A, W, bias, gate, rows, resid, swish)
        except AssertionError:          # no plan with this column count
            eng.profile_enable(False)
            continue
        r = eng.profile_read(reset=True)
        eng.profile_enable(False)
        us = r['expand_1x1'][1] / 4 * 1e3
        line += f' nt{nt}:{us:5.1f}'
        if nt and (best is None or us < best[1]):
            best = (nt, us)
    print(line + f' | best nt{best[0]} {best[1]:.1f}', flush=True)
eng.set_option('pw2_nt', 0)
