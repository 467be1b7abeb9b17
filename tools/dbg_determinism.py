"""Debug aid: repeats the trunk forward at the benchmark batch and reports which internal tensor of a block
(expanded activation, depthwise output, gate, block output) is not bit-identical between repeats.

    python tools/dbg_determinism.py                       # shipped kernels: tc_groups 1, 1, auto, auto
    OPTS="tc_tma=1;tc_tma=1,tc_dbg=4;tc_tma=1,tc_dbg=1" REPS=12 BLK=3 python tools/dbg_determinism.py
                                                          # the experimental TMA variant and its experiment bits
                                                          # (kernels_tc_tma.cuh), one engine per ';'-separated set
"""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import torch
from helpers import state_dict
from cosypose_b200 import _lib
if os.environ.get('NODUMP'):
    _lib._SIGS.pop('cosyb200_debug_dump', None)
from cosypose_b200.engine import Engine, _ptr
dev = torch.device('cuda', 0)
B, BLK = 64, int(os.environ.get('BLK', 3))
gen = torch.Generator().manual_seed(0)
crops = torch.rand((B, 3, 240, 320), generator=gen).to(dev)
renders = torch.rand((B, 3, 240, 320), generator=gen).to(dev)
print('lib', os.environ.get('COSYB200_LIB'))
OPT_SETS = ({'tc_groups': 1}, {'tc_groups': 1}, {}, {})
if os.environ.get('OPTS'):
    OPT_SETS = tuple({kv.split('=')[0]: int(kv.split('=')[1]) for kv in part.split(',') if kv}
                     for part in os.environ['OPTS'].split(';'))
for opts in OPT_SETS:
    eng = Engine(0, max_batch=B)
    eng.load_pose_model(0, state_dict(0))
    for k, v in opts.items():
        eng.set_option(k, v)
    e = torch.zeros(B * 60 * 80 * 192, device=dev); d = torch.zeros_like(e); gt = torch.zeros(B * 192, device=dev)
    if not os.environ.get('NODUMP'):
        _lib.check(eng._L.cosyb200_debug_dump(eng._h, BLK, _ptr(e), _ptr(d), _ptr(gt)))
    p0, t0 = eng.net_forward(0, crops, renders, taps=True)
    ref = dict(e=e.clone(), d=d.clone(), g=gt.clone(), y=t0[f'block{BLK}'].clone(), x=t0[f'block{BLK - 1}'].clone())
    for rep in range(int(os.environ.get('REPS', 4))):
        p1, t1 = eng.net_forward(0, crops, renders, taps=True)
        cur = dict(e=e, d=d, g=gt, y=t1[f'block{BLK}'], x=t1[f'block{BLK - 1}'])
        msg = []
        for k in ('x', 'e', 'd', 'g', 'y'):
            ne = (ref[k] != cur[k])
            if ne.any():
                idx = ne.flatten().nonzero().flatten()
                width = {'e': 192, 'd': 192, 'g': 192, 'y': 32, 'x': 32}[k]
                rows = torch.unique(idx // width)
                msg.append(f'{k}: {int(ne.sum())} elems, {rows.numel()} rows, first rows {rows[:6].tolist()} last {rows[-3:].tolist()}')
        if msg or rep == int(os.environ.get('REPS', 4)) - 1:
            print(opts, 'rep', rep, msg if msg else 'all identical')
    for k in opts:
        eng.set_option(k, 0)          # the GEMM options are process-wide
    eng.close()
