"""N-rank == 1-rank identity of the sharded single-view path on real GPUs: two
processes launched with torch.distributed.run, each refining its contiguous shard of ONE global detection table and
exchanging the records with one all-gather; every rank must return bit-for-bit what a single rank computes.
With >= 2 GPUs: one rank per GPU, NCCL, libcosyb200's cosyb200_allgather_candidates (in place).  On a 1-GPU box both
ranks share cuda:0 and the exchange goes through gloo (NCCL refuses two ranks on one device)."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

_WORKER = textwrap.dedent('''
    import os, sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + '/tests')
    from helpers import Workload, build_predictor
    from cosypose_b200.utils import tensor_collection as tc
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    n_gpu = torch.cuda.device_count()
    multi = n_gpu >= world
    dev_index = rank if multi else 0
    torch.cuda.set_device(dev_index)
    dev = torch.device('cuda', dev_index)
    dist.init_process_group('nccl' if multi else 'gloo', **(dict(device_id=dev) if multi else {}))
    w = Workload(3, 5, 7, 1, 2)                      # 15 hypotheses: uneven shards (8 + 7)
    from cosypose_b200.sharding import shard_bounds
    a, b = shard_bounds(w.n, rank, world)
    # single-rank result (full batch on this rank)
    pred, eng, views = build_predictor(w, dev_index, bsz_objects=8)
    det = tc.PandasTensorCollection(infos=w.infos(), bboxes=w.boxes.to(dev))
    full, full_preds = pred.get_predictions(w.images.to(dev), w.K.to(dev), detections=det, n_coarse_iterations=1,
                                            n_refiner_iterations=2)
    eng.close()
    # sharded: the renderer of each rank replays the views of ITS hypotheses
    w.views_c, w.views_r = w.views_c[:, a:b].contiguous(), w.views_r[:, a:b].contiguous()
    pred, eng, views = build_predictor(w, dev_index, bsz_objects=8)
    if multi:
        eng.nccl_init()
    out, preds = pred.get_predictions(w.images.to(dev), w.K.to(dev), detections=det, n_coarse_iterations=1,
                                      n_refiner_iterations=2, shard=True)
    torch.cuda.synchronize()
    assert len(out) == w.n and list(preds.keys()) == list(full_preds.keys())
    # chunking differs between the two runs (8+7 vs 8+... of 15), which changes GEMM tiling only: hypotheses are
    # independent, so the sharded result must equal the full-batch result of the same rows to rounding, and the
    # ranks must agree with each other bit for bit
    for k in full_preds:
        for f in ('poses', 'K_crop', 'boxes_rend', 'boxes_crop'):
            d = (getattr(preds[k], f) - getattr(full_preds[k], f)).abs().max().item()
            assert d < 1e-5, (k, f, d)
    mine = out.poses.cpu()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    if multi:
        g = [torch.empty_like(out.poses) for _ in range(world)]
        dist.all_gather(g, out.poses.contiguous())
        gathered = [x.cpu() for x in g]
    else:
        dist.all_gather(gathered, mine)
    assert all(torch.equal(gathered[0], x) for x in gathered)
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.write('rank%dok(%s) ' % (rank, 'nccl' if multi else 'gloo'))
    sys.stdout.flush()
''')


def test_sharded_equals_single_rank(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', OMP_NUM_THREADS='1')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29761', str(script), str(ROOT)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'rank0ok' in r.stdout and 'rank1ok' in r.stdout


def _forward(eng, dev, seed=3, B=4):
    gen = torch.Generator().manual_seed(seed)
    crops = torch.rand((B, 3, 240, 320), generator=gen).to(dev)
    renders = torch.rand((B, 3, 240, 320), generator=gen).to(dev)
    return eng.net_forward(0, crops, renders)


def test_two_handles_in_one_process_keep_their_own_state():
    """Handles are independent (include/cosyb200.h): options set on one handle do not leak into another, and - with
    two GPUs - a second handle on another device gets its own shared-memory opt-ins and SM count."""
    from helpers import state_dict
    from cosypose_b200.engine import Engine
    a, b = Engine(0, max_batch=4), Engine(0, max_batch=4)
    for e in (a, b):
        e.load_pose_model(0, state_dict(0))
    a.set_option('gemm_impl', 0)
    a.set_option('xdw', 0)
    ya1 = _forward(a, a.device)
    yb = _forward(b, b.device)            # default options: tcgen05 3xFP16 + fused kernel
    ya2 = _forward(a, a.device)
    assert torch.equal(ya1, ya2)          # b's launches did not change a's configuration
    assert not torch.equal(ya1, yb) and (ya1 - yb).abs().max() < 1e-4 * ya1.abs().max()   # raw head outputs, O(100)
    if torch.cuda.device_count() >= 2:
        c = Engine(1, max_batch=4)
        c.load_pose_model(0, state_dict(0))
        yc = _forward(c, c.device)
        assert torch.equal(yb.cpu(), yc.cpu())      # same kernels, same inputs, another device
        yb2 = _forward(b, b.device)
        assert torch.equal(yb, yb2)
        c.close()
    a.close()
    b.close()
