"""Multi-GPU layout of the path: hypotheses are independent through every refinement iteration
(no cross-hypothesis op in reference models/pose.py:89-132), so ranks take contiguous shards and
exchange nothing until the refined poses are collected with ONE all-gather.
The reference gathers per-rank predictions through pickle files on a shared filesystem
(utils/tensor_collection.py:142-163).  One process per GPU; NCCL on GPUs, gloo in CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, world_size):
    """Contiguous chunk [start, stop) of ceil(n / world) entries for `rank` (may be empty)."""
    per = -(-n // world_size)
    start = min(n, rank * per)
    return start, min(n, start + per)


def gather_poses(local, n_total=None):
    """All-gather fixed-size per-hypothesis records [n_local, ...] -> [n_total, ...] on every rank.
    Shards are padded to ceil(n_total / world) rows so a single equal-count collective suffices."""
    rank, ws = world()
    if ws == 1:
        return local
    n_local = local.shape[0]
    if n_total is None:
        per = n_local
        n_total = per * ws
    else:
        per = -(-n_total // ws)
    send = local
    if n_local < per:
        pad = torch.zeros((per - n_local,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send = torch.cat((local, pad), dim=0)
    out = torch.empty((ws * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, send.contiguous())
    return out[:n_total]
