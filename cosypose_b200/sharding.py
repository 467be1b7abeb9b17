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


# Per-hypothesis record exchanged after the last iteration: every tensor field of one iteration's collection.
RECORD_FIELDS = (('poses', 16), ('poses_input', 16), ('K_crop', 9), ('boxes_rend', 4), ('boxes_crop', 4))
RECORD_FLOATS = sum(n for _, n in RECORD_FIELDS)


def pack_records(collections):
    """[n_local, n_collections * RECORD_FLOATS] float32 from a list of per-iteration collections (missing fields,
    e.g. of an externally supplied initialisation, are zero)."""
    cols = []
    for c in collections:
        n = len(c)
        for name, width in RECORD_FIELDS:
            if hasattr(c, name):
                cols.append(getattr(c, name).reshape(n, width).float())
            else:
                cols.append(torch.zeros((n, width), dtype=torch.float32, device=c.poses.device))
    return torch.cat(cols, dim=1).contiguous()


def unpack_records(records, n_collections):
    """Inverse of pack_records: list of dicts field -> tensor."""
    out, col = [], 0
    shapes = dict(poses=(4, 4), poses_input=(4, 4), K_crop=(3, 3), boxes_rend=(4,), boxes_crop=(4,))
    for _ in range(n_collections):
        d = {}
        for name, width in RECORD_FIELDS:
            d[name] = records[:, col:col + width].reshape((-1,) + shapes[name]).contiguous()
            col += width
        out.append(d)
    return out


def gather_records(local, n_total, engine=None):
    """One all-gather of the ranks' record blocks -> [n_total, rec] on every rank.  On GPUs with an engine whose
    communicator is up the exchange is libcosyb200's ncclAllGather, written in place into the gathered buffer;
    otherwise (CPU tests, gloo) torch.distributed's all_gather_into_tensor."""
    rank, ws = world()
    if ws == 1:
        return local
    per = -(-n_total // ws)
    rec = local.shape[1]
    out = torch.zeros((ws * per, rec), dtype=local.dtype, device=local.device)
    out[rank * per: rank * per + local.shape[0]] = local
    if engine is not None and getattr(engine, 'nccl_ready', False):
        engine.allgather_candidates(out)                     # in place: this rank's rows are already there
    elif out.is_cuda and dist.get_backend() == 'gloo':          # test rigs without one GPU per rank
        host = out.cpu()
        dist.all_gather_into_tensor(host, host[rank * per:(rank + 1) * per].clone())
        out.copy_(host)
    else:
        dist.all_gather_into_tensor(out, out[rank * per:(rank + 1) * per].clone())
    return out[:n_total]
