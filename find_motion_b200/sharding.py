"""Static stream -> GPU map for one-process-per-GPU runs (bench.py under torchrun, multi-rank drivers):
stream s belongs to rank s mod G (streams are independent, so there is no data-path
collective -- the reference's analogue is one Pool worker per stream, find_motion.py:1071-1075).  The only
exchange is a gather of the small per-frame stats structs so that rank 0 can report motion flags for the
whole box (gloo on CPU, NCCL on GPUs).  File / camera jobs inside one process are spread over the GPUs
dynamically by find_motion_b200.jobs instead."""
from __future__ import annotations

import numpy as np


def shard_streams(n_streams: int, world: int, rank: int):
    """Round robin: rank r owns streams r, r + G, r + 2G, ...  (sizes differ by at most one)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return list(range(rank, n_streams, world))


def gather_stats(local_stats: np.ndarray, n_streams: int, dist=None, device=None):
    """All ranks pass their [s_local, T] structured stats array (rows in shard_streams order); every rank gets
    [n_streams, T] in global stream order.  Works with any torch.distributed backend."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_stats
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    T = local_stats.shape[1]
    counts = [len(shard_streams(n_streams, world, r)) for r in range(world)]
    width = local_stats.dtype.itemsize // 4
    pad = max(counts)
    buf = np.zeros((pad, T, width), np.int32)
    buf[:counts[rank]] = np.ascontiguousarray(local_stats).view(np.int32).reshape(counts[rank], T, width)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    mine = torch.from_numpy(buf).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.zeros((n_streams, T, width), np.int32)
    for r, p in enumerate(parts):
        out[shard_streams(n_streams, world, r)] = p.cpu().numpy()[:counts[r]]
    return np.ascontiguousarray(out).view(local_stats.dtype).reshape(n_streams, T)
