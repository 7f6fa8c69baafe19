"""Stream -> GPU sharding (the B200 replacement of run_pool's one-process-per-stream fan-out,
find_motion.py:1054-1122).  Streams are independent, so there is no data-path collective: rank r
owns a contiguous block of streams in its own context; the only exchange is a gather of the small
per-frame stats structs so that rank 0 can report for the whole box."""
from __future__ import annotations

import numpy as np


def shard_streams(n_streams: int, world: int, rank: int):
    """Contiguous, balanced blocks: the first (n % world) ranks get one extra stream."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_streams, world)
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))


def gather_stats(local_stats: np.ndarray, n_streams: int, dist=None):
    """All ranks pass their [s_local, T] structured stats array; every rank gets [n_streams, T] in
    global stream order.  Works with any torch.distributed backend (gloo tensors on CPU, nccl on GPU)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_stats
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    T = local_stats.shape[1]
    counts = [len(shard_streams(n_streams, world, r)) for r in range(world)]
    width = local_stats.dtype.itemsize // 4
    pad = max(counts)
    buf = np.zeros((pad, T, width), np.int32)
    buf[:counts[rank]] = local_stats.view(np.int32).reshape(counts[rank], T, width)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.from_numpy(buf).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.concatenate([p.cpu().numpy()[:counts[r]] for r, p in enumerate(parts)], axis=0)
    return np.ascontiguousarray(out).view(local_stats.dtype).reshape(n_streams, T)
