"""CPU: the N>1 host logic (stream sharding + stats gather) on a world-size-2 gloo group."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_streams_partitions():
    from find_motion_b200.sharding import shard_streams
    for n in (1, 7, 8, 64, 65):
        for world in (1, 2, 4, 8):
            got = sorted(s for r in range(world) for s in shard_streams(n, world, r))
            assert got == list(range(n))
            assert all(s % world == r for r in range(world) for s in shard_streams(n, world, r))
            sizes = [len(shard_streams(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_streams, T, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from find_motion_b200.engine import STATS_DTYPE
    from find_motion_b200.sharding import gather_stats, shard_streams
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = shard_streams(n_streams, world, rank)
    local = np.zeros((len(mine), T), STATS_DTYPE)
    for i, s in enumerate(mine):
        for t in range(T):
            local[i, t] = tuple(1000 * s + 10 * t + k for k in range(8))
    full = gather_stats(local, n_streams, dist)
    ok = full.shape == (n_streams, T) and all(
        tuple(full[s, t]) == tuple(1000 * s + 10 * t + k for k in range(8)) for s in range(n_streams) for t in range(T))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


@pytest.mark.timeout(120)
def test_gather_stats_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]
