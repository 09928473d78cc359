"""The N>1 path on CPU: two gloo ranks shard one batch by segment index (no data-path collective),
each decodes its range (with the oracle standing in for the device), and the concatenation in rank
order equals the whole batch.  Also checks the max-over-ranks timing reduction bench.py uses."""
import os
import socket

import numpy as np
import pytest

from python_zstandard_b200.sharding import rank_range, split_ranges
from tests import helpers


def test_split_ranges_properties():
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 17, 1000):
        lengths = rng.integers(1, 5000, n)
        for parts in (1, 2, 3, 8):
            r = split_ranges(lengths, parts)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(lo < hi for lo, hi in r)
            assert len(r) <= min(parts, n)
    # balanced by bytes, not by count
    r = split_ranges([1000, 1, 1, 1, 1, 1, 1, 1000], 2)
    assert r == [(0, 1), (1, 8)] or r == [(0, 2), (2, 8)] or r[0][1] <= 7


def _worker(rank, world, port, frames, raws, out_q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import Oracle
    from python_zstandard_b200.sharding import max_over_ranks
    orc = Oracle()
    lo, hi = rank_range([len(f) for f in frames], rank, world)
    mine = [orc.decompress(frames[i], len(raws[i])) for i in range(lo, hi)]
    dist.barrier()
    slow = max_over_ranks(1.0 + rank, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, mine))
    if rank == 0:
        out_q.put((slow, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    import torch.multiprocessing as mp
    vecs = [v for v in helpers.golden_vectors() if not v[3] and "nocs" not in v[0]]
    frames = [v[1] for v in vecs]
    raws = [v[2] for v in vecs]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, frames, raws, q)) for r in range(2)]
    for p in procs:
        p.start()
    slow, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert slow == 2.0                                   # max over ranks
    assert gathered[0][0] == 0 and gathered[0][1] == gathered[1][0] and gathered[1][1] == len(frames)
    joined = gathered[0][2] + gathered[1][2]
    assert joined == raws
