"""Batch sharding over devices / ranks.

A batch shards by segment index: contiguous ranges balanced by input bytes -- the reference's static
worker partition (c-ext/compressor.c:1127,1183-1200; c-ext/decompressor.c:1237,1290-1305) -- so output
order is preserved by concatenating the ranges in rank order.  No data-path collective is needed; ranks
only synchronise for timing (barrier + max)."""
import numpy as np


def split_ranges(lengths, parts):
    """[(lo, hi)] contiguous, non-empty, covering range(len(lengths)), balanced by sum(lengths)."""
    lengths = np.asarray(lengths, dtype=np.uint64)
    n = len(lengths)
    if parts <= 1 or n < 2:
        return [(0, n)]
    parts = min(parts, n)
    cum = np.cumsum(lengths, dtype=np.uint64)
    total = int(cum[-1])
    # one vectorised search for all targets (same dtype as `cum`: a Python int would make numpy convert the
    # whole array on every call)
    targets = np.array([total * p // parts for p in range(1, parts)], dtype=np.uint64)
    found = np.searchsorted(cum, targets, side="left")
    cuts = [0]
    for p in range(1, parts):
        k = int(found[p - 1]) + 1
        k = min(max(k, cuts[-1] + 1), n - (parts - p))
        cuts.append(k)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(parts) if cuts[i] < cuts[i + 1]]


def rank_range(lengths, rank, world):
    """The range of segment indices rank `rank` of `world` owns (possibly empty)."""
    r = split_ranges(lengths, world)
    return r[rank] if rank < len(r) else (len(lengths), len(lengths))


def max_over_ranks(value, dist=None):
    """Timing reduction used by bench.py: the slowest rank defines the step time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
