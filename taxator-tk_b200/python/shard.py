"""Sharding of a segment batch across GPUs/ranks: contiguous segment ranges balanced
by an estimate of the DP work (sum over candidates of span^2), candidates rebased per shard; results
are concatenated in rank order, which restores the input order.  There is no collective on the data
path -- ranks only exchange fixed-size result records at the end (host gather)."""
import numpy as np


def work_estimate(segs, cands):
    span = np.abs(cands["rstop"].astype(np.int64) - cands["rstart"].astype(np.int64)) + 1
    w = (span * span).astype(np.float64)
    csum = np.concatenate([[0.0], np.cumsum(w)])
    b = segs["cand_begin"].astype(np.int64)
    e = b + segs["cand_count"].astype(np.int64)
    return csum[e] - csum[b] + 1.0


def shard_bounds(segs, cands, world):
    """world+1 segment indices; shard r = segs[bounds[r]:bounds[r+1]]."""
    n = len(segs)
    if n == 0:
        return [0] * (world + 1)
    w = np.cumsum(work_estimate(segs, cands))
    total = w[-1]
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(w, total * r / world, side="left")) + 1 if total > 0 else n * r // world)
    bounds.append(n)
    bounds = [min(max(b, 0), n) for b in bounds]
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def take_shard(segs, cands, world, rank):
    b = shard_bounds(segs, cands, world)
    s = segs[b[rank]:b[rank + 1]].copy()
    if len(s) == 0:
        return s, cands[:0].copy(), b
    c0 = int(s["cand_begin"][0])
    c1 = int(s["cand_begin"][-1]) + int(s["cand_count"][-1])
    s["cand_begin"] -= c0
    return s, cands[c0:c1].copy(), b
