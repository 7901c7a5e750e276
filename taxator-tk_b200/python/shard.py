"""Sharding of a segment batch across GPUs/ranks: contiguous segment ranges balanced
by an estimate of the DP work (1 + sum over candidates of span^2), candidates rebased per shard; results
are concatenated in rank order, which restores the input order.  There is no collective on the data
path -- ranks only exchange fixed-size result records at the end (host gather)."""
import numpy as np


def work_estimate(segs, cands):
    """1 + sum of span^2 over a segment's records (numpy mirror of trpa_shard_bounds' weight; exact in int64
    up to 2^63, which no real table reaches)."""
    span = np.abs(cands["rstop"].astype(np.int64) - cands["rstart"].astype(np.int64)) + 1
    csum = np.concatenate([[0], np.cumsum(span * span)])
    b = segs["cand_begin"].astype(np.int64)
    e = b + segs["cand_count"].astype(np.int64)
    return csum[e] - csum[b] + 1


def shard_bounds(segs, cands, world):
    """world+1 segment indices; shard r = segs[bounds[r]:bounds[r+1]].  This is the library's host-only
    helper trpa_shard_bounds -- the same cuts the C++ host (RPAPredictionModelGPU::predictFlat) makes."""
    import rpa_b200
    return rpa_b200.shard_bounds(segs, cands, world)


def take_shard(segs, cands, world, rank):
    b = shard_bounds(segs, cands, world)
    s = segs[b[rank]:b[rank + 1]].copy()
    if len(s) == 0:
        return s, cands[:0].copy(), b
    c0 = int(s["cand_begin"][0])
    c1 = int(s["cand_begin"][-1]) + int(s["cand_count"][-1])
    s["cand_begin"] -= c0
    return s, cands[c0:c1].copy(), b
