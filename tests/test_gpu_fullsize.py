"""GPU: the hot path at BASELINE.json's full C2 size (100k 5-kb segments x 50 candidates) through
size-independent properties -- the placements must not depend on how the work is executed:
exact band vs full DP matrices, look-ahead budget, chunking of the batch (small staging arena),
number of concurrent sub-batch pipelines; and the reference binary agrees on a sample (bench.py
checks that on every run, here only the invariances)."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

sys.path.insert(0, ol.ROOT)
FIELDS = ["qrstart", "qrstop", "lower_node", "upper_node", "rtax_node", "support", "ival", "kind", "n_pass0", "n_pass1",
          "n_pass2", "cells"]


def _same(a, b):
    return all(np.array_equal(a[f], b[f]) for f in FIELDS)


def test_c2_full_size_execution_invariance(ctx):
    import bench
    fd = bench.Flat(bench.make_data("c2", 20261017))
    assert len(fd.segs) == 100000
    ctx.load_taxonomy(fd.parent, fd.left, fd.right, fd.depth, 0)
    ctx.load_store(0, 0, fd.q_chars, fd.q_off, fd.q_len)
    ctx.load_store(1, 0, fd.r_chars, fd.r_off, fd.r_len)
    try:
        ctx.profile_reset()
        base = ctx.predict_batch(fd.segs, fd.cands)
        prof = ctx.profile()
        assert (base["kind"] == 3).all()
        cells = int(base["cells"].sum())
        # the band executes a fraction of the reference's cells, and its thresholds are true upper bounds
        assert 0 < prof["cells_edit_distance"] < 0.3 * cells
        # thresholds from the mismatch profile are true upper bounds: the only re-runs are wedges whose certificate
        # failed (a fraction of a per mille of the pairs)
        assert prof["band_retries"] - prof["wedge_failures"] <= 10
        assert prof["wedge_failures"] <= 0.001 * prof["pairs"]
        # 1. no look-ahead, two pipelines, an arena that forces several chunks
        ctx.set_lookahead(0)
        ctx.set_tuning("pipes", 2)
        ctx.set_arena_bytes(3 << 30)
        alt = ctx.predict_batch(fd.segs, fd.cands)
        assert _same(base, alt)
        ctx.set_arena_bytes(0)
        ctx.set_tuning("pipes", 1)
        ctx.set_lookahead(-1)
        # 2. full DP matrices on a tenth of the batch (every 10th segment's candidates, re-based)
        pick = np.arange(0, len(fd.segs), 10)
        segs = fd.segs[pick].copy()
        cands = np.concatenate([fd.cands[s["cand_begin"]:s["cand_begin"] + s["cand_count"]] for s in fd.segs[pick]])
        segs["cand_begin"] = np.concatenate([[0], np.cumsum(segs["cand_count"])[:-1]])
        ctx.set_band(0)
        ctx.profile_reset()
        full = ctx.predict_batch(segs, cands)
        assert ctx.profile()["cells_edit_distance"] >= int(full["cells"].sum())
        assert _same(base[pick], full)
    finally:
        ctx.set_band(1)
        ctx.set_lookahead(-1)
        ctx.set_tuning("pipes", 1)
        ctx.set_arena_bytes(0)


@pytest.mark.parametrize("workload,n_queries", [("c4", 6000), ("c5", 400)])
def test_c4_c5_regime_band_wedge_vs_full_matrices(ctx, workload, n_queries):
    """BASELINE.json configs[3] (mixed 0.5-20 kb) and configs[4] (10-50 kb reads, 5 % substitutions + 10 %
    indels) on a sample: the planned band + certified wedge, the plain band and the full DP matrices give the
    same placements, counters and cells -- the wedge certificate and the verify-and-widen loop at the lengths and
    divergences where they matter."""
    import bench
    fd = bench.Flat(bench.make_data(workload, 20261017, n_queries=n_queries))
    ctx.load_taxonomy(fd.parent, fd.left, fd.right, fd.depth, 0)
    ctx.load_store(0, 0, fd.q_chars, fd.q_off, fd.q_len)
    ctx.load_store(1, 0, fd.r_chars, fd.r_off, fd.r_len)
    try:
        ctx.profile_reset()
        base = ctx.predict_batch(fd.segs, fd.cands)
        prof = ctx.profile()
        cells = int(base["cells"].sum())
        assert (base["kind"] == 3).mean() > 0.9 and cells > 0
        assert 0 < prof["cells_edit_distance"] < 0.5 * cells
        ctx.set_tuning("wedge", 0)
        ctx.profile_reset()
        plain = ctx.predict_batch(fd.segs, fd.cands)
        assert ctx.profile()["wedge_failures"] == 0
        assert _same(base, plain)
        ctx.set_band(0)
        ctx.profile_reset()
        full = ctx.predict_batch(fd.segs, fd.cands)
        assert ctx.profile()["cells_edit_distance"] >= cells
        assert _same(base, full)
    finally:
        ctx.set_band(1)
        ctx.set_tuning("wedge", 1)
