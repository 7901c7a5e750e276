"""CPU: the block-generated full-size batches of bench.py (BASELINE.json configs[3] / [4]) do not depend on the
number of ranks that generate them, and the shards are contiguous, re-based and balanced by estimated DP work."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol

sys.path.insert(0, ol.ROOT)


@pytest.mark.parametrize("name", ["c4", "c5"])
def test_sharded_batch_is_independent_of_world_size(name):
    import bench
    cat = {}
    for world in (2, 4):
        shards = [bench.make_shard(name, 11, rank, world, workers=2, scale=0.002) for rank in range(world)]
        for sh in shards:
            if len(sh.segs):
                assert sh.segs["cand_begin"][0] == 0
                assert int(sh.segs["cand_begin"][-1] + sh.segs["cand_count"][-1]) == len(sh.cands)
                assert (sh.segs["query_seq"] == np.arange(len(sh.segs))).all()
                assert int(sh.q_off[-1]) + int(sh.q_len[-1]) == len(sh.q_chars)
            assert sh.bounds[0] == 0 and sh.bounds[-1] == sh.total_segments
        assert sum(len(sh.segs) for sh in shards) == shards[0].total_segments
        work = [sh.est_work for sh in shards]
        assert max(work) / (sum(work) / len(work)) < 1.05
        cat[world] = [np.concatenate([getattr(sh, f) for sh in shards]) for f in ("q_len", "q_chars", "cands")]
    for a, b in zip(cat[2], cat[4]):
        assert (a == b).all()
