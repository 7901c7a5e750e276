"""N>1 path on CPU: two gloo ranks each place their shard of a segment batch (host-compiled state
machine + oracle alignments standing in for the GPU) and gather the fixed-size result records; the
concatenation must equal the single-rank result in input order."""
import os
import subprocess
import sys

import numpy as np

import golden_util as gu
import oracle_lib as ol
import shard

WORKER = r'''
import os, sys, pickle
import numpy as np
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import torch.distributed as dist
import golden_util as gu, oracle_lib as ol, shard
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % PORT, rank=RANK, world_size=2)
fd = ol.FlatData(gu.case_data("nt_small"))
segs, cands, bounds = shard.take_shard(fd.segs, fd.cands, 2, RANK)
fd.segs, fd.cands = segs, cands
res, _ = ol.host_machine_predict(fd)
out = [None, None]
dist.all_gather_object(out, res.tobytes())
if RANK == 0:
    open(OUT, "wb").write(b"".join(out))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_shards_equal_single_rank(tmp_path):
    fd = ol.FlatData(gu.case_data("nt_small"))
    full, _ = ol.host_machine_predict(fd)
    b = shard.shard_bounds(fd.segs, fd.cands, 2)
    assert b[0] == 0 and b[-1] == len(fd.segs) and 0 < b[1] < len(fd.segs)
    port = 29500 + os.getpid() % 2000
    out = str(tmp_path / "gathered.bin")
    procs = []
    for rank in range(2):
        code = "ROOT=%r; PORT=%r; RANK=%d; OUT=%r\n" % (ol.ROOT, str(port), rank, out) + WORKER
        procs.append(subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE))
    for p in procs:
        o, e = p.communicate(timeout=240)
        assert p.returncode == 0, e.decode()[-2000:]
    got = np.frombuffer(open(out, "rb").read(), dtype=full.dtype)
    assert len(got) == len(full)
    assert ol.results_equal(full, got) == []


def test_shard_bounds_properties():
    fd = ol.FlatData(gu.case_data("nt_1kb"))
    for world in (1, 2, 4, 8):
        b = shard.shard_bounds(fd.segs, fd.cands, world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == len(fd.segs)
        assert all(b[i] <= b[i + 1] for i in range(world))
        tot = 0
        for r in range(world):
            s, c, _ = shard.take_shard(fd.segs, fd.cands, world, r)
            tot += len(s)
            if len(s):
                assert s["cand_begin"][0] == 0
                assert int(s["cand_begin"][-1] + s["cand_count"][-1]) == len(c)
        assert tot == len(fd.segs)
    assert shard.shard_bounds(fd.segs[:0], fd.cands[:0], 4) == [0, 0, 0, 0, 0]


def test_shard_bounds_balance_work_not_counts():
    """Mixed-length table (the C4 regime): cuts follow sum span^2, so a shard of long segments holds far fewer
    segments than a shard of short ones, and every shard's work is within one segment of the ideal share."""
    import rpa_b200
    rng = np.random.default_rng(5)
    n, k = 4000, 12
    L = np.where(np.arange(n) < n // 2, 500, 20000)          # first half short, second half long
    segs = np.zeros(n, rpa_b200.SEG_DTYPE)
    segs["cand_count"] = k
    segs["cand_begin"] = np.arange(n) * k
    cands = np.zeros(n * k, rpa_b200.CAND_DTYPE)
    start = rng.integers(1, 1000, n * k)
    span = np.repeat(L, k)
    rev = rng.random(n * k) < 0.5
    cands["rstart"] = np.where(rev, start + span - 1, start)
    cands["rstop"] = np.where(rev, start, start + span - 1)
    w = shard.work_estimate(segs, cands)
    assert w[0] == 1 + k * 500 ** 2 and w[-1] == 1 + k * 20000 ** 2
    for world in (2, 4, 8):
        b = shard.shard_bounds(segs, cands, world)
        assert b[0] == 0 and b[-1] == n
        work = [int(w[b[r]:b[r + 1]].sum()) for r in range(world)]
        ideal = int(w.sum()) / world
        assert max(abs(x - ideal) for x in work) <= int(w.max()), (world, work)
        assert b[1] > n // 2          # the first shard swallows all the short segments
