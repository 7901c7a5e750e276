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
