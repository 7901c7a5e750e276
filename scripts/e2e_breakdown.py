"""Dev script: wall-time split of upload / run / download for a workload."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, rpa_b200, torch
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else None
d = bench.make_data(wl, 1, nq)
fd = bench.Flat(d)
ctx = rpa_b200.Context(0)
ctx.load_taxonomy(fd.parent, fd.left, fd.right, fd.depth, 0)
a = 1 if fd.protein else 0
ctx.load_store(0, a, fd.q_chars, fd.q_off, fd.q_len); ctx.load_store(1, a, fd.r_chars, fd.r_off, fd.r_len)
for it in range(4):
    t0 = time.perf_counter(); ctx.batch_upload(fd.segs, fd.cands); t1 = time.perf_counter()
    ctx.batch_run(); t2 = time.perf_counter()
    out = ctx.batch_download(); t3 = time.perf_counter()
    print("upload %.1f ms  run %.1f ms  download %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3), flush=True)
print(ctx.profile())
