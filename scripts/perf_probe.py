"""Dev script: raw kernel throughput probes (not a bench value)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import rpa_b200

ctx = rpa_b200.Context(0)
peak = ctx.int_alu_peak()
print("INT ALU lane-ops/s: %.3e" % peak, flush=True)
rng = np.random.default_rng(1)
alpha = np.frombuffer(b"ACGT", np.uint8)
for L, npairs in [(1000, 200000), (5000, 40000), (5000, 8000), (20000, 4000), (50000, 1200), (500, 400000)]:
    nseq = 512
    seqs = [alpha[rng.integers(0, 4, L)] for _ in range(nseq)]
    lens = np.full(nseq, L, np.uint32)
    off = (np.arange(nseq) * L).astype(np.uint64)
    chars = np.concatenate(seqs)
    pa = rng.integers(0, nseq, npairs).astype(np.uint32)
    pb = rng.integers(0, nseq, npairs).astype(np.uint32)
    out, ms = ctx.edit_distance_batch(chars, off, lens, pa, pb, repeat=3)
    cells = float(L) * L * npairs
    print("L=%d pairs=%d: %.2f ms  %.1f GCUPS" % (L, npairs, ms, cells / ms / 1e6), flush=True)
