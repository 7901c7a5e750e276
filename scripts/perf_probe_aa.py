"""Dev script: raw protein kernel throughput (not a bench value)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import rpa_b200
ctx = rpa_b200.Context(0)
rng = np.random.default_rng(1)
aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
for L, npairs in [(300, 200000), (100, 400000), (600, 50000), (900, 25000)]:
    nseq = 1024
    seqs = [aa[rng.integers(0, 20, L)] for _ in range(nseq)]
    lens = np.full(nseq, L, np.uint32); off = (np.arange(nseq) * L).astype(np.uint64)
    chars = np.concatenate(seqs)
    pa = rng.integers(0, nseq, npairs).astype(np.uint32); pb = rng.integers(0, nseq, npairs).astype(np.uint32)
    out, ms = ctx.protein_align_batch(chars, off, lens, pa, pb, repeat=3)
    print("L=%d pairs=%d: %.2f ms  %.1f GCUPS" % (L, npairs, ms, float(L) * L * npairs / ms / 1e6), flush=True)
