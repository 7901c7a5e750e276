"""Dev script: banded vs full-matrix edit-distance kernel on similar pairs (not a bench value)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import rpa_b200

ctx = rpa_b200.Context(0)
rng = np.random.default_rng(1)
alpha = np.frombuffer(b"ACGT", np.uint8)
cases = [(5000, 100000, 0.13), (5000, 20000, 0.13), (5000, 100000, 0.03), (1000, 200000, 0.15), (30000, 4000, 0.17),
         (5000, 40000, 0.75)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for L, npairs, div in cases:
    nseq = 256
    base = alpha[rng.integers(0, 4, L)]
    seqs = []
    for _ in range(nseq):
        s = base.copy()
        m = rng.random(L) < div / 2          # two derived sequences differ by ~div
        s[m] = alpha[(np.searchsorted(alpha, s[m]) + rng.integers(1, 4, int(m.sum()))) % 4]
        seqs.append(s)
    lens = np.full(nseq, L, np.uint32)
    off = (np.arange(nseq) * L).astype(np.uint64)
    chars = np.concatenate(seqs)
    pa = rng.integers(0, nseq, npairs).astype(np.uint32)
    pb = rng.integers(0, nseq, npairs).astype(np.uint32)
    cells = float(L) * L * npairs
    ref = None
    for name, band, wedge in (("full matrix", 0, 0), ("band", 1, 0), ("band + wedge", 1, 1)):
        ctx.set_tuning("wedge", wedge)
        ctx.set_band(band)
        ctx.profile_reset()
        out, ms = ctx.edit_distance_batch(chars, off, lens, pa, pb, repeat=2)
        p = ctx.profile()
        if ref is None:
            ref = out
        assert np.array_equal(out, ref), name
        ex = p["cells_edit_distance"] / 3.0   # warm-up + 2 timed runs
        print("L=%d pairs=%d div=%.2f %-12s %8.2f ms  %9.1f GCUPS algorithmic  executed %.3f  (%7.1f GCUPS executed) retries %d"
              % (L, npairs, div, name, ms, cells / ms / 1e6, ex / cells if ex else 1.0, (ex or cells) / ms / 1e6,
                 p["band_retries"] // 3), flush=True)
ctx.set_tuning("wedge", 1)
ctx.set_band(1)
