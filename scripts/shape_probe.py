"""Dev script: banded kernel time per forced (W, L) shape (not a bench value)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import rpa_b200

ctx = rpa_b200.Context(0)
rng = np.random.default_rng(1)
alpha = np.frombuffer(b"ACGT", np.uint8)
WS = [1, 2, 4, 8, 12, 16, 20, 24]
cases = [(5000, 100000, 0.13), (5000, 100000, 0.05), (5000, 20000, 0.13)]
for L, npairs, div in cases:
    nseq = 256
    base = alpha[rng.integers(0, 4, L)]
    seqs = []
    for _ in range(nseq):
        s = base.copy()
        m = rng.random(L) < div / 2
        s[m] = alpha[(np.searchsorted(alpha, s[m]) + rng.integers(1, 4, int(m.sum()))) % 4]
        seqs.append(s)
    lens = np.full(nseq, L, np.uint32)
    off = (np.arange(nseq) * L).astype(np.uint64)
    chars = np.concatenate(seqs)
    pa = rng.integers(0, nseq, npairs).astype(np.uint32)
    pb = rng.integers(0, nseq, npairs).astype(np.uint32)
    cells = float(L) * L * npairs
    ctx.set_tuning("force_shape", -1)
    ctx.profile_reset()
    ref, ms = ctx.edit_distance_batch(chars, off, lens, pa, pb, repeat=2)
    print("L=%d pairs=%d div=%.2f planner: %.2f ms executed %.3f" % (L, npairs, div, ms, ctx.profile()["cells_edit_distance"] / 3 / cells), flush=True)
    for li in range(0, 5):
        row = []
        for wi in range(1, 6):
            ctx.set_tuning("force_shape", li * 8 + wi)
            ctx.profile_reset()
            out, ms = ctx.edit_distance_batch(chars, off, lens, pa, pb, repeat=2)
            assert np.array_equal(out, ref)
            ex = ctx.profile()["cells_edit_distance"] / 3 / cells
            row.append("W%-2d %6.2f ms x%.3f %5.1fT" % (WS[wi], ms, ex, ex * cells / ms / 1e9))
        print("  L=%-2d " % (1 << li) + " | ".join(row), flush=True)
ctx.set_tuning("force_shape", -1)
