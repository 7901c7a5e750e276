"""Dev script: wedge vs plain band per length / divergence: time, executed cells, failed certificates."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import rpa_b200

ctx = rpa_b200.Context(0)
rng = np.random.default_rng(1)
alpha = np.frombuffer(b"ACGT", np.uint8)
for L in (1000, 2000, 3000, 5000, 8000, 12000, 20000):
    for div in (0.04, 0.1, 0.2):
        npairs = int(4e11 / (L * L * max(div, 0.05))) + 1000
        npairs = min(npairs, 300000)
        nseq = 128
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
        res = []
        for wedge in (0, 1):
            ctx.set_tuning("wedge", wedge)
            ctx.profile_reset()
            out, ms = ctx.edit_distance_batch(chars, off, lens, pa, pb, repeat=2)
            p = ctx.profile()
            res.append((ms, p["cells_edit_distance"] / 3 / cells, p["wedge_failures"] / 3 / npairs, out))
        assert np.array_equal(res[0][3], res[1][3])
        print("L=%5d div=%.2f pairs=%6d  band %7.2f ms x%.3f | wedge %7.2f ms x%.3f fail %.3f  speed-up %.2f"
              % (L, div, npairs, res[0][0], res[0][1], res[1][0], res[1][1], res[1][2], res[0][0] / res[1][0]), flush=True)
ctx.set_tuning("wedge", 1)
