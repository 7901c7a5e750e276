"""CPU: the executable model of protein3_kernel's schedule (scripts/protein3_model.py: lanes per pair, right-aligned
column blocks with padding, rows per step in skewed order, packed cells) against the oracle's NW with SeqAn's tie order."""
import ctypes
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol

sys.path.insert(0, os.path.join(ol.ROOT, "scripts"))
import protein3_model as pm


@pytest.mark.parametrize("lanes,C,R", [(8, 4, 4), (8, 8, 2), (16, 4, 4), (32, 2, 4), (8, 6, 3)])
def test_model_matches_oracle(lanes, C, R):
    rng = np.random.default_rng(lanes * 100 + C * 10 + R)
    O = ol.oracle()
    T = pm.blosum62()
    out6 = (ctypes.c_int * 6)()
    for it in range(60):
        alpha = 27 if it % 3 == 0 else 20
        n = int(rng.integers(1, lanes * C + 1))
        a = rng.integers(0, alpha, n).astype(np.uint8)
        if it % 2:
            b = a.copy()
            for _ in range(int(len(b) * 0.25)):
                b[rng.integers(0, len(b))] = rng.integers(0, alpha)
            cut = int(rng.integers(0, max(1, len(b) // 3)))
            b = b[cut:] if len(b) - cut >= 1 else b
        else:
            b = rng.integers(0, alpha, int(rng.integers(1, 60))).astype(np.uint8)
        O.orc_protein_align(ol.ptr(a, ol.u8p), len(a), ol.ptr(b, ol.u8p), len(b), out6)
        want = (out6[0], len(a) + len(b) - out6[2])
        assert pm.align(list(a), list(b), T, lanes, C, R) == want, (it, n, len(b))
