#!/usr/bin/env python3
"""Executable model of the protein3_kernel schedule (taxator-tk_b200/csrc/protein3.cu): a group of LANES lanes owns a
pair, C columns per lane RIGHT aligned (leading padding columns with profile 0), R rows per step walked in skewed
order, one packed integer per cell (score << 13 | priority << 11 | #gap columns of the traced path, per-row bias).
Lanes are simulated one after the other; a lane's left boundary comes from the lane before it (one step earlier)."""
import re, os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SH = 13; PRIO = 3 << 11; ROWBIAS = 8 << (SH - 1)
CV = -(1 << SH) + (1 << 11) + 1 + ROWBIAS      # vertical: score-1, priority 1, +1 gap, next row
CH = -(1 << SH) + 1                             # horizontal: score-1, priority 0, +1 gap
BND = -(1 << SH) + 1                            # boundary(k) = k * BND


def blosum62():
    txt = open(os.path.join(ROOT, "taxator-tk_b200", "csrc", "blosum62_table.h")).read()
    nums = [int(x) for x in re.findall(r"-?\d+", txt.split("{", 1)[1])]
    return np.array(nums[:27 * 32]).reshape(27, 32)


def align(a, b, T, lanes=8, C=8, R=4):
    """a = columns (ordinals), b = rows; returns (score, #diagonal steps of the traced path)."""
    n, m = len(a), len(b)
    assert 0 < n <= lanes * C and m > 0
    pad = lanes * C - n
    e = lambda r, bb: (2 * int(T[r][bb]) + 9) if r < 27 else 0
    col = [[None] * C for _ in range(lanes)]
    up = [[0] * C for _ in range(lanes)]
    dcarry = [0] * lanes
    for lp in range(lanes):
        v1 = lp * C - pad
        for c in range(C):
            v = v1 + c
            col[lp][c] = a[v] if v >= 0 else 27
            up[lp][c] = (v + 1) * BND if v >= 0 else 0
        dcarry[lp] = v1 * BND if v1 > 0 else 0
    last = [[0] * R for _ in range(lanes)]
    res = 0
    steps = (m + R - 1) // R + lanes - 1
    for t in range(1, steps + 1):
        recv = [[last[lp - 1][r] if lp > 0 else 0 for r in range(R)] for lp in range(lanes)]   # __shfl_up of the previous step
        for lp in range(lanes):
            j = t - lp
            i0 = R * (j - 1) + 1
            if j < 1 or i0 > m:
                continue
            br = [b[min(max(i0 + r, 1), m) - 1] for r in range(R)]
            left = [(i0 + r) * (BND + ROWBIAS) if lp == 0 else recv[lp][r] for r in range(R)]
            dg = [dcarry[lp]] + left[:R - 1]
            dcarry[lp] = left[R - 1]
            for k in range(C + R - 1):          # skewed walk
                for r in range(R):
                    c = k - r
                    if 0 <= c < C:
                        D = e(col[lp][c], br[r]) * (1 << (SH - 1)) + dg[r]
                        cell = max(D, up[lp][c] + CV, left[r] + CH) & ~PRIO
                        dg[r] = up[lp][c]; up[lp][c] = cell; left[r] = cell
            for r in range(R):
                last[lp][r] = left[r]
                if lp == lanes - 1 and i0 + r == m:
                    res = left[r]
    P = res - m * ROWBIAS
    return P >> SH, (n + m - (P & 0x7ff)) // 2
