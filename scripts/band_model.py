#!/usr/bin/env python
"""Executable model of the banded, strip-rotating bit-vector edit-distance kernel (csrc/myers3.cuh).

Python ints stand in for the W-word registers of a lane.  The model follows the kernel step by
step: a group of L lanes owns one pair; strip s (32*W pattern rows) runs on lane s % L; at group
step t strip s works on text block t - T0(s); lane l > 0 takes its top boundary from lane l-1's
output of the previous step (shuffle), lane 0 from the wrap-around scratch written by lane L-1.
Cells outside the Ukkonen band of threshold k are never computed; the boundaries of the computed
region are upper bounds (+1 deltas), so the result v >= d always and v == d whenever v <= k.

Run:  python scripts/band_model.py   (random self-check against a plain DP)
"""
import random
import sys


def dp_distance(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i] + [0] * len(b)
        for j, cb in enumerate(b, 1):
            cur[j] = min(prev[j - 1] + (ca != cb), prev[j] + 1, cur[j - 1] + 1)
        prev = cur
    return prev[len(b)]


def popc(x):
    return bin(x).count("1")


class Geometry:
    """Band geometry shared by the model and (re-implemented in) the kernel.

    a1 (optional) turns the band into a WEDGE: the half width shrinks linearly from a0 = (k - delta) / 2
    at the first pattern row to a1 at the last one (errors accumulate along the alignment, so far-off
    cells stop being useful towards the end).  A wedge result is only accepted with a certificate."""

    def __init__(self, m, n, k, W, L, a1=None, x0=0):
        assert m <= n and m > 0
        self.m, self.n, self.W, self.L = m, n, W, L
        self.R = 32 * W
        self.delta = n - m
        k = max(k, self.delta + 64)
        k = min(k, max(n, self.delta + 64))
        self.k = k
        self.a0 = (k - self.delta) // 2          # half band width at the first row (>= 32)
        self.a1 = self.a0 if a1 is None else max(32, min(self.a0, a1))
        self.row0 = self.a0 + x0                 # the wedge keeps the full width up to this row
        self.wedge = self.a1 < self.a0 and m > self.row0 + 64
        if not self.wedge:
            self.a1 = self.a0
        self.mwords = (m + 31) // 32
        self.nblk = (n + 31) // 32
        self.S = (self.mwords + W - 1) // W

    def a(self, s):
        # full width up to row0 (a cell at offset a has only seen row - a columns, and the number of errors
        # seen so far fluctuates), then linear
        row = min(s * self.R, self.m)
        if not self.wedge or row <= self.row0:
            return self.a0
        return self.a0 - ((self.a0 - self.a1) * (row - self.row0)) // (self.m - self.row0)

    def b0(self, s):
        return max(0, s * self.R - self.a(s)) // 32

    def b1(self, s):
        return min(self.n - 1, (s + 1) * self.R - 1 + self.delta + self.a(s)) // 32

    def round_gap(self, r):
        """Extra steps between round r and r+1 (Delta_r >= 1 when a next round exists)."""
        L = self.L
        g = 1
        for l in range(L):
            s = r * L + l
            if s + L < self.S:
                g = max(g, self.b1(s) - self.b0(s + L) + 1 - L)
        return g


def banded_distance(pat, txt, k, W, L, stats=None, a1=None, x0=0):
    """Returns (v, k_used, cert): v >= d always; v == d if v <= k_used and (no wedge or cert >= v).

    cert = min over the cells through which a path can leave the computed region of (value + remaining
    diagonal distance): a path that leaves costs at least that much, so cert >= v proves that no path
    outside the region beats v."""
    m, n = len(pat), len(txt)
    G = Geometry(m, n, k, W, L, a1, x0)
    R, S, nblk = G.R, G.S, G.nblk
    maskR = (1 << R) - 1
    sigma = sorted(set(pat) | set(txt))

    def peq(s):
        rows = pat[s * R:(s + 1) * R]
        return {ch: sum(1 << i for i, c in enumerate(rows) if c == ch) for ch in sigma}

    # per-lane state
    class Lane:
        pass
    lanes = [Lane() for _ in range(L)]
    scratch = {}              # block -> (hp, hn, c, botacc) written by lane L-1
    for l, ln in enumerate(lanes):
        ln.s = l
        ln.setup = False
        ln.out = (0, 0, 0, 0)  # (hpOut, hnOut, cOut, botacc) after the lane's latest block
        ln.off = 0             # offset_r of the lane's current round
        ln.r = 0
    result = None
    cert = 1 << 60
    delta = G.delta
    if G.b1(0) < nblk - 1:      # along row 0 past strip 0's range: D[0][j] = j, back to diagonal delta
        cert = min(cert, 2 * 32 * (G.b1(0) + 1) - delta)
    for s1 in range(1, S):      # down column 0 to the first strip that does not start at column 0
        if G.b0(s1) > 0:
            cert = min(cert, 2 * (s1 * R + 1) + delta)
            break
    t = 0
    blocks_done = 0
    while True:
        if all(ln.s >= S for ln in lanes):
            break
        prev_out = [ln.out for ln in lanes]      # what a shuffle at the top of step t sees
        for l, ln in enumerate(lanes):
            if ln.s >= S:
                continue
            s = ln.s
            if not ln.setup:
                ln.b0, ln.b1 = G.b0(s), G.b1(s)
                ln.peq = peq(s)
                ln.VP, ln.VN = maskR, 0
                ln.tsum = 0
                ln.corner = None
                ln.setup = True
            b = t - (ln.off + l)
            if b < ln.b0:
                continue
            assert b <= ln.b1
            # ---- top boundary of block b
            if s == 0:
                hp_in, hn_in, c_in, bot_in = 0xffffffff, 0, 0, None
                real = False
            elif b > G.b1(s - 1):
                hp_in, hn_in, c_in, bot_in = 0xffffffff, 0, 0, None   # virtual: +1 deltas
                real = False
            else:
                hp_in, hn_in, c_in, bot_in = prev_out[l - 1] if l > 0 else scratch[b]
                real = True
            if b == ln.b0:
                if s == 0:
                    ln.corner = 0
                else:
                    assert real
                    ln.corner = bot_in - (popc(hp_in) - popc(hn_in))
                ln.botacc = ln.corner + R
            ncols = min(32, n - 32 * b)
            last_strip = s == S - 1
            run = ncols if last_strip else 32      # non-last strips always run 32 columns
            if last_strip:
                vm = (0xffffffff << (32 - ncols)) & 0xffffffff
                ln.tsum += popc(hp_in & vm) - popc(hn_in & vm)
            hp_out = hn_out = c_out = 0
            for c in range(run):
                col = 32 * b + c
                ch = txt[col] if col < n else sigma[0]
                eq = ln.peq.get(ch, 0)
                hpb = (hp_in >> (31 - c)) & 1
                hnb = (hn_in >> (31 - c)) & 1
                cb = (c_in >> (31 - c)) & 1
                T = eq & ln.VP
                Ssum = T + ln.VP + cb
                co = Ssum >> R
                Ssum &= maskR
                D0 = (Ssum ^ ln.VP) | eq | ln.VN
                HP = ln.VN | (~(D0 | ln.VP) & maskR)
                HN = ln.VP & D0
                Xh = ((HP << 1) | hpb) & maskR
                HNs = ((HN << 1) | hnb) & maskR
                ln.VN = Xh & D0
                ln.VP = HNs | (~(Xh | D0) & maskR)
                hp_out = ((hp_out << 1) | (HP >> (R - 1))) & 0xffffffff
                hn_out = ((hn_out << 1) | (HN >> (R - 1))) & 0xffffffff
                c_out = ((c_out << 1) | co) & 0xffffffff
            if run < 32:   # keep the "column 0 at bit 31" convention (only the last strip gets here)
                hp_out <<= 32 - run; hn_out <<= 32 - run; c_out <<= 32 - run
            blocks_done += 1
            ln.botacc += popc(hp_out) - popc(hn_out)
            ln.out = (hp_out, hn_out, c_out, ln.botacc)
            if s + 1 < S:
                if b < G.b0(s + 1):       # bottom-row cells whose lower neighbours are not computed
                    term = ln.botacc + ((s + 1) * R - 32 * (b + 1)) + delta
                    if stats is not None and term < cert:
                        stats['argmin'] = ('bottom', s, b, ln.botacc, G.a(s), G.a(s + 1))
                    cert = min(cert, term)
                if b == ln.b1 and b < nblk - 1:   # right-column cells of the strip
                    term = ln.botacc + (32 * (b + 1) - (s + 1) * R) - delta
                    if stats is not None and term < cert:
                        stats['argmin'] = ('right', s, b, ln.botacc, G.a(s), G.a(s + 1))
                    cert = min(cert, term)
            if l == L - 1 and s + 1 < S:
                scratch[b] = ln.out
            if b == ln.b1:
                if last_strip:
                    rows = m - s * R
                    valid = (1 << rows) - 1
                    result = ln.corner + ln.tsum + popc(ln.VP & valid) - popc(ln.VN & valid)
                # next strip of this lane
                ln.off += L + G.round_gap(ln.r)
                ln.r += 1
                ln.s += L
                ln.setup = False
        t += 1
        assert t < 10 * (nblk + S + 10) * (S + 1), "schedule does not terminate"
    if stats is not None:
        stats["blocks"] = blocks_done
        stats["steps"] = t
        stats["full_blocks"] = S * nblk
    return result, G.k, (cert if G.wedge else None)


def exact_distance(pat, txt, k0, W, L, stats=None, a1=None, x0=0):
    """The retry loop of the kernel: widen until the banded / wedged result is provably exact."""
    if len(pat) > len(txt):
        pat, txt = txt, pat
    if len(pat) == 0:
        return len(txt)
    k = k0
    tries = 0
    wedge_fail = 0
    while True:
        v, ku, cert = banded_distance(pat, txt, k, W, L, stats, a1, x0)
        tries += 1
        if v <= ku and (cert is None or cert >= v):
            if stats is not None:
                stats["tries"] = tries
                stats["wedge_fail"] = wedge_fail
            return v
        if cert is not None:
            wedge_fail += 1
        a1 = None                      # second attempts use the plain band
        k = ku if v <= ku else min(v, 3 * ku)


def _rand_pair(rng, m, div, indel):
    a = [rng.choice("ACGT") for _ in range(m)]
    b = []
    for ch in a:
        r = rng.random()
        if r < indel / 2:
            continue
        if r < indel:
            b.append(rng.choice("ACGT"))
        b.append(rng.choice("ACGT") if rng.random() < div else ch)
    return "".join(a), "".join(b)


def main():
    rng = random.Random(5)
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    worst = 0.0
    for it in range(n_cases):
        m = rng.choice([1, 5, 31, 32, 33, 64, 100, 200, 257, 300, 511, 700])
        div = rng.choice([0.0, 0.02, 0.1, 0.3, 0.75])
        indel = rng.choice([0.0, 0.0, 0.05, 0.2])
        a, b = _rand_pair(rng, m, div, indel)
        if rng.random() < 0.2:
            b = b + "".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 200)))
        if rng.random() < 0.1:
            b = b[:rng.randrange(0, len(b) + 1)]
        W = rng.choice([1, 2, 4, 8])
        L = rng.choice([1, 2, 3, 4, 8, 32])
        k0 = rng.choice([0, 1, 10, 50, 100, 400, 10 ** 6])
        want = dp_distance(a, b)
        st = {}
        a1 = rng.choice([None, None, 32, 40, 64, 100, 200])
        got = exact_distance(a, b, k0, W, L, st, a1, rng.choice([0, 0, 64, 200]))
        assert got == want, (it, len(a), len(b), W, L, k0, a1, got, want)
        if st:
            worst = max(worst, st["blocks"] / max(1, st["full_blocks"]))
    print("band model ok: %d cases" % n_cases)


if __name__ == "__main__":
    main()
