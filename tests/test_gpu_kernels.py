"""GPU parity tests of the low-level C-ABI entry points against the oracle (bit-exact)."""
import ctypes

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def _table(seqs):
    lens = np.array([len(s) for s in seqs], np.uint32)
    off = np.zeros(len(seqs), np.uint64)
    off[1:] = np.cumsum(lens[:-1].astype(np.uint64))
    chars = np.concatenate([np.frombuffer(s, np.uint8) for s in seqs]) if seqs else np.zeros(0, np.uint8)
    return chars, off, lens


def _mutate(rng, s, rate, alpha):
    a = np.frombuffer(s, np.uint8).copy()
    r = rng.random(len(a))
    reps = np.ones(len(a), np.int64)
    reps[r < rate / 3] = 0
    reps[(r >= rate / 3) & (r < 2 * rate / 3)] = 2
    sub = (r >= 2 * rate / 3) & (r < rate)
    a[sub] = alpha[rng.integers(0, len(alpha), int(sub.sum()))]
    out = np.repeat(a, reps)
    ends = np.cumsum(reps)
    pos = ends[reps == 2] - 1
    out[pos] = alpha[rng.integers(0, len(alpha), len(pos))]
    return out.tobytes()


def _oracle_ed(a, b):
    O = ol.oracle()
    ca = ol.codes_of(np.frombuffer(a, np.uint8), False)
    cb = ol.codes_of(np.frombuffer(b, np.uint8), False)
    return O.orc_edit_distance(ol.ptr(ca, ol.u8p), len(ca), ol.ptr(cb, ol.u8p), len(cb))


@pytest.mark.parametrize("with_n", [False, True])
def test_edit_distance_lengths(ctx, with_n):
    rng = np.random.default_rng(7 + with_n)
    alpha = np.frombuffer(b"ACGTN" if with_n else b"ACGT", np.uint8)
    seqs, pa, pb = [], [], []
    lens = [0, 1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 255, 256, 257, 383, 384, 385, 500, 511, 512,
            513, 767, 768, 769, 1000, 1023, 1024, 1025, 1500, 2047, 2048, 2049, 3000, 4095, 4096, 4097, 5000, 6143, 6144,
            6145, 8000, 10000, 12288, 12289, 16384, 20000, 24576, 24577, 30000, 40000, 50000]
    for L in lens:
        a = alpha[rng.integers(0, len(alpha), L)].tobytes()
        for rate in (0.0, 0.05, 0.3):
            b = _mutate(rng, a, rate, alpha)
            seqs += [a, b]
            pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)
        # unrelated, different length
        c = alpha[rng.integers(0, len(alpha), int(rng.integers(0, max(2, L))))].tobytes()
        seqs.append(c)
        pa.append(len(seqs) - 1); pb.append(len(seqs) - 4)
    chars, off, ln = _table(seqs)
    got, _ = ctx.edit_distance_batch(chars, off, ln, pa, pb)
    for k in range(len(pa)):
        want = _oracle_ed(seqs[pa[k]], seqs[pb[k]])
        assert got[k] == want, (k, len(seqs[pa[k]]), len(seqs[pb[k]]), int(got[k]), want)


def test_edit_distance_many_random(ctx):
    rng = np.random.default_rng(11)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    seqs, pa, pb = [], [], []
    for _ in range(3000):
        L = int(rng.integers(1, 1500))
        a = alpha[rng.integers(0, 4, L)].tobytes()
        b = _mutate(rng, a, float(rng.uniform(0, 0.4)), alpha)
        seqs += [a, b]
        pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)
    chars, off, ln = _table(seqs)
    got, _ = ctx.edit_distance_batch(chars, off, ln, pa, pb)
    want = np.array([_oracle_ed(seqs[a], seqs[b]) for a, b in zip(pa, pb)], np.int32)
    assert np.array_equal(got, want), np.flatnonzero(got != want)[:10]


def test_edit_distance_lowercase_and_iupac(ctx):
    seqs = [b"acgtACGTnNRYKMuU--", b"ACGTACGTNNNNNNTT", b"", b"A"]
    chars, off, ln = _table(seqs)
    pa, pb = [0, 0, 2, 3, 2], [1, 0, 3, 1, 2]
    got, _ = ctx.edit_distance_batch(chars, off, ln, pa, pb)
    want = [_oracle_ed(seqs[a], seqs[b]) for a, b in zip(pa, pb)]
    assert got.tolist() == want


@pytest.mark.parametrize("maxlen", [1100, 700, 430, 90])
def test_protein_align(ctx, maxlen):
    """maxlen: longest sequence of the batch -- the kernels size their shared-memory profile from it, per lane width
    (8 / 16 / 32 lanes per pair)."""
    rng = np.random.default_rng(5 + maxlen)
    O = ol.oracle()
    full = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWYZX*", np.uint8)
    aa20 = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    lens_all = [1, 2, 5, 31, 32, 33, 100, 299, 300, 301, 321, 330, 400, 420, 511, 512, 513, 600, 641, 650, 1100]
    seqs, pa, pb = [], [], []
    for t in range(500 if maxlen == 1100 else 200):
        alpha = full if t % 4 == 0 else aa20
        L = int(rng.choice([x for x in lens_all if x <= maxlen]))
        a = alpha[rng.integers(0, len(alpha), L)].tobytes()
        if t % 2:
            b = _mutate(rng, a, float(rng.choice([0.0, 0.1, 0.4, 0.8])), alpha)
            if len(b) == 0:
                b = b"A"
        else:
            b = alpha[rng.integers(0, len(alpha), int(rng.integers(1, min(700, maxlen) + 1)))].tobytes()
        b = b[:maxlen]
        seqs += [a, b]
        pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)
    chars, off, ln = _table(seqs)
    got, _ = ctx.protein_align_batch(chars, off, ln, pa, pb)
    out6 = (ctypes.c_int * 6)()
    for k in range(len(pa)):
        ca = ol.codes_of(np.frombuffer(seqs[pa[k]], np.uint8), True)
        cb = ol.codes_of(np.frombuffer(seqs[pb[k]], np.uint8), True)
        O.orc_protein_align(ol.ptr(ca, ol.u8p), len(ca), ol.ptr(cb, ol.u8p), len(cb), out6)
        assert got[k].tolist() == [out6[0], out6[1], out6[2]], (k, len(ca), len(cb), got[k].tolist(), list(out6))


def _band_cases(rng, alpha, n_pairs, lens, rates):
    seqs, pa, pb = [], [], []
    for _ in range(n_pairs):
        L = int(rng.choice(lens))
        a = alpha[rng.integers(0, len(alpha), L)].tobytes()
        b = _mutate(rng, a, float(rng.choice(rates)), alpha)
        if rng.random() < 0.15:   # one side much longer / shorter
            b = b + alpha[rng.integers(0, len(alpha), int(rng.integers(0, 400)))].tobytes()
        if rng.random() < 0.1:
            b = b[int(rng.integers(0, max(1, len(b) // 3))):]
        seqs += [a, b]
        pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)
    return seqs, pa, pb


@pytest.mark.parametrize("with_n", [False, True])
def test_band_equals_full_matrix(ctx, with_n):
    """Ukkonen band + verify/widen loop returns the same integers as the full DP matrix and the oracle,
    for every (W, L) shape the planner can pick (plan_lanes moves it between lane-time and latency
    mode) and for initial thresholds that are far too small (band_k0 forces retries)."""
    rng = np.random.default_rng(23 + with_n)
    alpha = np.frombuffer(b"ACGTN" if with_n else b"ACGT", np.uint8)
    seqs, pa, pb = _band_cases(rng, alpha, 400, [40, 64, 65, 200, 700, 1500, 3000, 6000], [0.0, 0.01, 0.05, 0.15, 0.4, 0.9])
    chars, off, ln = _table(seqs)
    want = np.array([_oracle_ed(seqs[a], seqs[b]) for a, b in zip(pa, pb)], np.int32)
    try:
        ctx.set_band(0)
        ctx.profile_reset()
        full, _ = ctx.edit_distance_batch(chars, off, ln, pa, pb)
        cells_full = ctx.profile()["cells_edit_distance"]
        assert np.array_equal(full, want), np.flatnonzero(full != want)[:10]
        ctx.set_band(1)
        for lanes in (0, 32, 1 << 24):
            for k0 in (0, 64, 300):
                ctx.set_tuning("plan_lanes", lanes)
                ctx.set_tuning("band_k0", k0)
                ctx.profile_reset()
                got, _ = ctx.edit_distance_batch(chars, off, ln, pa, pb)
                prof = ctx.profile()
                assert np.array_equal(got, want), (lanes, k0, np.flatnonzero(got != want)[:10])
                if k0 == 0:
                    # the planned threshold is a true upper bound: only a failed wedge certificate re-runs a pair
                    assert prof["band_retries"] == prof["wedge_failures"]
                    assert prof["cells_edit_distance"] < cells_full
                else:
                    assert prof["band_retries"] > 0
    finally:
        ctx.set_band(1)
        ctx.set_tuning("plan_lanes", 0)
        ctx.set_tuning("band_k0", 0)


def test_band_long_pairs(ctx):
    """Long noisy pairs (the 10-50 kb regime): band with indels, several strips per lane."""
    rng = np.random.default_rng(31)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    seqs, pa, pb = _band_cases(rng, alpha, 24, [10000, 20000, 33000, 50000], [0.02, 0.15, 0.3])
    chars, off, ln = _table(seqs)
    want = np.array([_oracle_ed(seqs[a], seqs[b]) for a, b in zip(pa, pb)], np.int32)
    try:
        for lanes, k0 in ((0, 0), (1 << 24, 0), (1 << 24, 500), (0, 2000)):
            ctx.set_tuning("plan_lanes", lanes)
            ctx.set_tuning("band_k0", k0)
            got, _ = ctx.edit_distance_batch(chars, off, ln, pa, pb)
            assert np.array_equal(got, want), (lanes, k0, np.flatnonzero(got != want)[:10])
    finally:
        ctx.set_tuning("plan_lanes", 0)
        ctx.set_tuning("band_k0", 0)


def _subst(rng, a, rates):
    """substitutions only, rate per fifth of the sequence"""
    b = np.frombuffer(a, np.uint8).copy()
    alpha = np.frombuffer(b"ACGT", np.uint8)
    n = len(b)
    for q, r in enumerate(rates):
        lo, hi = n * q // len(rates), n * (q + 1) // len(rates)
        m = rng.random(hi - lo) < r
        idx = np.searchsorted(alpha, b[lo:hi][m])
        b[lo:hi][m] = alpha[(idx + rng.integers(1, 4, int(m.sum()))) % 4]
    return b.tobytes()


def test_wedge_is_exact_and_certified(ctx):
    """The narrowing band (wedge) must give the oracle's integers whether its certificate holds
    (uniform divergence) or fails and the pair is re-run (errors clustered at the end, an indel that
    moves the alignment off the diagonal late); and it must execute fewer cells than the plain band."""
    rng = np.random.default_rng(41)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    seqs, pa, pb = [], [], []
    profiles = [[0.1] * 5, [0.2] * 5, [0.03] * 5, [0.0, 0.0, 0.0, 0.0, 0.6], [0.5, 0.0, 0.0, 0.0, 0.0], [0.02, 0.3, 0.02, 0.3, 0.02],
                [0.15, 0.15, 0.0, 0.0, 0.0], [0.0, 0.0, 0.1, 0.2, 0.3]]
    for t in range(240):
        L = int(rng.choice([300, 1000, 2500, 5000, 9000]))
        a = alpha[rng.integers(0, 4, L)].tobytes()
        b = _subst(rng, a, profiles[t % len(profiles)])
        if t % 6 == 5:      # one indel late in the sequence: Hamming stays small, the path leaves the diagonal
            p = int(L * 0.9)
            b = b[:p] + alpha[rng.integers(0, 4, int(rng.integers(1, 40)))].tobytes() + b[p:]
        if t % 9 == 8:
            p = int(L * 0.95)
            b = b[:p] + b[p + int(rng.integers(1, 30)):]
        seqs += [a, b]
        pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)
    chars, off, ln = _table(seqs)
    want = np.array([_oracle_ed(seqs[a], seqs[b]) for a, b in zip(pa, pb)], np.int32)
    try:
        cells = {}
        for wedge in (0, 1, 2):     # 2: the planner asks for the narrowest wedge at once (test hook)
            for lanes in (0, 1 << 24):
                ctx.set_tuning("wedge", wedge)
                ctx.set_tuning("plan_lanes", lanes)
                ctx.profile_reset()
                got, _ = ctx.edit_distance_batch(chars, off, ln, pa, pb)
                prof = ctx.profile()
                assert np.array_equal(got, want), (wedge, lanes, np.flatnonzero(got != want)[:10])
                cells[(wedge, lanes)] = prof["cells_edit_distance"]
                if not wedge:
                    assert prof["band_retries"] == 0 and prof["wedge_failures"] == 0
                elif wedge == 1:
                    assert prof["band_retries"] == prof["wedge_failures"] < len(pa) // 4
                else:
                    # far too narrow: most certificates fail, the pairs are re-run with the plain band
                    assert prof["band_retries"] == prof["wedge_failures"] > len(pa) // 4
        assert cells[(1, 1 << 24)] < cells[(0, 1 << 24)]   # the 5 kb / 9 kb pairs of 8+ % divergence qualify
    finally:
        ctx.set_tuning("wedge", 1)
        ctx.set_tuning("plan_lanes", 0)
