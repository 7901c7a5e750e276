"""GPU parity of the whole batched RPA path (decide/stage/align rounds) against the oracle."""
import numpy as np
import pytest

import oracle_lib as ol
import synth

pytestmark = pytest.mark.gpu


def load(ctx, fd):
    ctx.load_taxonomy(fd.parent, fd.left, fd.right, fd.depth, 0)
    alpha = 1 if fd.protein else 0
    ctx.load_store(0, alpha, fd.q_chars, fd.q_off, fd.q_len)
    ctx.load_store(1, alpha, fd.r_chars, fd.r_off, fd.r_len)


CASES = [
    dict(seed=1, protein=False, n_genomes=60, genome_len=3000, n_queries=150, query_len=(200, 600), n_cand=25,
         levels=(2, 4, 6, 10, 20), multi_segment_frac=0.2, frac_n=0.002),
    dict(seed=2, protein=False, n_genomes=80, genome_len=8000, n_queries=120, query_len=(900, 2500), n_cand=30,
         levels=(2, 4, 6, 10, 20)),
    dict(seed=3, protein=True, n_genomes=60, genome_len=600, n_queries=150, query_len=(80, 300), n_cand=25,
         levels=(2, 4, 6, 10, 20), multi_segment_frac=0.2, frac_n=0.002),
    dict(seed=4, protein=False, n_genomes=40, genome_len=4000, n_queries=100, query_len=(300, 1200), n_cand=20,
         levels=(2, 3, 5, 8, 12), query_indel=0.1, query_sub=0.05),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "seed%d%s" % (c["seed"], "aa" if c["protein"] else "nt"))
def test_pipeline_matches_oracle(ctx, case):
    fd = ol.FlatData(synth.generate(synth.SynthConfig(**case)))
    want = ol.oracle_predict(fd)
    load(ctx, fd)
    ctx.set_params(0.5, 0.05)
    got = ctx.predict_batch(fd.segs, fd.cands)
    assert ol.results_equal(want, got) == []


def test_protein_full_alphabet(ctx):
    """All 27 SeqAn AminoAcid letters (and lower case / unknown characters) in both stores: the short-pair protein
    kernel keeps one profile row per residue the stores contain (here all of them)."""
    d = synth.generate(synth.SynthConfig(**CASES[2]))
    rng = np.random.default_rng(99)
    full = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWYZX*acdxz-", np.uint8)
    for seqs in (d.ref_seqs, d.q_seqs):
        for k in range(len(seqs)):
            s_ = seqs[k].copy()
            hit = rng.random(len(s_)) < 0.03
            s_[hit] = full[rng.integers(0, len(full), int(hit.sum()))]
            seqs[k] = np.ascontiguousarray(s_)
    fd = ol.FlatData(d)
    want = ol.oracle_predict(fd)
    load(ctx, fd)
    got = ctx.predict_batch(fd.segs, fd.cands)
    assert ol.results_equal(want, got) == []


def test_small_arena_chunks(ctx):
    case = CASES[0]
    fd = ol.FlatData(synth.generate(synth.SynthConfig(**case)))
    want = ol.oracle_predict(fd)
    load(ctx, fd)
    ctx.set_arena_bytes(64 * 1024)
    try:
        got = ctx.predict_batch(fd.segs, fd.cands)
    finally:
        ctx.set_arena_bytes(0)
    assert ol.results_equal(want, got) == []


@pytest.mark.parametrize("k", [0, 2, 64])
def test_lookahead_is_exact(ctx, k):
    for case in (CASES[1], CASES[2]):
        fd = ol.FlatData(synth.generate(synth.SynthConfig(**case)))
        want = ol.oracle_predict(fd)
        load(ctx, fd)
        ctx.set_lookahead(k)
        try:
            got = ctx.predict_batch(fd.segs, fd.cands)
        finally:
            ctx.set_lookahead(-1)
        assert ol.results_equal(want, got) == []


def test_fetch_and_lca(ctx):
    fd = ol.FlatData(synth.generate(synth.SynthConfig(**CASES[0])))
    load(ctx, fd)
    O = ol.oracle()
    rng = np.random.default_rng(3)
    n = 300
    ref = rng.integers(0, len(fd.r_len), n).astype(np.uint32)
    L = fd.r_len[ref].astype(np.int64)
    a = rng.integers(1, L + 50)
    b = rng.integers(1, L + 50)
    le = rng.integers(0, 60, n).astype(np.uint32)
    re_ = rng.integers(0, 60, n).astype(np.uint32)
    got = ctx.fetch_segments(ref, a.astype(np.uint32), b.astype(np.uint32), le, re_)
    buf = np.zeros(int(L.max()) + 400, np.uint8)
    for k in range(n):
        m = O.orc_fetch_segment(ol.ptr(fd.r_codes, ol.u8p), ol.ptr(fd.r_off, ol.u64p), ol.ptr(fd.r_len, ol.u32p),
                                len(fd.r_len), 0, int(ref[k]), int(a[k]), int(b[k]), int(le[k]), int(re_[k]),
                                ol.ptr(buf, ol.u8p))
        assert len(got[k]) == m and np.array_equal(got[k], buf[:m]), k
    x = rng.integers(0, len(fd.parent), 2000).astype(np.uint32)
    y = rng.integers(0, len(fd.parent), 2000).astype(np.uint32)
    g = ctx.lca_batch(x, y)
    O.orc_lca.restype = np.ctypeslib.ctypes.c_uint32
    for k in range(len(x)):
        w = O.orc_lca(ol.ptr(fd.parent, ol.u32p), ol.ptr(fd.left, ol.u32p), ol.ptr(fd.right, ol.u32p),
                      ol.ptr(fd.depth, ol.u8p), len(fd.parent), 0, int(x[k]), int(y[k]))
        assert g[k] == w


def test_packed_store_roundtrip(ctx):
    """trpa_export_store / trpa_load_store_packed: the HBM layout out and back in gives the same
    segments (fetch) and the same placements as loading the characters."""
    import edge_data
    for fd in (edge_data.n_rich(), edge_data.base(seed=31, protein=True)):
        alpha = 1 if fd.protein else 0
        ctx.load_taxonomy(fd.parent, fd.left, fd.right, fd.depth, 0)
        ctx.load_store(0, alpha, fd.q_chars, fd.q_off, fd.q_len)
        ctx.load_store(1, alpha, fd.r_chars, fd.r_off, fd.r_len)
        want = ctx.predict_batch(fd.segs, fd.cands)
        c = fd.cands[:200]
        ext = np.zeros(len(c), np.uint32)
        seg_want = ctx.fetch_segments(c["ref_seq"], c["rstart"], c["rstop"], ext, ext + 7)
        packs = [ctx.export_store(0), ctx.export_store(1)]
        assert packs[0][0] == alpha and len(packs[1][2]) == len(fd.r_len)
        # load something else in between, then the packed copies
        ctx.load_store(0, alpha, fd.q_chars[:1], fd.q_off[:1], np.ones(1, np.uint32))
        ctx.load_store(1, alpha, fd.r_chars[:1], fd.r_off[:1], np.ones(1, np.uint32))
        for which, (a, woff, lens, payload) in enumerate(packs):
            ctx.load_store_packed(which, a, woff, lens, payload)
        got = ctx.predict_batch(fd.segs, fd.cands)
        assert ol.results_equal(want, got) == []
        seg_got = ctx.fetch_segments(c["ref_seq"], c["rstart"], c["rstop"], ext, ext + 7)
        assert all(np.array_equal(x, y) for x, y in zip(seg_want, seg_got))
        # a payload that does not match its tables is rejected
        import rpa_b200
        a, woff, lens, payload = packs[1]
        with pytest.raises(rpa_b200.TrpaError):
            ctx.load_store_packed(1, a, woff + np.uint64(1), lens, payload)


def test_page_locked_host_tables(ctx):
    """trpa_host_alloc / trpa_host_free: tables in page-locked memory of the library go through trpa_predict_batch like any
    other host buffer (same results), and the memory is usable as ordinary host memory."""
    import ctypes
    import golden_util as gu
    import rpa_b200
    fd = ol.FlatData(gu.case_data("nt_small"))
    ctx.load_taxonomy(fd.parent, fd.left, fd.right, fd.depth, 0)
    ctx.load_store(0, 0, fd.q_chars, fd.q_off, fd.q_len)
    ctx.load_store(1, 0, fd.r_chars, fd.r_off, fd.r_len)
    want = ctx.predict_batch(fd.segs, fd.cands)
    L = ctx.L
    L.trpa_host_alloc.restype = ctypes.c_void_p
    L.trpa_host_alloc.argtypes = [ctypes.c_uint64]
    L.trpa_host_free.argtypes = [ctypes.c_void_p]
    L.trpa_host_free.restype = None
    bufs = []
    try:
        def pinned_copy(a):
            p = L.trpa_host_alloc(a.nbytes)
            assert p
            bufs.append(p)
            v = np.frombuffer((ctypes.c_uint8 * a.nbytes).from_address(p), dtype=a.dtype)
            v[:] = a
            return v
        segs = pinned_copy(np.ascontiguousarray(fd.segs, rpa_b200.SEG_DTYPE))
        cands = pinned_copy(np.ascontiguousarray(fd.cands, rpa_b200.CAND_DTYPE))
        out = pinned_copy(np.zeros(len(fd.segs), rpa_b200.RESULT_DTYPE))
        got = ctx.predict_batch_into(segs, cands, out)
        assert ol.results_equal(want, got) == []
    finally:
        for p in bufs:
            L.trpa_host_free(p)
