"""GPU: edge cases of the path against the oracle -- clipped / empty reference ranges, N-rich
sequences, single-record and fully identical segments, very uneven candidate counts."""
import pytest

import edge_data
import oracle_lib as ol

pytestmark = pytest.mark.gpu


def load(ctx, fd):
    ctx.load_taxonomy(fd.parent, fd.left, fd.right, fd.depth, 0)
    alpha = 1 if fd.protein else 0
    ctx.load_store(0, alpha, fd.q_chars, fd.q_off, fd.q_len)
    ctx.load_store(1, alpha, fd.r_chars, fd.r_off, fd.r_len)


def check(ctx, fd):
    want = ol.oracle_predict(fd)
    load(ctx, fd)
    got = ctx.predict_batch(fd.segs, fd.cands)
    assert ol.results_equal(want, got) == []
    return got


@pytest.mark.parametrize("protein", [False, True])
def test_ranges_past_sequence_ends(ctx, protein):
    check(ctx, edge_data.ranges_past_ends(protein))


def test_n_rich_sequences(ctx):
    check(ctx, edge_data.n_rich())


def test_single_record_identical_and_empty_segments(ctx):
    got = check(ctx, edge_data.special_segments())
    assert got["kind"][0] == 1 and got["kind"][1] == 0 and got["kind"][2] == 2 and got["kind"][3] == 3


def test_parameters_x_and_t(ctx):
    fd = edge_data.base(seed=24)
    load(ctx, fd)
    for x, t in [(0.0, 0.05), (0.9, 0.05), (0.5, 0.3), (0.5, 0.0)]:
        want = ol.oracle_predict(fd, exclude_factor=x, toppercent=t)
        ctx.set_params(x, t)
        try:
            got = ctx.predict_batch(fd.segs, fd.cands)
        finally:
            ctx.set_params(0.5, 0.05)
        assert ol.results_equal(want, got) == [], (x, t)


def test_many_candidates_one_segment(ctx):
    """One segment with hundreds of records next to tiny ones (insertion sorts, long loops)."""
    check(ctx, edge_data.many_candidates())


def test_stable_sort_with_ties(ctx):
    """SortFilter runs on the device (rank sort with the record position as tie-break)."""
    check(ctx, edge_data.score_ties())


@pytest.mark.parametrize("name", ["past_ends_nt", "past_ends_aa", "n_rich", "many_candidates"])
def test_edge_cases_against_the_real_reference(ctx, name):
    """GPU == GFF3 of the real reference on the edge cases it can run (tests/golden/edge_*.gff3)."""
    import golden_util as gu
    fd = dict(edge_data.reference_cases())[name]
    load(ctx, fd)
    got = ctx.predict_batch(fd.segs, fd.cands)
    assert gu.render_sorted(fd, got) == open(gu.os.path.join(gu.GOLDEN, "edge_%s.gff3" % name)).readlines()


@pytest.mark.parametrize("x,t", edge_data.PARAM_SETS)
def test_parameters_against_the_real_reference(ctx, x, t):
    import golden_util as gu
    fd = edge_data.base(seed=24)
    load(ctx, fd)
    try:
        ctx.set_params(float(x), float(t))
        got = ctx.predict_batch(fd.segs, fd.cands)
    finally:
        ctx.set_params(0.5, 0.05)
    assert gu.render_sorted(fd, got) == open(gu.os.path.join(gu.GOLDEN, "edge_params_x%s_t%s.gff3" % (x, t))).readlines()
