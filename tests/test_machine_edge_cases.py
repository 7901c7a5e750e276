"""CPU: the host-compiled product state machine on the edge-case inputs, against the oracle."""
import pytest

import edge_data
import oracle_lib as ol


def check(fd, **kw):
    want = ol.oracle_predict(fd, **kw)
    for k in (0, 5):
        got, _ = ol.host_machine_predict(fd, spec_k=k, **kw)
        assert ol.results_equal(want, got) == []
    return want


@pytest.mark.parametrize("protein", [False, True])
def test_ranges_past_sequence_ends(protein):
    check(edge_data.ranges_past_ends(protein))


def test_n_rich_sequences():
    check(edge_data.n_rich())


def test_special_segments():
    r = check(edge_data.special_segments())
    assert r["kind"][0] == 1 and r["kind"][1] == 0 and r["kind"][2] == 2 and r["kind"][3] == 3


def test_parameters_x_and_t():
    fd = edge_data.base(seed=24)
    for x, t in [(0.0, 0.05), (0.9, 0.05), (0.5, 0.3), (0.5, 0.0)]:
        check(fd, exclude_factor=x, toppercent=t)


def test_many_candidates_one_segment():
    check(edge_data.many_candidates())


def test_stable_sort_with_ties():
    check(edge_data.score_ties())


@pytest.mark.parametrize("name", ["past_ends_nt", "past_ends_aa", "n_rich", "many_candidates"])
def test_edge_cases_against_the_real_reference(name):
    """The edge cases the real reference can run (tests/golden/edge_*.gff3 from oracle/_ref/taxator,
    make_golden_edge.py): reference ranges clipped at the sequence ends, N-rich sequences, 300-record sets --
    oracle and host-compiled state machine reproduce its GFF3."""
    import golden_util as gu
    fd = dict(edge_data.reference_cases())[name]
    want = open(gu.os.path.join(gu.GOLDEN, "edge_%s.gff3" % name)).readlines()
    res = ol.oracle_predict(fd)
    assert gu.render_sorted(fd, res) == want
    got, _ = ol.host_machine_predict(fd)
    assert ol.results_equal(res, got) == []


@pytest.mark.parametrize("x,t", edge_data.PARAM_SETS)
def test_parameters_against_the_real_reference(x, t):
    """-x / -t of the real reference (tests/golden/edge_params_*.gff3)."""
    import golden_util as gu
    fd = edge_data.base(seed=24)
    want = open(gu.os.path.join(gu.GOLDEN, "edge_params_x%s_t%s.gff3" % (x, t))).readlines()
    res = ol.oracle_predict(fd, exclude_factor=float(x), toppercent=float(t))
    assert gu.render_sorted(fd, res) == want
    got, _ = ol.host_machine_predict(fd, exclude_factor=float(x), toppercent=float(t))
    assert ol.results_equal(res, got) == []
