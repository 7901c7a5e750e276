"""CPU tests: the oracle restatement and the host-compiled product state machine both reproduce
the GFF3 of the REAL reference (`oracle/_ref/taxator`, unmodified sources) committed under
tests/golden/, and the oracle's kernels reproduce the per-pair integers dumped from real SeqAn."""
import ctypes
import json
import os

import numpy as np
import pytest

import golden_util as gu
import oracle_lib as ol


@pytest.mark.parametrize("name", sorted(gu.CASES))
def test_oracle_matches_reference_gff3(name):
    fd = ol.FlatData(gu.case_data(name))
    res = ol.oracle_predict(fd)
    assert gu.render_sorted(fd, res) == gu.golden_lines(name)


@pytest.mark.parametrize("name", sorted(gu.CASES))
def test_host_machine_matches_reference_gff3(name):
    fd = ol.FlatData(gu.case_data(name))
    res, rounds = ol.host_machine_predict(fd)
    assert gu.render_sorted(fd, res) == gu.golden_lines(name)
    assert ol.results_equal(ol.oracle_predict(fd), res) == []
    assert rounds > 1
    # look-ahead (extra alignments requested per round) never changes a result, only the round count
    for k in ((1000,) if name in gu.HEAVY else (1, 4, 1000)):
        res_k, rounds_k = ol.host_machine_predict(fd, spec_k=k)
        assert ol.results_equal(res, res_k) == []
        assert rounds_k <= rounds


def test_oracle_kernels_match_seqan_vectors():
    O = ol.oracle()
    vec = json.load(open(os.path.join(gu.GOLDEN, "seqan_pairs.json")))
    for a, b, want in vec["edit_distance"]:
        ca = ol.codes_of(np.frombuffer(a.encode(), np.uint8), False)
        cb = ol.codes_of(np.frombuffer(b.encode(), np.uint8), False)
        assert O.orc_edit_distance(ol.ptr(ca, ol.u8p), len(ca), ol.ptr(cb, ol.u8p), len(cb)) == want
        assert O.orc_edit_distance_dp(ol.ptr(ca, ol.u8p), len(ca), ol.ptr(cb, ol.u8p), len(cb)) == want
    out = (ctypes.c_int * 6)()
    for a, b, want in vec["protein"]:
        ca = ol.codes_of(np.frombuffer(a.encode(), np.uint8), True)
        cb = ol.codes_of(np.frombuffer(b.encode(), np.uint8), True)
        O.orc_protein_align(ol.ptr(ca, ol.u8p), len(ca), ol.ptr(cb, ol.u8p), len(cb), out)
        assert list(out) == want


def test_oracle_edit_distance_matches_seqan_long_pairs():
    """4-50 kb pairs (lengths incl. multiples of 32, substitutions and indels up to 50 %, very different
    lengths): the fixture stores the generator seeds and the distance real SeqAn computed."""
    O = ol.oracle()
    for a, b, want in gu.long_pairs():
        ca, cb = ol.codes_of(a, False), ol.codes_of(b, False)
        assert O.orc_edit_distance(ol.ptr(ca, ol.u8p), len(ca), ol.ptr(cb, ol.u8p), len(cb)) == want


def test_alphabet_tables():
    O = ol.oracle()
    for c in range(128):
        assert O.orc_char2dna5(c) == int(ol.DNA_LUT[c])
        assert O.orc_char2aa(c) == int(ol.AA_LUT[c])
    R = ol.seqan_ref()
    if R is not None:  # real SeqAn available (build container)
        for c in range(128):
            assert R.ref_char2dna5(c) == O.orc_char2dna5(c)
            assert R.ref_char2aa(c) == O.orc_char2aa(c)


def test_special_cases_oracle_vs_machine():
    """n==1 records and the 100%-identity shortcut (hh:431-472): oracle vs machine."""
    import synth
    d = gu.case_data("nt_small")
    fd = ol.FlatData(d)
    segs, cands = fd.segs.copy(), fd.cands.copy()
    # segment 0 -> single record; segment 1 -> identical top hit with score ties
    segs[0]["cand_count"] = 1
    s1 = segs[1]
    b, n = int(s1["cand_begin"]), int(s1["cand_count"])
    qs, qe = cands["qstart"][b:b + n].min(), cands["qstop"][b:b + n].max()
    L = int(qe - qs + 1)
    top = b + int(np.argmax(cands["score"][b:b + n]))
    cands["qstart"][top], cands["qstop"][top] = qs, qe
    cands["alnlen"][top] = L; cands["identities"][top] = L; cands["score"][top] = 1e6
    if n > 2:
        others = [k for k in range(b, b + n) if k != top][:2]
        cands["score"][others[0]] = 1e6
        cands["score"][others[1]] = 5e5
    fd.segs, fd.cands = segs, cands
    a = ol.oracle_predict(fd)
    m, _ = ol.host_machine_predict(fd)
    assert ol.results_equal(a, m) == []
    assert a["kind"][0] == 1 and a["kind"][1] == 2
