"""The alignment-free placement models (simple-lca, megan-lca / ic-megan-lca, n-best-lca, dummy;
core/src/taxonpredictionmodel.hh:57-259) -- SURVEY.md 8 f4.
CPU: the oracle restatement reproduces the GFF3 of the REAL reference binary (tests/golden/lca_*.gff3, generated
by tests/golden/make_golden_lca.py).  GPU: trpa_predict_lca_batch equals the oracle on the golden cases and on
adversarial random tables, and the drop-in CLI reproduces the reference's files byte for byte."""
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
import oracle_lib as ol
import synth

LCA_FIELDS = ["qrstart", "qrstop", "lower_node", "upper_node", "rtax_node", "support", "kind", "ival"]


def _render(d, parent, depth, segs, res):
    import gff3
    q_len = np.array([len(s) for s in d.q_seqs], np.uint32)
    return sorted(gff3.render(res, segs, d.q_names, q_len, parent, depth, [str(t) for t in d.tax_ids]))


def _case(case):
    d, evalue, named = gu.lca_case_data(case)
    parent, left, right, depth = d.nested_set()
    segs, cands, ev = gu.lca_flat(d, evalue)
    uncl = gu.lca_unclassified_flags(d, named)
    return d, (parent, left, right, depth), segs, cands, ev, uncl


@pytest.mark.parametrize("case", gu.LCA_CASES)
def test_oracle_matches_reference_gff3(case):
    d, tax, segs, cands, ev, uncl = _case(case)
    for variant, (_, kw) in gu.LCA_VARIANTS.items():
        res = ol.oracle_predict_lca(*tax, segs, cands, ev, uncl, **kw)
        assert _render(d, tax[0], tax[3], segs, res) == gu.lca_golden_lines(case, variant), variant


def _distinct_case():
    d, evalue, named = gu.lca_distinct_case_data()
    parent, left, right, depth = d.nested_set()
    segs, cands, ev = gu.lca_flat(d, evalue)
    return d, (parent, left, right, depth), segs, cands, ev, gu.lca_unclassified_flags(d, named)


def _distinct_golden(variant):
    return open(os.path.join(gu.GOLDEN, "lca_distinct_%s.gff3" % variant)).readlines()


def test_oracle_matches_reference_with_min_support():
    """megan-lca -c 2 / 3 / 4 pinned against the real binary on records with pairwise distinct query ranges (where the
    reference's record order does not depend on heap addresses)."""
    d, tax, segs, cands, ev, uncl = _distinct_case()
    for variant, (_, kw) in gu.LCA_C_VARIANTS.items():
        res = ol.oracle_predict_lca(*tax, segs, cands, ev, uncl, **kw)
        assert _render(d, tax[0], tax[3], segs, res) == _distinct_golden(variant), variant


@pytest.mark.gpu
def test_gpu_matches_reference_with_min_support(ctx):
    d, tax, segs, cands, ev, uncl = _distinct_case()
    ctx.load_taxonomy(*tax, 0)
    for variant, (_, kw) in gu.LCA_C_VARIANTS.items():
        got, _ = ctx.predict_lca_batch(kw["model"], segs, cands, ev, uncl, **{k: v for k, v in kw.items() if k != "model"})
        assert _render(d, tax[0], tax[3], segs, got) == _distinct_golden(variant), variant


def _random_tables(rng, n_nodes, n_segs):
    """Adversarial record sets: ties, reversed query ranges, empty and > 32-record sets, scores below / above 0."""
    counts = rng.choice([0, 1, 2, 3, 7, 31, 32, 33, 64, 65, 150], n_segs)
    n = int(counts.sum())
    cands = np.zeros(n, synth.CAND_DTYPE)
    cands["score"] = rng.choice(np.array([-8, 0, 8, 16, 40, 40, 48, 56, 120, 128, 500], np.float32), n)
    a = rng.integers(1, 5000, n); b = rng.integers(1, 5000, n)
    cands["qstart"], cands["qstop"] = a, b            # either orientation
    cands["node"] = rng.integers(0, n_nodes, n)
    ev = np.power(10.0, -rng.integers(-4, 30, n).astype(np.float64))
    segs = np.zeros(n_segs, synth.SEG_DTYPE)
    segs["cand_count"] = counts
    segs["cand_begin"] = np.concatenate([[0], np.cumsum(counts)[:-1]])
    return segs, cands, ev


PARAM_SETS = [dict(model=0), dict(model=1), dict(model=2), dict(model=2, toppercent=1.0, minsupport=1),
              dict(model=2, toppercent=0.3, minscore=16.0, maxevalue=1e-3, minsupport=2),
              dict(model=2, toppercent=0.0, minscore=-100.0, maxevalue=1e-20, minsupport=1, ignore_unclassified=True),
              dict(model=2, toppercent=0.5, minsupport=4, ignore_unclassified=True),
              dict(model=3, nbest=0), dict(model=3, nbest=1), dict(model=3, nbest=2), dict(model=3, nbest=50)]


@pytest.mark.gpu
def test_gpu_equals_oracle_random(ctx):
    d, tax, _, _, _, uncl = _case("nt_small")
    ctx.load_taxonomy(*tax, 0)
    rng = np.random.default_rng(77)
    segs, cands, ev = _random_tables(rng, len(tax[0]), 400)
    for kw in PARAM_SETS:
        for use_ev, use_un in ((True, True), (False, False)):
            want = ol.oracle_predict_lca(*tax, segs, cands, ev if use_ev else None, uncl if use_un else None, **kw)
            got, _ = ctx.predict_lca_batch(kw["model"], segs, cands, ev if use_ev else None, uncl if use_un else None,
                                           **{k: v for k, v in kw.items() if k != "model"})
            bad = ol.results_equal(got, want, LCA_FIELDS)
            assert not bad, (kw, use_ev, bad)
    # empty batch, bad node
    assert len(ctx.predict_lca_batch(1, segs[:0], cands[:0])[0]) == 0
    import rpa_b200
    bad = cands.copy(); bad["node"][3] = 10 ** 6
    with pytest.raises(rpa_b200.TrpaError):
        ctx.predict_lca_batch(1, segs, bad)


@pytest.mark.gpu
@pytest.mark.parametrize("case", gu.LCA_CASES)
def test_gpu_matches_reference_gff3(ctx, case):
    d, tax, segs, cands, ev, uncl = _case(case)
    ctx.load_taxonomy(*tax, 0)
    for variant, (_, kw) in gu.LCA_VARIANTS.items():
        got, _ = ctx.predict_lca_batch(kw["model"], segs, cands, ev, uncl, **{k: v for k, v in kw.items() if k != "model"})
        assert _render(d, tax[0], tax[3], segs, got) == gu.lca_golden_lines(case, variant), variant


@pytest.mark.gpu
def test_cli_matches_reference_gff3(tmp_path):
    exe = os.path.join(ol.ROOT, "taxator-tk_b200", "bin", "taxator-b200")
    assert os.path.exists(exe), "build first: make -C taxator-tk_b200"
    case = "nt_small"
    d, evalue, named = gu.lca_case_data(case)
    tmp = str(tmp_path / case)
    gu.lca_write_files(d, evalue, named, tmp)
    env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
    # the block-parallel ingest (default, small blocks so that several are in flight) and the record-at-a-time path
    for extra in (["--batch-bytes", "20000"], ["--legacy-ingest", "--batch-segments", "50"]):
        for variant, (args, _) in gu.LCA_VARIANTS.items():
            with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin:
                out = subprocess.run([exe] + args + ["-g", "mapping.tax", "-p", "1", "-o", "0"] + extra, cwd=tmp,
                                     env=env, stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout.decode()
            assert out.startswith("##gff-version 3\n")
            lines = sorted(l + "\n" for l in out.splitlines() if not l.startswith("##"))
            assert lines == gu.lca_golden_lines(case, variant), (variant, extra)


def _build_host_lca_harness():
    host = os.path.join(ol.ROOT, "taxator-tk_b200", "host")
    os.makedirs(ol.BUILD_DIR, exist_ok=True)
    exe = os.path.join(ol.BUILD_DIR, "host_lca_harness")
    src = os.path.join(ol.ROOT, "tests", "host_lca_harness.cpp")
    deps = [src] + [os.path.join(host, f) for f in os.listdir(host)]
    oracle_so = ol.build_oracle()
    if not ol._newer(exe, *deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ol.ROOT, "include"), "-o", exe, src,
                               os.path.join(host, "taxonomy.cpp"), os.path.join(host, "seqstore.cpp"),
                               os.path.join(host, "records.cpp"), os.path.join(host, "rpa_model.cpp"),
                               os.path.join(host, "ingest.cpp"), oracle_so,
                               os.path.join(ol.ROOT, "taxator-tk_b200", "lib", "libtaxator_rpa_b200.so"),
                               "-Wl,-rpath," + ol.ORACLE_DIR,
                               "-Wl,-rpath," + os.path.join(ol.ROOT, "taxator-tk_b200", "lib"), "-lz", "-lpthread"])
    return exe


@pytest.mark.parametrize("case", gu.LCA_CASES)
def test_host_path_matches_reference_gff3(tmp_path, case):
    """CPU: taxonomy loader ('unclassified' flags), parser, store-less block ingest with e-values and the GFF3
    formatter of the CLI, with the oracle in place of the GPU, on the reference's own input files."""
    exe = _build_host_lca_harness()
    d, evalue, named = gu.lca_case_data(case)
    tmp = str(tmp_path / case)
    gu.lca_write_files(d, evalue, named, tmp)
    env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
    for variant, (_, kw) in gu.LCA_VARIANTS.items():
        args = [str(kw["model"]), repr(kw.get("toppercent", 0.05)), repr(kw.get("minscore", 0.0)), repr(kw.get("maxevalue", 1000.0)),
                str(kw.get("minsupport", 1)), str(kw.get("nbest", 1)), "1" if kw.get("ignore_unclassified") else "0"]
        for block in ("134217728", "30000"):
            with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin:
                p = subprocess.run([exe] + args + ["mapping.tax", block], cwd=tmp, env=env, stdin=fin, stdout=subprocess.PIPE,
                                   stderr=subprocess.PIPE)
            assert p.returncode == 0, p.stderr.decode()
            out = p.stdout.decode()
            assert out.startswith("##gff-version 3\n")
            lines = sorted(l + "\n" for l in out.splitlines() if not l.startswith("##"))
            assert lines == gu.lca_golden_lines(case, variant), (variant, block)
