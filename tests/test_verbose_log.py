"""The verbose per-alignment log (core/src/taxonpredictionmodelsequence.hh:341-838, `taxator -l`).
CPU: the product's log writer (host/verbose_log.cpp) fed with the alignment trace of the host-compiled state machine
reproduces the log of the REAL reference (tests/golden/log_*.log.gz, generator make_golden_log.py) block for block --
every ID / PASS / +ALN / current ... node / EXT / SCORE / NUMALN / RANGE / STATS line and, for protein, SeqAn's
rendering of every alignment; only the three CPU-time columns of STATS are masked.
GPU: the same through taxator-b200 -l (trace recorded by the decide kernel)."""
import os
import subprocess

import pytest

import log_util as lu
import oracle_lib as ol


@pytest.mark.parametrize("case", sorted(lu.LOG_CASES))
@pytest.mark.parametrize("spec_k", [0, 1000])
def test_log_writer_reproduces_reference_log(case, spec_k, tmp_path):
    fd = ol.FlatData(lu.log_case_data(case))
    res, _, trace = ol.host_machine_predict(fd, spec_k=spec_k, want_trace=True)
    out = str(tmp_path / "ours.log")
    lu.write_log(fd, res, trace, out)
    ours = lu.blocks_of(open(out).read())
    want = lu.golden_blocks(case)
    assert len(ours) == len(want)
    assert sorted(ours) == sorted(want)
    if fd.protein:
        assert any("        " in b and "|" in b for b in ours)    # alignment renderings present


def test_corner_cases_against_reference_log_and_gff3(tmp_path):
    """n == 1 and n == 0 through masked records, the 100 % top-hit shortcut with a score tie, a 100 % hit that is not the
    best score: GFF3 and log of the REAL reference (tests/golden/special.gff3, log_special.log.gz) vs the host-compiled
    state machine + the product's log writer."""
    import golden_util as gu
    d, masked = lu.special_case()
    fd = lu.flat_with_masks(d, masked)
    res, _, trace = ol.host_machine_predict(fd, want_trace=True)
    assert set(int(k) for k in res["kind"]) >= {0, 1, 2, 3}
    out = str(tmp_path / "ours.log")
    lu.write_log(fd, res, trace, out)
    ours, want = lu.blocks_of(open(out).read()), lu.golden_blocks("special")
    assert sorted(ours) == sorted(want)
    text = "\n".join(ours)
    assert "*ALN" in text and "current ref/lower node" in text and "NUMREF\t1\n" in text and "NUMREF\t0\n" in text
    want_gff = [l for l in open(os.path.join(gu.GOLDEN, "special.gff3")).readlines() if not l.startswith("##")]
    import gff3
    taxids = [str(t) for t in d.tax_ids]
    got_gff = gff3.render(res, fd.segs, d.q_names, fd.q_len, fd.parent, fd.depth, taxids)
    assert got_gff == want_gff          # in input order: the n == 0 set inherits the ival of the record before it


def test_log_writer_detects_a_wrong_trace(tmp_path):
    fd = ol.FlatData(lu.log_case_data("nt_small"))
    res, _, trace = ol.host_machine_predict(fd, want_trace=True)
    bad = trace.copy()
    bad["a"][len(bad) // 2] = 77          # not the pair the control flow asks for at this point
    with pytest.raises(AssertionError, match="diverged"):
        lu.write_log(fd, res, bad, str(tmp_path / "x.log"))
    with pytest.raises(AssertionError, match="diverged"):
        lu.write_log(fd, res, trace[:-1], str(tmp_path / "y.log"))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["nt_small", "aa_300", "nt_mixed"])
def test_gpu_trace_equals_host_machine_trace(ctx, case):
    """trpa_set_trace / trpa_batch_trace: the decide kernel records exactly the alignments (pair, raw integers, lengths,
    self scores) that the host-compiled state machine consumes with the oracle's alignments, in the same order -- with
    and without look-ahead, and tracing never changes a result."""
    import golden_util as gu
    import numpy as np
    fd = ol.FlatData(gu.case_data(case))
    want_res, _, want = ol.host_machine_predict(fd, want_trace=True)
    ctx.load_taxonomy(fd.parent, fd.left, fd.right, fd.depth, 0)
    alpha = 1 if fd.protein else 0
    ctx.load_store(0, alpha, fd.q_chars, fd.q_off, fd.q_len)
    ctx.load_store(1, alpha, fd.r_chars, fd.r_off, fd.r_len)
    plain = ctx.predict_batch(fd.segs, fd.cands)
    try:
        ctx.set_trace(1)
        for la in (-1, 0):
            ctx.set_lookahead(la)
            got_res = ctx.predict_batch(fd.segs, fd.cands)
            got = ctx.batch_trace()
            assert ol.results_equal(plain, got_res) == [] and ol.results_equal(want_res, got_res) == []
            assert len(got) == len(want) == int((got_res["n_pass0"] + got_res["n_pass1"] + got_res["n_pass2"]).sum())
            for f in got.dtype.names:
                assert np.array_equal(got[f], want[f]), f
    finally:
        ctx.set_trace(0)
        ctx.set_lookahead(-1)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(lu.LOG_CASES))
def test_cli_verbose_log_matches_reference(case, tmp_path):
    exe = os.path.join(ol.ROOT, "taxator-tk_b200", "bin", "taxator-b200")
    data = lu.log_case_data(case)
    d = str(tmp_path)
    data.write_files(d)
    env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=d)
    cmd = [exe, "-a", "rpa", "-g", "mapping.tax", "-q", "query.fna", "-f", "ref.fna", "-i", "ref.fna.fai", "-l", "ours.log"]
    if data.cfg.protein:
        cmd += ["-b", "protein"]
    outs = []
    for extra in ([], ["--batch-bytes", "8192"]):     # one block / many blocks (the trace is per block)
        if os.path.exists(os.path.join(d, "ours.log")):
            os.remove(os.path.join(d, "ours.log"))
        with open(os.path.join(d, "alignments.tsv"), "rb") as fin:
            p = subprocess.run(cmd + extra, cwd=d, env=env, stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode == 0, p.stderr.decode()
        outs.append(p.stdout)
        ours = lu.blocks_of(open(os.path.join(d, "ours.log")).read())
        assert sorted(ours) == sorted(lu.golden_blocks(case))
    assert outs[0] == outs[1]


@pytest.mark.gpu
def test_cli_corner_cases_match_reference(tmp_path):
    """taxator-b200 on an alignment file with masked records (n == 1, n == 0), the 100 % shortcut and a *ALN record:
    GFF3 byte-identical to the real reference (incl. the ival an n == 0 set inherits) and the same log."""
    import golden_util as gu
    exe = os.path.join(ol.ROOT, "taxator-tk_b200", "bin", "taxator-b200")
    data, masked = lu.special_case()
    d = str(tmp_path)
    lu.write_case_files(data, masked, d)
    env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=d)
    cmd = [exe, "-a", "rpa", "-g", "mapping.tax", "-q", "query.fna", "-f", "ref.fna", "-i", "ref.fna.fai", "-l", "ours.log"]
    with open(os.path.join(d, "alignments.tsv"), "rb") as fin:
        p = subprocess.run(cmd, cwd=d, env=env, stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()
    assert p.stdout == open(os.path.join(gu.GOLDEN, "special.gff3"), "rb").read()
    assert sorted(lu.blocks_of(open(os.path.join(d, "ours.log")).read())) == sorted(lu.golden_blocks("special"))
