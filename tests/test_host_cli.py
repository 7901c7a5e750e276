"""CPU tests of the host-side C++ (taxator-tk_b200/host): taxonomy / FASTA+.fai / mapping loaders,
alignment parser, record-set segmentation, candidate flattening, driver loop and GFF3 printing --
with the GPU predictor replaced by the oracle (tests/host_cli_harness.cpp).  The GFF3 must equal the
committed output of the REAL reference on the same files."""
import os
import subprocess
import tempfile

import pytest

import golden_util as gu
import oracle_lib as ol

HOST = os.path.join(ol.ROOT, "taxator-tk_b200", "host")


def build_harness():
    os.makedirs(ol.BUILD_DIR, exist_ok=True)
    exe = os.path.join(ol.BUILD_DIR, "host_cli_harness")
    src = os.path.join(ol.ROOT, "tests", "host_cli_harness.cpp")
    deps = [src] + [os.path.join(HOST, f) for f in os.listdir(HOST)]
    oracle_so = ol.build_oracle()
    if not ol._newer(exe, *deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ol.ROOT, "include"), "-o", exe, src,
                               os.path.join(HOST, "taxonomy.cpp"), os.path.join(HOST, "seqstore.cpp"),
                               os.path.join(HOST, "records.cpp"), os.path.join(HOST, "rpa_model.cpp"),
                               os.path.join(HOST, "ingest.cpp"),
                               oracle_so, os.path.join(ol.ROOT, "taxator-tk_b200", "lib", "libtaxator_rpa_b200.so"),
                               "-Wl,-rpath," + ol.ORACLE_DIR,
                               "-Wl,-rpath," + os.path.join(ol.ROOT, "taxator-tk_b200", "lib"), "-lz", "-lpthread"])
    return exe


def run_harness(data, batch=100000, extra_lines=None, sorted_input=False, fast_bytes=0, split=True, expect_fail=False):
    exe = build_harness()
    with tempfile.TemporaryDirectory() as tmp:
        data.write_files(tmp)
        aln = open(os.path.join(tmp, "alignments.tsv")).read()
        if extra_lines:
            aln = extra_lines(aln)
        env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
        p = subprocess.run([exe, "protein" if data.cfg.protein else "nucleotide", "mapping.tax", "query.fna", "ref.fna",
                            "ref.fna.fai", str(batch), "1" if split else "0", "1" if sorted_input else "0", str(fast_bytes)],
                           cwd=tmp, env=env, input=aln.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if expect_fail:
            assert p.returncode != 0
            return p.stdout.decode().splitlines(keepends=True)[1:], p.stderr.decode()
        assert p.returncode == 0, p.stderr.decode()
        out = p.stdout.decode().splitlines(keepends=True)
    assert out[0] == "##gff-version 3\n"
    return out[1:]


@pytest.mark.parametrize("name", ["nt_small", "aa_small", "nt_indel"])
def test_host_pipeline_matches_reference(name):
    lines = run_harness(gu.case_data(name))
    assert sorted(lines) == gu.golden_lines(name)


def test_small_batches_and_comments_keep_order():
    data = gu.case_data("nt_small")
    a = run_harness(data)
    b = run_harness(data, batch=7, extra_lines=lambda s: "# a comment line\n" + s)
    assert a == b


def test_masked_records_and_gz_taxonomy():
    """'*' lines stay in the record set for segmentation but are not candidates (alignmentrecord.hh:98-104)."""
    data = gu.case_data("nt_small")

    def mask_some(s):
        lines = s.splitlines(keepends=True)
        return "".join(("*" + l) if i % 7 == 3 else l for i, l in enumerate(lines))

    out = run_harness(data, extra_lines=mask_some)
    assert len(out) > 0
    # fully masked query -> unclassified root record over the whole query
    first_q = data.q_names[0]

    def mask_first(s):
        return "".join(("*" + l) if l.startswith(first_q + "\t") else l for l in s.splitlines(keepends=True))

    out2 = run_harness(data, extra_lines=mask_first)
    hit = [l for l in out2 if l.startswith(first_q + "\t")]
    assert hit and all("tax=1;rtax=1" in l for l in hit)


@pytest.mark.parametrize("name", ["nt_small", "aa_small", "nt_indel"])
def test_fast_ingest_matches_reference(name):
    """The block-parallel ingest path of the CLI (host/ingest.cpp) against the real reference's GFF3."""
    lines = run_harness(gu.case_data(name), fast_bytes=1 << 20)
    assert sorted(lines) == gu.golden_lines(name)


def test_fast_ingest_equals_record_at_a_time_path():
    """Same bytes out for every block size (blocks are cut at query boundaries, carried text, several
    parser threads), with comments, masked lines, oddly formatted numbers and no trailing newline."""
    data = gu.case_data("nt_small")

    def decorate(s):
        lines = s.splitlines(keepends=True)
        out = ["# header comment\n"]
        for i, l in enumerate(lines):
            if i % 11 == 5:
                l = "*" + l
            if i % 13 == 7:   # boost::lexical_cast<float> accepts an explicit plus sign
                f = l.split("\t")
                f[7] = "+" + f[7]
                l = "\t".join(f)
            if i % 17 == 3:
                out.append("# in between\n")
            out.append(l)
        return "".join(out).rstrip("\n")

    ref = run_harness(data, extra_lines=decorate)
    for fb in (4096, 5000, 20000, 1 << 16, 1 << 24):
        assert run_harness(data, extra_lines=decorate, fast_bytes=fb) == ref, fb
    ref_nosplit = run_harness(data, extra_lines=decorate, split=False)
    assert run_harness(data, extra_lines=decorate, fast_bytes=8192, split=False) == ref_nosplit
    assert ref_nosplit != ref


def test_fast_ingest_reports_errors_like_the_record_at_a_time_path():
    data = gu.case_data("nt_small")

    def break_line(s):
        lines = s.splitlines(keepends=True)
        f = lines[40].split("\t")
        f[2] = "12x"
        lines[40] = "\t".join(f)
        return "".join(lines)

    _, err_a = run_harness(data, extra_lines=break_line, expect_fail=True)
    _, err_b = run_harness(data, extra_lines=break_line, fast_bytes=4096, expect_fail=True)
    assert "line 41" in err_a and err_a == err_b

    def unknown_ref(s):
        lines = s.splitlines(keepends=True)
        f = lines[10].split("\t")
        f[4] = "NOSUCHREF"
        lines[10] = "\t".join(f)
        return "".join(lines)

    # boost::lexical_cast accepts neither leading blanks nor hexadecimal floats (strtoull / strtof would)
    for col, text in ((1, " 12"), (7, " 55.0"), (7, "0x1p4"), (8, " 0")):
        def blank(s, col=col, text=text):
            lines = s.splitlines(keepends=True)
            f = lines[7].split("\t")
            f[col] = text
            lines[7] = "\t".join(f)
            return "".join(lines)
        _, err_a = run_harness(data, extra_lines=blank, expect_fail=True)
        _, err_b = run_harness(data, extra_lines=blank, fast_bytes=4096, expect_fail=True)
        assert "line 8" in err_a and err_a == err_b, (col, text, err_a, err_b)

    _, err_a = run_harness(data, extra_lines=unknown_ref, expect_fail=True)
    _, err_b = run_harness(data, extra_lines=unknown_ref, fast_bytes=4096, expect_fail=True)
    assert "NOSUCHREF" in err_a and err_a == err_b
