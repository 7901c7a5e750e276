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
                               oracle_so, os.path.join(ol.ROOT, "taxator-tk_b200", "lib", "libtaxator_rpa_b200.so"),
                               "-Wl,-rpath," + ol.ORACLE_DIR,
                               "-Wl,-rpath," + os.path.join(ol.ROOT, "taxator-tk_b200", "lib"), "-lz", "-lpthread"])
    return exe


def run_harness(data, batch=100000, extra_lines=None, sorted_input=False):
    exe = build_harness()
    with tempfile.TemporaryDirectory() as tmp:
        data.write_files(tmp)
        aln = open(os.path.join(tmp, "alignments.tsv")).read()
        if extra_lines:
            aln = extra_lines(aln)
        env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
        p = subprocess.run([exe, "protein" if data.cfg.protein else "nucleotide", "mapping.tax", "query.fna", "ref.fna",
                            "ref.fna.fai", str(batch), "1", "1" if sorted_input else "0"], cwd=tmp, env=env,
                           input=aln.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
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
