"""Regenerates the binner fixtures from the REAL reference (build container only): for every case of
binner_util.CASES the committed GFF3 of the reference taxator (tests/golden/<case>.gff3) is fed to
oracle/_ref/binner (unmodified core/binner.cpp + Boost shims, shipped flags: with -DNDEBUG the reference's regex
match sits inside an assert and is compiled out) with the option sets of binner_util.VARIANTS.
Output: tests/golden/binner_<case>_<variant>.tsv, byte for byte what the reference printed.
Run:  python tests/golden/make_golden_binner.py
"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import binner_util as bu  # noqa: E402
import golden_util as gu  # noqa: E402


def main():
    binary = os.path.join(ROOT, "oracle", "_ref", "binner")
    for case in bu.CASES:
        data = gu.case_data(case)
        with tempfile.TemporaryDirectory() as tmp:
            bu.deep_taxonomy_files(data, tmp)
            gff = "##gff-version 3\n" + "".join(gu.golden_lines(case))
            env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
            for variant, (args, _) in bu.VARIANTS.items():
                p = subprocess.run([binary, "-n", "sample_" + case, "-l", os.path.join(tmp, "binning.log")] + args, cwd=tmp, env=env,
                                   input=gff.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                assert p.returncode == 0, p.stderr.decode()[-500:]
                open(bu.golden_path(case, variant), "wb").write(p.stdout)
                print(case, variant, p.stdout.count(b"\n"), "lines;", p.stderr.decode().replace("\n", " | ")[:160])


def pipeline():
    """`taxator -p 1 | binner`: the reference's two programs piped (record order = the order taxator prints)."""
    taxator = os.path.join(ROOT, "oracle", "_ref", "taxator")
    binary = os.path.join(ROOT, "oracle", "_ref", "binner")
    for case in ("nt_small", "nt_1kb"):
        data = gu.case_data(case)
        with tempfile.TemporaryDirectory() as tmp:
            bu.deep_taxonomy_files(data, tmp)
            env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
            with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin:
                gff = subprocess.run([taxator, "-a", "rpa", "-g", "mapping.tax", "-q", "query.fna", "-f", "ref.fna", "-i", "ref.fna.fai",
                                      "-p", "1", "-x", "0.5", "-o", "0"], cwd=tmp, env=env, stdin=fin, stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, check=True).stdout
            for variant in ("default", "glob10"):
                p = subprocess.run([binary, "-n", "sample_" + case, "-l", os.path.join(tmp, "binning.log")] + bu.VARIANTS[variant][0],
                                   cwd=tmp, env=env, input=gff, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                assert p.returncode == 0, p.stderr.decode()[-500:]
                open(bu.golden_path(case, "pipeline_" + variant), "wb").write(p.stdout)
                print("pipeline", case, variant, p.stdout.count(b"\n"), "lines")


if __name__ == "__main__":
    main()
    pipeline()
