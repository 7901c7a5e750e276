"""Regenerates the golden GFF3 of the alignment-free models (simple-lca, megan-lca, ic-megan-lca, n-best-lca,
dummy) from the REAL reference binary oracle/_ref/taxator (build container only):
    tests/golden/lca_<case>_<variant>.gff3   sorted output lines, cases/variants of tests/golden_util.py
Run:  python tests/golden/make_golden_lca.py
"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import golden_util as gu  # noqa: E402


def main():
    binary = os.path.join(ROOT, "oracle", "_ref", "taxator")
    for case in gu.LCA_CASES:
        d, evalue, named = gu.lca_case_data(case)
        with tempfile.TemporaryDirectory() as tmp:
            gu.lca_write_files(d, evalue, named, tmp)
            env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
            for variant, (args, _) in gu.LCA_VARIANTS.items():
                with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin:
                    out = subprocess.run([binary] + args + ["-g", "mapping.tax", "-p", "1", "-o", "0"], cwd=tmp, env=env, stdin=fin,
                                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout.decode()
                lines = sorted(l + "\n" for l in out.splitlines() if not l.startswith("##"))
                with open(os.path.join(HERE, "lca_%s_%s.gff3" % (case, variant)), "w") as f:
                    f.writelines(lines)
                print(case, variant, len(lines), "lines,", sum("tax=1:" in l or "tax=1;" in l for l in lines), "at the root")


def distinct():
    """-c >= 2 on the input whose records have pairwise distinct query ranges (golden_util.LCA_C_VARIANTS)."""
    binary = os.path.join(ROOT, "oracle", "_ref", "taxator")
    d, evalue, named = gu.lca_distinct_case_data()
    with tempfile.TemporaryDirectory() as tmp:
        gu.lca_write_files(d, evalue, named, tmp)
        env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
        for variant, (args, _) in gu.LCA_C_VARIANTS.items():
            outs = []
            for threads in ("1", "3"):     # the result must not depend on the allocator / thread interleaving here
                with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin:
                    out = subprocess.run([binary] + args + ["-g", "mapping.tax", "-p", threads, "-o", "0"], cwd=tmp, env=env, stdin=fin,
                                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout.decode()
                outs.append(sorted(l + "\n" for l in out.splitlines() if not l.startswith("##")))
            assert outs[0] == outs[1]
            with open(os.path.join(HERE, "lca_distinct_%s.gff3" % variant), "w") as f:
                f.writelines(outs[0])
            print("distinct", variant, len(outs[0]), "lines,", sum("tax=1:" in l or "tax=1;" in l for l in outs[0]), "at the root")


if __name__ == "__main__":
    main()
    distinct()
