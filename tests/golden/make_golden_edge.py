"""Regenerates the edge-case fixtures from the REAL reference (build container only): tests/edge_data.py builds flat
tables with reference ranges past the sequence ends, N-rich sequences and 300-record sets; they are written out as the
files the reference reads and run through oracle/_ref/taxator.  Output: tests/golden/edge_<case>.gff3 (sorted).
Run:  python tests/golden/make_golden_edge.py
"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import edge_data  # noqa: E402


def main():
    binary = os.path.join(ROOT, "oracle", "_ref", "taxator")
    cases = [(name, fd, "0.5", "0.05") for name, fd in edge_data.reference_cases()]
    # the two parameters of the model: -x (exclude_alignments_factor) and -t (re-evaluation band), hh:329-339
    cases += [("params_x%s_t%s" % (x, t), edge_data.base(seed=24), x, t) for x, t in edge_data.PARAM_SETS]
    for name, fd, x, t in cases:
        with tempfile.TemporaryDirectory() as tmp:
            edge_data.write_files_from_flat(fd, tmp)
            env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
            cmd = [binary, "-a", "rpa", "-g", "mapping.tax", "-q", "query.fna", "-f", "ref.fna", "-i", "ref.fna.fai", "-p", "4",
                   "-x", x, "-t", t, "-o", "0"]
            if fd.protein:
                cmd += ["-b", "protein"]
            with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin:
                p = subprocess.run(cmd, cwd=tmp, env=env, stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            if p.returncode != 0:
                print(name, "REFERENCE FAILED:", p.returncode, p.stderr.decode()[-300:])
                continue
            lines = sorted(l + "\n" for l in p.stdout.decode().splitlines() if not l.startswith("##"))
            open(os.path.join(HERE, "edge_%s.gff3" % name), "w").writelines(lines)
            print(name, len(lines), "lines")


if __name__ == "__main__":
    main()
