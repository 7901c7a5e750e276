"""Regenerates the verbose-log fixtures from the REAL reference (build container only): `oracle/_ref/taxator -l`
(-p 1) on the first queries of two golden cases.  Output: tests/golden/log_<case>.log.gz.
Run:  python tests/golden/make_golden_log.py
"""
import gzip
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import log_util as lu  # noqa: E402


def main():
    binary = os.path.join(ROOT, "oracle", "_ref", "taxator")
    for case, nq in lu.LOG_CASES.items():
        data = lu.log_case_data(case)
        with tempfile.TemporaryDirectory() as tmp:
            data.write_files(tmp)
            env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
            cmd = [binary, "-a", "rpa", "-g", "mapping.tax", "-q", "query.fna", "-f", "ref.fna", "-i", "ref.fna.fai", "-p", "1",
                   "-x", "0.5", "-o", "0", "-l", "ref.log"]
            if data.cfg.protein:
                cmd += ["-b", "protein"]
            with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin:
                subprocess.run(cmd, cwd=tmp, env=env, stdin=fin, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
            raw = open(os.path.join(tmp, "ref.log"), "rb").read()
        with gzip.GzipFile(lu.golden_log_path(case), "wb", mtime=0) as f:
            f.write(raw)
        print(case, len(raw), "bytes,", raw.count(b"\nID\t") + raw.startswith(b"ID\t"), "segments")


if __name__ == "__main__":
    main()
