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


def mask_timers(raw):
    """The three CPU-time columns of the STATS lines differ from run to run: stored as 0 (the tests mask them anyway)."""
    out = []
    for line in raw.split(b"\n"):
        if line.startswith(b"STATS\t"):
            f = line.split(b"\t")
            if len(f) >= 11:
                f[7] = f[8] = f[9] = b"0"
            line = b"\t".join(f)
        out.append(line)
    return b"\n".join(out)


def special(binary):
    """The corner cases of predict() (log_util.special_case): log AND GFF3 of the real reference."""
    data, masked = lu.special_case()
    with tempfile.TemporaryDirectory() as tmp:
        lu.write_case_files(data, masked, tmp)
        env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
        cmd = [binary, "-a", "rpa", "-g", "mapping.tax", "-q", "query.fna", "-f", "ref.fna", "-i", "ref.fna.fai", "-p", "1",
               "-x", "0.5", "-o", "0", "-l", "ref.log"]
        with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin:
            out = subprocess.run(cmd, cwd=tmp, env=env, stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        raw = open(os.path.join(tmp, "ref.log"), "rb").read()
    with gzip.GzipFile(lu.golden_log_path("special"), "wb", mtime=0) as f:
        f.write(mask_timers(raw))
    open(os.path.join(HERE, "special.gff3"), "wb").write(out)
    print("special", len(raw), "bytes,", raw.count(b"*ALN"), "*ALN,", raw.count(b"current ref/lower node"), "shortcut lines,",
          raw.count(b"NUMREF\t1\n"), "n==1,", raw.count(b"NUMREF\t0\n"), "n==0")


def main():
    binary = os.path.join(ROOT, "oracle", "_ref", "taxator")
    special(binary)
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
            f.write(mask_timers(raw))
        print(case, len(raw), "bytes,", raw.count(b"\nID\t") + raw.startswith(b"ID\t"), "segments")


if __name__ == "__main__":
    main()
