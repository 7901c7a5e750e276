"""Dev script: wall time of the taxator-b200 CLI on a C2-shaped file set (not a bench value)."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import bench  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
workload = sys.argv[2] if len(sys.argv) > 2 else "c2"
extra = sys.argv[3:]
d = bench.make_data(workload, 20261017, n_queries=nq)
tmp = tempfile.mkdtemp(prefix="trpa_cli_")
t0 = time.time()
d.write_files(tmp)
print("files written in %.1f s: %s" % (time.time() - t0, {f: os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp)}), flush=True)
env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
binary = os.path.join(ROOT, "taxator-tk_b200", "bin", "taxator-b200")
cmd = [binary, "-a", "rpa", "-g", "mapping.tax", "-q", "query.fna", "-f", "ref.fna", "-i", "ref.fna.fai", "-x", "0.5", "-o", "0",
       "--timing"] + extra
for rep in range(2):
    with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin, open(os.path.join(tmp, "out.gff3"), "wb") as fout:
        t0 = time.perf_counter()
        p = subprocess.run(cmd, cwd=tmp, env=env, stdin=fin, stdout=fout, stderr=subprocess.PIPE)
        dt = time.perf_counter() - t0
    print("run %d: rc=%d wall %.2f s (%.0f segments/s)  %s" % (rep, p.returncode, dt, nq / dt, "\n   ".join(l for l in p.stderr.decode().strip().splitlines() if l.startswith("taxator-b200"))), flush=True)
