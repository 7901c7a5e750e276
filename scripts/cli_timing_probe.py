"""Dev script: wall clock and --timing lines of taxator-b200 on n C2 segments, repeated (CUDA start-up variance, pinned
ingest blocks A/B with TRPA_NO_PINNED=1)."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
d = bench.make_data("c2", 20261017, n_queries=n)
tmp = tempfile.mkdtemp(prefix="cli_probe_")
d.write_files(tmp)
exe = os.path.join(ROOT, "taxator-tk_b200", "bin", "taxator-b200")
cmd = [exe, "-a", "rpa", "-g", "mapping.tax", "-q", "query.fna", "-f", "ref.fna", "-i", "ref.fna.fai", "-x", "0.5", "-o", "0", "--timing"]
for label, extra in (("pinned", {}), ("pinned", {}), ("plain", {"TRPA_NO_PINNED": "1"}), ("plain", {"TRPA_NO_PINNED": "1"}), ("pinned", {})):
    env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp, **extra)
    with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin:
        t0 = time.perf_counter()
        p = subprocess.run(cmd, cwd=tmp, env=env, stdin=fin, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        dt = time.perf_counter() - t0
    print(label, "wall %.3f s" % dt, flush=True)
    print("   " + p.stderr.decode().replace("\n", "\n   "), flush=True)
