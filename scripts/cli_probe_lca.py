"""Dev script: wall time of `taxator-b200 -a megan-lca` against the reference binary on a C2-shaped alignment file
(not a bench value).  usage: cli_probe_lca.py [n_queries]"""
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
d = bench.make_data("c2", 20261017, n_queries=nq)
tmp = tempfile.mkdtemp(prefix="trpa_cli_lca_")
t0 = time.time()
d.write_files(tmp)
size = os.path.getsize(os.path.join(tmp, "alignments.tsv"))
print("files written in %.1f s, alignments.tsv = %.1f MB" % (time.time() - t0, size / 1e6), flush=True)
env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
outs = {}
for name, binary, extra in (("reference -p 16", os.path.join(ROOT, "oracle", "_ref", "taxator"), ["-p", str(os.cpu_count() or 1)]),
                            ("taxator-b200", os.path.join(ROOT, "taxator-tk_b200", "bin", "taxator-b200"), []),
                            ("taxator-b200 --legacy-ingest", os.path.join(ROOT, "taxator-tk_b200", "bin", "taxator-b200"), ["--legacy-ingest"])):
    if not os.path.exists(binary):
        print(name, "missing")
        continue
    cmd = [binary, "-a", "megan-lca", "-g", "mapping.tax", "-o", "0"] + extra
    for rep in range(2):
        with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin:
            t0 = time.perf_counter()
            p = subprocess.run(cmd, cwd=tmp, env=env, stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            dt = time.perf_counter() - t0
        print("%-30s run %d: rc=%d wall %.2f s (%.0f segments/s)" % (name, rep, p.returncode, dt, nq / dt), flush=True)
    outs[name] = sorted(p.stdout.decode().splitlines())
names = list(outs)
for n in names[1:]:
    print("output of %s identical to %s: %s" % (n, names[0], outs[n] == outs[names[0]]))
