"""Regenerates the golden fixtures from the REAL reference (build container only):
  * <case>.gff3       sorted GFF3 of oracle/_ref/taxator (unmodified reference sources + Boost
                      shims, -DNDEBUG) on the synthetic data of cases.json[<case>]
  * seqan_pairs.json  per-pair integers from the unmodified vendored SeqAn
                      (oracle/_ref/libseqan_ref.so): edit distances and protein
                      (mutual, self, traced length) triples
Run:  python tests/golden/make_golden.py
"""
import ctypes
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import synth  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import long_pair  # noqa: E402

CASES = {
    "nt_small": dict(seed=101, protein=False, n_genomes=60, genome_len=3000, n_queries=150, query_len=[200, 600],
                     n_cand=25, levels=[2, 4, 6, 10, 20], multi_segment_frac=0.2, frac_n=0.002),
    "nt_1kb": dict(seed=102, protein=False, n_genomes=120, genome_len=8000, n_queries=200, query_len=[1000, 1000],
                   n_cand=36, levels=[3, 6, 10, 20, 50]),
    "nt_indel": dict(seed=103, protein=False, n_genomes=40, genome_len=4000, n_queries=100, query_len=[300, 1200],
                     n_cand=20, levels=[2, 3, 5, 8, 12], query_indel=0.1),
    "nt_5kb": dict(seed=104, protein=False, n_genomes=80, genome_len=12000, n_queries=40, query_len=[4000, 6000],
                   n_cand=40, levels=[2, 4, 8, 16, 30]),
    "aa_small": dict(seed=105, protein=True, n_genomes=60, genome_len=600, n_queries=150, query_len=[80, 300],
                     n_cand=25, levels=[2, 4, 6, 10, 20], multi_segment_frac=0.2, frac_n=0.002),
    "aa_300": dict(seed=106, protein=True, n_genomes=100, genome_len=400, n_queries=120, query_len=[300, 300],
                   n_cand=30, levels=[3, 6, 10, 20, 40]),
    # BASELINE.json configs[3] regime: mixed 0.5-20 kb segments in one batch (length-bucketed shapes, wedge)
    "nt_mixed": dict(seed=107, protein=False, n_genomes=60, genome_len=30000, n_queries=48, query_len=[500, 20000],
                     n_cand=20, levels=[2, 4, 6, 10, 20]),
    # configs[4] regime: long noisy reads, 5 % substitutions + 10 % indels (wide bands, band retries)
    "nt_longread": dict(seed=108, protein=False, n_genomes=40, genome_len=40000, n_queries=16, query_len=[10000, 20000],
                        n_cand=12, levels=[2, 3, 5, 8, 12], query_sub=0.05, query_indel=0.10),
}
# cases whose CPU checks are run once (no look-ahead sweep) to keep the CPU suite short
HEAVY = ("nt_mixed", "nt_longread")


def cfg_of(case):
    d = dict(case)
    d["query_len"] = tuple(d["query_len"])
    d["levels"] = tuple(d["levels"])
    return synth.SynthConfig(**d)


def run_reference(data, binary, threads=1):
    with tempfile.TemporaryDirectory() as tmp:
        data.write_files(tmp)
        env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
        cmd = [binary, "-a", "rpa", "-g", "mapping.tax", "-q", "query.fna", "-f", "ref.fna", "-i", "ref.fna.fai",
               "-p", str(threads), "-x", "0.5", "-o", "0"]
        if data.cfg.protein:
            cmd += ["-b", "protein"]
        with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin:
            out = subprocess.run(cmd, cwd=tmp, env=env, stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                 check=True).stdout.decode()
    lines = [l + "\n" for l in out.splitlines() if not l.startswith("##")]
    return sorted(lines)


LONG_PAIRS = [dict(seed=9000 + 37 * i, L=L, rate=rate, indel=indel, n_over_m=nm)
              for i, (L, rate, indel, nm) in enumerate(
                  [(L, rate, indel, 1.0) for L in (4096, 5000, 5024, 20000, 20480, 49984, 50000)
                   for rate, indel in ((0.02, False), (0.15, True), (0.30, True))] +
                  [(5000, 0.10, True, 4.0), (20000, 0.05, False, 2.5), (4992, 0.0, False, 1.0), (8192, 0.5, True, 1.0)])]


def main():
    binary = os.path.join(ROOT, "oracle", "_ref", "taxator")
    json.dump(CASES, open(os.path.join(HERE, "cases.json"), "w"), indent=1, sort_keys=True)
    json.dump({"heavy": list(HEAVY)}, open(os.path.join(HERE, "cases_meta.json"), "w"))
    only = set(sys.argv[1:])
    for name, case in CASES.items():
        if only and name not in only:
            continue
        data = synth.generate(cfg_of(case))
        lines = run_reference(data, binary, threads=os.cpu_count() or 1)
        with open(os.path.join(HERE, name + ".gff3"), "w") as f:
            f.writelines(lines)
        print(name, len(lines), "lines")
    # kernel-level vectors
    R = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libseqan_ref.so"))
    rng = np.random.default_rng(2024)

    def rnd(n, alpha):
        return "".join(alpha[i] for i in rng.integers(0, len(alpha), n))

    def mutate(s, rate, alpha):
        out = []
        for ch in s:
            r = rng.random()
            if r < rate / 3:
                continue
            if r < 2 * rate / 3:
                out.append(alpha[int(rng.integers(0, len(alpha)))])
            elif r < rate:
                out.append(ch)
                out.append(alpha[int(rng.integers(0, len(alpha)))])
            else:
                out.append(ch)
        return "".join(out)

    nt = []
    for L in [1, 2, 5, 31, 32, 33, 63, 64, 65, 96, 100, 127, 128, 129, 200, 255, 256, 257, 500, 777, 1024, 1500]:
        for rate in (0.0, 0.05, 0.3):
            for alpha in ("ACGT", "ACGTN"):
                a = rnd(L, alpha)
                b = mutate(a, rate, alpha) or "A"
                nt.append([a, b, R.ref_edit_distance(a.encode(), len(a), b.encode(), len(b))])
        a = rnd(L, "ACGT"); b = rnd(int(rng.integers(1, 2 * L + 2)), "ACGT")
        nt.append([a, b, R.ref_edit_distance(a.encode(), len(a), b.encode(), len(b))])
    nt.append(["acgtACGTnNRYKMuU--", "ACGTACGTNNNNNNTT", R.ref_edit_distance(b"acgtACGTnNRYKMuU--", 18, b"ACGTACGTNNNNNNTT", 16)])
    aa = []
    out = (ctypes.c_int * 6)()
    for L in [1, 2, 5, 30, 31, 32, 33, 64, 100, 150, 299, 300, 301, 400]:
        for rate in (0.0, 0.1, 0.4, 0.8):
            for alpha in ("ACDEFGHIKLMNPQRSTVWY", "ABCDEFGHIJKLMNOPQRSTUVWYZX*"):
                a = rnd(L, alpha)
                b = mutate(a, rate, alpha) or "A"
                R.ref_protein_align(a.encode(), len(a), b.encode(), len(b), out)
                aa.append([a, b, list(out)])
    if not only:
        json.dump({"edit_distance": nt, "protein": aa}, open(os.path.join(HERE, "seqan_pairs.json"), "w"))
    print("pairs:", len(nt), len(aa))
    # long pairs (4-50 kb, lengths incl. multiples of 32): seeds + the distance real SeqAn computes
    longs = []
    for spec in LONG_PAIRS:
        a, b = long_pair(**spec)
        dist = R.ref_edit_distance(a.tobytes(), len(a), b.tobytes(), len(b))
        longs.append(dict(spec, m=len(a), n=len(b), dist=int(dist)))
    json.dump(longs, open(os.path.join(HERE, "seqan_long_pairs.json"), "w"), indent=0)
    print("long pairs:", len(longs))


if __name__ == "__main__":
    main()
