#!/usr/bin/env python
"""Benchmark of the RPA hot path (BASELINE.json metric: query segments/s and GCUPS).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c5|tiny] [--impl ours|reference]

One "step" = one pass of the whole batched RPA path (decide / stage / align rounds until every
segment is placed) over one batch of synthetic query segments.  Default workload = BASELINE.json
configs[1]: 100k 5-kb nucleotide segments x ~50 candidate references on one B200.
 * value  : segments/s with the batch already resident in HBM (trpa_batch_run only)
 * e2e    : segments/s through the C-ABI call a host makes (trpa_predict_batch: host buffers in, host
            results out, H2D/D2H inside the timed region); the sequence stores are loaded once at
            start-up, exactly like the reference loads its stores before the first predict()
 * roofline: the edit-distance kernel against the measured INT32 ALU-pipe peak (integer DP is ALU
            bound), plus roofline_hbm for the segment staging kernel
 * cpu_baseline: the REAL reference (oracle/_ref/taxator, unmodified sources) on a bounded sample
With --impl reference only the reference binary runs (on the host cores).
Under torchrun every rank owns one GPU and its own shard of segments (weak scaling, no collective
on the data path); time = max over ranks.

Besides the headline (C2, weak scaling) the default run adds, in the same JSON line:
 * "c4": BASELINE.json configs[3], ONE 2M-segment mixed-length batch cut into world shards by estimated DP work,
         host tables in, result records gathered on rank 0 inside the timed region (strong scaling; at N=1 the
         one-eighth shard a GPU of an 8-GPU box owns)
 * "c5": configs[4] at full size, 200k 10-50 kb noisy reads, sharded the same way
 * "parity": GFF3 of this library == GFF3 of the unmodified reference binary on a sample of all five configs
 * "e2e_cli": wall clock of taxator-b200 and of the reference binary on the SAME files (N=1, rank 0)
(--extras none skips them).
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import synth  # noqa: E402
import gff3  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]
    "c2": dict(desc="nucleotide RPA, 100k 5-kb contig segments x ~50 candidate refs", protein=False,
               cfg=dict(n_genomes=400, genome_len=50000, n_queries=100000, query_len=(5000, 5000), n_cand=50,
                        levels=(4, 8, 16, 40, 100)), cpu_sample=600),
    # configs[0] (the reference's own CPU-runnable case)
    "c1": dict(desc="nucleotide RPA, 1k 1-kb contigs, 200-genome refpack", protein=False,
               cfg=dict(n_genomes=200, genome_len=20000, n_queries=1000, query_len=(1000, 1000), n_cand=36,
                        levels=(4, 8, 16, 40, 100)), cpu_sample=1000),
    # configs[2]
    "c3": dict(desc="protein RPA (BLOSUM62, linear gap), 50k 300-aa segments", protein=True,
               cfg=dict(n_genomes=400, genome_len=400, n_queries=50000, query_len=(300, 300), n_cand=30,
                        levels=(4, 8, 16, 40, 100)), cpu_sample=600),
    # configs[3]: the per-GPU shard (1/8) of the 2M-segment batch, mixed 0.5-20 kb, length-bucketed on the device
    "c4": dict(desc="nucleotide RPA, 250k segments per GPU (1/8 of 2M), mixed 0.5-20 kb, ~30 candidate refs", protein=False,
               cfg=dict(n_genomes=400, genome_len=100000, n_queries=250000, query_len=(500, 20000), n_cand=30,
                        levels=(4, 8, 16, 40, 100)), cpu_sample=100),
    # configs[4], scaled to one GPU-minute per step
    "c5": dict(desc="long-read nucleotide RPA, 10-50 kb noisy reads (15% error)", protein=False,
               cfg=dict(n_genomes=100, genome_len=200000, n_queries=4000, query_len=(10000, 50000), n_cand=20,
                        levels=(2, 4, 8, 16, 40), query_sub=0.05, query_indel=0.10), cpu_sample=16),
    "tiny": dict(desc="smoke-sized nucleotide RPA", protein=False,
                 cfg=dict(n_genomes=60, genome_len=6000, n_queries=400, query_len=(1000, 1000), n_cand=25,
                          levels=(2, 4, 6, 10, 20)), cpu_sample=200),
}

# ALU-pipe instructions per 32-cell word-step: the MINIMUM of the bit-vector formulation (7 LOP3 + 2 SHF +
# 1 IADD3.X, myers3_column); what the column loop really issues is 10.9 at W = 4 and 10.2 at W = 20
# (scripts/sass_loop.py).  Rounds r01-r02 quoted the roofline on the 10.4 counted for the W = 20 kernel of r01
# (profiles/r01_sass_mix.md); that figure is still reported as frac_r01_constant.
ALU_OPS_PER_WORDSTEP = 10.0
ALU_OPS_PER_WORDSTEP_R01 = 10.4


# stdout carries exactly ONE JSON line: everything else a library prints there (NCCL's version banner, ...) is
# sent to stderr by pointing fd 1 at fd 2 for the whole run; emit() writes to the saved real stdout.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(text):
    out = _REAL_STDOUT or sys.stdout
    out.write(text + "\n")
    out.flush()


def make_data(workload, seed, n_queries=None, rank=0):
    """Rank 0: synth.generate_fast(seed).  Other ranks (weak scaling): the SAME refpack and taxonomy (drawn from
    `seed`) with their own query block -- every rank then owns statistically the same work; with one seed per rank
    the refpacks differ (tree shape, divergence) and the slowest rank's luck shows up as lost scaling efficiency."""
    w = WORKLOADS[workload]
    cfg = dict(w["cfg"])
    if n_queries is not None:
        cfg["n_queries"] = n_queries
    c = synth.SynthConfig(seed=seed, protein=w["protein"], **cfg)
    if rank == 0:
        return synth.generate_fast(c)
    d, ref = synth.build_reference_fast(c, np.random.default_rng(c.seed))
    d.q_seqs, d.rec = synth.generate_queries_fast(c, ref, synth.block_rng(c, 1000003 + rank), c.n_queries)
    d.q_names = ["Q%06d" % i for i in range(c.n_queries)]
    return d


class Flat:
    def __init__(self, d):
        self.d = d
        self.protein = bool(d.cfg.protein)
        self.parent, self.left, self.right, self.depth = d.nested_set()
        self.q_chars, self.q_off, self.q_len = d.store_arrays(d.q_seqs)
        self.r_chars, self.r_off, self.r_len = d.store_arrays(d.ref_seqs)
        self.segs, self.cands = synth.segments_fast(d)


def subset(d, n_queries):
    """First n_queries queries (and their records) of a SynthData, same refpack/taxonomy."""
    s = synth.SynthData(cfg=d.cfg)
    s.tax_ids, s.tax_parent, s.tax_rank = d.tax_ids, d.tax_parent, d.tax_rank
    s.ref_names, s.ref_seqs, s.ref_taxnode = d.ref_names, d.ref_seqs, d.ref_taxnode
    s.q_names, s.q_seqs = d.q_names[:n_queries], d.q_seqs[:n_queries]
    m = d.rec["q"] < n_queries
    s.rec = {k: v[m] for k, v in d.rec.items()}
    return s


def run_reference_binary(d, threads, keep_dir=None, model_args=None):
    """Times oracle/_ref/taxator (real reference, -DNDEBUG) on the files of d; returns (seconds, sorted GFF3 lines)."""
    binary = os.path.join(ROOT, "oracle", "_ref", "taxator")
    if not os.path.exists(binary):
        return None, None
    tmp = keep_dir or tempfile.mkdtemp(prefix="trpa_ref_")
    try:
        d.write_files(tmp)
        env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
        cmd = [binary, "-a", "rpa", "-g", "mapping.tax", "-q", "query.fna", "-f", "ref.fna", "-i", "ref.fna.fai",
               "-p", str(threads), "-x", "0.5", "-o", "0"]
        if d.cfg.protein:
            cmd += ["-b", "protein"]
        if model_args:   # the alignment-free models read no sequence files
            cmd = [binary] + list(model_args) + ["-g", "mapping.tax", "-p", str(threads), "-o", "0"]
        with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin:
            t0 = time.perf_counter()
            p = subprocess.run(cmd, cwd=tmp, env=env, stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            dt = time.perf_counter() - t0
        if p.returncode != 0:
            raise RuntimeError("reference binary failed: " + p.stderr.decode()[-400:])
        lines = sorted(l + "\n" for l in p.stdout.decode().splitlines() if not l.startswith("##"))
        return dt, lines
    finally:
        if keep_dir is None:
            shutil.rmtree(tmp, ignore_errors=True)


class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation on the host cores (rank 0 only)."""
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    sample = w["cpu_sample"]
    d = make_data(args.workload, args.seed, n_queries=max(sample, 1))
    n_seg = len(synth.segments_fast(d)[0])
    tmp = tempfile.mkdtemp(prefix="trpa_refarm_")
    times = []
    try:
        for it in range(args.warmup + args.steps):
            dt, _ = run_reference_binary(d, cores, keep_dir=tmp)
            if dt is None:
                emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/taxator was not built"}))
                return
            if it >= args.warmup:
                times.append(dt)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    ms = 1e3 * sum(times) / len(times)
    value = n_seg / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "query segments/sec (taxator -a rpa hot path)", "value": value,
        "unit": "segments/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32" if not w["protein"] else "int32+f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "segments_per_step": n_seg,
                   "note": "bounded sample of the workload, wall time of the whole binary incl. start-up"},
        "cpu_baseline": {"value": value, "unit": "segments/s", "cores": cores, "kind": "reference",
                         "sample": "%d segments of the workload, oracle/_ref/taxator -DNDEBUG -p %d" % (n_seg, cores)},
        "e2e": {"value": value, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


def lca_bench(args, rank, local_rank, world):
    """the alignment-free models (megan-lca with the reference's defaults) on the record tables of
    C2 (100k segments x 50 records).  HBM-bound streaming kernel: 44 B read per record, 72 B per segment."""
    w = WORKLOADS["c2"]
    d = make_data("c2", args.seed, n_queries=args.segments, rank=rank)
    segs, cands = synth.segments_fast(d)
    parent, left, right, depth = d.nested_set()
    n_seg, n_cand = len(segs), len(cands)
    evalue = np.power(10.0, -np.round(cands["score"].astype(np.float64) / 12.0))
    kw = dict(toppercent=0.05, minscore=0.0, maxevalue=1000.0, minsupport=1)
    cores = os.cpu_count() or 1
    if args.impl == "reference":
        if rank != 0:
            return
        sub = subset(d, min(5000, len(d.q_names)))
        times = []
        for it in range(args.warmup + args.steps):
            dt, _ = run_reference_binary(sub, cores, model_args=["-a", "megan-lca"])
            if dt is None:
                emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/taxator was not built"}))
                return
            if it >= args.warmup:
                times.append(dt)
        nseg_sub = len(synth.segments_fast(sub)[0])
        v = nseg_sub / (sum(times) / len(times))
        emit(json.dumps({"impl": "reference", "metric": "query segments/sec (taxator -a megan-lca)", "value": v, "unit": "segments/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
                          "config": {"workload": "lca: megan-lca on the record tables of c2", "segments_per_step": nseg_sub},
                          "cpu_baseline": {"value": v, "unit": "segments/s", "cores": cores, "kind": "reference",
                                           "sample": "%d segments, oracle/_ref/taxator -a megan-lca -p %d, wall incl. start-up" % (nseg_sub, cores)},
                          "e2e": {"value": v, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    import torch
    import rpa_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def reduce(x, op):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    ctx = rpa_b200.Context(local_rank, torch.cuda.current_stream().cuda_stream)
    ctx.load_taxonomy(parent, left, right, depth, 0)
    # pinned host tables (the e2e contract: inputs come from pinned host memory every step)
    pins = [torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).copy()).pin_memory() for a in (segs, cands, evalue)]
    segs = pins[0].numpy().view(rpa_b200.SEG_DTYPE)
    cands = pins[1].numpy().view(rpa_b200.CAND_DTYPE)
    evalue = pins[2].numpy().view(np.float64)
    for _ in range(max(args.warmup, 3)):
        ctx.predict_lca_batch(rpa_b200.MODEL_MEGAN_LCA, segs, cands, evalue, None, **kw)
    sampler = ClockSampler(local_rank)
    sampler.start()
    if dist is not None:
        dist.barrier()
    # resident: the call uploads once, runs one untimed pass, then `steps` timed launches (CUDA events in the library)
    res, k_ms = ctx.predict_lca_batch(rpa_b200.MODEL_MEGAN_LCA, segs, cands, evalue, None, repeat=max(args.steps, 2), **kw)
    k_ms = reduce(k_ms, torch.distributed.ReduceOp.MAX if dist else None)
    # end to end: host tables in, host results out, every step
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.predict_lca_batch(rpa_b200.MODEL_MEGAN_LCA, segs, cands, evalue, None, **kw)
    torch.cuda.synchronize()
    ms_e2e = reduce(1e3 * (time.perf_counter() - t0) / args.steps, torch.distributed.ReduceOp.MAX if dist else None)
    clocks = sampler.finish()
    total = reduce(float(n_seg), torch.distributed.ReduceOp.SUM if dist else None)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    h2d = n_cand * (36 + 8) + n_seg * 16
    alg_bytes = h2d + n_seg * 56
    gbs = alg_bytes / (k_ms / 1e3) / 1e9
    line = {"metric": "query segments/sec (taxator -a megan-lca)", "value": total / (k_ms / 1e3), "unit": "segments/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": k_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
            "config": {"workload": "lca: megan-lca (-t 0.05 -m 0 -e 1000 -c 1) on the record tables of c2", "segments_per_gpu": n_seg,
                       "candidates_per_gpu": n_cand, "l2": "tables (%.0f MB) larger than L2" % (alg_bytes / 1e6), "seed": args.seed},
            "e2e": {"value": total / (ms_e2e / 1e3), "unit": "segments/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(n_seg * 56)},
            "roofline": {"bound": "hbm", "kernel": "lca_models_kernel", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": gbs / hbm_peak, "traffic": None,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_launch": alg_bytes},
            "gpu_launches": int(args.steps), "clocks": clocks}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sub = subset(d, min(5000, len(d.q_names)))
        dt, ref_lines = run_reference_binary(sub, cores, model_args=["-a", "megan-lca"])
        if dt is not None:
            nsub = len(synth.segments_fast(sub)[0])
            q_len = np.array([len(x) for x in d.q_seqs], np.uint32)
            ours = sorted(gff3.render(res[:nsub], segs[:nsub], d.q_names, q_len, parent, depth, [str(t) for t in d.tax_ids]))
            line["cpu_baseline"] = {"value": nsub / dt, "unit": "segments/s", "cores": cores, "kind": "reference",
                                    "sample": "first %d segments; oracle/_ref/taxator -a megan-lca -p %d, wall %.2f s incl. start-up and parsing" % (nsub, cores, dt),
                                    "gff3_identical_to_gpu": ours == ref_lines}
    if rank == 0:
        emit(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# Full-size sharded workloads (BASELINE.json configs[3] and [4]): one fixed batch, block-wise generated
# (synth.generate_blocks), cut into `world` contiguous shards by the estimated DP work n_cand * L^2
# (SURVEY.md 8e), host tables in, 56-byte result records gathered on rank 0 inside the timed region.
BIG = {
    # 2M segments, 64 blocks of 31250; at N=1 only the first eighth (what one GPU of an 8-GPU box owns)
    "c4": dict(workload="c4", total_queries=2000000, block_queries=31250, n1_blocks=8, warmup=1, steps=2),
    # 200k reads, 64 blocks of 3125; runs in full at every N
    "c5": dict(workload="c5", total_queries=200000, block_queries=3125, n1_blocks=64, warmup=1, steps=1),
}


class Shard:
    """This rank's part of a block-generated batch (+ the replicated refpack / taxonomy)."""
    pass


def make_shard(name, seed, rank, world, workers, scale=1.0):
    """CPU only (forks worker processes): call before CUDA is initialised."""
    spec = BIG[name]
    w = WORKLOADS[spec["workload"]]
    cfg = synth.SynthConfig(seed=seed, protein=w["protein"], **dict(w["cfg"], n_queries=0))
    t0 = time.time()
    d, ref = synth.build_reference_fast(cfg, np.random.default_rng(cfg.seed))
    bq = max(1, int(spec["block_queries"] * scale))
    n_blocks = spec["total_queries"] // spec["block_queries"] if world > 1 else spec["n1_blocks"]
    L = synth.block_query_lengths(cfg, ref, n_blocks, bq).astype(np.float64)
    cw = np.cumsum(ref.K * L * L)
    bounds = [0] + [int(np.searchsorted(cw, cw[-1] * r / world, side="left")) + 1 for r in range(1, world)] + [len(L)]
    lo, hi = bounds[rank], bounds[rank + 1]
    b_lo, b_hi = lo // bq, (hi + bq - 1) // bq
    chars, qlen, segs, cands = synth.generate_blocks(cfg, d, ref, range(b_lo, b_hi), bq, workers=workers)
    # slice the generated blocks down to [lo, hi): one segment per query in these configs
    assert len(segs) == (b_hi - b_lo) * bq and (segs["query_seq"] == np.arange(len(segs))).all()
    s0, s1 = lo - b_lo * bq, hi - b_lo * bq
    qoff = np.concatenate([[0], np.cumsum(qlen.astype(np.uint64))])
    sh = Shard()
    sh.name, sh.cfg, sh.d, sh.protein = name, cfg, d, bool(w["protein"])
    sh.total_segments, sh.bounds, sh.lo, sh.hi = len(L), bounds, lo, hi
    c0 = int(segs["cand_begin"][s0]) if s1 > s0 else 0
    c1 = int(segs["cand_begin"][s1 - 1] + segs["cand_count"][s1 - 1]) if s1 > s0 else 0
    sh.segs = segs[s0:s1].copy()
    sh.segs["cand_begin"] -= c0
    sh.segs["query_seq"] -= s0
    sh.cands = np.ascontiguousarray(cands[c0:c1])
    sh.q_len = np.ascontiguousarray(qlen[s0:s1])
    sh.q_chars = chars[int(qoff[s0]):int(qoff[s1])]
    sh.q_off = (qoff[s0:s1] - qoff[s0]).astype(np.uint64)
    sh.parent, sh.left, sh.right, sh.depth = d.nested_set()
    sh.r_chars, sh.r_off, sh.r_len = d.store_arrays(d.ref_seqs)
    sh.est_work = float(cw[hi - 1] - (cw[lo - 1] if lo else 0.0)) if hi > lo else 0.0
    sh.t_gen = time.time() - t0
    return sh


class _DevBytes:
    """Device memory of the library as a CUDA array (torch.as_tensor view, no copy)."""
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def run_shard(sh, local_rank, rank, world, dist, spec, tune):
    """Timed end to end at N ranks: pinned host tables -> trpa_batch_upload -> trpa_batch_run -> result records of
    all ranks gathered GPU to GPU (NCCL send/recv over NVLink) into rank 0's buffer -> one D2H copy on rank 0.
    No collective touches the DP; the gather moves 56 B per segment."""
    import torch
    import rpa_b200
    rs = rpa_b200.RESULT_DTYPE.itemsize
    ctx = rpa_b200.Context(local_rank, torch.cuda.current_stream().cuda_stream)
    try:
        t0 = time.time()
        ctx.load_taxonomy(sh.parent, sh.left, sh.right, sh.depth, 0)
        ctx.load_store(0, 0, sh.q_chars, sh.q_off, sh.q_len)
        ctx.load_store(1, 0, sh.r_chars, sh.r_off, sh.r_len)
        torch.cuda.synchronize()
        t_load = time.time() - t0
        for k, v in tune:
            ctx.set_tuning(k, v)
        n_seg = len(sh.segs)
        segs_t = torch.from_numpy(sh.segs.view(np.uint8).copy()).pin_memory()
        cands_t = torch.from_numpy(sh.cands.view(np.uint8).copy()).pin_memory()
        segs_p, cands_p = segs_t.numpy().view(rpa_b200.SEG_DTYPE), cands_t.numpy().view(rpa_b200.CAND_DTYPE)
        total = sh.total_segments
        out_t = torch.zeros((total if rank == 0 else 1) * rs, dtype=torch.uint8).pin_memory()
        big = torch.empty(total * rs, dtype=torch.uint8, device="cuda") if (rank == 0 and world > 1) else None

        def sync_all():
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        times, compute = [], []
        for it in range(spec["warmup"] + spec["steps"]):
            if it == spec["warmup"]:
                ctx.profile_reset()
            sync_all()
            t0 = time.perf_counter()
            if world == 1:
                res = ctx.predict_batch_into(segs_p, cands_p, out_t.numpy().view(rpa_b200.RESULT_DTYPE))
                t_c = time.perf_counter() - t0
            else:
                ctx.batch_upload(segs_p, cands_p)
                ctx.batch_run()
                torch.cuda.synchronize()
                t_c = time.perf_counter() - t0
                ptr, n = ctx.batch_results_dev()
                mine = torch.as_tensor(_DevBytes(ptr, n * rs), device="cuda") if n else torch.empty(0, dtype=torch.uint8, device="cuda")
                if rank == 0:
                    ops = [dist.P2POp(dist.irecv, big[sh.bounds[r] * rs:sh.bounds[r + 1] * rs], r)
                           for r in range(1, world) if sh.bounds[r + 1] > sh.bounds[r]]
                    big[:n * rs].copy_(mine)
                    for req in (dist.batch_isend_irecv(ops) if ops else []):
                        req.wait()
                    out_t.copy_(big, non_blocking=True)
                    torch.cuda.synchronize()
                    res = out_t.numpy().view(rpa_b200.RESULT_DTYPE)
                elif n:
                    for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine, 0)]):
                        req.wait()
                    torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it >= spec["warmup"]:
                times.append(dt); compute.append(t_c)
        prof = ctx.profile()
        return dict(res=res if rank == 0 else None, ms=1e3 * sum(times) / len(times), ms_compute=1e3 * sum(compute) / len(compute),
                    prof=prof, t_load=t_load, h2d=int(segs_t.numel() + cands_t.numel()), n_seg=n_seg, steps=len(times))
    finally:
        ctx.close()


def big_workload(name, sh, args, rank, local_rank, world, dist, alu_peak):
    """Runs one BIG workload over all ranks and returns its JSON object (rank 0) or None."""
    import torch
    spec = BIG[name]
    tune = [(kv.split("=")[0], int(kv.split("=")[1])) for kv in args.tune]
    r = run_shard(sh, local_rank, rank, world, dist, spec, tune)

    def allreduce(x, op):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def gather_floats(x):
        if dist is None:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    MAX = torch.distributed.ReduceOp.MAX if dist is not None else None
    ms = allreduce(r["ms"], MAX)
    per_rank_compute = gather_floats(r["ms_compute"])
    per_rank_kernel = gather_floats(r["prof"]["ms_edit_distance"] / max(1, r["steps"]))
    per_rank_exec = gather_floats(float(r["prof"]["cells_edit_distance"]) / max(1, r["steps"]))
    per_rank_segs = gather_floats(float(r["n_seg"]))
    gen_s = allreduce(sh.t_gen, MAX)
    if rank != 0:
        return None
    res = r["res"]
    total = sh.total_segments
    cells = float(res["cells"].sum())
    pairs = float((res["n_pass0"].astype(np.int64) + res["n_pass1"] + res["n_pass2"]).sum())
    kernel_ms = max(per_rank_kernel)
    peak = alu_peak * 32.0 / ALU_OPS_PER_WORDSTEP / 1e9
    exec_rate = [e / (k / 1e3) / 1e9 if k > 0 else 0.0 for e, k in zip(per_rank_exec, per_rank_kernel)]
    w = WORKLOADS[spec["workload"]]
    full = spec["total_queries"] // spec["block_queries"] * spec["block_queries"]
    obj = {
        "workload": name + ": " + w["desc"],
        "segments": total, "of_full_batch": total / float(full),
        "scaling": "strong" if world > 1 or total == full else "one-eighth shard (what one GPU of an 8-GPU box owns)",
        "n_gpus": world, "steps": r["steps"], "warmup": spec["warmup"],
        "value": total / (ms / 1e3), "unit": "segments/s", "ms_per_step": ms,
        "gcups": cells / (ms / 1e3) / 1e9, "cells_per_step": cells, "alignments_per_step": pairs,
        "timed_region": "pinned host tables -> H2D -> all rounds -> result records gathered on rank 0 (NCCL send/recv of "
                        "56 B per segment, no collective on the DP) -> D2H on rank 0; max over ranks",
        "h2d_bytes_per_step_rank0": r["h2d"], "d2h_bytes_per_step": int(total * 56),
        "segments_per_rank": [int(x) for x in per_rank_segs],
        "compute_ms_per_rank": [round(x, 1) for x in per_rank_compute],
        "imbalance_max_over_mean": max(per_rank_compute) / (sum(per_rank_compute) / len(per_rank_compute)),
        "gather_ms": ms - max(per_rank_compute),
        "roofline_frac_per_rank": [round(x / peak, 4) for x in exec_rate],
        "kernel_share_of_step": kernel_ms / ms,
        "band_retries_per_step": r["prof"]["band_retries"] / max(1, r["steps"]),
        "wedge_failures_per_step": r["prof"]["wedge_failures"] / max(1, r["steps"]),
        "generate_s": gen_s, "load_stores_s": r["t_load"],
        "results_in_input_order": bool((res["kind"] == 3).mean() > 0.9 and len(res) == total),
    }
    return obj


# ---------------------------------------------------------------------------------------------------------
# Parity against the REAL reference binary on a sample of every BASELINE.json config (N=1, rank 0)
PARITY_SAMPLES = {"c1": 1000, "c2": 1000, "c3": 1000, "c4": 300, "c5": 40}


def parity_all(args, local_rank, skip_c2_data=None):
    import rpa_b200
    import torch
    cores = os.cpu_count() or 1
    out = {}
    for wl, nq in PARITY_SAMPLES.items():
        w = WORKLOADS[wl]
        d = make_data(wl, args.seed + 7, n_queries=nq)
        fd = Flat(d)
        ctx = rpa_b200.Context(local_rank, torch.cuda.current_stream().cuda_stream)
        try:
            ctx.load_taxonomy(fd.parent, fd.left, fd.right, fd.depth, 0)
            alpha = 1 if fd.protein else 0
            ctx.load_store(0, alpha, fd.q_chars, fd.q_off, fd.q_len)
            ctx.load_store(1, alpha, fd.r_chars, fd.r_off, fd.r_len)
            res = ctx.predict_batch(fd.segs, fd.cands)
        finally:
            ctx.close()
        dt, ref_lines = run_reference_binary(d, cores)
        if dt is None:
            out[wl] = {"gff3_identical": None, "note": "oracle/_ref/taxator missing"}
            continue
        taxids = [str(t) for t in d.tax_ids]
        ours = sorted(gff3.render(res, fd.segs, d.q_names, fd.q_len, fd.parent, fd.depth, taxids))
        cells = float(res["cells"].sum())
        out[wl] = {"gff3_identical": ours == ref_lines, "segments": len(fd.segs),
                   "alignments": int((res["n_pass0"].astype(np.int64) + res["n_pass1"] + res["n_pass2"]).sum()),
                   "reference_wall_s": round(dt, 2), "reference_segments_per_s": len(fd.segs) / dt,
                   "reference_gcups": cells / dt / 1e9, "cores": cores}
    out["all_identical"] = all(v.get("gff3_identical") is True for v in out.values())
    out["how"] = ("per config: seeded sample -> this library (trpa_predict_batch) rendered as GFF3 vs oracle/_ref/taxator "
                  "(unmodified reference sources, -DNDEBUG, -p all cores) on the same files, both sorted")
    return out


def e2e_cli(args, d_c2, n_queries=5000):
    """Same files, wall clock, start-up included on both sides: taxator-b200 (files -> GFF3) vs the reference binary."""
    exe = os.path.join(ROOT, "taxator-tk_b200", "bin", "taxator-b200")
    sub = subset(d_c2, min(n_queries, len(d_c2.q_names)))
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix="trpa_cli_")
    try:
        dt_ref, ref_lines = run_reference_binary(sub, cores, keep_dir=tmp)
        if dt_ref is None or not os.path.exists(exe):
            return {"note": "reference binary or taxator-b200 missing"}
        env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=tmp)
        cmd = [exe, "-a", "rpa", "-g", "mapping.tax", "-q", "query.fna", "-f", "ref.fna", "-i", "ref.fna.fai", "-x", "0.5",
               "-o", "0", "--timing"]
        walls, timing, lines = [], "", None
        for _ in range(2):
            with open(os.path.join(tmp, "alignments.tsv"), "rb") as fin:
                t0 = time.perf_counter()
                p = subprocess.run(cmd, cwd=tmp, env=env, stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                walls.append(time.perf_counter() - t0)
            if p.returncode != 0:
                return {"note": "taxator-b200 failed: " + p.stderr.decode()[-300:]}
            timing = p.stderr.decode()
            lines = sorted(l + "\n" for l in p.stdout.decode().splitlines() if not l.startswith("##"))
        nseg = len(ref_lines)
        predict_s = load_s = None
        for ln in timing.splitlines():
            if ln.startswith("taxator-b200: load ") and "predict" in ln:
                f = ln.replace(",", "").split()
                load_s, predict_s = float(f[2]), float(f[5])
        return {"segments": nseg, "same_files": True, "gff3_identical": lines == ref_lines,
                "reference_wall_s": dt_ref, "reference_segments_per_s": nseg / dt_ref, "reference_cores": cores,
                "ours_wall_s": min(walls), "ours_wall_s_first_run": walls[0], "ours_segments_per_s": nseg / min(walls),
                "ours_load_s": load_s, "ours_predict_s": predict_s,
                "wall_ratio": dt_ref / min(walls),
                "note": "both sides: one process, files in, GFF3 out, start-up (file parsing, FASTA/.fai loading, CUDA context "
                        "creation on our side) inside the wall clock; ours_load_s / ours_predict_s split our wall"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["lca"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--segments", type=int, default=None, help="override segments per GPU (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lookahead", type=int, default=-1, help="look-ahead budget (-1: automatic)")
    ap.add_argument("--tune", action="append", default=[], help="key=value tuning hook (trpa_set_tuning)")
    ap.add_argument("--band", type=int, default=1, help="1: exact Ukkonen band (default), 0: full DP matrices (A/B)")
    ap.add_argument("--extras", default="auto", help="auto (all of c4,c5,parity,cli with the default workload), none, or a comma list")
    ap.add_argument("--big-scale", type=float, default=1.0, help="shrink the c4 / c5 batches (debug)")
    args = ap.parse_args()
    capture_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.workload == "lca":
        lca_bench(args, rank, local_rank, world)
        return
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    extras = set()
    if args.extras == "auto":
        extras = {"c4", "c5", "parity", "cli"} if (args.workload == "c2" and args.segments is None) else set()
    elif args.extras != "none":
        extras = set(x for x in args.extras.split(",") if x)
    w = WORKLOADS[args.workload]

    # ---- synthetic data (CPU, before CUDA is initialised: the block generator forks worker processes).
    # Headline workload: every rank owns its own shard (weak scaling).
    t0 = time.time()
    d = make_data(args.workload, args.seed, n_queries=args.segments, rank=rank)
    fd = Flat(d)
    n_seg, n_cand = len(fd.segs), len(fd.cands)
    t_gen = time.time() - t0
    workers = max(1, min(16, (os.cpu_count() or 1) // max(1, world)))
    shards = {name: make_shard(name, args.seed, rank, world, workers, scale=args.big_scale)
              for name in ("c4", "c5") if name in extras}

    import torch
    import rpa_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG=VERSION (set on some boxes) prints a banner on stdout; stdout carries the one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    stream = torch.cuda.current_stream().cuda_stream
    ctx = rpa_b200.Context(local_rank, stream)
    t0 = time.time()
    ctx.load_taxonomy(fd.parent, fd.left, fd.right, fd.depth, 0)
    alpha = 1 if fd.protein else 0
    ctx.load_store(0, alpha, fd.q_chars, fd.q_off, fd.q_len)
    ctx.load_store(1, alpha, fd.r_chars, fd.r_off, fd.r_len)
    torch.cuda.synchronize()
    t_load = time.time() - t0
    alu_peak = ctx.int_alu_peak()
    ctx.set_band(args.band)
    ctx.set_lookahead(args.lookahead)
    for kv in args.tune:
        k, v = kv.split("=")
        ctx.set_tuning(k, int(v))

    # pinned host buffers for the e2e path
    segs_t = torch.from_numpy(fd.segs.view(np.uint8).copy()).pin_memory()
    cands_t = torch.from_numpy(fd.cands.view(np.uint8).copy()).pin_memory()
    out_t = torch.zeros(n_seg * rpa_b200.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    segs_p = segs_t.numpy().view(rpa_b200.SEG_DTYPE)
    cands_p = cands_t.numpy().view(rpa_b200.CAND_DTYPE)
    out_p = out_t.numpy().view(rpa_b200.RESULT_DTYPE)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident timing
    ctx.batch_upload(segs_p, cands_p)
    for _ in range(args.warmup):
        ctx.batch_run()
    # one sampler per job (rank 0's GPU): nvidia-smi polling takes driver locks that every rank's launches wait on
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ctx.profile_reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ctx.batch_run()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    prof = ctx.profile()
    res = ctx.batch_download(out_p).copy()
    cells_step = float(res["cells"].sum())
    pairs_step = float((res["n_pass0"] + res["n_pass1"] + res["n_pass2"]).sum())

    # ---- end-to-end timing through the host-buffer C-ABI call
    import ctypes
    vp = ctypes.c_void_p

    def predict_host():
        return ctx.L.trpa_predict_batch(ctx.h, vp(segs_p.ctypes.data), ctypes.c_uint32(n_seg), vp(cands_p.ctypes.data),
                                        ctypes.c_uint32(n_cand), vp(out_p.ctypes.data))

    for _ in range(min(args.warmup, 1)):
        predict_host()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        rc = predict_host()
        assert rc == 0, ctx.L.trpa_last_error()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)))
    clocks = sampler.finish() if sampler else None

    total_segs = sum_over_ranks(float(n_seg))
    total_cells = sum_over_ranks(cells_step)
    ms_step = ms_total / args.steps
    value = total_segs / (ms_step / 1e3)
    e2e_value = total_segs / (ms_e2e / args.steps / 1e3)

    # ---- roofline of the dominant kernel
    if fd.protein:
        k_ms, k_launch = prof["ms_protein"], prof["launches_protein"]
        # protein3_kernel issues 3 alu-pipe instructions per DP cell in two of three columns (2 VIADDMNMX + 1 LOP3; the
        # multiply-add of the diagonal runs on the fma pipe, the profile lookup on the LSU) and 2 in every third column
        # (both additions as IMAD on the fma pipe + VIMNMX3 + LOP3): 8/3 per cell -- the mix at which the alu pipe and the
        # issue slots (16/3 per cell) bound the kernel at the same rate.  Round 1 quoted 3 per cell: frac_r01_constant.
        ALU_OPS_PER_CELL_AA = 8.0 / 3.0
        peak = alu_peak / ALU_OPS_PER_CELL_AA / 1e9
        kname = "protein3_kernel"
    else:
        k_ms, k_launch = prof["ms_edit_distance"], prof["launches_edit_distance"]
        peak = alu_peak * 32.0 / ALU_OPS_PER_WORDSTEP / 1e9
        kname = "myers3_kernel"
    algorithmic = cells_step * args.steps / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
    # The edit-distance kernel only EXECUTES the cells inside an exact Ukkonen band (same integers as
    # the full matrix); the roofline fraction is quoted on the executed cells (what the ALU pipe did),
    # the algorithmic rate (cells of the reference's full matrices / time) is given beside it.
    executed_cells = float(prof["cells_edit_distance"]) if not fd.protein else cells_step * args.steps
    achieved = executed_cells / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
    try:
        ncu_traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        ncu_traffic = {}
    tr = ncu_traffic.get(kname, {})
    roofline = {"bound": "int32_alu", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GCUPS",
                "frac": achieved / peak if peak else None, "traffic": tr.get("bytes_per_launch"),
                "traffic_source": tr.get("source"),
                "achieved_algorithmic": algorithmic,
                "executed_cell_fraction": executed_cells / (cells_step * args.steps) if cells_step else None,
                "band": bool(args.band) and not fd.protein, "band_retries_per_step": prof["band_retries"] / args.steps,
                "peak_source": "own probe trpa_int_alu_peak (%.3e lane-ops/s) / %.1f ALU ops per 32-cell word-step"
                               % (alu_peak, ALU_OPS_PER_WORDSTEP) if not fd.protein else "own probe trpa_int_alu_peak / (8/3) alu ops per cell",
                "peak_lane_ops": alu_peak, "alu_ops_per_unit": ALU_OPS_PER_CELL_AA if fd.protein else ALU_OPS_PER_WORDSTEP,
                "executed_cells_note": "executed cells = 32x32-cell word-blocks the kernel walked: a pair's partial last text block "
                                       "counts as 32 columns, and blocks of re-run attempts (band_retries) are included",
                "kernel_ms_per_step": k_ms / args.steps, "kernel_share_of_step": (k_ms / args.steps) / ms_step}
    if fd.protein and peak:
        roofline["frac_r01_constant"] = roofline["frac"] * 3.0 / ALU_OPS_PER_CELL_AA   # against lane-ops/s / 3 (round 1)
    if not fd.protein and peak:
        roofline["frac_r01_constant"] = roofline["frac"] * ALU_OPS_PER_WORDSTEP_R01 / ALU_OPS_PER_WORDSTEP
        # Round 2 narrows the wedge further (fewer executed cells for the same integers), which LOWERS `frac` (per-step
        # overheads are paid on fewer cells) while the step gets faster.  For a like-for-like reading against round 1:
        # the cells round 1's band executed on this workload (0.1333 of the algorithmic cells on C2) over
        # this build's kernel time.
        if args.workload == "c2" and cells_step:
            r1_cells = 0.1333 * cells_step * args.steps
            roofline["frac_at_round1_band"] = r1_cells / (k_ms / 1e3) / 1e9 / peak
            roofline["frac_note"] = ("frac = executed cells / kernel time / peak; this build executes %.4f of the algorithmic cells "
                                     "(round 1: 0.1333), frac_at_round1_band counts round 1's executed cells over this build's time"
                                     % (executed_cells / (cells_step * args.steps)))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline_hbm = None
    if prof["ms_stage"] > 0:
        # algorithmic bytes of staging: packed store bits read + staged bits written
        # (nt: 3 planes x 4 B per 32-base word read, those + 8 B of column codes written; aa: 5 bit read +
        # 1 byte written per residue)
        gbs = prof["bytes_stage"] / (prof["ms_stage"] / 1e3) / 1e9
        roofline_hbm = {"bound": "hbm", "kernel": "stage_kernel", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                        "frac": gbs / hbm_peak, "traffic": ncu_traffic.get("stage_kernel", {}).get("bytes_per_launch") if not fd.protein else None,
                        "traffic_source": ncu_traffic.get("stage_kernel", {}).get("source") if not fd.protein else None,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                        "bytes_per_step": prof["bytes_stage"] / args.steps,
                        "ms_per_step": prof["ms_stage"] / args.steps}

    line = {
        "metric": "query segments/sec (taxator -a rpa hot path)", "value": value, "unit": "segments/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32" if not fd.protein else "int32+f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "segments_per_gpu": n_seg, "candidates_per_gpu": n_cand,
                   "alignments_per_step": pairs_step, "cells_per_step": cells_step,
                   "l2": "inputs larger than L2 (staged segments + refpack + candidate tables per step >> 126 MB)",
                   "seed": args.seed},
        "gcups": total_cells / (ms_step / 1e3) / 1e9,
        "e2e": {"value": e2e_value, "unit": "segments/s", "ms_per_step": ms_e2e / args.steps,
                "gcups": total_cells / (ms_e2e / args.steps / 1e3) / 1e9,
                "h2d_bytes_per_step": int(segs_t.numel() + cands_t.numel()), "d2h_bytes_per_step": int(out_t.numel())},
        "roofline": roofline,
        "gpu_launches": int(prof["launches_edit_distance"] + prof["launches_protein"] + prof["launches_stage"] +
                            prof["launches_decide"] + prof["launches_other"]),
        "rounds_per_step": prof["rounds"] / args.steps,
        "pairs_launched_per_step": prof["pairs"] / args.steps,
        "phase_ms_per_step": {"align": k_ms / args.steps, "stage": prof["ms_stage"] / args.steps,
                              "decide": prof["ms_decide"] / args.steps, "bucket": prof["ms_other"] / args.steps},
        "clocks": clocks,
        "setup_s": {"generate": t_gen, "load_stores": t_load},
    }
    if roofline_hbm:
        line["roofline_hbm"] = roofline_hbm

    # ---- CPU baseline: the real reference on a bounded sample (rank 0, N=1 only)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        nq = min(w["cpu_sample"], len(d.q_seqs))
        sub = subset(d, nq)
        dt, ref_lines = run_reference_binary(sub, cores)
        if dt is not None:
            nseg_sub = int((fd.segs["query_seq"] < nq).sum())
            sel = fd.segs["query_seq"] < nq
            taxids = [str(t) for t in d.tax_ids]
            ours = sorted(gff3.render(res[sel], fd.segs[sel], d.q_names, fd.q_len, fd.parent, fd.depth, taxids))
            cells_sub = float(res["cells"][sel].sum())
            line["cpu_baseline"] = {"value": nseg_sub / dt, "unit": "segments/s", "cores": cores, "kind": "reference",
                                    "gcups": cells_sub / dt / 1e9,
                                    "sample": "first %d segments of the workload; oracle/_ref/taxator (unmodified "
                                              "reference, -DNDEBUG) -p %d, wall %.2f s incl. start-up" % (nseg_sub, cores, dt),
                                    "gff3_identical_to_gpu": ours == ref_lines}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "segments/s", "cores": cores, "kind": "reference",
                                    "sample": "oracle/_ref/taxator missing"}
    ctx.close()
    del segs_t, cands_t, out_t
    torch.cuda.empty_cache()

    # ---- the configs north_star names beyond the headline (see the module docstring)
    for name in ("c4", "c5"):
        if name in shards:
            try:
                obj = big_workload(name, shards.pop(name), args, rank, local_rank, world, dist, alu_peak)
            except Exception as e:   # the headline must survive a failure here; say so in the line
                obj = {"error": repr(e)[:300]}
            if rank == 0:
                line[name] = obj
    if rank == 0 and world == 1:
        if "parity" in extras:
            line["parity"] = parity_all(args, local_rank)
            if "cpu_baseline" in line and "c2" in line["parity"]:
                line["parity"]["c2_cpu_baseline_sample"] = line["cpu_baseline"].get("gff3_identical_to_gpu")
        if "cli" in extras:
            line["e2e_cli"] = e2e_cli(args, d)
    if rank == 0:
        emit(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
