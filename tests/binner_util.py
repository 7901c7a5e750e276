"""Test helpers for the binner: GFF3 prediction lines -> the flat tables of trpa_bin_batch /
orc_binner (a Python mirror of the host parser, test infrastructure), the oracle call, rendering of the
Bioboxes body lines, and the input files the reference binner reads."""
import ctypes
import os
import re

import numpy as np

import oracle_lib as ol

BIN_RECORD = np.dtype([("lower_node", "<u4"), ("upper_node", "<u4"), ("support_begin", "<u4"), ("query_length", "<u4"),
                       ("query_id", "<u4"), ("reserved", "<u4")])
BIN_RESULT = np.dtype([("node", "<u4"), ("support", "<u4"), ("length", "<u4"), ("mode", "<u4"), ("lower_node", "<u4"),
                       ("upper_node", "<u4"), ("lower_support", "<u4"), ("upper_support", "<u4")])


class BinParams(ctypes.Structure):
    _fields_ = [("signal_majority", ctypes.c_float), ("min_support_per_sequence", ctypes.c_uint32),
                ("min_support_in_sample", ctypes.c_uint32), ("min_support_in_sample_fraction", ctypes.c_float),
                ("n_ranks", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class BinStats(ctypes.Structure):
    _fields_ = [("nested_taxa", ctypes.c_uint64), ("root_support", ctypes.c_uint64), ("pruned_taxa", ctypes.c_uint64),
                ("min_support_found", ctypes.c_uint64)]


# option sets of the golden runs: (command-line arguments of the reference binner, keyword arguments here)
VARIANTS = {
    "default": ([], {}),
    "glob10": (["-g", r"(Q000\d\d)\d"], dict(glob=r"(Q000\d\d)\d")),
    "globall": (["-g", r"(Q).*"], dict(glob=r"(Q).*")),
    "noglob": (["-g", ""], dict(glob="")),
    "strict": (["-g", r"(Q000\d\d)\d", "-s", "500", "-j", "0.9"], dict(glob=r"(Q000\d\d)\d", min_support=500, majority=0.9)),
    "noise_count": (["-g", r"(Q000\d\d)\d", "-m", "3000"], dict(glob=r"(Q000\d\d)\d", sample_min="3000")),
    "noise_frac": (["-m", "0.02"], dict(sample_min="0.02")),
    "pid": (["-g", r"(Q000\d\d)\d", "-i", "species:0.8", "-i", "genus:0.5", "-i", "family:0.2"],
            dict(glob=r"(Q000\d\d)\d", pid={"species": 0.8, "genus": 0.5, "family": 0.2})),
}
CASES = ("nt_small", "nt_1kb", "nt_5kb")


def deep_taxonomy_files(data, outdir):
    """The reference sizes a per-depth table by the depth of the UNPRUNED taxonomy (binner.cpp:219,
    taxontree.hh:222) and indexes it by pruned depth: with a taxonomy made only of ranked nodes it overruns that
    table by one.  Real NCBI dumps are far deeper than the 7 ranks kept; the synthetic dump gets a chain of
    unranked nodes under the root (deleted by the pruning) to be like that."""
    data.write_files(outdir)
    with open(os.path.join(outdir, "nodes.dmp"), "a") as f:
        parent = 1
        for k in range(12):
            f.write("%d\t|\t%d\t|\tno rank\t|\t\t|\n" % (900001 + k, parent))
            parent = 900001 + k
    with open(os.path.join(outdir, "names.dmp"), "a") as f:
        for k in range(12):
            f.write("%d\t|\tdeep%d\t|\t\t|\tscientific name\t|\n" % (900001 + k, k))


def flat_tables(lines, data, glob="(.+)"):
    """(records, supports, group_begin, group names) from GFF3 lines; groups in first-appearance order."""
    parent, left, right, depth = data.nested_set()
    node_of = {str(t): i for i, t in enumerate(data.tax_ids)}
    groups, order = {}, []
    qids = {}
    for line in lines:
        line = line.rstrip("\n")
        if not line or line.startswith("#"):
            continue
        f = line.split("\t")
        begin, end = int(f[3]), int(f[4])
        kv = dict(x.split("=", 1) for x in f[8].split(";") if x)
        toks = kv["tax"].split("-")
        first = toks[0].split(":")
        support = int(first[1]) if len(first) > 1 and first[1] else end - begin + 1
        last = node_of[first[0]]
        lower = last
        rev = []
        for t in toks[1:]:
            ts = t.split(":")
            node = node_of[ts[0]]
            assert right[node] > left[last] and left[node] < left[last]
            x = last
            while x != node:
                rev.append(support)
                x = int(parent[x])
            if len(ts) > 1 and ts[1]:
                support = int(ts[1])
            last = node
        rev.append(support)
        name = "consensus_sequence" if glob == "" else re.fullmatch(glob, f[0]).group(1)
        if name not in groups:
            groups[name] = []
            order.append(name)
        qid = qids.setdefault(f[0], len(qids))
        groups[name].append((lower, last, rev[::-1], int(kv["seqlen"]), qid))
    recs, sup, gb = [], [], [0]
    for name in order:
        for lower, upper, s, qlen, qid in groups[name]:
            recs.append((lower, upper, len(sup), qlen, qid, 0))
            sup.extend(s)
        gb.append(len(recs))
    return (np.array(recs, BIN_RECORD) if recs else np.zeros(0, BIN_RECORD), np.array(sup, np.uint32),
            np.array(gb, np.uint32), order)


def params_of(data, majority=0.7, min_support=50, sample_min="0", pid=None, **_):
    pp = BinParams(float(majority), int(min_support), 0, 0.0, 0, 0)
    if "." in sample_min:
        pp.min_support_in_sample_fraction = float(sample_min)
    else:
        pp.min_support_in_sample = int(sample_min)
    rank_of_node = pid_per_rank = None
    if pid:
        names = sorted(set(data.tax_rank))
        rank_of_node = np.array([names.index(r) for r in data.tax_rank], np.uint8)
        pid_per_rank = np.array([pid.get(r, -1.0) for r in names] + [-1.0], np.float32)
        pp.n_ranks = len(pid_per_rank)
    return pp, rank_of_node, pid_per_rank


_hb = None


def host_binner():
    """tests/host_binner_harness.cpp: the product's kernel bodies (csrc/binner_core.h) compiled for the host."""
    global _hb
    if _hb is None:
        import subprocess
        os.makedirs(ol.BUILD_DIR, exist_ok=True)
        so = os.path.join(ol.BUILD_DIR, "libhost_binner.so")
        src = os.path.join(ol.ROOT, "tests", "host_binner_harness.cpp")
        csrc = os.path.join(ol.ROOT, "taxator-tk_b200", "csrc")
        deps = [src, os.path.join(csrc, "binner_core.h"), os.path.join(csrc, "machine.h"), os.path.join(ol.ROOT, "include", "taxator_rpa_b200.h")]
        if not ol._newer(so, *deps):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src])
        _hb = ctypes.CDLL(so)
    return _hb


def oracle_bin(data, recs, sup, gb, pp, rank_of_node, pid_per_rank, lib=None):
    O = lib or ol.oracle()
    if lib is not None:
        O.orc_binner = lib.hb_binner
    parent, left, right, depth = data.nested_set()
    parent = np.ascontiguousarray(parent, np.uint32)
    out = np.zeros(len(gb) - 1, BIN_RESULT)
    st = BinStats()
    vp = ctypes.c_void_p
    rc = O.orc_binner(vp(parent.ctypes.data), vp(depth.ctypes.data), ctypes.c_uint32(len(parent)), ctypes.c_uint32(0),
                      ctypes.byref(pp), vp(recs.ctypes.data), ctypes.c_uint32(len(recs)), vp(sup.ctypes.data), vp(gb.ctypes.data),
                      ctypes.c_uint32(len(gb) - 1), vp(rank_of_node.ctypes.data) if rank_of_node is not None else None,
                      vp(pid_per_rank.ctypes.data) if pid_per_rank is not None else None, vp(out.ctypes.data), ctypes.byref(st))
    assert rc == 0
    return out, st


def body_lines(names, res, data):
    return ["%s\t%d\t%d\t%d\n" % (n, data.tax_ids[int(r["node"])], int(r["support"]), int(r["length"]))
            for n, r in zip(names, res) if r["mode"] != 0]


def golden_path(case, variant):
    return os.path.join(ol.ROOT, "tests", "golden", "binner_%s_%s.tsv" % (case, variant))
