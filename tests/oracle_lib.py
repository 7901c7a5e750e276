"""ctypes access to the test-infrastructure libraries: the oracle restatement
(oracle/librpa_oracle.so), the real SeqAn harness (oracle/_ref/libseqan_ref.so, when it was built)
and the CPU harness that single-steps the product's state machine (tests/_build/libhost_machine.so).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "taxator-tk_b200", "python"))
import synth  # noqa: E402

ORACLE_DIR = os.path.join(ROOT, "oracle")
BUILD_DIR = os.path.join(ROOT, "tests", "_build")

u8p = ctypes.POINTER(ctypes.c_uint8)
u32p = ctypes.POINTER(ctypes.c_uint32)
u64p = ctypes.POINTER(ctypes.c_uint64)


def _newer(target, *sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_oracle():
    so = os.path.join(ORACLE_DIR, "librpa_oracle.so")
    src = os.path.join(ORACLE_DIR, "rpa_oracle.cpp")
    if not _newer(so, src, os.path.join(ORACLE_DIR, "blosum62_table.h"), os.path.join(ORACLE_DIR, "binner_oracle.cpp"),
                  os.path.join(ROOT, "include", "taxator_rpa_b200.h")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "librpa_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def build_host_machine():
    os.makedirs(BUILD_DIR, exist_ok=True)
    so = os.path.join(BUILD_DIR, "libhost_machine.so")
    src = os.path.join(ROOT, "tests", "host_machine_harness.cpp")
    csrc = os.path.join(ROOT, "taxator-tk_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("machine.h", "shapes.h", "types.h", "hostprep.h")]
    oracle_so = build_oracle()
    if not _newer(so, *deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src,
                               oracle_so, "-Wl,-rpath," + ORACLE_DIR])
    return so


_oracle = None
_seqan = None
_hm = None


def oracle():
    global _oracle
    if _oracle is None:
        _oracle = ctypes.CDLL(build_oracle())
    return _oracle


def seqan_ref():
    """The real SeqAn harness, or None when oracle/_ref was not built (no /root/reference)."""
    global _seqan
    if _seqan is None:
        p = os.path.join(ORACLE_DIR, "_ref", "libseqan_ref.so")
        if not os.path.exists(p):
            return None
        _seqan = ctypes.CDLL(p)
    return _seqan


def host_machine():
    global _hm
    if _hm is None:
        _hm = ctypes.CDLL(build_host_machine(), mode=ctypes.RTLD_GLOBAL)
    return _hm


DNA_LUT = np.full(256, 4, np.uint8)
for i, ch in enumerate(b"ACGT"):
    DNA_LUT[ch] = i
    DNA_LUT[ch + 32] = i
DNA_LUT[ord("U")] = 3
DNA_LUT[ord("u")] = 3
AA_ORDER = b"ABCDEFGHIJKLMNOPQRSTUVWYZX*"
AA_LUT = np.full(256, 25, np.uint8)
for i, ch in enumerate(AA_ORDER):
    AA_LUT[ch] = i
    if 65 <= ch <= 90:
        AA_LUT[ch + 32] = i


def codes_of(chars: np.ndarray, protein: bool) -> np.ndarray:
    return (AA_LUT if protein else DNA_LUT)[chars]


def ptr(a, t):
    return a.ctypes.data_as(t)


class FlatData:
    """Flat arrays of a SynthData for the oracle / harness / C-ABI."""

    def __init__(self, d: "synth.SynthData"):
        self.d = d
        self.protein = bool(d.cfg.protein)
        self.parent, self.left, self.right, self.depth = d.nested_set()
        self.q_chars, self.q_off, self.q_len = d.store_arrays(d.q_seqs)
        self.r_chars, self.r_off, self.r_len = d.store_arrays(d.ref_seqs)
        self.q_codes = codes_of(self.q_chars, self.protein)
        self.r_codes = codes_of(self.r_chars, self.protein)
        self.segs, self.cands = d.segments()


def oracle_predict(fd: FlatData, exclude_factor=0.5, toppercent=0.05, want_pairs=False):
    O = oracle()
    n = len(fd.segs)
    res = np.zeros(n, dtype=synth.RESULT_DTYPE)
    pairlog_dtype = np.dtype([("pass", "<u4"), ("i", "<u4"), ("j", "<u4"), ("la", "<u4"), ("lb", "<u4"),
                              ("dist", "<f4"), ("sim", "<f4")])
    logs = []
    f = O.orc_predict_segment
    f.restype = ctypes.c_int
    for s in range(n):
        sg = fd.segs[s]
        c = np.ascontiguousarray(fd.cands[sg["cand_begin"]:sg["cand_begin"] + sg["cand_count"]])
        cap = 4 * int(sg["cand_count"]) + 8 if want_pairs else 0
        plog = np.zeros(max(cap, 1), dtype=pairlog_dtype)
        pn = ctypes.c_uint32(0)
        rc = f(ptr(fd.parent, u32p), ptr(fd.left, u32p), ptr(fd.right, u32p), ptr(fd.depth, u8p),
               ctypes.c_uint32(len(fd.parent)), ctypes.c_uint32(0),
               ptr(fd.q_codes, u8p), ptr(fd.q_off, u64p), ptr(fd.q_len, u32p), ctypes.c_uint32(len(fd.q_len)),
               ptr(fd.r_codes, u8p), ptr(fd.r_off, u64p), ptr(fd.r_len, u32p), ctypes.c_uint32(len(fd.r_len)),
               ctypes.c_int(int(fd.protein)), ctypes.c_float(exclude_factor), ctypes.c_float(toppercent),
               ctypes.c_uint32(int(sg["query_seq"])), c.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(len(c)),
               res[s:s + 1].ctypes.data_as(ctypes.c_void_p),
               plog.ctypes.data_as(ctypes.c_void_p) if want_pairs else None, ctypes.c_uint32(cap), ctypes.byref(pn))
        assert rc == 0
        if want_pairs:
            logs.append(plog[:min(pn.value, cap)].copy())
    return (res, logs) if want_pairs else res


def host_machine_predict(fd: FlatData, exclude_factor=0.5, toppercent=0.05, spec_k=0, want_trace=False):
    H = host_machine()
    n = len(fd.segs)
    res = np.zeros(n, dtype=synth.RESULT_DTYPE)
    rounds = ctypes.c_uint32(0)
    trace_dtype = np.dtype([("seg", "<u4"), ("a", "<u4"), ("b", "<u4"), ("r0", "<i4"), ("r1", "<i4"), ("len_a", "<u4"),
                            ("len_b", "<u4"), ("self", "<u4")])
    trace = np.zeros(len(fd.cands) * 8 + n * 64 + 16 if want_trace else 1, trace_dtype)
    trace_n = ctypes.c_uint32(0)
    f = H.hm_predict_batch
    f.restype = ctypes.c_int
    rc = f(ptr(fd.parent, u32p), ptr(fd.left, u32p), ptr(fd.right, u32p), ptr(fd.depth, u8p),
           ctypes.c_uint32(len(fd.parent)), ctypes.c_uint32(0),
           ptr(fd.q_codes, u8p), ptr(fd.q_off, u64p), ptr(fd.q_len, u32p), ctypes.c_uint32(len(fd.q_len)),
           ptr(fd.r_codes, u8p), ptr(fd.r_off, u64p), ptr(fd.r_len, u32p), ctypes.c_uint32(len(fd.r_len)),
           ctypes.c_int(int(fd.protein)), ctypes.c_float(exclude_factor), ctypes.c_float(toppercent),
           fd.segs.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(n),
           fd.cands.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(len(fd.cands)),
           res.ctypes.data_as(ctypes.c_void_p), ctypes.byref(rounds), ctypes.c_uint32(spec_k),
           trace.ctypes.data_as(ctypes.c_void_p) if want_trace else None, ctypes.c_uint32(len(trace) if want_trace else 0),
           ctypes.byref(trace_n))
    assert rc == 0, rc
    if want_trace:
        assert trace_n.value <= len(trace)
        t = trace[:trace_n.value]
        return res, rounds.value, t[np.argsort(t["seg"], kind="stable")]
    return res, rounds.value


RESULT_FIELDS = ["qrstart", "qrstop", "lower_node", "upper_node", "rtax_node", "support", "ival", "kind",
                 "n_pass0", "n_pass1", "n_pass2", "cells"]


def results_equal(a, b, fields=RESULT_FIELDS):
    bad = []
    for f in fields:
        x, y = a[f], b[f]
        if f == "ival":
            ne = ~((x == y) | ((a["kind"] == 0) & (b["kind"] == 0)))
        else:
            ne = x != y
        if ne.any():
            bad.append((f, np.flatnonzero(ne)[:5].tolist()))
    return bad


def oracle_predict_lca(parent, left, right, depth, segs, cands, evalue=None, unclassified=None, model=1, toppercent=0.05,
                       minscore=0.0, maxevalue=1000.0, minsupport=1, nbest=1, ignore_unclassified=False):
    """The alignment-free models (oracle/rpa_oracle.cpp orc_predict_lca_model), one call per segment."""
    O = oracle()
    f = O.orc_predict_lca_model
    f.restype = ctypes.c_int
    n = len(segs)
    res = np.zeros(n, dtype=synth.RESULT_DTYPE)
    ev = None if evalue is None else np.ascontiguousarray(evalue, np.float64)
    un = None if unclassified is None else np.ascontiguousarray(unclassified, np.uint8)
    for s in range(n):
        b, c = int(segs[s]["cand_begin"]), int(segs[s]["cand_count"])
        cc = np.ascontiguousarray(cands[b:b + c])
        evs = None if ev is None else np.ascontiguousarray(ev[b:b + c])
        rc = f(ptr(parent, u32p), ptr(left, u32p), ptr(right, u32p), ptr(depth, u8p), ctypes.c_uint32(len(parent)),
               ctypes.c_uint32(0), ctypes.c_uint32(int(model)), ctypes.c_float(toppercent), ctypes.c_float(minscore),
               ctypes.c_double(maxevalue), ctypes.c_uint32(int(minsupport)), ctypes.c_uint32(int(nbest)),
               ctypes.c_int(int(bool(ignore_unclassified))), cc.ctypes.data_as(ctypes.c_void_p),
               None if evs is None else evs.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(c),
               None if un is None else un.ctypes.data_as(ctypes.c_void_p), res[s:s + 1].ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
    return res
