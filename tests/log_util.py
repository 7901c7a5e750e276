"""Test helpers for the verbose per-alignment log: the cases, masking of the reference's
CPU-time columns, block splitting, and the host harness that writes the log from a trace."""
import ctypes
import gzip
import os
import subprocess

import numpy as np

import golden_util as gu
import oracle_lib as ol

LOG_CASES = {"nt_small": 60, "aa_small": 40, "nt_indel": 30}   # first N queries of the golden case

TRACE_DTYPE = np.dtype([("seg", "<u4"), ("a", "<u4"), ("b", "<u4"), ("r0", "<i4"), ("r1", "<i4"), ("len_a", "<u4"),
                        ("len_b", "<u4"), ("self", "<u4")])


def golden_log_path(case):
    return os.path.join(gu.GOLDEN, "log_%s.log.gz" % case)


def log_case_data(case):
    """The golden case cut down to its first LOG_CASES[case] queries (same refpack / taxonomy)."""
    import bench_subset
    return bench_subset.subset(gu.case_data(case), LOG_CASES[case])


def blocks_of(text):
    """The log split into per-segment blocks ('ID\\t...' to the STATS line), CPU-time columns of STATS masked."""
    out, cur = [], []
    for line in text.splitlines():
        if line.startswith("ID\t") and cur:
            out.append("\n".join(cur)); cur = []
        if line.startswith("STATS\t"):
            f = line.split("\t")
            if len(f) >= 11:
                f[7] = f[8] = f[9] = "0"
            line = "\t".join(f)
        cur.append(line)
    if cur:
        out.append("\n".join(cur))
    return [b.rstrip("\n") for b in out]


def golden_blocks(case):
    return blocks_of(gzip.open(golden_log_path(case), "rb").read().decode())


_hl = None


def host_log():
    global _hl
    if _hl is None:
        os.makedirs(ol.BUILD_DIR, exist_ok=True)
        so = os.path.join(ol.BUILD_DIR, "libhost_log.so")
        host = os.path.join(ol.ROOT, "taxator-tk_b200", "host")
        srcs = [os.path.join(ol.ROOT, "tests", "host_log_harness.cpp"), os.path.join(host, "verbose_log.cpp"), os.path.join(host, "taxonomy.cpp")]
        deps = srcs + [os.path.join(host, "verbose_log.h"), os.path.join(ol.ROOT, "include", "taxator_rpa_b200.h")]
        if not ol._newer(so, *deps):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I" + os.path.join(ol.ROOT, "include"), "-I" + host,
                                   "-o", so] + srcs + ["-lz"])
        _hl = ctypes.CDLL(so)
    return _hl


def write_log(fd, res, trace, out_path, exclude_factor=0.5, toppercent=0.05):
    """host/verbose_log.cpp over the flat data of fd, results and a trace (TRACE_DTYPE, ordered by segment)."""
    H = host_log()
    d = fd.d
    names = [("node%d" % t).encode() for t in d.tax_ids]
    names_off = np.concatenate([[0], np.cumsum([len(x) for x in names])]).astype(np.uint32)
    qids = [d.q_names[int(s["query_seq"])].encode() for s in fd.segs]
    qid_off = np.concatenate([[0], np.cumsum([len(x) for x in qids])]).astype(np.uint32)
    vp = ctypes.c_void_p
    err = ctypes.create_string_buffer(512)
    parent = np.ascontiguousarray(fd.parent, np.uint32)
    trace = np.ascontiguousarray(trace, TRACE_DTYPE)
    res = np.ascontiguousarray(res)
    rc = H.hl_write_log(vp(parent.ctypes.data), vp(fd.left.ctypes.data), vp(fd.right.ctypes.data), vp(fd.depth.ctypes.data),
                        ctypes.c_uint32(len(parent)), b"".join(names), vp(names_off.ctypes.data),
                        vp(fd.q_chars.ctypes.data), vp(fd.q_off.ctypes.data), vp(fd.q_len.ctypes.data), ctypes.c_uint32(len(fd.q_len)),
                        vp(fd.r_chars.ctypes.data), vp(fd.r_off.ctypes.data), vp(fd.r_len.ctypes.data), ctypes.c_uint32(len(fd.r_len)),
                        ctypes.c_int(int(fd.protein)), ctypes.c_float(exclude_factor), ctypes.c_float(toppercent),
                        vp(fd.segs.ctypes.data), ctypes.c_uint32(len(fd.segs)), vp(fd.cands.ctypes.data), vp(res.ctypes.data),
                        b"".join(qids), vp(qid_off.ctypes.data), vp(trace.ctypes.data), ctypes.c_uint32(len(trace)),
                        out_path.encode(), err, ctypes.c_uint32(512))
    assert rc == 0, err.value.decode()
