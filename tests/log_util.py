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


def special_case():
    """(SynthData, masked flag per record in FILE order): the first 24 queries of nt_small with the corner cases of
    predict() forced onto four single-segment queries -- n == 1 and n == 0 through masked records (hh:350-388), a
    full-length 100 % top hit with a score tie (the shortcut of hh:431-472) and a 100 % hit that is not the best score
    (the *ALN branch of hh:511-517)."""
    d = log_case_data_n("nt_small", 24)
    r = d.rec
    n = len(r["q"])
    masked = np.zeros(n, bool)
    segs, _ = d.segments()
    per_q = {}
    for sg in segs:
        per_q.setdefault(int(sg["query_seq"]), 0)
        per_q[int(sg["query_seq"])] += 1
    single = [q for q in sorted(per_q) if per_q[q] == 1 and int((r["q"] == q).sum()) >= 5][:4]
    assert len(single) == 4
    q0, q1, q2, q3 = single
    idx = np.flatnonzero(r["q"] == q0)
    masked[idx] = True
    masked[idx[int(np.argmax(r["score"][idx]))]] = False            # n == 1
    masked[r["q"] == q1] = True                                       # n == 0
    for q, top in ((q2, True), (q3, False)):
        idx = np.flatnonzero(r["q"] == q)
        qs, qe = int(r["qstart"][idx].min()), int(r["qstop"][idx].max())
        L = qe - qs + 1
        order = idx[np.argsort(-r["score"][idx], kind="stable")]
        k = order[0] if top else order[-1]
        r["qstart"][k], r["qstop"][k] = qs, qe
        r["alnlen"][k] = L; r["ident"][k] = L
        if top:
            r["score"][k] = 1e6
            r["score"][order[1]] = 1e6                                # tie with the identical hit
            r["score"][order[2]] = r["score"][order[3]]               # two records share the next score (upper node loop)
    return d, masked


def log_case_data_n(case, n):
    import bench_subset
    return bench_subset.subset(gu.case_data(case), n)


def write_case_files(d, masked, outdir):
    """The reference's input files with '*'-masked alignment lines (AlignmentRecord::isFiltered)."""
    d.write_files(outdir)
    r = d.rec
    with open(os.path.join(outdir, "alignments.tsv"), "w") as f:
        for k in range(len(r["q"])):
            qi = r["q"][k]
            f.write("%s%s\t%d\t%d\t%d\t%s\t%d\t%d\t%s\t0\t%d\t%d\n" % (
                "*" if masked[k] else "", d.q_names[qi], r["qstart"][k], r["qstop"][k], len(d.q_seqs[qi]),
                d.ref_names[r["r"][k]], r["rstart"][k], r["rstop"][k], repr(float(r["score"][k])), r["ident"][k], r["alnlen"][k]))


def flat_with_masks(d, masked):
    """ol.FlatData whose record sets are formed from ALL records and whose candidate table holds the unmasked ones."""
    fd = ol.FlatData(d)
    r = d.rec
    nrec = len(r["q"])
    order = np.lexsort((np.arange(nrec), r["qstop"], r["qstart"], r["q"]))
    keep = ~masked[order]
    csum = np.concatenate([[0], np.cumsum(keep)])
    segs = fd.segs.copy()
    for i, sg in enumerate(fd.segs):
        b, c = int(sg["cand_begin"]), int(sg["cand_count"])
        segs[i]["cand_begin"] = csum[b]
        segs[i]["cand_count"] = csum[b + c] - csum[b]
    fd.segs, fd.cands = segs, np.ascontiguousarray(fd.cands[keep])
    return fd


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
