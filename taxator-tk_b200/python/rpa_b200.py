"""ctypes binding of the C ABI (include/taxator_rpa_b200.h) for tests and bench.py.

This is a thin mirror of the header: every method maps 1:1 to an exported symbol.  There is no
CPU path here; creating a context without a CUDA device raises."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(HERE), "lib", "libtaxator_rpa_b200.so")

CAND_DTYPE = np.dtype([("ref_seq", "<u4"), ("rstart", "<u4"), ("rstop", "<u4"), ("qstart", "<u4"), ("qstop", "<u4"),
                       ("score", "<f4"), ("identities", "<u4"), ("alnlen", "<u4"), ("node", "<u4")])
SEG_DTYPE = np.dtype([("query_seq", "<u4"), ("cand_begin", "<u4"), ("cand_count", "<u4"), ("reserved", "<u4")])
RESULT_DTYPE = np.dtype([("qrstart", "<u4"), ("qrstop", "<u4"), ("lower_node", "<u4"), ("upper_node", "<u4"),
                         ("rtax_node", "<u4"), ("support", "<u4"), ("ival", "<f4"), ("signal", "<f4"),
                         ("n_pass0", "<u4"), ("n_pass1", "<u4"), ("n_pass2", "<u4"), ("kind", "<u4"),
                         ("cells", "<u8")])
TRACE_DTYPE = np.dtype([("seg", "<u4"), ("a", "<u4"), ("b", "<u4"), ("r0", "<i4"), ("r1", "<i4"), ("len_a", "<u4"),
                        ("len_b", "<u4"), ("self", "<u4")])
assert CAND_DTYPE.itemsize == 36 and SEG_DTYPE.itemsize == 16 and RESULT_DTYPE.itemsize == 56 and TRACE_DTYPE.itemsize == 32


class Profile(ctypes.Structure):
    _fields_ = [("ms_edit_distance", ctypes.c_double), ("launches_edit_distance", ctypes.c_uint64),
                ("cells_edit_distance", ctypes.c_uint64),
                ("ms_protein", ctypes.c_double), ("launches_protein", ctypes.c_uint64), ("cells_protein", ctypes.c_uint64),
                ("ms_stage", ctypes.c_double), ("launches_stage", ctypes.c_uint64), ("bytes_stage", ctypes.c_uint64),
                ("ms_decide", ctypes.c_double), ("launches_decide", ctypes.c_uint64),
                ("ms_other", ctypes.c_double), ("launches_other", ctypes.c_uint64),
                ("rounds", ctypes.c_uint64), ("pairs", ctypes.c_uint64), ("band_retries", ctypes.c_uint64), ("wedge_failures", ctypes.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


MODEL_DUMMY, MODEL_SIMPLE_LCA, MODEL_MEGAN_LCA, MODEL_NBEST_LCA = 0, 1, 2, 3
KIND_NONE, KIND_SINGLE, KIND_IDENTICAL, KIND_PLACED, KIND_LCA = 0, 1, 2, 3, 4


class LcaParams(ctypes.Structure):     # trpa_lca_params
    _fields_ = [("model", ctypes.c_uint32), ("toppercent", ctypes.c_float), ("minscore", ctypes.c_float),
                ("maxevalue", ctypes.c_float), ("minsupport", ctypes.c_uint32), ("nbest", ctypes.c_uint32),
                ("ignore_unclassified", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


EXPORTS = ["trpa_abi_version", "trpa_last_error", "trpa_create", "trpa_destroy", "trpa_set_params",
           "trpa_set_arena_bytes", "trpa_set_lookahead", "trpa_set_band", "trpa_set_tuning", "trpa_profile_reset", "trpa_profile_get", "trpa_load_taxonomy", "trpa_load_store", "trpa_store_info", "trpa_export_store", "trpa_load_store_packed",
           "trpa_predict_batch", "trpa_batch_upload", "trpa_batch_run", "trpa_batch_download",
           "trpa_predict_lca_batch", "trpa_edit_distance_batch", "trpa_protein_align_batch", "trpa_fetch_segments", "trpa_lca_batch",
           "trpa_int_alu_peak", "trpa_shard_bounds", "trpa_batch_results_dev", "trpa_bin_batch", "trpa_set_trace", "trpa_batch_trace", "trpa_host_alloc", "trpa_host_free"]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("CUDA extension %s is missing: run `make -C taxator-tk_b200` (or __graft_entry__.build())"
                               % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.trpa_last_error.restype = ctypes.c_char_p
        L.trpa_create.restype = ctypes.c_void_p
        L.trpa_create.argtypes = [ctypes.c_int, ctypes.c_void_p]
        L.trpa_destroy.argtypes = [ctypes.c_void_p]
        L.trpa_destroy.restype = None
        _lib = L
    return _lib


class TrpaError(RuntimeError):
    pass


def shard_bounds(segs, cands, world):
    """trpa_shard_bounds: world+1 segment indices cutting the table into shards of about equal DP work
    (host-only helper of the library; needs no GPU)."""
    L = lib()
    segs = np.ascontiguousarray(segs, SEG_DTYPE); cands = np.ascontiguousarray(cands, CAND_DTYPE)
    out = np.zeros(int(world) + 1, np.uint32)
    rc = L.trpa_shard_bounds(segs.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(len(segs)),
                             cands.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(len(cands)),
                             ctypes.c_uint32(int(world)), out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise TrpaError("rc=%d: %s" % (rc, L.trpa_last_error().decode()))
    return [int(x) for x in out]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Context:
    def __init__(self, device=0, stream=None):
        L = lib()
        self.L = L
        self.h = L.trpa_create(int(device), ctypes.c_void_p(stream) if stream else None)
        if not self.h:
            raise TrpaError(L.trpa_last_error().decode())
        self.h = ctypes.c_void_p(self.h)
        self._keep = []

    def close(self):
        if self.h:
            self.L.trpa_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise TrpaError("rc=%d: %s" % (rc, self.L.trpa_last_error().decode()))

    def set_params(self, exclude_factor=0.5, toppercent=0.05):
        self._ck(self.L.trpa_set_params(self.h, ctypes.c_float(exclude_factor), ctypes.c_float(toppercent)))

    def set_arena_bytes(self, nbytes):
        self._ck(self.L.trpa_set_arena_bytes(self.h, ctypes.c_uint64(int(nbytes))))

    def set_band(self, on):
        self._ck(self.L.trpa_set_band(self.h, ctypes.c_int(int(on))))

    def set_tuning(self, key, value):
        self._ck(self.L.trpa_set_tuning(self.h, ctypes.c_char_p(key.encode()), ctypes.c_int64(int(value))))

    def set_lookahead(self, k):
        self._ck(self.L.trpa_set_lookahead(self.h, int(k)))

    def load_taxonomy(self, parent, left, right, depth, root=0):
        parent = np.ascontiguousarray(parent, np.uint32); left = np.ascontiguousarray(left, np.uint32)
        right = np.ascontiguousarray(right, np.uint32); depth = np.ascontiguousarray(depth, np.uint8)
        self._ck(self.L.trpa_load_taxonomy(self.h, _p(parent), _p(left), _p(right), _p(depth),
                                           ctypes.c_uint32(len(parent)), ctypes.c_uint32(root)))

    def load_store(self, store, alphabet, chars, off, lens):
        chars = np.ascontiguousarray(chars, np.uint8); off = np.ascontiguousarray(off, np.uint64)
        lens = np.ascontiguousarray(lens, np.uint32)
        self._ck(self.L.trpa_load_store(self.h, int(store), int(alphabet), _p(chars), _p(off), _p(lens),
                                        ctypes.c_uint32(len(lens))))

    def export_store(self, store):
        """(alphabet, woff[n+1], len[n], payload bytes) of a loaded store in its packed HBM layout."""
        alpha = ctypes.c_int(0); n_seq = ctypes.c_uint32(0); n_words = ctypes.c_uint64(0)
        self._ck(self.L.trpa_store_info(self.h, int(store), ctypes.byref(alpha), ctypes.byref(n_seq), ctypes.byref(n_words)))
        woff = np.zeros(n_seq.value + 1, np.uint64); lens = np.zeros(max(n_seq.value, 1), np.uint32)
        payload = np.zeros(max(1, n_words.value * (12 if alpha.value == 0 else 4)), np.uint8)
        self._ck(self.L.trpa_export_store(self.h, int(store), _p(woff), _p(lens), _p(payload)))
        return alpha.value, woff, lens[:n_seq.value], payload[:n_words.value * (12 if alpha.value == 0 else 4)]

    def load_store_packed(self, store, alphabet, woff, lens, payload):
        woff = np.ascontiguousarray(woff, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
        payload = np.ascontiguousarray(payload, np.uint8)
        self._ck(self.L.trpa_load_store_packed(self.h, int(store), int(alphabet), _p(woff), _p(lens), ctypes.c_uint32(len(lens)),
                                               _p(payload), ctypes.c_uint64(int(woff[-1]) if len(woff) else 0)))

    def predict_batch(self, segs, cands):
        segs = np.ascontiguousarray(segs, SEG_DTYPE); cands = np.ascontiguousarray(cands, CAND_DTYPE)
        out = np.zeros(len(segs), RESULT_DTYPE)
        self._ck(self.L.trpa_predict_batch(self.h, _p(segs), ctypes.c_uint32(len(segs)), _p(cands),
                                           ctypes.c_uint32(len(cands)), _p(out)))
        return out

    def predict_batch_into(self, segs, cands, out):
        """trpa_predict_batch with caller-owned (e.g. pinned) buffers; no copies on the Python side."""
        assert segs.dtype == SEG_DTYPE and cands.dtype == CAND_DTYPE and out.dtype == RESULT_DTYPE and len(out) >= len(segs)
        self._ck(self.L.trpa_predict_batch(self.h, _p(segs), ctypes.c_uint32(len(segs)), _p(cands),
                                           ctypes.c_uint32(len(cands)), _p(out)))
        return out[:len(segs)]

    def batch_upload(self, segs, cands):
        segs = np.ascontiguousarray(segs, SEG_DTYPE); cands = np.ascontiguousarray(cands, CAND_DTYPE)
        self._n_segs = len(segs)
        self._ck(self.L.trpa_batch_upload(self.h, _p(segs), ctypes.c_uint32(len(segs)), _p(cands),
                                          ctypes.c_uint32(len(cands))))

    def batch_run(self):
        self._ck(self.L.trpa_batch_run(self.h))

    def batch_download(self, out=None):
        if out is None:
            out = np.zeros(self._n_segs, RESULT_DTYPE)
        self._ck(self.L.trpa_batch_download(self.h, _p(out)))
        return out

    def set_trace(self, on):
        self._ck(self.L.trpa_set_trace(self.h, ctypes.c_int(int(on))))

    def batch_trace(self):
        """The alignments the last batch_run consumed (TRACE_DTYPE), ordered by segment, reference order inside."""
        n = ctypes.c_uint64(0)
        self._ck(self.L.trpa_batch_trace(self.h, None, ctypes.c_uint64(0), ctypes.byref(n)))
        out = np.zeros(n.value, TRACE_DTYPE)
        if n.value:
            self._ck(self.L.trpa_batch_trace(self.h, _p(out), ctypes.c_uint64(n.value), ctypes.byref(n)))
        return out

    def batch_results_dev(self):
        """(device pointer, n_segs) of the result table of the last batch_run (valid until the next upload)."""
        ptr = ctypes.c_void_p(0); n = ctypes.c_uint32(0)
        self._ck(self.L.trpa_batch_results_dev(self.h, ctypes.byref(ptr), ctypes.byref(n)))
        return ptr.value, n.value

    def profile_reset(self):
        self._ck(self.L.trpa_profile_reset(self.h))

    def profile(self):
        p = Profile()
        self._ck(self.L.trpa_profile_get(self.h, ctypes.byref(p)))
        return p.as_dict()

    def edit_distance_batch(self, chars, off, lens, pa, pb, repeat=1):
        chars = np.ascontiguousarray(chars, np.uint8); off = np.ascontiguousarray(off, np.uint64)
        lens = np.ascontiguousarray(lens, np.uint32)
        pa = np.ascontiguousarray(pa, np.uint32); pb = np.ascontiguousarray(pb, np.uint32)
        out = np.zeros(len(pa), np.int32)
        ms = ctypes.c_double(0)
        self._ck(self.L.trpa_edit_distance_batch(self.h, _p(chars), _p(off), _p(lens), ctypes.c_uint32(len(lens)),
                                                 _p(pa), _p(pb), ctypes.c_uint32(len(pa)), _p(out), int(repeat),
                                                 ctypes.byref(ms)))
        return out, ms.value

    def protein_align_batch(self, chars, off, lens, pa, pb, repeat=1):
        chars = np.ascontiguousarray(chars, np.uint8); off = np.ascontiguousarray(off, np.uint64)
        lens = np.ascontiguousarray(lens, np.uint32)
        pa = np.ascontiguousarray(pa, np.uint32); pb = np.ascontiguousarray(pb, np.uint32)
        out = np.zeros((len(pa), 3), np.int32)
        ms = ctypes.c_double(0)
        self._ck(self.L.trpa_protein_align_batch(self.h, _p(chars), _p(off), _p(lens), ctypes.c_uint32(len(lens)),
                                                 _p(pa), _p(pb), ctypes.c_uint32(len(pa)), _p(out), int(repeat),
                                                 ctypes.byref(ms)))
        return out, ms.value

    def fetch_segments(self, ref_seq, start, stop, left_ext, right_ext):
        arrs = [np.ascontiguousarray(a, np.uint32) for a in (ref_seq, start, stop, left_ext, right_ext)]
        n = len(arrs[0])
        span = np.abs(arrs[1].astype(np.int64) - arrs[2].astype(np.int64)) + 1 + arrs[3] + arrs[4]
        cap = int(span.sum()) + 16
        out = np.zeros(cap, np.uint8)
        ooff = np.zeros(n, np.uint64); olen = np.zeros(n, np.uint32)
        self._ck(self.L.trpa_fetch_segments(self.h, *[_p(a) for a in arrs], ctypes.c_uint32(n), _p(out),
                                            ctypes.c_uint64(cap), _p(ooff), _p(olen)))
        return [out[int(o):int(o) + int(l)] for o, l in zip(ooff, olen)]

    def predict_lca_batch(self, model, segs, cands, evalue=None, unclassified=None, toppercent=0.05, minscore=0.0,
                          maxevalue=1000.0, minsupport=1, nbest=1, ignore_unclassified=False, repeat=1):
        """The alignment-free models (MODEL_DUMMY / SIMPLE_LCA / MEGAN_LCA / NBEST_LCA) on the loaded taxonomy.
        Returns (results, device ms per run)."""
        segs = np.ascontiguousarray(segs, SEG_DTYPE); cands = np.ascontiguousarray(cands, CAND_DTYPE)
        pp = LcaParams(int(model), float(toppercent), float(minscore), float(np.float32(maxevalue)), int(minsupport),
                       int(nbest), int(bool(ignore_unclassified)), 0)
        ev = None if evalue is None else np.ascontiguousarray(evalue, np.float64)
        un = None if unclassified is None else np.ascontiguousarray(unclassified, np.uint8)
        out = np.zeros(len(segs), RESULT_DTYPE)
        ms = ctypes.c_double(0)
        self._ck(self.L.trpa_predict_lca_batch(self.h, ctypes.byref(pp), _p(segs), ctypes.c_uint32(len(segs)), _p(cands),
                                               ctypes.c_uint32(len(cands)), None if ev is None else _p(ev),
                                               None if un is None else _p(un), _p(out), ctypes.c_int(int(repeat)),
                                               ctypes.byref(ms)))
        return out, ms.value

    def lca_batch(self, a, b):
        a = np.ascontiguousarray(a, np.uint32); b = np.ascontiguousarray(b, np.uint32)
        out = np.zeros(len(a), np.uint32)
        self._ck(self.L.trpa_lca_batch(self.h, _p(a), _p(b), ctypes.c_uint32(len(a)), _p(out)))
        return out

    def int_alu_peak(self):
        v = ctypes.c_double(0)
        self._ck(self.L.trpa_int_alu_peak(self.h, ctypes.byref(v)))
        return v.value
