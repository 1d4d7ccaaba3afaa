"""ctypes binding of libpctsea_b200.so (include/pctsea_b200.h). Thin: numpy in, numpy out, error codes -> exceptions.

There is no CPU fallback anywhere in this package: if the CUDA library is missing or no device is present, the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libpctsea_b200.so")

PCTSEA_OK, PCTSEA_EINVAL, PCTSEA_ECUDA, PCTSEA_ENCCL, PCTSEA_ETOOFEW, PCTSEA_EMINGENES, PCTSEA_ENOMEM = 0, -1, -2, -3, -4, -5, -6
METHOD_PEARSONS_CORRELATION, METHOD_SIMPLE_SCORE = 0, 1
JITTER_NAN, JITTER_PHILOX = 0, 1
PERM_REPLAY, PERM_PHILOX = 0, 1
KS_EXACT, KS_FAST = 0, 1

DEFAULT_FEISTEL_ROUNDS = 5   # PCTSEA_DEFAULT_FEISTEL_ROUNDS of include/pctsea_b200.h

# pctsea_ctx_set_option: testing / tuning switches (include/pctsea_b200.h)
(OPT_OBS_OVERLAP, OPT_SCORE_PATH, OPT_GM_UNROLL, OPT_SCORE_WARPS, OPT_PERM_CHUNK, OPT_NULL_THREADS, OPT_NULL_CTAS_PER_SM,
 OPT_NULL_MIN_THREADS, OPT_DEBUG_GAPS, OPT_SHARD_SCORE, OPT_OBS_CHAIN_MB, OPT_NULL_FINE_ENTRIES, OPT_SHARD_OBSERVED) = range(1, 14)

EXPORTS = [
    "pctsea_abi_version", "pctsea_create", "pctsea_destroy", "pctsea_last_error", "pctsea_get_timings",
    "pctsea_ctx_set_option",
    "pctsea_stream", "pctsea_matrix_load_csr", "pctsea_matrix_wrap_device", "pctsea_matrix_set_labels", "pctsea_matrix_build_gene_index",
    "pctsea_matrix_free", "pctsea_score", "pctsea_rank", "pctsea_enrich", "pctsea_reduce_null",
    "pctsea_reduce_null_dev", "pctsea_analyze", "pctsea_analyze_batch", "pctsea_last_null_dev", "pctsea_last_type_counts",
    "pctsea_last_observed_curve", "pctsea_comm_unique_id", "pctsea_comm_init_rank", "pctsea_comm_destroy",
]
COMM_ID_BYTES = 128   # PCTSEA_COMM_ID_BYTES


class PctseaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.msg = msg


class ScoreParams(C.Structure):
    _fields_ = [("method", C.c_int32), ("min_genes_cells", C.c_int32), ("min_corr", C.c_double),
                ("has_min_corr", C.c_int32), ("jitter_mode", C.c_int32), ("seed", C.c_uint64)]


class NullParams(C.Structure):
    _fields_ = [("n_perm", C.c_int32), ("perm_begin", C.c_int32), ("perm_count", C.c_int32),
                ("perm_mode", C.c_int32), ("ks_mode", C.c_int32), ("feistel_rounds", C.c_int32),
                ("seed", C.c_uint64), ("mean_policy", C.c_int32), ("min_pass", C.c_int32)]


class TypeResult(C.Structure):
    _fields_ = [("ews", C.c_float), ("dstat", C.c_float), ("supX", C.c_int32), ("norm_supX", C.c_double),
                ("sizeA", C.c_int32), ("sizeB", C.c_int32), ("raw_sup", C.c_float), ("norm_ews", C.c_float),
                ("p_emp", C.c_double), ("fdr", C.c_double), ("sec_ews", C.c_float), ("sec_supX", C.c_int32),
                ("ks_d", C.c_double)]


class Timings(C.Structure):
    _fields_ = [("setup_ms", C.c_float), ("score_ms", C.c_float), ("rank_ms", C.c_float),
                ("observed_ms", C.c_float), ("labels_ms", C.c_float), ("null_ms", C.c_float),
                ("reduce_ms", C.c_float), ("total_ms", C.c_float), ("n_pass", C.c_int64), ("n_used", C.c_int64),
                ("launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("gather_ms", C.c_float), ("reserved0", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


TYPE_RESULT_DTYPE = np.dtype([("ews", "<f4"), ("dstat", "<f4"), ("supX", "<i4"), ("_pad0", "<i4"),
                              ("norm_supX", "<f8"), ("sizeA", "<i4"), ("sizeB", "<i4"), ("raw_sup", "<f4"),
                              ("norm_ews", "<f4"), ("p_emp", "<f8"), ("fdr", "<f8"), ("sec_ews", "<f4"),
                              ("sec_supX", "<i4"), ("ks_d", "<f8")])
assert TYPE_RESULT_DTYPE.itemsize == C.sizeof(TypeResult) == 72

_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree CUDA library. Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: run `python -m pctsea_parent_b200.build` (nvcc, sm_100a). "
                          "pctsea_parent_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.pctsea_abi_version.restype = C.c_int
    L.pctsea_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.pctsea_destroy.argtypes = [vp]
    L.pctsea_last_error.argtypes = [vp]
    L.pctsea_last_error.restype = C.c_char_p
    L.pctsea_get_timings.argtypes = [vp, C.POINTER(Timings)]
    L.pctsea_ctx_set_option.argtypes = [vp, i32, i64]
    L.pctsea_stream.argtypes = [vp]
    L.pctsea_stream.restype = vp
    L.pctsea_matrix_load_csr.argtypes = [vp, i64, i32, vp, vp, vp, C.POINTER(vp)]
    L.pctsea_matrix_wrap_device.argtypes = [vp, i64, i32, vp, vp, vp, C.POINTER(vp)]
    L.pctsea_matrix_set_labels.argtypes = [vp, vp, vp, i32]
    L.pctsea_matrix_build_gene_index.argtypes = [vp, vp]
    L.pctsea_matrix_free.argtypes = [vp, vp]
    L.pctsea_score.argtypes = [vp, vp, i32, vp, vp, C.POINTER(ScoreParams), vp, vp, vp, vp]
    L.pctsea_rank.argtypes = [vp, i64, vp, vp, f64, vp, C.POINTER(i64)]
    L.pctsea_enrich.argtypes = [vp, i64, vp, i64, vp, vp, i32, C.POINTER(NullParams), vp, vp, vp, vp]
    L.pctsea_reduce_null.argtypes = [vp, i32, i32, vp, i32, vp]
    L.pctsea_reduce_null_dev.argtypes = [vp, i32, i32, vp, i32, vp]
    L.pctsea_analyze.argtypes = [vp, vp, i32, vp, vp, C.POINTER(ScoreParams), f64, vp, i32, C.POINTER(NullParams),
                                 vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(i64)]
    L.pctsea_analyze_batch.argtypes = [vp, vp, i32, vp, vp, vp, C.POINTER(ScoreParams), f64, vp, i32,
                                       C.POINTER(NullParams), vp, vp, vp, vp, vp]
    L.pctsea_last_null_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(i32)]
    L.pctsea_last_type_counts.argtypes = [vp, i32, vp, vp, C.POINTER(i64)]
    L.pctsea_last_observed_curve.argtypes = [vp, i32, i64, vp, vp]
    L.pctsea_comm_unique_id.argtypes = [vp]
    L.pctsea_comm_init_rank.argtypes = [vp, i32, i32, vp]
    L.pctsea_comm_destroy.argtypes = [vp]
    for name in EXPORTS:
        if name not in ("pctsea_last_error", "pctsea_stream"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Matrix:
    def __init__(self, ctx: "Context", handle, n_cells: int, n_genes: int, keepalive=None):
        self.ctx, self.handle, self.n_cells, self.n_genes = ctx, handle, n_cells, n_genes
        self.n_types = 0
        self.gene_index = False
        self._keepalive = keepalive

    def set_labels(self, label, n_types: int):
        label = _c(label, np.int32)
        assert label.size == self.n_cells
        self.ctx._check(self.ctx.L.pctsea_matrix_set_labels(self.ctx.h, self.handle, _ptr(label), n_types))
        self.n_types = n_types

    def build_gene_index(self):
        """Gene-major copy of the matrix (once per dataset): scoring then reads only the list's genes."""
        self.ctx._check(self.ctx.L.pctsea_matrix_build_gene_index(self.ctx.h, self.handle))
        self.gene_index = True
        return self

    def free(self):
        if self.handle is not None:
            self.ctx.L.pctsea_matrix_free(self.ctx.h, self.handle)
            self.handle = None


class Context:
    """One pctsea_ctx: bound to one device, one analysis at a time."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.pctsea_create(device, C.byref(h))
        if rc != 0:
            raise PctseaError(rc, "pctsea_create failed: no usable CUDA device (there is no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None) is not None:
            self.L.pctsea_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise PctseaError(rc, (self.L.pctsea_last_error(self.h) or b"").decode())

    def timings(self) -> dict:
        t = Timings()
        self._check(self.L.pctsea_get_timings(self.h, C.byref(t)))
        return t.as_dict()

    def set_option(self, option: int, value: int):
        """Testing / tuning switch of this ctx (OPT_* constants; none is needed in production)."""
        self._check(self.L.pctsea_ctx_set_option(self.h, int(option), int(value)))

    # ---- multi-GPU: one Context per GPU, joined by an NCCL communicator inside the library
    @staticmethod
    def comm_unique_id() -> bytes:
        """128 bytes from rank 0 (ncclGetUniqueId); hand them to every rank, then call comm_init_rank everywhere."""
        buf = C.create_string_buffer(COMM_ID_BYTES)
        rc = load_library().pctsea_comm_unique_id(buf)
        if rc != 0:
            raise PctseaError(rc, "pctsea_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def comm_init_rank(self, nranks: int, rank: int, unique_id: bytes):
        """Collective over the ranks. Afterwards analyze() shards the permutations over the ranks and returns the
        complete results (and null) on every rank."""
        assert len(unique_id) == COMM_ID_BYTES
        self._check(self.L.pctsea_comm_init_rank(self.h, int(nranks), int(rank), C.c_char_p(unique_id)))
        self.comm_nranks, self.comm_rank = int(nranks), int(rank)

    def comm_destroy(self):
        self._check(self.L.pctsea_comm_destroy(self.h))
        self.comm_nranks, self.comm_rank = 1, 0

    def stream(self) -> int:
        return int(self.L.pctsea_stream(self.h) or 0)

    # ---- matrix ---------------------------------------------------------------------------------
    def load_csr(self, row_ptr, col, val, n_genes: int) -> Matrix:
        row_ptr, col, val = _c(row_ptr, np.int64), _c(col, np.int32), _c(val, np.float32)
        n = row_ptr.size - 1
        h = C.c_void_p()
        self._check(self.L.pctsea_matrix_load_csr(self.h, n, n_genes, _ptr(row_ptr), _ptr(col), _ptr(val), C.byref(h)))
        return Matrix(self, h, n, n_genes)

    def wrap_device_csr(self, row_ptr_dev: int, col_dev: int, val_dev: int, n_cells: int, n_genes: int,
                        keepalive=None) -> Matrix:
        h = C.c_void_p()
        self._check(self.L.pctsea_matrix_wrap_device(self.h, n_cells, n_genes, C.c_void_p(row_ptr_dev),
                                                     C.c_void_p(col_dev), C.c_void_p(val_dev), C.byref(h)))
        return Matrix(self, h, n_cells, n_genes, keepalive)

    # ---- stages -----------------------------------------------------------------------------------
    def score(self, m: Matrix, gene_idx, abundance, method=METHOD_PEARSONS_CORRELATION, min_genes_cells=4,
              min_corr=0.2, has_min_corr=True, jitter_mode=JITTER_PHILOX, seed=0):
        gene_idx, abundance = _c(gene_idx, np.int32), _c(abundance, np.float32)
        prm = ScoreParams(method, min_genes_cells, float(min_corr), int(bool(has_min_corr)), jitter_mode, seed)
        n = m.n_cells
        sc, used = np.empty(n, np.float64), np.empty(n, np.int32)
        nz, kept = np.empty(n, np.int32), np.empty(n, np.uint8)
        self._check(self.L.pctsea_score(self.h, m.handle, gene_idx.size, _ptr(gene_idx), _ptr(abundance),
                                        C.byref(prm), _ptr(sc), _ptr(used), _ptr(nz), _ptr(kept)))
        return sc, used, nz, kept

    def rank(self, score, kept=None, min_score=0.0):
        score = _c(score, np.float64)
        kept = None if kept is None else _c(kept, np.uint8)
        order = np.empty(max(score.size, 1), np.int32)
        npass = C.c_int64()
        self._check(self.L.pctsea_rank(self.h, score.size, _ptr(score), _ptr(kept), float(min_score), _ptr(order),
                                       C.byref(npass)))
        return order[:npass.value].copy(), int(npass.value)

    @staticmethod
    def null_params(n_perm, perm_mode=PERM_PHILOX, ks_mode=KS_FAST, seed=0, perm_begin=0, perm_count=0,
                    feistel_rounds=0, mean_policy=0, min_pass=0) -> NullParams:
        return NullParams(n_perm, perm_begin, perm_count, perm_mode, ks_mode, feistel_rounds, seed, mean_policy,
                          min_pass)

    def enrich(self, score, order, label, n_types, nprm: NullParams, replay=None, want_null=True):
        score, order, label = _c(score, np.float64), _c(order, np.int32), _c(label, np.int32)
        replay = None if replay is None else _c(replay, np.int32)
        cnt = nprm.perm_count or (nprm.n_perm - nprm.perm_begin)
        res = np.zeros(n_types, TYPE_RESULT_DTYPE)
        ne = np.empty((n_types, cnt), np.float32) if want_null else None
        nd = np.empty((n_types, cnt), np.float32) if want_null else None
        self._check(self.L.pctsea_enrich(self.h, score.size, _ptr(score), order.size, _ptr(order), _ptr(label),
                                         n_types, C.byref(nprm), _ptr(replay), _ptr(res), _ptr(ne), _ptr(nd)))
        return res, ne, nd

    def reduce_null(self, results, null_ews, mean_policy=0):
        res = np.array(results, dtype=TYPE_RESULT_DTYPE, copy=True)
        null_ews = _c(null_ews, np.float32)
        T, P = null_ews.shape
        self._check(self.L.pctsea_reduce_null(self.h, T, P, _ptr(null_ews), mean_policy, _ptr(res)))
        return res

    def reduce_null_dev(self, results, null_dev_ptr: int, n_types: int, n_perm: int, mean_policy=0):
        res = np.array(results, dtype=TYPE_RESULT_DTYPE, copy=True)
        self._check(self.L.pctsea_reduce_null_dev(self.h, n_types, n_perm, C.c_void_p(null_dev_ptr), mean_policy,
                                                  _ptr(res)))
        return res

    def analyze(self, m: Matrix, gene_idx, abundance, sprm: ScoreParams, min_score, n_types, nprm: NullParams,
                label=None, replay=None, want_null=False, want_scores=False, want_order=False):
        gene_idx, abundance = _c(gene_idx, np.int32), _c(abundance, np.float32)
        label = None if label is None else _c(label, np.int32)
        replay = None if replay is None else _c(replay, np.int32)
        cnt = nprm.perm_count or (nprm.n_perm - nprm.perm_begin)   # with a communicator: the complete null, n_perm
        n = m.n_cells
        res = np.zeros(n_types, TYPE_RESULT_DTYPE)
        ne = np.empty((n_types, cnt), np.float32) if want_null else None
        nd = np.empty((n_types, cnt), np.float32) if want_null else None
        sc = np.empty(n, np.float64) if want_scores else None
        used = np.empty(n, np.int32) if want_scores else None
        kept = np.empty(n, np.uint8) if want_scores else None
        order = np.empty(n, np.int32) if want_order else None
        npass = C.c_int64()
        self._check(self.L.pctsea_analyze(self.h, m.handle, gene_idx.size, _ptr(gene_idx), _ptr(abundance),
                                          C.byref(sprm), float(min_score), _ptr(label), n_types, C.byref(nprm),
                                          _ptr(replay), _ptr(res), _ptr(ne), _ptr(nd), _ptr(sc), _ptr(used),
                                          _ptr(kept), _ptr(order), C.byref(npass)))
        out = {"results": res, "n_pass": int(npass.value)}
        if want_null:
            out["null"], out["nullD"] = ne, nd
        if want_scores:
            out["score"], out["n_genes_used"], out["kept"] = sc, used, kept
        if want_order:
            out["order"] = order[:npass.value].copy()
        return out

    def analyze_batch(self, m: Matrix, lists, sprm: ScoreParams, min_score, n_types, nprm: NullParams, label=None,
                      want_null=False):
        """Many input lists against one resident matrix (BASELINE config E). `lists` = sequence of
        (gene_idx, abundance). Returns results [n_lists][n_types], n_pass [n_lists], status [n_lists] (0, or the
        PCTSEA_ETOOFEW / PCTSEA_EMINGENES code of a list that failed the reference's input checks)."""
        lens = [len(g) for g, _ in lists]
        ptr = np.zeros(len(lists) + 1, np.int64)
        ptr[1:] = np.cumsum(lens)
        gi = _c(np.concatenate([np.asarray(g, np.int32) for g, _ in lists]) if lists else [], np.int32)
        ab = _c(np.concatenate([np.asarray(a, np.float32) for _, a in lists]) if lists else [], np.float32)
        label = None if label is None else _c(label, np.int32)
        cnt = nprm.perm_count or (nprm.n_perm - nprm.perm_begin)
        res = np.zeros((len(lists), n_types), TYPE_RESULT_DTYPE)
        npass = np.zeros(len(lists), np.int64)
        status = np.zeros(len(lists), np.int32)
        ne = np.empty((len(lists), n_types, cnt), np.float32) if want_null else None
        nd = np.empty((len(lists), n_types, cnt), np.float32) if want_null else None
        self._check(self.L.pctsea_analyze_batch(self.h, m.handle, len(lists), _ptr(ptr), _ptr(gi), _ptr(ab),
                                                C.byref(sprm), float(min_score), _ptr(label), n_types,
                                                C.byref(nprm), _ptr(res), _ptr(npass), _ptr(status), _ptr(ne),
                                                _ptr(nd)))
        out = {"results": res, "n_pass": npass, "status": status}
        if want_null:
            out["null"], out["nullD"] = ne, nd
        return out

    def last_type_counts(self, n_types: int):
        """(numCellsOfType[T], numCellsOfTypePassing[T], N kept) of the last analyze — hypergeometric inputs."""
        a, b = np.zeros(n_types, np.int32), np.zeros(n_types, np.int32)
        n = C.c_int64()
        self._check(self.L.pctsea_last_type_counts(self.h, n_types, _ptr(a), _ptr(b), C.byref(n)))
        return a, b, int(n.value)

    def last_observed_curve(self, type_id: int, n_used: int):
        """a / b running-sum series of one cell type from the last analyze / enrich call (fp32 [n_used] each)."""
        a, b = np.empty(n_used, np.float32), np.empty(n_used, np.float32)
        self._check(self.L.pctsea_last_observed_curve(self.h, int(type_id), int(n_used), _ptr(a), _ptr(b)))
        return a, b

    def last_null_dev(self):
        a, b = C.c_void_p(), C.c_void_p()
        t, p = C.c_int32(), C.c_int32()
        self._check(self.L.pctsea_last_null_dev(self.h, C.byref(a), C.byref(b), C.byref(t), C.byref(p)))
        return int(a.value or 0), int(b.value or 0), t.value, p.value
