"""ctypes wrapper over oracle/libpctsea_oracle.so — the CPU restatement of the PCTSEA hot path.

TEST INFRASTRUCTURE ONLY (see oracle/pctsea_oracle.h). PARITY UNPINNED: the reference has no numeric tests
and cannot run here. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under pctsea_parent_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpctsea_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pctsea_oracle.c")
    hdr = os.path.join(_HERE, "pctsea_oracle.h")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpctsea_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


class ScoreParams(C.Structure):
    _fields_ = [("method", C.c_int32), ("min_genes_cells", C.c_int32), ("min_corr", C.c_double),
                ("has_min_corr", C.c_int32), ("jitter_mode", C.c_int32), ("seed", C.c_uint64)]


class TypeResult(C.Structure):
    _fields_ = [("ews", C.c_float), ("dstat", C.c_float), ("supX", C.c_int32), ("norm_supX", C.c_double),
                ("sizeA", C.c_int32), ("sizeB", C.c_int32), ("raw_sup", C.c_float), ("norm_ews", C.c_float),
                ("p_emp", C.c_double), ("fdr", C.c_double), ("sec_ews", C.c_float), ("sec_supX", C.c_int32),
                ("ks_d", C.c_double)]


TYPE_RESULT_DTYPE = np.dtype([("ews", "<f4"), ("dstat", "<f4"), ("supX", "<i4"), ("_pad0", "<i4"),
                              ("norm_supX", "<f8"), ("sizeA", "<i4"), ("sizeB", "<i4"), ("raw_sup", "<f4"),
                              ("norm_ews", "<f4"), ("p_emp", "<f8"), ("fdr", "<f8"), ("sec_ews", "<f4"),
                              ("sec_supX", "<i4"), ("ks_d", "<f8")])
assert TYPE_RESULT_DTYPE.itemsize == C.sizeof(TypeResult)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_jrandom_next_int.restype = C.c_int32
        L.orc_jrandom_next_int_bound.restype = C.c_int32
        L.orc_jrandom_next_double.restype = C.c_double
        L.orc_java_binary_search_f32.restype = C.c_int32
        L.orc_java_double_compare.restype = C.c_int
        L.orc_jitter_uniform.restype = C.c_double
        L.orc_jitter_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_pearson.restype = C.c_double
        L.orc_pearson_strided32.restype = C.c_double
        L.orc_pearson_shifted.restype = C.c_double
        L.orc_ks_statistic.restype = C.c_double
        L.orc_score2.restype = C.c_int
        L.orc_rank.restype = C.c_int64
        L.orc_score.restype = C.c_int
        L.orc_fast_frac_bits.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class JRandom:
    """java.util.Random"""

    class _S(C.Structure):
        _fields_ = [("seed", C.c_uint64)]

    def __init__(self, seed: int):
        self._s = JRandom._S()
        lib().orc_jrandom_init(C.byref(self._s), C.c_int64(seed))

    def next_int(self, bound: int | None = None) -> int:
        if bound is None:
            return lib().orc_jrandom_next_int(C.byref(self._s))
        return lib().orc_jrandom_next_int_bound(C.byref(self._s), C.c_int32(bound))

    def next_double(self) -> float:
        return lib().orc_jrandom_next_double(C.byref(self._s))

    def shuffle(self, a: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.int32).copy()
        lib().orc_java_shuffle_i32(_p(a, C.c_int32), C.c_int64(a.size), C.byref(self._s))
        return a


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def jitter_uniform(seed, cell, vec, i):
    return lib().orc_jitter_uniform(seed, cell, vec, i)


DEFAULT_FEISTEL_ROUNDS = 5   # PCTSEA_DEFAULT_FEISTEL_ROUNDS of include/pctsea_b200.h


def feistel_perm(seed: int, p: int, n: int, rounds: int = DEFAULT_FEISTEL_ROUNDS) -> np.ndarray:
    out = np.empty(max(n, 0), dtype=np.int32)
    lib().orc_feistel_perm(C.c_uint64(seed), C.c_uint32(p), C.c_int64(n), C.c_int(rounds), _p(out, C.c_int32))
    return out


def perm_labels(seed: int, p: int, lab, T: int, rounds: int = DEFAULT_FEISTEL_ROUNDS) -> np.ndarray:
    """Permuted labels of PHILOX mode: the size-sorted arrangement of `lab` read through the Feistel bijection of
    permutation p (global index)."""
    lab = np.ascontiguousarray(lab, dtype=np.int32)
    out = np.empty(lab.size, dtype=np.int32)
    lib().orc_perm_labels(C.c_uint64(seed), C.c_uint32(p), C.c_int64(lab.size), _p(lab, C.c_int32), C.c_int32(T),
                          C.c_int(rounds), _p(out, C.c_int32))
    return out


def binary_search_f32(a: np.ndarray, key: float) -> int:
    a = np.ascontiguousarray(a, dtype=np.float32)
    return lib().orc_java_binary_search_f32(_p(a, C.c_float), C.c_int32(a.size), C.c_float(key))


def java_sort_f32(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32).copy()
    lib().orc_java_sort_f32(_p(a, C.c_float), C.c_int64(a.size))
    return a


def double_compare(a: float, b: float) -> int:
    return lib().orc_java_double_compare(C.c_double(a), C.c_double(b))


def pearson(x, y) -> float:
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    return lib().orc_pearson(_p(x, C.c_double), _p(y, C.c_double), C.c_int32(x.size))


def ks_statistic(x, y) -> float:
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    return lib().orc_ks_statistic(_p(x, C.c_double), C.c_int64(x.size), _p(y, C.c_double), C.c_int64(y.size))


def pearson_shifted(x, y) -> float:
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    return lib().orc_pearson_shifted(_p(x, C.c_double), _p(y, C.c_double), C.c_int32(x.size))


def pearson_strided32(x, y) -> float:
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    return lib().orc_pearson_strided32(_p(x, C.c_double), _p(y, C.c_double), C.c_int32(x.size))


def score(row_ptr, col, val, n_genes, gene_idx, abundance, method=0, min_genes_cells=4, min_corr=0.2,
          has_min_corr=True, jitter_mode=1, seed=0, canonical=True):
    """canonical=True (or 1): the CSR-scan kernel's reduction order; canonical=2: the gene-major kernel's (pairs
    in ascending gene order, shifted sequential sums) — the GPU is bit-identical to the one of the path it runs;
    canonical=False: the reference-faithful sequential SimpleRegression update."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float32)
    gene_idx = np.ascontiguousarray(gene_idx, dtype=np.int32)
    abundance = np.ascontiguousarray(abundance, dtype=np.float32)
    n = row_ptr.size - 1
    prm = ScoreParams(method, min_genes_cells, float(min_corr), int(bool(has_min_corr)), jitter_mode, seed)
    sc = np.empty(n, dtype=np.float64)
    used = np.empty(n, dtype=np.int32)
    nz = np.empty(n, dtype=np.int32)
    kept = np.empty(n, dtype=np.uint8)
    rc = lib().orc_score2(C.c_int64(n), C.c_int32(n_genes), _p(row_ptr, C.c_int64), _p(col, C.c_int32),
                          _p(val, C.c_float), C.c_int32(gene_idx.size), _p(gene_idx, C.c_int32),
                          _p(abundance, C.c_float), C.byref(prm), C.c_int(int(canonical)),
                          _p(sc, C.c_double), _p(used, C.c_int32), _p(nz, C.c_int32), _p(kept, C.c_uint8))
    if rc != 0:
        raise ValueError(f"orc_score rc={rc}")
    return sc, used, nz, kept


def rank(score_arr, kept, min_score=0.0):
    score_arr = np.ascontiguousarray(score_arr, dtype=np.float64)
    n = score_arr.size
    kept_p = None
    if kept is not None:
        kept = np.ascontiguousarray(kept, dtype=np.uint8)
        kept_p = _p(kept, C.c_uint8)
    order = np.empty(max(n, 1), dtype=np.int32)
    npass = lib().orc_rank(C.c_int64(n), _p(score_arr, C.c_double), kept_p, C.c_double(min_score),
                           _p(order, C.c_int32))
    return order[:npass].copy(), int(npass)


def ks_observed(s, lab, T):
    s = np.ascontiguousarray(s, dtype=np.float64)
    lab = np.ascontiguousarray(lab, dtype=np.int32)
    out = np.zeros(T, dtype=TYPE_RESULT_DTYPE)
    lib().orc_ks_observed(C.c_int64(s.size), _p(s, C.c_double), _p(lab, C.c_int32), C.c_int32(T),
                          out.ctypes.data_as(C.c_void_p))
    return out


def null(s, lab, T, P, perm_mode=0, seed=0, perm_begin=0, rounds=DEFAULT_FEISTEL_ROUNDS, nthreads=1, want_replay=False):
    s = np.ascontiguousarray(s, dtype=np.float64)
    lab = np.ascontiguousarray(lab, dtype=np.int32)
    nu = s.size
    ne = np.empty((T, P), dtype=np.float32)
    nd = np.empty((T, P), dtype=np.float32)
    rep = np.empty((P, nu), dtype=np.int32) if want_replay else None
    lib().orc_null(C.c_int64(nu), _p(s, C.c_double), _p(lab, C.c_int32), C.c_int32(T), C.c_int32(P),
                   C.c_int(perm_mode), C.c_uint64(seed), C.c_int32(perm_begin), C.c_int(rounds),
                   C.c_int(nthreads), _p(rep, C.c_int32) if rep is not None else None,
                   _p(ne, C.c_float), _p(nd, C.c_float))
    return (ne, nd, rep) if want_replay else (ne, nd)


def null_replay(s, lab, T, replay):
    s = np.ascontiguousarray(s, dtype=np.float64)
    lab = np.ascontiguousarray(lab, dtype=np.int32)
    replay = np.ascontiguousarray(replay, dtype=np.int32)
    P = replay.shape[0]
    ne = np.empty((T, P), dtype=np.float32)
    nd = np.empty((T, P), dtype=np.float32)
    lib().orc_null_replay(C.c_int64(s.size), _p(s, C.c_double), _p(lab, C.c_int32), C.c_int32(T),
                          C.c_int32(P), _p(replay, C.c_int32), _p(ne, C.c_float), _p(nd, C.c_float))
    return ne, nd


def reduce(results, null_ews, mean_policy=0, want_norm_null=False):
    res = np.array(results, dtype=TYPE_RESULT_DTYPE, copy=True)
    null_ews = np.ascontiguousarray(null_ews, dtype=np.float32)
    T, P = null_ews.shape
    nn = np.empty((T, P), dtype=np.float32) if want_norm_null else None
    lib().orc_reduce(C.c_int32(T), C.c_int32(P), _p(null_ews, C.c_float), C.c_int(mean_policy),
                     res.ctypes.data_as(C.c_void_p), _p(nn, C.c_float) if nn is not None else None)
    return (res, nn) if want_norm_null else res


def fast_frac_bits(s) -> int:
    s = np.ascontiguousarray(s, dtype=np.float64)
    return lib().orc_fast_frac_bits(C.c_int64(s.size), _p(s, C.c_double))


def ks_null_fast_one(s, lab_perm, T, frac_bits=None):
    s = np.ascontiguousarray(s, dtype=np.float64)
    lab_perm = np.ascontiguousarray(lab_perm, dtype=np.int32)
    if frac_bits is None:
        frac_bits = fast_frac_bits(s)
    out = np.empty(T, dtype=np.float32)
    lib().orc_ks_null_fast_one(C.c_int64(s.size), _p(s, C.c_double), _p(lab_perm, C.c_int32), C.c_int32(T),
                               C.c_int(frac_bits), _p(out, C.c_float))
    return out


def ks_null_one(s, lab_perm, T):
    s = np.ascontiguousarray(s, dtype=np.float64)
    lab_perm = np.ascontiguousarray(lab_perm, dtype=np.int32)
    ne = np.empty(T, dtype=np.float32)
    nd = np.empty(T, dtype=np.float32)
    lib().orc_ks_null_one(C.c_int64(s.size), _p(s, C.c_double), _p(lab_perm, C.c_int32), C.c_int32(T),
                          _p(ne, C.c_float), _p(nd, C.c_float))
    return ne, nd
