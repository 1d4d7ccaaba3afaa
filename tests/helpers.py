import numpy as np


def bits64(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def bits32(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bits_equal_f64(a, b, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"{what}: NaN pattern differs at {np.flatnonzero(na != nb)[:10]}"
    ok = bits64(a)[~na] == bits64(b)[~nb]
    assert ok.all(), f"{what}: {np.count_nonzero(~ok)} of {ok.size} values differ in bits; first {np.flatnonzero(~ok)[:5]}"


def assert_bits_equal_f32(a, b, what=""):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"{what}: NaN pattern differs"
    ok = bits32(a)[~na] == bits32(b)[~nb]
    assert ok.all(), f"{what}: {np.count_nonzero(~ok)} of {ok.size} values differ in bits"


RESULT_EXACT_F32 = ("ews", "dstat", "raw_sup", "norm_ews", "sec_ews")
RESULT_EXACT_F64 = ("norm_supX", "p_emp", "fdr", "ks_d")
RESULT_EXACT_INT = ("supX", "sizeA", "sizeB", "sec_supX")


def assert_results_equal(a, b, fields=None, what=""):
    for f in RESULT_EXACT_INT:
        if fields is None or f in fields:
            assert np.array_equal(a[f], b[f]), f"{what}.{f}: {a[f]} vs {b[f]}"
    for f in RESULT_EXACT_F32:
        if fields is None or f in fields:
            assert_bits_equal_f32(a[f], b[f], f"{what}.{f}")
    for f in RESULT_EXACT_F64:
        if fields is None or f in fields:
            assert_bits_equal_f64(a[f], b[f], f"{what}.{f}")


OBSERVED_FIELDS = ("ews", "dstat", "raw_sup", "supX", "norm_supX", "sizeA", "sizeB", "sec_ews", "sec_supX", "ks_d")
