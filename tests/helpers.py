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


# ---------------------------------------------------------------------------------------------------------------
# Hand-derived vectors (every value is a dyadic rational, so each float operation of the reference is exact and the
# expected results below were worked out on paper from the cited Java lines; both the oracle and the CUDA library
# are checked against them).
# ---------------------------------------------------------------------------------------------------------------
def hand_vector_secondary():
    """EnrichmentWeigthedScoreParallel.java:140-227 (main loop, compensation), :492-512 (lookForPriorNegativeSupremum,
    which keeps the LAST negative difference), :514-533 (lookForSecondaryEnrichmentIndex), :244-312 (second walk).
    11 ranked cells, type 0 = H, other = M, denomA = 16, denomB = 4:
        i      0     1     2     3     4    5     6     7     8    9    10
        cell   H     H     M     H     H    H     H     H     H    M    M
        score  2     2     2     2     2    2     2     2     2    1    1
        a     .125  .25   .25  .375   .5  .625  .75  .875   1    1    1
        b      0     0    .5    .5    .5   .5    .5   .5    .5  .75   1
        diff  .125  .25  -.25  -.125   0  .125  .25  .375   .5  .25   0
    supremum .5 at x = 9; negatives at i = 2, 3, the loop keeps the last one (-.125): ES = .5 - .125 = .375;
    going down from i = 8 the first positive difference that follows a negative one is i = 1 -> secondary index 1;
    the second walk over i = 0..1 gives a - b = .125, .25 -> secondary supremum .25 at x = 2.
    KS D of {2 x8} against {2, 1, 1}: |F_x - F_y| = 2/3 at z = 1 -> 16 / 24."""
    s = np.array([2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1], np.float64)
    lab = np.array([0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 1], np.int32)
    exp = {"ews": np.float32(0.375), "raw_sup": np.float32(0.5), "supX": 9, "norm_supX": 9.0 / 11.0, "sizeA": 8, "sizeB": 3,
           "dstat": np.float32(0.5) * np.float32(np.sqrt((8 * 3 * 1.0) / (1.0 * (8 + 3)))), "sec_ews": np.float32(0.25),
           "sec_supX": 2, "ks_d": 16.0 / 24.0}
    return s, lab, 2, exp


def hand_vector_secondary_index_zero():
    """The `secondaryEnrichmentIndex != 0` quirk (:246): the only positive -> negative transition sits at index 0, so
    the reference reports no secondary enrichment. 13 cells, denomA = 16, denomB = 8:
        cell   H     M     H    H     H    H     H    H     H    M    M    M     M
        score  2     2     2    2     2    2     2    2     2    2    2    1     1
        diff  .125 -.125   0  .125  .25  .375   .5  .625  .75   .5  .25  .125   0
    supremum .75 at x = 9, last negative -.125: ES = .625."""
    s = np.array([2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1], np.float64)
    lab = np.array([0, 1, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1], np.int32)
    exp = {"ews": np.float32(0.625), "raw_sup": np.float32(0.75), "supX": 9, "norm_supX": 9.0 / 13.0, "sizeA": 8, "sizeB": 5,
           "dstat": np.float32(0.75) * np.float32(np.sqrt((8 * 5 * 1.0) / (1.0 * (8 + 5)))), "sec_ews": np.float32(0.0),
           "sec_supX": 0}
    return s, lab, 2, exp


def hand_vector_fdr_duplicates():
    """a10-a12 with duplicate keys (PCTSEA.java:1449-1549, CellTypeClassification.java:407-441). T = 3, P = 4:
        ES = [2, 2, 1];  null0 = [2, 2, 0, 0] (zeros count as positives in a11, not in a10), null1 = [2, .5, .5, 1],
        null2 = [2, 2, -1, -1]
    expected |mean of the positives| = 1, 1, 2 -> normalised ES = [2, 2, .5]; normalised nulls [2,2,0,0], [2,.5,.5,1],
    [1,1] (negatives dropped). R sorted = [.5, 2, 2]; Z sorted = [0,0,.5,.5,1,1,1,2,2,2].
    key 2: Arrays.binarySearch(R) probes mid = 1 and returns 1 (not the last duplicate) -> sobs = 3 - 2 = 1; in Z it
    probes 4, then 7 and returns 7 -> snull = 10 - 8 = 2 -> fdr = (2 / 1) * (3 / 10) = 0.6 for BOTH types 0 and 1.
    key .5: R -> index 0, sobs = 2; Z probes 4, 1, 2 -> index 2, snull = 7 -> fdr = 3.5 * 0.3.
    a10: type 0: positives {2, 2}, none > 2 -> 0/2; type 1: 0/4; type 2: positives {2, 2}, both > 1 -> 2/2."""
    ews = np.array([2.0, 2.0, 1.0], np.float32)
    null = np.array([[2, 2, 0, 0], [2, .5, .5, 1], [2, 2, -1, -1]], np.float32)
    exp = {"norm_ews": np.array([2.0, 2.0, 0.5], np.float32), "p_emp": np.array([0.0, 0.0, 1.0]),
           "fdr": np.array([(1.0 * 2 / 1) * (1.0 * 3 / 10), (1.0 * 2 / 1) * (1.0 * 3 / 10), (1.0 * 7 / 2) * (1.0 * 3 / 10)])}
    return ews, null, exp


def hand_vectors_rank():
    """order_out of pctsea_rank / orc_rank = the list the KS stage walks, then the dropped cell. From
    ScoreThreshold.getSingleCellsPassingThresholdSortedByScore (scoring/ScoreThreshold.java:57-75), subList(0, size - 1)
    (PCTSEA.java:421-424) and the stable re-sort PCTSEAUtils.sortByScoreDescending (PCTSEA.java:2497):
    threshold < 0 keeps score < threshold, sorts ASCENDING by Double.compare, drops the last cell (the one closest to
    the threshold) and re-sorts the rest descending; threshold >= 0 keeps score >= threshold (-0.0 passes 0.0), sorts
    descending by Double.compare (0.0 before -0.0) and drops the last. Ties keep DB order in every sort.
      1. [-.5, .3, -.7, -.2, -.7, NaN, -.1], threshold -.15: ascending -.7(2) -.7(4) -.5(0) -.2(3); drop 3;
         descending: -.5(0) -.7(2) -.7(4)
      2. [-0.0, 0.0, .5, 0.0], threshold 0: .5(2) 0.0(1) 0.0(3) -0.0(0); drop 0
      3. [.25, NaN, .2, .25, .1999], threshold .2: .25(0) .25(3) .2(2); drop 2"""
    nan = float("nan")
    return [
        (np.array([-0.5, 0.3, -0.7, -0.2, -0.7, nan, -0.1]), -0.15, [0, 2, 4, 3]),
        (np.array([-0.0, 0.0, 0.5, 0.0]), 0.0, [2, 1, 3, 0]),
        (np.array([0.25, nan, 0.2, 0.25, 0.1999]), 0.2, [0, 3, 2]),
    ]
