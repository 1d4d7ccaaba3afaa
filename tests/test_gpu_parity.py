"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): scores <= 1e-9 relative (here: bit-exact, canonical pair order), ranking and
observed enrichment bit-exact, REPLAY null bit-exact, PHILOX null bit-exact against the oracle's restatement
of the same keyed permutation, FAST arithmetic within a stated tolerance of EXACT, p/FDR exact given the null.
"""
import numpy as np
import pytest

import pctsea_parent_b200 as P
from pctsea_parent_b200 import synth
from tests.helpers import (OBSERVED_FIELDS, assert_bits_equal_f32, assert_bits_equal_f64, assert_results_equal)

pytestmark = pytest.mark.gpu


def _load(ctx, d):
    m = ctx.load_csr(d["row_ptr"], d["col"], d["val"], d["cfg"].n_genes)
    m.set_labels(d["label"], d["cfg"].n_types)
    return m


# ---------------------------------------------------------------------------------------------------
# stage 1
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("jitter", [0, 1])
def test_score_bit_exact(ctx, oracle, tiny, method, jitter):
    d = tiny
    m = _load(ctx, d)
    kw = dict(method=method, min_genes_cells=4, min_corr=0.2 if method == 0 else -1.0, has_min_corr=True,
              jitter_mode=jitter, seed=1234)
    g = ctx.score(m, d["gene_idx"], d["abundance"], **kw)
    o = oracle.score(d["row_ptr"], d["col"], d["val"], d["cfg"].n_genes, d["gene_idx"], d["abundance"], **kw)
    assert np.array_equal(g[1], o[1]), "n_genes_used"
    assert np.array_equal(g[2], o[2]), "n_nonzero"
    assert np.array_equal(g[3], o[3]), "kept"
    assert_bits_equal_f64(g[0], o[0], "score")
    # north_star tolerance, stated: <= 1e-9 relative against the reference-faithful sequential
    # SimpleRegression update (the bit-exact check above is against the canonical two-pass strided order).
    # Zero-variance cells are excluded from the relative bar: the reference adds unseeded Math.random()/1e6
    # to them (model/SingleCell.java:433-448), their correlation is noise of relative condition ~1e7, and
    # any two summation orders differ there by ~1e-5 relative / ~1e-9 absolute on the same jitter stream.
    f = oracle.score(d["row_ptr"], d["col"], d["val"], d["cfg"].n_genes, d["gene_idx"], d["abundance"],
                     canonical=False, **kw)
    assert np.array_equal(np.isnan(f[0]), np.isnan(g[0])), "NaN pattern vs sequential oracle"
    nj = oracle.score(d["row_ptr"], d["col"], d["val"], d["cfg"].n_genes, d["gene_idx"], d["abundance"],
                      canonical=False, method=0, min_genes_cells=4, min_corr=-1.0, jitter_mode=0, seed=1234)
    zero_var = (nj[3] == 1) & (nj[1] >= 2) & np.isnan(nj[0])  # kept, scorable, but NaN without jitter
    ok = ~np.isnan(f[0])
    regular = ok & ~zero_var
    assert regular.sum() > 100
    assert np.all(np.abs(g[0][regular] - f[0][regular]) <= 1e-9 * np.maximum(1.0, np.abs(f[0][regular])))
    assert np.all(np.abs(g[0][ok] - f[0][ok]) <= 1e-7)
    m.free()


def test_score_edge_cases(ctx, oracle):
    # empty rows, a row with NaN / zero / negative values, list genes absent from the matrix (-1),
    # non-positive abundance, zero-variance cells, a cell with exactly min_genes pairs
    G = 50
    rows = [
        [],                                                        # empty cell
        [(1, 1.0), (2, 1.0), (3, 1.0), (4, 1.0)],                 # zero variance in the cell
        [(1, 2.0), (2, 5.0), (3, 1.0), (4, 7.0), (9, 3.0)],       # regular
        [(1, float("nan")), (2, 5.0), (3, 0.0), (4, -2.0), (5, 1.0), (6, 2.0), (7, 9.0)],
        [(1, 3.0), (2, 1.0)],                                      # fewer than min_genes pairs
        [(10, 4.0), (11, 4.0), (12, 2.0), (13, 8.0)],             # list genes 10..13 share one abundance
        [(40, 1.0), (41, 2.0)],                                    # no list gene at all
    ]
    row_ptr = np.cumsum([0] + [len(r) for r in rows]).astype(np.int64)
    col = np.array([c for r in rows for c, _ in r], np.int32)
    val = np.array([v for r in rows for _, v in r], np.float32)
    gene_idx = np.array([1, 2, 3, 4, 5, 6, 7, 9, -1, 10, 11, 12, 13, 20], np.int32)
    abundance = np.array([3, 1, 4, 1.5, 9, 2, 6, 5, 7, 2, 2, 2, 2, 0.0], np.float32)
    m = ctx.load_csr(row_ptr, col, val, G)
    for method in (0, 1):
        for jitter in (0, 1):
            kw = dict(method=method, min_genes_cells=3, min_corr=-1.0, has_min_corr=True, jitter_mode=jitter, seed=7)
            g = ctx.score(m, gene_idx, abundance, **kw)
            o = oracle.score(row_ptr, col, val, G, gene_idx, abundance, **kw)
            assert np.array_equal(g[1], o[1]) and np.array_equal(g[2], o[2]) and np.array_equal(g[3], o[3])
            assert_bits_equal_f64(g[0], o[0], f"score m{method} j{jitter}")
    # min_genes_cells > L raises like model/SingleCell.java:347-352
    with pytest.raises(P.PctseaError) as e:
        ctx.score(m, gene_idx, abundance, min_genes_cells=gene_idx.size + 1)
    assert e.value.code == -5
    # duplicate matrix column in the list is rejected
    with pytest.raises(P.PctseaError):
        ctx.score(m, np.array([1, 1, 2, 3], np.int32), np.ones(4, np.float32), min_genes_cells=1)
    m.free()


def test_score_long_list_small_buffer(ctx, oracle):
    # dense cells against a long list: forces the flush-and-rescan path of the pair buffer
    rng = np.random.default_rng(5)
    N, G, L = 300, 3000, 2500
    dense = rng.random((N, G)) < 0.6
    vals = rng.integers(1, 9, size=(N, G)).astype(np.float32) * dense
    row_ptr = np.zeros(N + 1, np.int64)
    row_ptr[1:] = np.cumsum(dense.sum(1))
    r, c = np.nonzero(dense)
    col, val = c.astype(np.int32), vals[r, c]
    gene_idx = rng.permutation(G)[:L].astype(np.int32)
    abundance = rng.integers(1, 50, L).astype(np.float32)
    m = ctx.load_csr(row_ptr, col, val, G)
    g = ctx.score(m, gene_idx, abundance, min_genes_cells=10, min_corr=-1.0)
    o = oracle.score(row_ptr, col, val, G, gene_idx, abundance, min_genes_cells=10, min_corr=-1.0)
    assert np.array_equal(g[1], o[1])
    assert_bits_equal_f64(g[0], o[0], "score")
    m.free()


def _random_csr(rng, N, G, mean_nnz, unsorted_rows=False):
    nnz_c = np.clip(rng.poisson(mean_nnz, N), 0, G)
    nnz_c[rng.integers(0, N, max(1, N // 50))] = 0            # a few empty rows
    row_ptr = np.zeros(N + 1, np.int64)
    row_ptr[1:] = np.cumsum(nnz_c)
    col = np.empty(int(row_ptr[-1]), np.int32)
    for i in range(N):
        c = rng.choice(G, int(nnz_c[i]), replace=False)
        col[row_ptr[i]:row_ptr[i + 1]] = c if unsorted_rows else np.sort(c)
    val = rng.integers(1, 12, col.size).astype(np.float32)
    return row_ptr, col, val


@pytest.mark.parametrize("G,L,unsorted_rows", [
    (120_000, 900, False),   # gene table too large for one byte per gene: two-bit membership table
    (6_000, 5_000, True),    # list longer than the on-chip rank directory: abundances gathered from L2
    (40_000, 64, False),     # short list, byte table, many warps per CTA
])
def test_score_kernel_table_variants(ctx, oracle, G, L, unsorted_rows):
    rng = np.random.default_rng(G + L)
    N = 1500
    row_ptr, col, val = _random_csr(rng, N, G, 700 if L >= 900 else 4000, unsorted_rows)
    # bias the list towards genes that occur, so that cells reach min_genes_cells
    present = np.unique(col)
    gene_idx = rng.permutation(present)[:L].astype(np.int32) if present.size >= L else rng.permutation(G)[:L].astype(np.int32)
    gene_idx[rng.integers(0, L, 3)] = -1                       # proteins without a matrix gene
    abundance = rng.integers(0, 40, L).astype(np.float32)      # includes zero abundances
    m = ctx.load_csr(row_ptr, col, val, G)
    for method, jitter in ((0, 1), (1, 0)):
        kw = dict(method=method, min_genes_cells=3, min_corr=-1.0, has_min_corr=True, jitter_mode=jitter, seed=99)
        g = ctx.score(m, gene_idx, abundance, **kw)
        o = oracle.score(row_ptr, col, val, G, gene_idx, abundance, **kw)
        assert np.array_equal(g[1], o[1]), "n_genes_used"
        assert np.array_equal(g[2], o[2]), "n_nonzero"
        assert np.array_equal(g[3], o[3]), "kept"
        assert_bits_equal_f64(g[0], o[0], "score")
        assert np.isfinite(g[0]).sum() > 20
    m.free()


def test_score_row_alignment_and_tails(ctx, oracle):
    # rows of every length 0..300 back to back: every (row begin mod 4, row length mod 128) combination that the
    # 128-bit head / tail handling of the scoring kernel sees, including the very last quad of the arrays
    rng = np.random.default_rng(11)
    G = 2000
    lens = np.concatenate([np.arange(0, 301), rng.integers(0, 700, 400)])
    N = lens.size
    row_ptr = np.zeros(N + 1, np.int64)
    row_ptr[1:] = np.cumsum(lens)
    col = np.empty(int(row_ptr[-1]), np.int32)
    for i in range(N):
        col[row_ptr[i]:row_ptr[i + 1]] = np.sort(rng.choice(G, int(lens[i]), replace=False))
    val = rng.integers(1, 6, col.size).astype(np.float32)
    gene_idx = rng.permutation(G)[:600].astype(np.int32)
    abundance = rng.integers(1, 30, 600).astype(np.float32)
    for trim in (0, 1, 2, 3):   # move the end of the arrays through all four alignments
        n_rows = N - trim
        rp = row_ptr[:n_rows + 1]
        nnz = int(rp[-1]) - trim if rp[-1] - rp[-2] >= trim else int(rp[-1])
        rp = rp.copy(); rp[-1] = nnz
        m = ctx.load_csr(rp, col[:nnz], val[:nnz], G)
        g = ctx.score(m, gene_idx, abundance, min_genes_cells=2, min_corr=-1.0, seed=3)
        o = oracle.score(rp, col[:nnz], val[:nnz], G, gene_idx, abundance, min_genes_cells=2, min_corr=-1.0, seed=3)
        assert np.array_equal(g[1], o[1]) and np.array_equal(g[2], o[2]) and np.array_equal(g[3], o[3])
        assert_bits_equal_f64(g[0], o[0], f"score trim{trim}")
        m.free()


def test_matrix_validation(ctx):
    row_ptr = np.array([0, 2, 4], np.int64)
    val = np.ones(4, np.float32)
    with pytest.raises(P.PctseaError) as e:   # column outside [0, n_genes)
        ctx.load_csr(row_ptr, np.array([0, 1, 2, 7], np.int32), val, 5)
    assert e.value.code == -1
    with pytest.raises(P.PctseaError) as e:   # negative column
        ctx.load_csr(row_ptr, np.array([0, 1, -2, 3], np.int32), val, 5)
    assert e.value.code == -1
    with pytest.raises(P.PctseaError) as e:   # row_ptr not monotone
        ctx.load_csr(np.array([0, 3, 2], np.int64), np.array([0, 1, 2], np.int32), val[:3], 5)
    assert e.value.code == -1
    # a row that stores a column twice (adjacent, or apart in an unsorted row): the reference holds one value per
    # (cell, gene), model/SingleCell.java:80, and the two scoring kernels would disagree on such a row
    for cols in ([0, 1, 3, 3], [0, 4, 1, 4]):
        with pytest.raises(P.PctseaError) as e:
            ctx.load_csr(np.array([0, 1, 4], np.int64), np.array(cols, np.int32), val, 5)
        assert e.value.code == -1 and "twice" in e.value.msg
    rng = np.random.default_rng(5)
    rp, col, v = _random_csr(rng, 300, 70_000, 40, unsorted_rows=True)   # unsorted rows, many genes: accepted
    ctx.load_csr(rp, col, v, 70_000).free()
    col2 = col.copy()
    col2[rp[200] + 3] = col2[rp[200]]   # one repeated column deep inside the matrix
    if rp[201] - rp[200] > 3:
        with pytest.raises(P.PctseaError):
            ctx.load_csr(rp, col2, v, 70_000)
    m = ctx.load_csr(row_ptr, np.array([0, 1, 2, 4], np.int32), val, 5)   # the context still works afterwards
    m.free()


# ---------------------------------------------------------------------------------------------------
# stage 2
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("thr", [0.0, 0.2, 0.5, -0.1])
def test_rank_exact(ctx, oracle, thr):
    rng = np.random.default_rng(11)
    n = 50_000
    s = np.round(rng.uniform(-1, 1, n), 3)  # heavy ties
    s[rng.random(n) < 0.05] = np.nan
    s[:7] = [0.0, -0.0, 0.0, -0.0, 0.5, 0.5, 0.2]
    kept = (rng.random(n) < 0.9).astype(np.uint8)
    go, gn = ctx.rank(s, kept, thr)
    oo, on = oracle.rank(s, kept, thr)
    assert gn == on
    assert np.array_equal(go, oo)


def test_rank_edge_cases(ctx, oracle):
    for s in (np.array([]), np.array([np.nan, np.nan]), np.array([0.3]), np.full(5000, 0.25),
              np.array([1e-300, 5e-324, 0.0, -0.0, 1e300, np.inf])):
        go, gn = ctx.rank(s, None, 0.0)
        oo, on = oracle.rank(s, None, 0.0)
        assert gn == on and np.array_equal(go, oo)


def test_rank_sortedness_large(ctx):
    # size-independent property at a BASELINE-scale N: order is a permutation of the passing cells and
    # (score desc, index asc)
    rng = np.random.default_rng(3)
    n = 2_000_000
    s = rng.uniform(0, 1, n)
    s[rng.random(n) < 0.3] = np.nan
    o, npass = ctx.rank(s, None, 0.2)
    assert npass == int(np.count_nonzero(s >= 0.2))
    so = s[o]
    assert np.all(so[:-1] >= so[1:])
    tie = so[:-1] == so[1:]
    assert np.all(o[:-1][tie] < o[1:][tie])
    assert np.unique(o).size == npass


# ---------------------------------------------------------------------------------------------------
# stage 3
# ---------------------------------------------------------------------------------------------------
def _ranked(n_used, T, seed=0, simple=False):
    s, lab = synth.synthetic_ranked(n_used, T, seed, simple)
    order = np.arange(n_used, dtype=np.int32)
    return s, lab, order


@pytest.mark.parametrize("n_used,T", [(5000, 12), (20_000, 40), (3000, 300), (64, 3)])
def test_observed_bit_exact(ctx, oracle, n_used, T):
    s, lab, order = _ranked(n_used, T, seed=n_used)
    prm = ctx.null_params(0, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_EXACT)
    res, _, _ = ctx.enrich(s, order, lab, T, prm)
    ores = oracle.ks_observed(s, lab, T)
    assert_results_equal(res, ores, OBSERVED_FIELDS, "observed")


def test_hand_derived_vectors(ctx):
    """The CUDA library against results worked out on paper from the reference's lines (tests/helpers.py)."""
    from tests import helpers as H
    from tests.test_oracle_cpu import _check_observed
    prm = ctx.null_params(0, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_EXACT)
    for make in (H.hand_vector_secondary, H.hand_vector_secondary_index_zero):
        s, lab, T, exp = make()
        res, _, _ = ctx.enrich(s, np.arange(s.size, dtype=np.int32), lab, T, prm)
        _check_observed(res, 0, exp, make.__name__)
    ews, null, exp = H.hand_vector_fdr_duplicates()
    res = np.zeros(3, P.TYPE_RESULT_DTYPE)
    res["ews"] = ews
    red = ctx.reduce_null(res, null)
    assert np.array_equal(red["norm_ews"], exp["norm_ews"]) and np.array_equal(red["p_emp"], exp["p_emp"])
    assert np.array_equal(red["fdr"], exp["fdr"]), red["fdr"]
    for score, thr, order in H.hand_vectors_rank():
        got, npass = ctx.rank(score, None, thr)
        assert npass == len(order) and got.tolist() == order, (thr, got)


def test_observed_survey_micro_vector(ctx):
    # SURVEY.md Appendix B.1 (hand-derived bit patterns)
    s = np.array([0.9, 0.85, 0.8, 0.75, 0.7, 0.65, 0.6, 0.55, 0.5, 0.45])
    lab = np.array([1, 0, 0, 0, 1, 1, 1, 1, 1, 1], np.int32)
    res, _, _ = ctx.enrich(s, np.arange(10, dtype=np.int32), lab, 2, ctx.null_params(0, ks_mode=P.KS_EXACT))
    f = lambda x: np.float32(x).view(np.uint32)
    assert f(res["raw_sup"][0]) == 0x3F4B08D4 and res["supX"][0] == 4
    assert f(res["ews"][0]) == 0x3F1611A8 and f(res["dstat"][0]) == 0x3F931CCA
    assert f(res["ews"][1]) == 0xBFCB08D4 and f(res["dstat"][1]) == 0xBF931CCA
    assert (res["sizeA"][0], res["sizeB"][0]) == (3, 7)


def test_observed_int32_overflow_dstat(ctx, oracle):
    # sizeA*sizeB >= 2^31 wraps negative in the reference -> sqrt(negative) = NaN dStat (a8 step 4)
    n_used, T = 100_000, 2
    rng = np.random.default_rng(1)
    s = np.sort(rng.uniform(0.2, 1, n_used))[::-1].copy()
    lab = (rng.random(n_used) < 0.5).astype(np.int32)
    res, _, _ = ctx.enrich(s, np.arange(n_used, dtype=np.int32), lab, T, ctx.null_params(0, ks_mode=P.KS_EXACT))
    ores = oracle.ks_observed(s, lab, T)
    assert np.isnan(ores["dstat"]).all()
    assert_results_equal(res, ores, OBSERVED_FIELDS, "overflow")


@pytest.mark.parametrize("n_used,T,Pn", [(4000, 12, 24), (1500, 300, 6)])
def test_null_replay_bit_exact(ctx, oracle, n_used, T, Pn):
    s, lab, order = _ranked(n_used, T, seed=1)
    one, ond, rep = oracle.null(s, lab, T, Pn, perm_mode=0, seed=42, want_replay=True)
    prm = ctx.null_params(Pn, perm_mode=P.PERM_REPLAY, ks_mode=P.KS_EXACT)
    res, ne, nd = ctx.enrich(s, order, lab, T, prm, replay=rep)
    assert_bits_equal_f32(ne, one, "null ews")
    assert_bits_equal_f32(nd, ond, "null dstat")
    ored = oracle.reduce(oracle.ks_observed(s, lab, T), one)
    assert_results_equal(res, ored, None, "replay end-to-end")


@pytest.mark.parametrize("n_used,T,Pn", [(5000, 12, 16), (2500, 300, 5), (777, 5, 9)])
def test_null_philox_exact_bit_exact(ctx, oracle, n_used, T, Pn):
    s, lab, order = _ranked(n_used, T, seed=2)
    one, ond = oracle.null(s, lab, T, Pn, perm_mode=1, seed=99)
    prm = ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_EXACT, seed=99)
    res, ne, nd = ctx.enrich(s, order, lab, T, prm)
    assert_bits_equal_f32(ne, one, "null ews")
    assert_bits_equal_f32(nd, ond, "null dstat")


@pytest.mark.parametrize("n_used,T,Pn,simple", [(5000, 12, 16, False), (2500, 300, 5, False), (9000, 40, 8, True),
                                                 (140_000, 20, 3, False),
                                                 # more types than one pass of the lane kernel holds: type slices,
                                                 # with 8-bit (T <= 254) and 16-bit labels
                                                 (6000, 230, 3, False), (7000, 300, 3, False), (9000, 1000, 2, True),
                                                 (5000, 218, 3, False)])
def test_null_philox_fast_matches_oracle_fast(ctx, oracle, n_used, T, Pn, simple):
    s, lab, order = _ranked(n_used, T, seed=3, simple=simple)
    prm = ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=5)
    res, ne, nd = ctx.enrich(s, order, lab, T, prm)
    F = oracle.fast_frac_bits(s)
    for p in range(Pn):
        exp = oracle.ks_null_fast_one(s, oracle.perm_labels(5, p, lab, T), T, F)
        assert_bits_equal_f32(ne[:, p], exp, f"fast null perm {p}")


def _label_layouts():
    rng = np.random.default_rng(17)
    out = {}
    out["37 cells, 5 types"] = (rng.integers(0, 5, 37).astype(np.int32), 5)
    out["4 cells, 2 types"] = (np.array([0, 1, 1, 0], np.int32), 2)
    out["one type owns every cell"] = (np.full(5000, 2, np.int32), 6)
    lab = np.concatenate([np.arange(40), np.full(1400, 47), np.full(2000, 44), np.full(1560, -1)]).astype(np.int32)
    out["forty one-cell types, two big ones, 31 % unknown"] = (rng.permutation(lab), 50)
    w = 1.0 / np.arange(1, 121) ** 2
    sizes = np.maximum(3, np.floor(w / w.sum() * (70_000 - 3 * 120)).astype(np.int64))
    lab = np.repeat(np.arange(120, dtype=np.int32), sizes)[:70_000]
    lab = np.concatenate([lab, np.zeros(70_000 - lab.size, np.int32)])
    out["long tail: 120 types, sizes ~ 1 / k^2, floor 3"] = (rng.permutation(lab), 120)
    lab = np.repeat(np.arange(400, dtype=np.int32), 25)
    out["400 types of 25 cells (sliced, every class narrower than a bucket)"] = (rng.permutation(lab), 400)
    return out


@pytest.mark.parametrize("layout", list(_label_layouts()))
def test_null_class_table_edge_cases(ctx, oracle, layout):
    """The class of a permuted index comes from a two-region bucket table of the size-sorted classes whose split is
    chosen on the device: label layouts that stress it (tiny rankings, empty and one-cell classes, a single class,
    long tails, more tiny classes than table entries), FAST and EXACT arithmetic against the oracle."""
    lab, T = _label_layouts()[layout]
    n_used = lab.size
    s = synth.synthetic_ranked(n_used, max(T, 2), seed=5)[0]
    order = np.arange(n_used, dtype=np.int32)
    Pn = 3
    res, ne, nd = ctx.enrich(s, order, lab, T, ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=13))
    F = oracle.fast_frac_bits(s)
    for p in range(Pn):
        exp = oracle.ks_null_fast_one(s, oracle.perm_labels(13, p, lab, T), T, F)
        assert_bits_equal_f32(ne[:, p], exp, f"{layout}: fast null perm {p}")
    one, _ = oracle.null(s, lab, T, Pn, perm_mode=1, seed=13)
    ex = ctx.enrich(s, order, lab, T, ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_EXACT, seed=13))[1]
    assert_bits_equal_f32(ex, one, f"{layout}: exact null")
    assert_results_equal(res, oracle.ks_observed(s, lab, T), OBSERVED_FIELDS, f"{layout}: observed")


@pytest.mark.parametrize("rounds", [2, 4, 6, 9])
@pytest.mark.parametrize("n_used,T", [(5000, 12), (9000, 300), (500, 4)])
def test_null_philox_explicit_round_counts(ctx, oracle, rounds, n_used, T):
    """feistel_rounds other than the default (5, unrolled): the generic-loop instantiation of the null kernel with 8- and
    16-bit labels, and the row-major label kernel of the EXACT null."""
    s, lab, order = _ranked(n_used, T, seed=8)
    prm = ctx.null_params(3, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=31, feistel_rounds=rounds)
    res, ne, nd = ctx.enrich(s, order, lab, T, prm)
    F = oracle.fast_frac_bits(s)
    for p in range(3):
        exp = oracle.ks_null_fast_one(s, oracle.perm_labels(31, p, lab, T, rounds), T, F)
        assert_bits_equal_f32(ne[:, p], exp, f"rounds {rounds} perm {p}")
    one, _ = oracle.null(s, lab, T, 2, perm_mode=1, seed=31, rounds=rounds)
    ex = ctx.enrich(s, order, lab, T, ctx.null_params(2, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_EXACT, seed=31, feistel_rounds=rounds))[1]
    assert_bits_equal_f32(ex, one, f"rounds {rounds} exact")


def test_fast_vs_exact_tolerance(ctx):
    # FAST (exact int32 fixed-point sums, fp32 candidates) vs EXACT (the reference's fp32 running sums) on identical
    # labels. Stated tolerance: ||ES_fast| - |ES_exact|| <= 2e-4 (ES in [-1,1]) for rankings of up to 250 k cells;
    # it is EXACT that carries the error (its fp32 accumulators drift by ~1e-5 relative over 1e5 steps, FAST stays
    # within 2e-5 of the real-number value, tests/test_oracle_cpu.py). The sign can differ only when the positive
    # and the negative excursion are within that tolerance of each other.
    n_used, T, Pn = 60_000, 40, 12
    s, lab, order = _ranked(n_used, T, seed=4)
    a = ctx.enrich(s, order, lab, T, ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_EXACT, seed=8))[1]
    b = ctx.enrich(s, order, lab, T, ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=8))[1]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    assert np.max(np.abs(np.abs(a[ok]) - np.abs(b[ok]))) <= 2e-4
    assert np.mean(np.sign(a[ok]) == np.sign(b[ok])) > 0.97


def test_null_slices_concatenate(ctx):
    # permutation index is global: slices computed separately equal the one-shot null (multi-GPU invariance)
    n_used, T, Pn = 6000, 15, 20
    s, lab, order = _ranked(n_used, T, seed=6)
    for ks in (P.KS_EXACT, P.KS_FAST):
        full = ctx.enrich(s, order, lab, T, ctx.null_params(Pn, ks_mode=ks, seed=3))[1]
        parts = []
        for b, c in ((0, 7), (7, 6), (13, 7)):
            r, ne, _ = ctx.enrich(s, order, lab, T, ctx.null_params(Pn, ks_mode=ks, seed=3, perm_begin=b, perm_count=c))
            assert np.isnan(r["p_emp"]).all()  # slices do not reduce
            parts.append(ne)
        assert_bits_equal_f32(np.concatenate(parts, axis=1), full, "sliced null")


@pytest.mark.parametrize("T,Pn", [(12, 50), (300, 7), (40, 1000)])
def test_reduce_exact(ctx, oracle, T, Pn):
    rng = np.random.default_rng(T + Pn)
    null = rng.normal(0.1, 0.3, (T, Pn)).astype(np.float32)
    null[rng.random((T, Pn)) < 0.02] = 0.0
    null = np.round(null, 2)  # duplicates: exercises the binarySearch landing index
    res = np.zeros(T, P.TYPE_RESULT_DTYPE)
    res["ews"] = np.round(rng.normal(0.2, 0.4, T), 2).astype(np.float32)
    res["ews"][0] = np.nan
    null[0, :] = np.nan
    if T > 3:
        null[2, :] = -np.abs(null[2, :]) - 0.01  # no positive null at all -> NaN p-value / norm
        res["ews"][3] = 0.0
    for pol in (0, 1):
        g = ctx.reduce_null(res, null, pol)
        o = oracle.reduce(res, null, pol)
        assert_results_equal(g, o, ("norm_ews", "p_emp", "fdr"), f"reduce policy {pol}")


def test_analyze_end_to_end_vs_oracle(ctx, oracle, small):
    d = small
    cfg = d["cfg"]
    m = _load(ctx, d)
    sprm = P.ScoreParams(0, 4, 0.2, 1, P.JITTER_PHILOX, 77)
    Pn = 12
    nprm = ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_EXACT, seed=123, min_pass=100)
    out = ctx.analyze(m, d["gene_idx"], d["abundance"], sprm, 0.0, cfg.n_types, nprm, want_null=True,
                      want_scores=True, want_order=True)
    sc, used, nz, kept = oracle.score(d["row_ptr"], d["col"], d["val"], cfg.n_genes, d["gene_idx"], d["abundance"],
                                      method=0, min_genes_cells=4, min_corr=0.2, jitter_mode=1, seed=77)
    assert_bits_equal_f64(out["score"], sc, "score")
    order, npass = oracle.rank(sc, kept, 0.0)
    assert out["n_pass"] == npass and np.array_equal(out["order"], order)
    nu = npass - 1
    s_r, lab_r = sc[order[:nu]], d["label"][order[:nu]]
    ores = oracle.ks_observed(s_r, lab_r, cfg.n_types)
    one, ond = oracle.null(s_r, lab_r, cfg.n_types, Pn, perm_mode=1, seed=123)
    ores = oracle.reduce(ores, one)
    assert_bits_equal_f32(out["null"], one, "null")
    assert_results_equal(out["results"], ores, None, "analyze")
    # too few passing cells -> the reference's IllegalArgumentException (PCTSEA.java:389-395)
    with pytest.raises(P.PctseaError) as e:
        ctx.analyze(m, d["gene_idx"], d["abundance"], sprm, 0.0, cfg.n_types,
                    ctx.null_params(2, min_pass=10**9))
    assert e.value.code == -4
    m.free()


@pytest.mark.parametrize("gene_major", [False, True])
def test_host_api_mirror(ctx, small, tmp_path, gene_major):
    d = small
    cfg = d["cfg"]
    ds = P.SingleCellDataset.from_arrays(ctx, d["row_ptr"], d["col"], d["val"], cfg.n_genes, d["label"], cfg.n_types,
                                         gene_major=gene_major)
    ip = P.InputParameters()
    ip.addScoringSchema(P.ScoringMethod.PEARSONS_CORRELATION, 0.0, 4)
    ip.setNumPermutations(20)
    ip.setMinCorr(0.2)
    p = P.PCTSEA(ip, ds)
    p.MIN_CELLS_PASSING_SCORE_TRESHOLD = 100
    p.setInputList([f"G{g}" for g in d["gene_idx"]], d["abundance"])
    r = p.run()
    types = r.cellTypesPerRound[0]
    assert len(types) == cfg.n_types
    assert all(0.0 <= c.getHypergeometricPValue() <= 1.0 and c.getNumCellsOfType() >= c.getNumCellsOfTypePassingCorrelationThreshold() for c in types)
    es = [c.getEnrichmentScore() for c in types if not np.isnan(c.getEnrichmentScore())]
    assert es == sorted(es, reverse=True)
    assert all(c.getEnrichmentFDR() <= 0.05 for c in r.getSignificantTypes())
    # BH-corrected KS values (PCTSEA.java:1551-1570) and the per-type report files (:325-349)
    adj = np.array([c.getKSTestCorrectedPvalue() for c in types])
    raw = np.array([c.getKSTestPvalue() for c in types])
    ok = ~np.isnan(raw)
    assert np.all(adj[ok] >= raw[ok] - 1e-15) and np.all(adj[ok] <= 1.0)
    p.setPrefix("run1")
    p.outputFolder = str(tmp_path)
    p.writeCellTypeFiles = True
    r2 = p.run()
    import os
    written = sorted(os.listdir(tmp_path))
    pos = [c for c in r2.cellTypesPerRound[0] if not np.isnan(c.getEnrichmentScore()) and c.getEnrichmentScore() >= 0]
    assert pos and all(f"run1_{P.ScoringMethod.PEARSONS_CORRELATION.getScoreName()}_{c.getName()}_ews.txt" in written for c in pos)
    assert sum(f.endswith("_corr.txt") for f in written) == len(pos) == sum(f.endswith("_genes_per_cell_hist.txt") for f in written)
    c0 = pos[0]
    corr = open(os.path.join(tmp_path, f"run1_{P.ScoringMethod.PEARSONS_CORRELATION.getScoreName()}_{c0.getName()}_corr.txt")).read().splitlines()
    assert len(corr) - 1 == c0.getSizeA()   # one score per cell of the type in the ranked list


def test_against_committed_golden_fixture(ctx, tiny):
    # tests/golden/golden.json "regression": produced by the oracle with tests/golden/make_golden.py
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))["regression"]
    d, cfg = tiny, tiny["cfg"]
    m = _load(ctx, d)
    sprm = P.ScoreParams(0, 4, 0.2, 1, P.JITTER_PHILOX, 1)
    out = ctx.analyze(m, d["gene_idx"], d["abundance"], sprm, 0.0, cfg.n_types,
                      ctx.null_params(8, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_EXACT, seed=42, min_pass=100),
                      want_null=True, want_scores=True, want_order=True)
    sc = out["score"]
    assert out["n_pass"] == g["n_pass"] and int(out["kept"].sum()) == g["kept_sum"]
    assert sc[~np.isnan(sc)][:32].view(np.uint64).tolist() == g["score_bits_first32_nonnan"]
    assert out["order"][:32].tolist() == g["order_first32"]
    r = out["results"]
    assert r["ews"].view(np.uint32).tolist() == g["ews_bits"] and r["supX"].tolist() == g["supX"]
    assert out["null"].view(np.uint32).ravel().tolist() == g["null_feistel_seed42_bits"]
    assert r["norm_ews"].view(np.uint32).tolist() == g["norm_ews_bits"]
    assert r["p_emp"].view(np.uint64).tolist() == g["p_emp_bits"] and r["fdr"].view(np.uint64).tolist() == g["fdr_bits"]
    fast = ctx.analyze(m, d["gene_idx"], d["abundance"], sprm, 0.0, cfg.n_types,
                       ctx.null_params(8, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=42, min_pass=100),
                       want_null=True)
    assert fast["null"][:, 0].view(np.uint32).tolist() == g["fast_null_p0_bits"]
    m.free()


def test_full_size_properties(ctx):
    """BASELINE-scale stage 3 (600k ranked cells, 100 types): size-independent properties instead of an oracle run.
    (1) slices of the permutation range concatenate to the one-shot null; (2) the label permutation preserves type
    counts, so a type is NaN in the null iff it is NaN observed; (3) |ES| <= 1 and dStat = ES*sqrt(nA*nB/(nA+nB));
    (4) FAST stays within the stated tolerance of EXACT on the same permutations."""
    n_used, T, Pn = 600_000, 100, 6
    s, lab = synth.synthetic_ranked(n_used, T, seed=9)
    order = np.arange(n_used, dtype=np.int32)
    fast, nf, df = ctx.enrich(s, order, lab, T, ctx.null_params(Pn, ks_mode=P.KS_FAST, seed=1))
    exact, ne, de = ctx.enrich(s, order, lab, T, ctx.null_params(Pn, ks_mode=P.KS_EXACT, seed=1))
    nan_obs = np.isnan(exact["ews"])
    assert np.array_equal(np.isnan(ne).all(axis=1), nan_obs) and np.array_equal(np.isnan(nf), np.isnan(ne))
    ok = ~np.isnan(ne)
    # the supremum is the larger of a positive and a negative excursion: when the two are within the tolerance
    # of each other the winner (hence the sign) can differ, so the bound is on the magnitude
    # (stated tolerance: 2e-4 up to 250k ranked cells, growing linearly with the ranking above that — what EXACT
    # replays is the reference's fp32 accumulators, whose rounding drift grows with the number of additions)
    assert np.all(np.abs(ne[ok]) <= 1.0) and np.max(np.abs(np.abs(ne[ok]) - np.abs(nf[ok]))) <= 2e-4 * n_used / 250_000
    assert np.mean(np.sign(ne[ok]) == np.sign(nf[ok])) > 0.97
    a = ctx.enrich(s, order, lab, T, ctx.null_params(Pn, ks_mode=P.KS_FAST, seed=1, perm_begin=0, perm_count=2))[1]
    b = ctx.enrich(s, order, lab, T, ctx.null_params(Pn, ks_mode=P.KS_FAST, seed=1, perm_begin=2, perm_count=4))[1]
    assert_bits_equal_f32(np.concatenate([a, b], axis=1), nf, "sliced FAST null at full size")
    nA = exact["sizeA"].astype(np.int64)
    nB = exact["sizeB"].astype(np.int64)
    for t in np.flatnonzero(~nan_obs)[:10]:
        prod = int(np.array([(nA[t] * nB[t]) & 0xFFFFFFFF], dtype=np.uint32).view(np.int32)[0])  # int32 wrap
        with np.errstate(invalid="ignore"):
            exp = np.float32(np.sqrt(np.float64(prod) / np.float64(nA[t] + nB[t]))) * ne[t]
        assert_bits_equal_f32(de[t], exp.astype(np.float32), "dStat")


def test_fast_null_is_independent_of_the_kernel_geometry(ctx, oracle):
    """The FAST null carries exact integer sums, so how the kernel splits the ranking into per-thread segments (threads
    per permutation), how many CTAs share an SM and whether the cell types are processed in slices cannot change a
    bit — and every geometry agrees with the oracle's restatement of the FAST definition."""
    n_used, T, Pn = 50_000, 30, 7
    s, lab, order = _ranked(n_used, T, seed=12)
    prm = ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=21)
    outs = {}
    try:
        for name, opts in (("default", {}), ("64 threads", {P.OPT_NULL_THREADS: 64}), ("160 threads, 1 CTA/SM", {P.OPT_NULL_THREADS: 160, P.OPT_NULL_CTAS_PER_SM: 1}),
                           ("sliced", {P.OPT_NULL_MIN_THREADS: 512, P.OPT_NULL_THREADS: 256}), ("no overlap", {P.OPT_OBS_OVERLAP: 0}),
                           # class table: no fine region (every small class resolved by walking the cumulative counts),
                           # a fine region of 8 and of 100 entries
                           ("coarse class table", {P.OPT_NULL_FINE_ENTRIES: 1}), ("8 fine entries", {P.OPT_NULL_FINE_ENTRIES: 9}),
                           ("100 fine entries, sliced", {P.OPT_NULL_FINE_ENTRIES: 101, P.OPT_NULL_MIN_THREADS: 512, P.OPT_NULL_THREADS: 256})):
            for o in (P.OPT_NULL_THREADS, P.OPT_NULL_CTAS_PER_SM, P.OPT_NULL_MIN_THREADS, P.OPT_NULL_FINE_ENTRIES):
                ctx.set_option(o, 0)
            ctx.set_option(P.OPT_OBS_OVERLAP, 1)
            for o, v in opts.items():
                ctx.set_option(o, v)
            outs[name] = ctx.enrich(s, order, lab, T, prm)
    finally:
        for o in (P.OPT_NULL_THREADS, P.OPT_NULL_CTAS_PER_SM, P.OPT_NULL_MIN_THREADS, P.OPT_NULL_FINE_ENTRIES):
            ctx.set_option(o, 0)
        ctx.set_option(P.OPT_OBS_OVERLAP, 1)
    ref = outs["default"]
    for name, (res, ne, nd) in outs.items():
        assert_bits_equal_f32(ne, ref[1], f"null ES, geometry '{name}'")
        assert_bits_equal_f32(nd, ref[2], f"null dStat, geometry '{name}'")
        assert res.tobytes() == ref[0].tobytes(), f"results, geometry '{name}'"
    F = oracle.fast_frac_bits(s)
    for p in (0, Pn - 1):
        exp = oracle.ks_null_fast_one(s, oracle.perm_labels(21, p, lab, T), T, F)
        assert_bits_equal_f32(ref[1][:, p], exp, f"null kernel vs oracle, perm {p}")


def test_exact_null_chunking_is_invisible(ctx):
    """The EXACT (validation) null processes its permutations in chunks when the label matrix would not fit: chunk sizes
    of 1, 7 and unchunked give bit-identical nulls, for generated and for replayed permutations."""
    n_used, T, Pn = 6000, 15, 20
    s, lab, order = _ranked(n_used, T, seed=6)
    rng = np.random.default_rng(3)
    replay = np.stack([rng.permutation(n_used).astype(np.int32) for _ in range(Pn)])
    try:
        for mode, rp in ((P.PERM_PHILOX, None), (P.PERM_REPLAY, replay)):
            prm = ctx.null_params(Pn, perm_mode=mode, ks_mode=P.KS_EXACT, seed=3)
            outs = []
            for chunk in (0, 1, 7):
                ctx.set_option(P.OPT_PERM_CHUNK, chunk)
                outs.append(ctx.enrich(s, order, lab, T, prm, replay=rp))
            for o in outs[1:]:
                assert_bits_equal_f32(o[1], outs[0][1], "chunked null ES")
                assert_bits_equal_f32(o[2], outs[0][2], "chunked null dStat")
                assert o[0].tobytes() == outs[0][0].tobytes()
    finally:
        ctx.set_option(P.OPT_PERM_CHUNK, 0)


def test_load_random_checkpoint(ctx, small, tmp_path):
    """-load_random (PCTSEA.java:1385-1387): a second run that finds the null-distribution file reads it instead
    of permuting, and reproduces the first run's normalised scores, p-values and FDR exactly."""
    d, cfg = small, small["cfg"]
    ds = P.SingleCellDataset.from_arrays(ctx, d["row_ptr"], d["col"], d["val"], cfg.n_genes, d["label"], cfg.n_types)

    def run(load):
        ip = P.InputParameters()
        ip.addScoringSchema(P.ScoringMethod.PEARSONS_CORRELATION, 0.0, 4)
        ip.setNumPermutations(25)
        ip.setMinCorr(0.2)
        ip.setOutputPrefix("chk")
        ip.setLoadRandom(load)
        p = P.PCTSEA(ip, ds)
        p.outputFolder = str(tmp_path)
        p.MIN_CELLS_PASSING_SCORE_TRESHOLD = 100
        p.setInputList([f"G{g}" for g in d["gene_idx"]], d["abundance"])
        return {c.getName(): c for c in p.run().cellTypesPerRound[0]}

    first = run(False)
    import os
    f = os.path.join(str(tmp_path), "chk_correlation_25_random_scores.txt")
    assert os.path.exists(f)
    # poison the generator path: if the second run permuted again with another seed the numbers would move
    second = run(True)
    for name, a in first.items():
        b = second[name]
        for get in ("getEnrichmentScore", "getNormalizedEnrichmentScore", "getEnrichmentSignificance", "getEnrichmentFDR"):
            x, y = getattr(a, get)(), getattr(b, get)()
            assert (np.isnan(x) and np.isnan(y)) or x == y, (name, get, x, y)
        if a.randomEnrichmentScores is not None and not np.isnan(a.getEnrichmentScore()):
            assert np.array_equal(np.asarray(a.randomEnrichmentScores).view(np.uint32),
                                  np.asarray(b.randomEnrichmentScores).view(np.uint32))


def test_type_counts_for_hypergeometric_statistics(ctx, oracle, small):
    """§8 f4: numCellsOfType / numCellsOfTypeWithPositiveCorrelation / N of PCTSEA.java:2169-2264 from the GPU."""
    d, cfg = small, small["cfg"]
    m = _load(ctx, d)
    sprm = P.ScoreParams(0, 4, 0.2, 1, P.JITTER_PHILOX, 3)
    out = ctx.analyze(m, d["gene_idx"], d["abundance"], sprm, 0.3, cfg.n_types, ctx.null_params(0, min_pass=10),
                      want_scores=True)
    a, b, n_kept = ctx.last_type_counts(cfg.n_types)
    sc, used, nz, kept = oracle.score(d["row_ptr"], d["col"], d["val"], cfg.n_genes, d["gene_idx"], d["abundance"],
                                      method=0, min_genes_cells=4, min_corr=0.2, jitter_mode=1, seed=3)
    lab = d["label"]
    keep = kept == 1
    passing = keep & (sc >= 0.3)
    assert n_kept == int(keep.sum()) and out["n_pass"] == int(passing.sum())
    for t in range(cfg.n_types):
        assert a[t] == int((keep & (lab == t)).sum()) and b[t] == int((passing & (lab == t)).sum())
    m.free()


def test_negative_threshold_mode(ctx, oracle):
    """min_score < 0: cells with score < threshold pass, all ranked scores are negative (scoring/ScoreThreshold.java:
    66-74 + PCTSEA.java:2497). Accumulators, denominators and the observed extras run on negative sums."""
    rng = np.random.default_rng(31)
    n, T = 9000, 9
    score = rng.uniform(-1.0, 0.4, n)
    score[rng.random(n) < 0.05] = np.nan
    label = rng.integers(-1, T, n).astype(np.int32)
    thr = -0.2
    go, gn = ctx.rank(score, None, thr)
    oo, on = oracle.rank(score, None, thr)
    assert gn == on and np.array_equal(go, oo)
    nu = gn - 1
    order = go[:nu]
    s_r, lab_r = score[order], label[order]
    assert np.all(s_r < 0)
    Pn = 6
    res, ne, nd = ctx.enrich(score, order, label, T, ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_EXACT, seed=4))
    ores = oracle.ks_observed(s_r, lab_r, T)
    one, ond = oracle.null(s_r, lab_r, T, Pn, perm_mode=1, seed=4)
    assert_bits_equal_f32(ne, one, "null (negative scores)")
    assert_results_equal(res, oracle.reduce(ores, one), None, "negative-threshold analysis")
    fast = ctx.enrich(score, order, label, T, ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=4))[1]
    F = oracle.fast_frac_bits(s_r)
    for p in range(Pn):
        exp = oracle.ks_null_fast_one(s_r, oracle.perm_labels(4, p, lab_r, T), T, F)
        assert_bits_equal_f32(fast[:, p], exp, f"fast null (negative scores) perm {p}")


def test_ties_zeros_and_many_types(ctx, oracle):
    """Heavily tied scores with a block of exact zeros at the end (the accumulators stop moving, tie groups matter for
    the KS D statistic), 200 types (lane kernel at its smallest thread count) and 1000 types (u16 labels)."""
    for n_used, T in ((30_000, 200), (6000, 1000)):
        rng = np.random.default_rng(n_used)
        s = np.round(np.sort(rng.uniform(0.0, 1.0, n_used))[::-1], 2)
        s[-n_used // 5:] = 0.0
        lab = rng.integers(-1, T, n_used).astype(np.int32)
        order = np.arange(n_used, dtype=np.int32)
        Pn = 3
        res, ne, _ = ctx.enrich(s, order, lab, T, ctx.null_params(Pn, ks_mode=P.KS_EXACT, seed=6))
        ores = oracle.ks_observed(s, lab, T)
        one, _ = oracle.null(s, lab, T, Pn, perm_mode=1, seed=6)
        assert_bits_equal_f32(ne, one, f"null T={T}")
        assert_results_equal(res, oracle.reduce(ores, one), None, f"ties/zeros T={T}")
        fast = ctx.enrich(s, order, lab, T, ctx.null_params(Pn, ks_mode=P.KS_FAST, seed=6))[1]
        F = oracle.fast_frac_bits(s)
        exp = oracle.ks_null_fast_one(s, oracle.perm_labels(6, 1, lab, T), T, F)
        assert_bits_equal_f32(fast[:, 1], exp, f"fast null T={T}")


def test_degenerate_shapes(ctx, oracle):
    s, lab = synth.synthetic_ranked(500, 4, seed=2)
    order = np.arange(500, dtype=np.int32)
    # no permutations: observed only, a10-a12 on an empty null are NaN
    res, ne, nd = ctx.enrich(s, order, lab, 4, ctx.null_params(0))
    assert ne.shape == (4, 0) and np.isnan(res["p_emp"]).all() and np.isnan(res["norm_ews"]).all()
    assert_results_equal(res, oracle.reduce(oracle.ks_observed(s, lab, 4), np.zeros((4, 0), np.float32)), None, "P=0")
    # a single permutation, two ranked cells, all labels unknown
    res, ne, _ = ctx.enrich(s[:2], order[:2], np.array([-1, -1], np.int32), 3, ctx.null_params(1, ks_mode=P.KS_FAST))
    assert np.isnan(res["ews"]).all() and np.isnan(ne).all() and (res["supX"] == -1).all()
    # bad arguments come back as errors, not crashes
    with pytest.raises(P.PctseaError):
        ctx.enrich(s, order, lab, 4, ctx.null_params(3, perm_mode=P.PERM_REPLAY, ks_mode=P.KS_EXACT))   # no replay array
    with pytest.raises(P.PctseaError):
        ctx.enrich(s, order, lab, 4, ctx.null_params(3, perm_begin=2, perm_count=5))                   # slice out of range
    with pytest.raises(P.PctseaError):
        ctx.enrich(s, np.array([0, 1, 10**6], np.int32), lab, 4, ctx.null_params(1))                   # order out of range


def test_observed_curves_and_cell_type_files(ctx, tmp_path):
    # pctsea_last_observed_curve: the a / b series of EnrichmentWeigthedScoreParallel.java:138-186, bit for bit against a
    # sequential float32 emulation of the reference loop; then the "<type>_ews" writer on top of it
    from pctsea_parent_b200 import api
    n_used, T = 3000, 5
    s, lab, order = _ranked(n_used, T, seed=12)
    res, _, _ = ctx.enrich(s, order, lab, T, ctx.null_params(0), want_null=False)
    sr, lr = s[order[:n_used]], lab[order[:n_used]]
    for t in range(T):
        a, b = ctx.last_observed_curve(t, n_used)
        numA = np.float32(0.0); numB = np.float32(0.0)
        ra = np.zeros(n_used, np.float32); rb = np.zeros(n_used, np.float32)
        anyA = anyB = False
        for i in range(n_used):
            if lr[i] == t:
                numA = np.float32(np.float64(numA) + sr[i]); anyA = True
            else:
                numB = np.float32(np.float64(numB) + sr[i]); anyB = True
            ra[i] = numA if anyA else np.float32(0.0)
            rb[i] = numB if anyB else np.float32(0.0)
        ea = np.where(np.arange(n_used) >= np.argmax(lr == t) if (lr == t).any() else False, ra / numA, np.float32(0.0)).astype(np.float32)
        eb = np.where(np.arange(n_used) >= np.argmax(lr != t), rb / numB, np.float32(0.0)).astype(np.float32)
        assert_bits_equal_f32(a, ea, f"a series type {t}")
        assert_bits_equal_f32(b, eb, f"b series type {t}")
        # the supremum the kernel reported is the first position of max |a - b|
        d = np.abs(a - b)
        assert int(res["supX"][t]) == int(np.argmax(d)) + 1
    with pytest.raises(P.PctseaError):
        ctx.last_observed_curve(T, n_used)
    with pytest.raises(P.PctseaError):
        ctx.last_observed_curve(0, n_used - 1)
    a, b = ctx.last_observed_curve(0, n_used)
    path = tmp_path / "t0_ews.txt"
    api.write_score_calculation_file(str(path), "type0", a, b, int(res["supX"][0]), int(res["sec_supX"][0]))
    lines = path.read_text().splitlines()
    assert lines[0] == "-\tcell #\tCumulative Probability [Fn(X)]"
    assert lines[1].startswith("type0\t1\t")
    sup = [l for l in lines if l.startswith("supremum\t")]
    assert len(sup) == 2 and sup[0].split("\t")[1] == str(int(res["supX"][0]))
    assert float(np.float32(sup[0].split("\t")[2])) == float(a[int(res["supX"][0]) - 1])


def test_stage3_random_shapes_sweep(ctx, oracle):
    # Random ranking sizes and type counts around the switch points of the stage-3 kernels (4096 cells: cooperative
    # vs lane kernel; ~218 / 254 types: plain vs sliced lane kernel, 8- vs 16-bit labels; segment and macro-step
    # remainders), each checked bit for bit: observed statistics, FAST null vs the oracle's definition, reduction.
    rng = np.random.default_rng(2026)
    shapes = [(4095, 7), (4096, 7), (4097, 219), (4100, 255), (4128, 256), (12_345, 110), (8191, 33), (33_000, 3),
              (5003, 500), (6001, 1), (4099, 254)]
    shapes += [(int(rng.integers(4096, 30_000)), int(rng.integers(1, 400))) for _ in range(6)]
    for n_used, T in shapes:
        s, lab, order = _ranked(n_used, T, seed=n_used + T, simple=bool(n_used & 1))
        Pn = 3
        prm = ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=n_used)
        res, ne, nd = ctx.enrich(s, order, lab, T, prm)
        ores = oracle.ks_observed(s, lab, T)
        assert_results_equal(res, ores, OBSERVED_FIELDS, f"observed n={n_used} T={T}")
        F = oracle.fast_frac_bits(s)
        for p in range(Pn):
            exp = oracle.ks_null_fast_one(s, oracle.perm_labels(n_used, p, lab, T), T, F)
            assert_bits_equal_f32(ne[:, p], exp, f"fast null n={n_used} T={T} perm {p}")
        ored = oracle.reduce(ores, ne)
        assert_results_equal(res, ored, ("norm_ews", "p_emp", "fdr"), f"reduce n={n_used} T={T}")


def test_two_contexts_on_one_device(ctx, oracle):
    """Several ctxs may share a device (one per PCTSEA instance of the web front-end's pool): interleaved calls with
    different shapes — hence different dynamic shared-memory sizes of the same kernels — stay correct."""
    other = P.Context(0)
    try:
        for (c, n_used, T) in ((ctx, 9000, 40), (other, 5000, 150), (ctx, 140_000, 20), (other, 9000, 40), (ctx, 5000, 150)):
            s, lab, order = _ranked(n_used, T, seed=4)
            prm = c.null_params(3, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=5)
            res, ne, nd = c.enrich(s, order, lab, T, prm)
            F = oracle.fast_frac_bits(s)
            exp = oracle.ks_null_fast_one(s, oracle.perm_labels(5, 1, lab, T), T, F)
            assert_bits_equal_f32(ne[:, 1], exp, f"ctx {'A' if c is ctx else 'B'} Nu={n_used} T={T}")
    finally:
        other.close()


def test_pooled_contexts_share_one_matrix(ctx, small):
    """Config E the way the reference's web front-end runs it (an executor pool of PCTSEA instances,
    AnalyzeView.java:163): two contexts on one GPU, one host thread each, the SAME resident matrix, list i on context
    i mod 2 (dist.analyze_batch_pooled). Every list's results equal the single-context batch bit for bit."""
    from pctsea_parent_b200 import dist as pdist
    d = small
    cfg = d["cfg"]
    m = _load(ctx, d)
    m.set_labels(d["label"], cfg.n_types)
    m.build_gene_index()
    sprm = P.ScoreParams(0, 4, 0.2, 1, P.JITTER_PHILOX, 77)
    nprm = ctx.null_params(16, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=9, min_pass=100)
    lists = [tuple(t.numpy() for t in synth.make_list(cfg, "cpu", s)) for s in range(7)]
    ref = ctx.analyze_batch(m, lists, sprm, cfg.min_score, cfg.n_types, nprm, want_null=True)
    other = P.Context(0)
    try:
        for _ in range(3):   # a few rounds: the interleaving of the two threads differs from run to run
            out = pdist.analyze_batch_pooled([ctx, other], m, lists, sprm, cfg.min_score, cfg.n_types, nprm, want_null=True)
            assert np.array_equal(out["n_pass"], ref["n_pass"]) and np.array_equal(out["status"], ref["status"])
            for i in range(len(lists)):
                assert_results_equal(out["results"][i], ref["results"][i], None, f"pooled list {i}")
            assert out["results"].tobytes() == ref["results"].tobytes()   # padding included
            assert_bits_equal_f32(out["null"], ref["null"], "pooled null")
    finally:
        other.close()
        m.free()


@pytest.mark.parametrize("gene_major", [False, True])
def test_analyze_batch_equals_single_analyses(ctx, oracle, small, gene_major):
    """BASELINE config E in miniature: many lists against one resident matrix. Every list's results are
    bit-identical to pctsea_analyze on that list alone; a list that fails the reference's own input checks fails
    alone; one list is checked end to end against the oracle."""
    d = small
    cfg = d["cfg"]
    m = _load(ctx, d)
    m.set_labels(d["label"], cfg.n_types)
    if gene_major:
        m.build_gene_index()
    sprm = P.ScoreParams(0, 4, 0.2, 1, P.JITTER_PHILOX, 77)
    Pn = 24
    nprm = ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=9, min_pass=100)
    lists = []
    for seed in range(5):
        g, a = synth.make_list(cfg, "cpu", seed)
        lists.append((g.numpy(), a.numpy()))
    lists.insert(2, (lists[0][0][:3], lists[0][1][:3]))                          # 3 genes < min_genes_cells 4
    lists.insert(4, (np.array([-1, -1, -1, -1, -1], np.int32), np.ones(5, np.float32)))  # no gene in the dataset
    out = ctx.analyze_batch(m, lists, sprm, 0.0, cfg.n_types, nprm, want_null=True)
    tm = ctx.timings()
    assert list(out["status"]) == [0, 0, -5, 0, -4, 0, 0]
    n_ok = 0
    for i, (g, a) in enumerate(lists):
        if out["status"][i] != 0:
            assert np.isnan(out["results"][i]["ews"]).all() and np.isnan(out["results"][i]["fdr"]).all()
            assert (out["results"][i]["supX"] == -1).all() and np.isnan(out["null"][i]).all()
            with pytest.raises(P.PctseaError) as e:
                ctx.analyze(m, g, a, sprm, 0.0, cfg.n_types, nprm)
            assert e.value.code == out["status"][i]
            continue
        one = ctx.analyze(m, g, a, sprm, 0.0, cfg.n_types, nprm, want_null=True)
        assert one["n_pass"] == out["n_pass"][i]
        assert_results_equal(out["results"][i], one["results"], None, f"list {i}")
        assert_bits_equal_f32(out["null"][i], one["null"], f"list {i} null")
        assert_bits_equal_f32(out["nullD"][i], one["nullD"], f"list {i} nullD")
        n_ok += 1
    assert n_ok == 5 and tm["launches"] > 5 * 20 and tm["score_ms"] > 0
    # resident labels and labels passed with the call give the same batch
    out2 = ctx.analyze_batch(m, lists[:2], sprm, 0.0, cfg.n_types, nprm, label=d["label"])
    assert_results_equal(out2["results"][1], out["results"][1], None, "labels per call")
    # list 3 against the oracle, end to end (FAST arithmetic on the Philox permutations)
    g, a = lists[3]
    sc, used, nz, kept = oracle.score(d["row_ptr"], d["col"], d["val"], cfg.n_genes, g, a, method=0, min_genes_cells=4,
                                      min_corr=0.2, jitter_mode=1, seed=77, canonical=2 if gene_major else 1)
    order, npass = oracle.rank(sc, kept, 0.0)
    assert npass == out["n_pass"][3]
    nu = npass - 1
    s_r, lab_r = sc[order[:nu]], d["label"][order[:nu]]
    F = oracle.fast_frac_bits(s_r)
    one = np.stack([oracle.ks_null_fast_one(s_r, oracle.perm_labels(9, p, lab_r, cfg.n_types), cfg.n_types, F)
                    for p in range(Pn)], axis=1)
    ores = oracle.reduce(oracle.ks_observed(s_r, lab_r, cfg.n_types), one)
    assert_bits_equal_f32(out["null"][3], one, "batch list 3 null vs oracle")
    assert_results_equal(out["results"][3], ores, None, "batch list 3 vs oracle")
    # empty batch
    empty = ctx.analyze_batch(m, [], sprm, 0.0, cfg.n_types, nprm)
    assert empty["results"].shape == (0, cfg.n_types)
    m.free()


# ---------------------------------------------------------------------------------------------------
# stage 1 over the gene-major copy (pctsea_matrix_build_gene_index): lane-per-cell kernel, pair order = ascending
# gene, shifted-data sums; bit-exact against the oracle's canonical form 2
# ---------------------------------------------------------------------------------------------------
def _check_gm(ctx, oracle, row_ptr, col, val, G, gene_idx, abundance, what, **kw):
    m = ctx.load_csr(row_ptr, col, val, G).build_gene_index()
    g = ctx.score(m, gene_idx, abundance, **kw)
    o = oracle.score(row_ptr, col, val, G, gene_idx, abundance, canonical=2, **kw)
    assert np.array_equal(g[1], o[1]), f"{what}: n_genes_used"
    assert np.array_equal(g[2], o[2]), f"{what}: n_nonzero"
    assert np.array_equal(g[3], o[3]), f"{what}: kept"
    assert_bits_equal_f64(g[0], o[0], f"{what}: score")
    m.free()
    return g


@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("jitter", [0, 1])
def test_score_gene_major_bit_exact(ctx, oracle, tiny, method, jitter):
    d = tiny
    kw = dict(method=method, min_genes_cells=4, min_corr=0.2 if method == 0 else -1.0, has_min_corr=True,
              jitter_mode=jitter, seed=1234)
    g = _check_gm(ctx, oracle, d["row_ptr"], d["col"], d["val"], d["cfg"].n_genes, d["gene_idx"], d["abundance"],
                  "tiny", **kw)
    # <= 1e-9 relative against the reference-faithful sequential SimpleRegression update on every cell whose
    # vectors are not constant (those get jitter: noise of relative condition ~1e7, see test_score_bit_exact)
    f = oracle.score(d["row_ptr"], d["col"], d["val"], d["cfg"].n_genes, d["gene_idx"], d["abundance"],
                     canonical=False, **kw)
    assert np.array_equal(np.isnan(f[0]), np.isnan(g[0])), "NaN pattern vs sequential oracle"
    nj = oracle.score(d["row_ptr"], d["col"], d["val"], d["cfg"].n_genes, d["gene_idx"], d["abundance"],
                      canonical=False, method=0, min_genes_cells=4, min_corr=-1.0, jitter_mode=0, seed=1234)
    zero_var = (nj[3] == 1) & (nj[1] >= 2) & np.isnan(nj[0])
    ok = ~np.isnan(f[0])
    regular = ok & ~zero_var
    assert regular.sum() > 100
    assert np.all(np.abs(g[0][regular] - f[0][regular]) <= 1e-9 * np.maximum(1.0, np.abs(f[0][regular])))
    assert np.all(np.abs(g[0][ok] - f[0][ok]) <= 1e-7)
    # and the CSR-scan kernel (a matrix without the index) agrees to the same bar
    m = ctx.load_csr(d["row_ptr"], d["col"], d["val"], d["cfg"].n_genes)
    c = ctx.score(m, d["gene_idx"], d["abundance"], **kw)
    assert np.array_equal(c[1], g[1]) and np.array_equal(c[2], g[2]) and np.array_equal(c[3], g[3])
    assert np.all(np.abs(g[0][regular] - c[0][regular]) <= 1e-9 * np.maximum(1.0, np.abs(c[0][regular])))
    m.free()


def test_score_gene_major_edge_cases(ctx, oracle):
    G = 50
    rows = [
        [],                                                        # empty cell
        [(1, 1.0), (2, 1.0), (3, 1.0), (4, 1.0)],                 # zero variance in the cell
        [(9, 3.0), (1, 2.0), (4, 7.0), (2, 5.0), (3, 1.0)],       # regular, row stored out of gene order
        [(1, float("nan")), (2, 5.0), (3, 0.0), (4, -2.0), (5, 1.0), (6, 2.0), (7, 9.0)],
        [(1, 3.0), (2, 1.0)],                                      # fewer than min_genes pairs
        [(10, 4.0), (11, 4.0), (12, 2.0), (13, 8.0)],             # list genes 10..13 share one abundance
        [(40, 1.0), (41, 2.0)],                                    # no list gene at all
        [(10, 5.0), (11, 5.0), (12, 5.0)],                        # both vectors constant
    ] + [[(1, 2.0), (2, float(i % 5 + 1)), (3, 1.0), (9, float(i % 3 + 1))] for i in range(70)]   # > 2 groups of 32
    row_ptr = np.cumsum([0] + [len(r) for r in rows]).astype(np.int64)
    col = np.array([c for r in rows for c, _ in r], np.int32)
    val = np.array([v for r in rows for _, v in r], np.float32)
    gene_idx = np.array([1, 2, 3, 4, 5, 6, 7, 9, -1, 10, 11, 12, 13, 20], np.int32)
    abundance = np.array([3, 1, 4, 1.5, 9, 2, 6, 5, 7, 2, 2, 2, 2, 0.0], np.float32)
    for method in (0, 1):
        for jitter in (0, 1):
            kw = dict(method=method, min_genes_cells=3, min_corr=-1.0, has_min_corr=True, jitter_mode=jitter, seed=7)
            _check_gm(ctx, oracle, row_ptr, col, val, G, gene_idx, abundance, f"edge m{method} j{jitter}", **kw)


@pytest.mark.parametrize("G,L,unsorted_rows,N,all_positive", [
    (6_000, 5_000, True, 1500, False), (40_000, 64, False, 1501, False), (3_000, 33, True, 31, False),
    (2_000, 600, False, 701, False),
    # every stored value > 0: the kernel variant that counts expressed genes and pairs per chunk from the bit sets
    # (the list below still has zero abundances and proteins without a matrix gene)
    (2_000, 600, False, 701, True), (6_000, 1_000, True, 333, True),
    # more genes than one pass of the index build holds in shared memory (50 k): two and three gene ranges
    (70_000, 900, False, 300, False), (120_000, 2_000, True, 200, True)])
def test_score_gene_major_shapes(ctx, oracle, G, L, unsorted_rows, N, all_positive):
    rng = np.random.default_rng(G + L)
    row_ptr, col, val = _random_csr(rng, N, G, 700 if G >= 6000 else 300, unsorted_rows)
    if not all_positive:
        val[rng.integers(0, val.size, 20)] = 0.0          # explicit zeros: not expressed
        val[rng.integers(0, val.size, 5)] = np.nan        # NaN: expressed, never a pair
    present = np.unique(col)
    gene_idx = rng.permutation(present)[:L].astype(np.int32) if present.size >= L else rng.permutation(G)[:L].astype(np.int32)
    gene_idx[rng.integers(0, L, 3)] = -1
    abundance = rng.integers(0, 40, L).astype(np.float32)
    for method, jitter in ((0, 1), (1, 0)):
        kw = dict(method=method, min_genes_cells=3, min_corr=-1.0, has_min_corr=True, jitter_mode=jitter, seed=99)
        g = _check_gm(ctx, oracle, row_ptr, col, val, G, gene_idx, abundance, f"G{G} L{L}", **kw)
        assert np.isfinite(g[0]).sum() > 5


def test_analyze_with_gene_index_vs_oracle(ctx, oracle, small):
    d = small
    cfg = d["cfg"]
    m = _load(ctx, d).build_gene_index()
    sprm = P.ScoreParams(0, 4, 0.2, 1, P.JITTER_PHILOX, 77)
    Pn = 12
    nprm = ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_EXACT, seed=123, min_pass=100)
    out = ctx.analyze(m, d["gene_idx"], d["abundance"], sprm, 0.0, cfg.n_types, nprm, want_null=True,
                      want_scores=True, want_order=True)
    sc, used, nz, kept = oracle.score(d["row_ptr"], d["col"], d["val"], cfg.n_genes, d["gene_idx"], d["abundance"],
                                      method=0, min_genes_cells=4, min_corr=0.2, jitter_mode=1, seed=77, canonical=2)
    assert_bits_equal_f64(out["score"], sc, "score")
    assert np.array_equal(out["n_genes_used"], used) and np.array_equal(out["kept"], kept)
    order, npass = oracle.rank(sc, kept, 0.0)
    assert out["n_pass"] == npass and np.array_equal(out["order"], order)
    nu = npass - 1
    s_r, lab_r = sc[order[:nu]], d["label"][order[:nu]]
    one, _ = oracle.null(s_r, lab_r, cfg.n_types, Pn, perm_mode=1, seed=123)
    ores = oracle.reduce(oracle.ks_observed(s_r, lab_r, cfg.n_types), one)
    assert_bits_equal_f32(out["null"], one, "null")
    assert_results_equal(out["results"], ores, None, "analyze (gene index)")
    m.free()


@pytest.mark.parametrize("gene_major", [False, True])
def test_baseline_config_A_end_to_end(ctx, oracle, gene_major):
    """BASELINE configs[0] at full size — 20k cells x 20k genes, 40 cell types, 500-protein list, PEARSONS_CORRELATION,
    -perm 50, min_genes_cells 4, min_corr 0.2 (the README example the reference runs on CPU) — through pctsea_analyze
    with the reference's own guard of 1000 passing cells, against the oracle: scores, ranking, observed statistics,
    the null in validation mode (REPLAY of the oracle's JDK Collections.shuffle indices, EXACT arithmetic), p-values,
    normalised scores and FDR, all bit-exact; then the production mode (PHILOX permutations, FAST arithmetic) against
    the oracle's restatement of the same definition. Both scoring kernels."""
    cfg = synth.CONFIGS["A"]
    d = synth.make_dataset(cfg, "cpu").numpy()
    m = ctx.load_csr(d["row_ptr"], d["col"], d["val"], cfg.n_genes)
    m.set_labels(d["label"], cfg.n_types)
    if gene_major:
        m.build_gene_index()
    kw = dict(method=cfg.method, min_genes_cells=cfg.min_genes_cells, min_corr=cfg.min_corr, jitter_mode=1, seed=11)
    sc, used, nz, kept = oracle.score(d["row_ptr"], d["col"], d["val"], cfg.n_genes, d["gene_idx"], d["abundance"],
                                      canonical=2 if gene_major else 1, **kw)
    order, npass = oracle.rank(sc, kept, cfg.min_score)
    assert npass >= 1000
    nu = npass - 1
    s_r, lab_r = sc[order[:nu]], d["label"][order[:nu]]
    one, ond, replay = oracle.null(s_r, lab_r, cfg.n_types, cfg.n_perm, perm_mode=0, seed=2026, want_replay=True)
    ores = oracle.reduce(oracle.ks_observed(s_r, lab_r, cfg.n_types), one)
    sprm = P.ScoreParams(cfg.method, cfg.min_genes_cells, cfg.min_corr, 1, P.JITTER_PHILOX, 11)
    nprm = ctx.null_params(cfg.n_perm, perm_mode=P.PERM_REPLAY, ks_mode=P.KS_EXACT, min_pass=1000)
    out = ctx.analyze(m, d["gene_idx"], d["abundance"], sprm, cfg.min_score, cfg.n_types, nprm, replay=replay,
                      want_null=True, want_scores=True, want_order=True)
    assert_bits_equal_f64(out["score"], sc, "score")
    assert np.array_equal(out["n_genes_used"], used) and np.array_equal(out["kept"], kept)
    assert out["n_pass"] == npass and np.array_equal(out["order"], order)
    assert_bits_equal_f32(out["null"], one, "REPLAY null")
    assert_bits_equal_f32(out["nullD"], ond, "REPLAY null dStat")
    assert_results_equal(out["results"], ores, None, "config A")
    assert np.isfinite(out["results"]["fdr"]).sum() >= 3 and (out["results"]["fdr"] <= 0.05).sum() >= 1   # planted types are found
    # production mode
    nprm = ctx.null_params(cfg.n_perm, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=7, min_pass=1000)
    fast = ctx.analyze(m, d["gene_idx"], d["abundance"], sprm, cfg.min_score, cfg.n_types, nprm, want_null=True)
    F = oracle.fast_frac_bits(s_r)
    exp = np.stack([oracle.ks_null_fast_one(s_r, oracle.perm_labels(7, p, lab_r, cfg.n_types), cfg.n_types, F)
                    for p in range(cfg.n_perm)], axis=1)
    assert_bits_equal_f32(fast["null"], exp, "PHILOX FAST null")
    assert_results_equal(fast["results"], oracle.reduce(oracle.ks_observed(s_r, lab_r, cfg.n_types), exp), None,
                         "config A, production mode")
    m.free()


def test_scoring_paths_agree_at_scale(ctx):
    """HCL-shaped matrix (config B's generator, 100k cells x 25k genes, ~135M non-zeros, generated on the device):
    no oracle run at this size, so size-independent properties — the CSR scan and the gene-major kernel count the
    same pairs / expressed genes / kept cells for every cell and agree on every regular score to 1e-9 relative; each
    is reproducible bit for bit; the ranking of either is sorted by (score desc, index asc)."""
    import dataclasses
    import torch
    cfg = dataclasses.replace(synth.CONFIGS["B"], n_cells=100_000)
    ds = synth.make_dataset(cfg, torch.device("cuda", 0))
    gi, ab = ds.gene_idx.cpu().numpy(), ds.abundance.cpu().numpy()
    m = ctx.wrap_device_csr(ds.row_ptr.data_ptr(), ds.col.data_ptr(), ds.val.data_ptr(), cfg.n_cells, cfg.n_genes, keepalive=ds)
    m.set_labels(ds.label.cpu().numpy(), cfg.n_types)
    kw = dict(method=0, min_genes_cells=4, min_corr=0.2, has_min_corr=True, jitter_mode=P.JITTER_NAN, seed=3)
    a = ctx.score(m, gi, ab, **kw)
    m.build_gene_index()
    b = ctx.score(m, gi, ab, **kw)
    b2 = ctx.score(m, gi, ab, **kw)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert np.array_equal(np.isnan(a[0]), np.isnan(b[0]))
    ok = ~np.isnan(a[0])
    assert ok.sum() > 10_000
    assert np.max(np.abs(a[0][ok] - b[0][ok]) / np.maximum(1.0, np.abs(a[0][ok]))) <= 1e-9
    assert np.array_equal(b[0].view(np.uint64), b2[0].view(np.uint64))
    order, npass = ctx.rank(b[0], b[3], 0.0)
    s = b[0][order]
    assert npass == int(np.count_nonzero((b[3] == 1) & (b[0] >= 0.0)))
    assert np.all((s[:-1] > s[1:]) | ((s[:-1] == s[1:]) & (order[:-1] < order[1:])))
    m.free()


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configurations at FULL size against the oracle (the configurations the published numbers are quoted on)
# ---------------------------------------------------------------------------------------------------
def _device_dataset(ctx, cfg, gene_index=True):
    """Config generated in HBM (as bench.py does), wrapped as a resident matrix; host copy for the oracle."""
    import torch
    ds = synth.make_dataset(cfg, torch.device("cuda", 0))
    d = ds.numpy()
    m = ctx.wrap_device_csr(ds.row_ptr.data_ptr(), ds.col.data_ptr(), ds.val.data_ptr(), cfg.n_cells, cfg.n_genes, keepalive=ds)
    m.set_labels(d["label"], cfg.n_types)
    if gene_index:
        m.build_gene_index()
    return ds, d, m


@pytest.fixture(scope="module")
def config_B(ctx):
    cfg = synth.CONFIGS["B"]
    ds, d, m = _device_dataset(ctx, cfg)
    yield cfg, ds, d, m
    m.free()


def _score_kw(cfg, seed):
    return dict(method=cfg.method, min_genes_cells=cfg.min_genes_cells, min_corr=cfg.min_corr, has_min_corr=True,
                jitter_mode=1, seed=seed)


def test_baseline_config_B_full_size_vs_oracle(ctx, oracle, config_B):
    """BASELINE configs[1] exactly as bench.py runs it — 600 000 cells x 25 000 genes (~5 %), 100 cell types,
    1000-protein list, PEARSONS_CORRELATION, -perm 1000, gene-major scoring, PHILOX permutations, FAST arithmetic —
    against the oracle at full size: scores of all 600 000 cells bit for bit (canonical form of the gene-major kernel)
    and <= 1e-9 relative to the sequential SimpleRegression restatement on 2000 cells, the ranking, every observed
    field, the FAST null of permutations 0, 1, 500 and 999 bit for bit, and a10-a12 recomputed from the GPU's null."""
    from tests.parity import check_analysis
    cfg, ds, d, m = config_B
    sprm = P.ScoreParams(cfg.method, cfg.min_genes_cells, cfg.min_corr, 1, P.JITTER_PHILOX, 1)
    nprm = ctx.null_params(cfg.n_perm, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=20261019, min_pass=1000)
    out = ctx.analyze(m, d["gene_idx"], d["abundance"], sprm, cfg.min_score, cfg.n_types, nprm,
                      want_null=True, want_scores=True, want_order=True)
    rep, _ = check_analysis(oracle, d, cfg.n_genes, cfg.n_types, out, score_params=_score_kw(cfg, 1),
                            min_score=cfg.min_score, canonical=2, seed=20261019,
                            perm_indices=(0, 1, 500, cfg.n_perm - 1), what="config B")
    assert rep["n_used"] > 200_000 and rep["max_rel_vs_sequential"] <= 1e-9


def test_baseline_config_B_csr_scan_and_exact_null_vs_oracle(ctx, oracle, config_B):
    """The same matrix through the CSR-scan scoring kernel (north_star's warp-per-cell kernel; canonical form 1) and
    the validation arithmetic (PHILOX permutations, EXACT fp32/fp64 recurrence) for 3 permutations, at full size."""
    from tests.parity import check_analysis
    cfg, ds, d, m = config_B
    ctx.set_option(P.OPT_SCORE_PATH, 1)
    try:
        sprm = P.ScoreParams(cfg.method, cfg.min_genes_cells, cfg.min_corr, 1, P.JITTER_PHILOX, 1)
        nprm = ctx.null_params(3, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_EXACT, seed=5, min_pass=1000)
        out = ctx.analyze(m, d["gene_idx"], d["abundance"], sprm, cfg.min_score, cfg.n_types, nprm,
                          want_null=True, want_scores=True, want_order=True)
    finally:
        ctx.set_option(P.OPT_SCORE_PATH, 0)
    check_analysis(oracle, d, cfg.n_genes, cfg.n_types, out, score_params=_score_kw(cfg, 1), min_score=cfg.min_score,
                   canonical=1, seed=5, perm_indices=(0, 1, 2), ks_fast=False, sequential_sample=0,
                   what="config B, CSR scan + EXACT")


def test_baseline_config_E_batch_lists_vs_oracle(ctx, oracle, config_B):
    """BASELINE configs[4]'s shape: several input lists against the resident 600 000-cell matrix through
    pctsea_analyze_batch, -perm 1000 each. Three of the lists are checked in full against the oracle (scores through
    a single analysis of the same list — the batch must equal it bit for bit —, observed fields, FAST null columns,
    a10-a12)."""
    from tests.parity import check_analysis
    cfg, ds, d, m = config_B
    lists = [tuple(t.cpu().numpy() for t in synth.make_list(cfg, "cpu", list_seed=s)) for s in range(1, 6)]
    sprm = P.ScoreParams(cfg.method, cfg.min_genes_cells, cfg.min_corr, 1, P.JITTER_PHILOX, 1)
    nprm = ctx.null_params(cfg.n_perm, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=77, min_pass=1000)
    batch = ctx.analyze_batch(m, lists, sprm, cfg.min_score, cfg.n_types, nprm, want_null=True)
    for i in (0, 2, 4):
        if batch["status"][i] != 0:
            assert batch["status"][i] == -4   # PCTSEA_ETOOFEW: the oracle must agree that fewer than 1000 cells pass
            sc, used, nz, kept = oracle.score(d["row_ptr"], d["col"], d["val"], cfg.n_genes, lists[i][0], lists[i][1],
                                              canonical=2, **_score_kw(cfg, 1))
            assert oracle.rank(sc, kept, cfg.min_score)[1] < 1000
            continue
        one = ctx.analyze(m, lists[i][0], lists[i][1], sprm, cfg.min_score, cfg.n_types, nprm,
                          want_null=True, want_scores=True, want_order=True)
        assert one["n_pass"] == batch["n_pass"][i]
        assert_bits_equal_f32(batch["null"][i], one["null"], f"list {i}: batch null vs single analysis")
        assert_results_equal(batch["results"][i], one["results"], None, f"list {i}: batch vs single analysis")
        di = dict(d, gene_idx=lists[i][0], abundance=lists[i][1])
        check_analysis(oracle, di, cfg.n_genes, cfg.n_types, one, score_params=_score_kw(cfg, 1),
                       min_score=cfg.min_score, canonical=2, seed=77, perm_indices=(0, 333, cfg.n_perm - 1),
                       sequential_sample=0, what=f"config E, list {i}")


def test_baseline_config_D_shape_full_size_vs_oracle(ctx, oracle):
    """BASELINE configs[3]'s shape: 30 000 genes, 300 cell types (16-bit labels, the null kernel's type-sliced form),
    SIMPLE_SCORE, ~1.2 M ranked cells (60 % of D's 2 M cells: the host copy the oracle needs stays at 14 GB) —
    everything against the oracle as for config B."""
    import dataclasses
    from tests.parity import check_analysis
    cfg = dataclasses.replace(synth.CONFIGS["D"], n_cells=1_200_000, n_perm=64)
    ds, d, m = _device_dataset(ctx, cfg)
    try:
        sprm = P.ScoreParams(cfg.method, cfg.min_genes_cells, cfg.min_corr, 1, P.JITTER_PHILOX, 3)
        nprm = ctx.null_params(cfg.n_perm, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=4, min_pass=1000)
        out = ctx.analyze(m, d["gene_idx"], d["abundance"], sprm, cfg.min_score, cfg.n_types, nprm,
                          want_null=True, want_scores=True, want_order=True)
    finally:
        m.free()
    del ds
    rep, _ = check_analysis(oracle, d, cfg.n_genes, cfg.n_types, out, score_params=_score_kw(cfg, 3),
                            min_score=cfg.min_score, canonical=2, seed=4, perm_indices=(0, 31, cfg.n_perm - 1),
                            what="config D shape")
    assert rep["n_used"] >= 1_000_000


def test_philox_null_matches_jdk_shuffle_null_at_config_B_size(ctx, oracle):
    """north_star: 'in the Philox mode [nulls] must be within a stated statistical tolerance (KS test on the null
    distributions)'. At config B's size (219 530 ranked cells, 100 cell types), 2000 permutations each, all on the GPU:
        A  REPLAY of java.util.Collections.shuffle index arrays (the reference's permutation source) + EXACT arithmetic
        B  PHILOX (keyed Feistel on the size-sorted arrangement)                                    + EXACT arithmetic
        C  PHILOX                                                                                    + FAST arithmetic
    Stated tolerance, per cell type with Bonferroni over the types: two-sample KS statistic below the alpha = 0.001
    critical value; the numbers of |ES| values of A and of B / C above the pooled 99th and 99.9th percentiles within
    4.5 sigma of each other (Binomial(total, 1/2); p-values and FDR live in that tail), per type and summed over the
    types; the mean over the types of log(sd_B / sd_A) within 1.5 %; and the empirical p-value of the observed score
    from B / C within 4 * sqrt(2 p (1 - p) / P) + 1 / P of the one from A."""
    n_used, T, Pn = 219_530, 100, 2000
    s, lab, order = _ranked(n_used, T, seed=5)
    rng_seed = 424242
    replay = np.empty((Pn, n_used), np.int32)
    ident = np.arange(n_used, dtype=np.int32)
    jr = oracle.JRandom(rng_seed)          # one Random for all permutations, as PCTSEA.java:1398-1405
    for p in range(Pn):
        replay[p] = jr.shuffle(ident)
    prmA = ctx.null_params(Pn, perm_mode=P.PERM_REPLAY, ks_mode=P.KS_EXACT)
    resA, nullA, _ = ctx.enrich(s, order, lab, T, prmA, replay=replay)
    del replay
    prmB = ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_EXACT, seed=99)
    resB, nullB, _ = ctx.enrich(s, order, lab, T, prmB)
    prmC = ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=99)
    resC, nullC, _ = ctx.enrich(s, order, lab, T, prmC)
    # spot check of A against the oracle's own walk of the same shuffled labels (first and last permutation)
    jr = oracle.JRandom(rng_seed)
    first = jr.shuffle(ident)
    assert_bits_equal_f32(nullA[:, 0], oracle.ks_null_one(s, lab[first], T)[0], "REPLAY null, permutation 0")

    alpha = 0.001 / T
    d_crit = np.sqrt(-np.log(alpha / 2.0) / 2.0) * np.sqrt(2.0 / Pn)
    z = 4.5
    worst = {"ks": 0.0, "dp": 0.0}
    above = {("PHILOX+EXACT", 0.99): [0, 0], ("PHILOX+EXACT", 0.999): [0, 0], ("PHILOX+FAST", 0.99): [0, 0], ("PHILOX+FAST", 0.999): [0, 0]}
    log_sd = {"PHILOX+EXACT": [], "PHILOX+FAST": []}
    for t in range(T):
        a = np.sort(nullA[t][~np.isnan(nullA[t])])
        if a.size < Pn:
            continue
        for name, other in (("PHILOX+EXACT", nullB[t]), ("PHILOX+FAST", nullC[t])):
            b = np.sort(other)
            allv = np.concatenate([a, b])
            d = np.max(np.abs(np.searchsorted(a, allv, side="right") / a.size - np.searchsorted(b, allv, side="right") / b.size))
            worst["ks"] = max(worst["ks"], d)
            assert d < d_crit, f"type {t} {name}: two-sample KS {d:.4f} >= {d_crit:.4f}"
            # the upper tail of |ES| (p-values and FDR live there): both samples against the POOLED quantile; given
            # the total above it, the count from one sample is Binomial(total, 1/2)
            aa, bb = np.abs(a), np.abs(b)
            for q in (0.99, 0.999):
                thr = np.quantile(np.concatenate([aa, bb]), q)
                na, nb = int(np.count_nonzero(aa > thr)), int(np.count_nonzero(bb > thr))
                assert abs(na - nb) <= z * np.sqrt(na + nb) + 1, f"type {t} {name}: {na} vs {nb} values above the pooled {q} quantile"
                above[(name, q)][0] += na
                above[(name, q)][1] += nb
            log_sd[name].append(np.log(np.std(b) / np.std(a)))
            # empirical p of the observed score (a10: positive side, strict comparisons)
            real = resA["ews"][t]
            if real == real and real >= 0:
                pa = np.count_nonzero(a > real) / max(np.count_nonzero(a > 0), 1)
                pb = np.count_nonzero(b > real) / max(np.count_nonzero(b > 0), 1)
                pm = 0.5 * (pa + pb)
                worst["dp"] = max(worst["dp"], abs(pa - pb))
                assert abs(pa - pb) <= 4.0 * np.sqrt(2.0 * pm * (1.0 - pm) / Pn) + 1.0 / Pn, f"type {t} {name}: p_emp {pa} vs {pb}"
    # over all types together the tail counts and the spread of the null must agree much more closely: a keyed
    # permutation with too few rounds inflates the null's standard deviation by 2 % (profiles/feistel_rounds_check.txt),
    # which moves the mean log ratio of the standard deviations 9 standard errors away from 0 here
    for (name, q), (na, nb) in above.items():
        assert abs(na - nb) <= z * np.sqrt(na + nb) + 1, f"{name}: {na} vs {nb} values above the pooled {q} quantiles, all types"
    for name, v in log_sd.items():
        assert len(v) >= T // 2 and abs(np.mean(v)) <= 0.015, f"{name}: mean log sd ratio {np.mean(v):+.4f} over {len(v)} types"
    assert_results_equal(resA, resB, OBSERVED_FIELDS, "observed fields do not depend on the permutation source")
    print(f"worst two-sample KS {worst['ks']:.4f} (critical {d_crit:.4f}), worst |dp| {worst['dp']:.4f}, "
          f"mean log sd ratio {np.mean(log_sd['PHILOX+EXACT']):+.4f} / {np.mean(log_sd['PHILOX+FAST']):+.4f}, tail counts {above}")


def test_observed_history_slices_are_invisible(ctx, oracle):
    """The observed accumulator history (8 bytes per ranked cell and type) is kept for a slice of the types at a time
    (PCTSEA_OPT_OBS_CHAIN_MB; an atlas-scale ranking does not fit otherwise). With a 1 MiB budget a 20 000-cell, 40-type
    ranking goes through 7 slices: every observed field equals the one-slice run and the oracle, the null is unaffected,
    and pctsea_last_observed_curve returns the same series for a type of an evicted slice (it is walked again) as for
    one of the resident slice."""
    n_used, T = 20_000, 40
    s, lab, order = _ranked(n_used, T, seed=21)
    prm = ctx.null_params(8, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=3)
    ref, ref_null, _ = ctx.enrich(s, order, lab, T, prm)
    curves = {t: ctx.last_observed_curve(t, n_used) for t in (0, 17, T - 1)}
    ctx.set_option(P.OPT_OBS_CHAIN_MB, 1)
    try:
        res, null, _ = ctx.enrich(s, order, lab, T, prm)
        assert_results_equal(res, ref, None, "sliced history")
        assert_results_equal(res, oracle.ks_observed(s, lab, T), OBSERVED_FIELDS, "sliced history vs oracle")
        assert_bits_equal_f32(null, ref_null, "null beside a sliced observed stage")
        for t in (T - 1, 0, 17, 0):   # resident slice first, then evicted ones (each walk makes its type resident)
            a, b = ctx.last_observed_curve(t, n_used)
            assert_bits_equal_f32(a, curves[t][0], f"a series, type {t}")
            assert_bits_equal_f32(b, curves[t][1], f"b series, type {t}")
    finally:
        ctx.set_option(P.OPT_OBS_CHAIN_MB, 0)


def test_null_is_reproducible_run_to_run(ctx):
    """compute-sanitizer (racecheck) is not available on the GPU pool this repository is developed on, so the
    data-race question for the fused null kernel — per-thread shared-memory read-modify-writes without barriers, per-type
    maxima read stale through a volatile load and raised with atomicMax, CTAs taking permutations from a ticket counter
    in whatever order they run — is asked of the results instead: every outcome that a race could change would show
    up as a difference between runs or between geometries. Ten runs of the same 1000-permutation null (long-tailed
    types, 20 000 cells) and the same null with 128 and 512 threads per permutation are bit-identical."""
    n_used, T, Pn = 20_000, 60, 1000
    s, lab, order = _ranked(n_used, T, seed=33)
    rng = np.random.default_rng(1)
    w = 1.0 / np.arange(1, T + 1) ** 2
    lab = rng.choice(T, size=n_used, p=w / w.sum()).astype(np.int32)   # many types of a few dozen cells
    lab[rng.random(n_used) < 0.002] = -1
    prm = ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=8)
    _, ref, refd = ctx.enrich(s, order, lab, T, prm)
    for _ in range(9):
        _, n, nd = ctx.enrich(s, order, lab, T, prm)
        assert np.array_equal(n.view(np.uint32), ref.view(np.uint32)) and np.array_equal(nd.view(np.uint32), refd.view(np.uint32))
    for threads in (128, 512):
        ctx.set_option(P.OPT_NULL_THREADS, threads)
        try:
            _, n, _ = ctx.enrich(s, order, lab, T, prm)
        finally:
            ctx.set_option(P.OPT_NULL_THREADS, 0)
        assert np.array_equal(n.view(np.uint32), ref.view(np.uint32)), f"{threads} threads per permutation"
