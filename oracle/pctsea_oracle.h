/*
 * pctsea_oracle.h — CPU restatement of the PCTSEA hot path. TEST INFRASTRUCTURE ONLY.
 *
 * This is the checker, not the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it. Nothing under pctsea_parent_b200/ links or calls it.
 *
 * PARITY UNPINNED: the reference (proteomicsyates/pctsea-parent) is Java 8 + MongoDB, ships no numeric
 * test, golden vector or fixture for this path, and cannot be compiled or run in this image (no JDK, four
 * private -SNAPSHOT dependencies). Every function below restates the cited reference lines by reading
 * them; the JDK pieces (java.util.Random, Collections.shuffle, Arrays.binarySearch) are pinned by the
 * published JDK known answers in tests/golden/. commons-math3 3.x SimpleRegression/PearsonsCorrelation and
 * edu.scripps.yates Maths.mean/var are un-vendored third-party code restated from their published
 * algorithm (see DESIGN.md "Oracle").
 *
 * Paths cited as CORE/… are relative to
 * /root/reference/pctsea-core/src/main/java/edu/scripps/yates/pctsea/.
 */
#ifndef PCTSEA_ORACLE_H
#define PCTSEA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- java.util.Random (JDK 8) ------------------------------------------------------------- */
typedef struct { uint64_t seed; } orc_jrandom;
void    orc_jrandom_init(orc_jrandom* r, int64_t seed);
int32_t orc_jrandom_next(orc_jrandom* r, int bits);
int32_t orc_jrandom_next_int(orc_jrandom* r);
int32_t orc_jrandom_next_int_bound(orc_jrandom* r, int32_t bound);
double  orc_jrandom_next_double(orc_jrandom* r);
/* Collections.shuffle(list, rnd) on an int32 array (RandomAccess branch). */
void    orc_java_shuffle_i32(int32_t* a, int64_t n, orc_jrandom* r);
/* Arrays.binarySearch(float[], key) */
int32_t orc_java_binary_search_f32(const float* a, int32_t n, float key);
/* Arrays.sort(float[]) total order (-0.0 < 0.0, NaN last) */
void    orc_java_sort_f32(float* a, int64_t n);
/* Double.compare */
int     orc_java_double_compare(double a, double b);

/* ---- Philox4x32-10 and the keyed Feistel label permutation (the repo's own PHILOX mode) --- */
void     orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* U[0,1) with 53 bits, as a pure function of (seed, cell, vec, i) — the deterministic stand-in for
 * Math.random() in CORE/model/SingleCell.java:436-447. */
double   orc_jitter_uniform(uint64_t seed, uint64_t cell, uint32_t vec, uint32_t i);
/* perm_out[i] = pi_p(i), a bijection of [0,n) built from `rounds` Feistel rounds with arithmetic round
 * functions, one Philox-generated 32-bit key per round, keyed by (seed, p). */
void     orc_feistel_perm(uint64_t seed, uint32_t p, int64_t n, int rounds, int32_t* perm_out);
void     orc_feistel_keys(uint64_t seed, uint32_t p, int rounds, uint32_t* keys_out);
/* Permuted labels of PHILOX mode: lab_perm_out[x] = class at index pi_p(x) of the size-sorted arrangement of
 * lab[0..Nu) (types 0..T-1 ascending, unknown-type cells (-1) last). */
void     orc_set_mix_variant(int v);   /* profiles/feistel_mix_check.py only */
void     orc_perm_labels(uint64_t seed, uint32_t p, int64_t Nu, const int32_t* lab, int32_t T, int rounds,
                         int32_t* lab_perm_out);

/* ---- stage 1: per-cell scoring (a1-a3) ----------------------------------------------------- */
typedef struct {
    int32_t  method;          /* 0 = PEARSONS_CORRELATION, 1 = SIMPLE_SCORE */
    int32_t  min_genes_cells; /* a2 */
    double   min_corr;        /* a3: r < min_corr -> NaN */
    int32_t  has_min_corr;    /* Double minCorr != null */
    int32_t  jitter_mode;     /* 0 = zero-variance cells score NaN; 1 = Philox jitter U[0,1)/1e6 */
    uint64_t seed;
} orc_score_params;

/* CSR: row_ptr int64[N+1], col int32[nnz], val fp32[nnz]. gene_idx[L] = matrix column of list entry l
 * (-1 = not in the dataset), abundance[L] fp32. Pair order = CSR storage order of the row.
 * Outputs: score fp64[N] (NaN if not kept), n_genes_used int32[N] (#pairs x>0&&y>0),
 * n_nonzero int32[N] (#list genes with x != 0), kept u8[N]. Returns 0, or -5 if
 * min_genes_cells > L (CORE/model/SingleCell.java:347-352). */
int orc_score(int64_t n_cells, int32_t n_genes, const int64_t* row_ptr, const int32_t* col,
              const float* val, int32_t L, const int32_t* gene_idx, const float* abundance,
              const orc_score_params* prm, double* score, int32_t* n_genes_used,
              int32_t* n_nonzero, uint8_t* kept);

/* Same with canonical == 1: the CSR-scan kernel's canonical reduction order (orc_pearson_strided32);
 * canonical == 2: the gene-major kernel's (orc_pearson_shifted over the pairs in ascending gene order). */
int orc_score2(int64_t n_cells, int32_t n_genes, const int64_t* row_ptr, const int32_t* col,
               const float* val, int32_t L, const int32_t* gene_idx, const float* abundance,
               const orc_score_params* prm, int canonical, double* score, int32_t* n_genes_used,
               int32_t* n_nonzero, uint8_t* kept);

/* commons-math3 PearsonsCorrelation.correlation(x, y) via SimpleRegression (restated; sequential,
 * reference-faithful). */
double orc_pearson(const double* x, const double* y, int32_t n);
/* Canonical Pearson of the accelerated path: two-pass centred sums over 32 strided sub-streams
 * (pair k -> k % 32) added pairwise in a fixed tree, then SimpleRegression's getR() formula. */
double orc_pearson_strided32(const double* x, const double* y, int32_t n);
/* canonical == 2 in orc_score2: pairs in ascending gene order, sequential sums shifted by the first pair. */
double orc_pearson_shifted(const double* x, const double* y, int32_t n);

/* ---- stage 2: threshold + rank (a4, a5, a7) ------------------------------------------------ */
/* order_out[n_pass]: the first n_pass-1 entries are the KS list in its final order (score desc by
 * Double.compare, ties by cell index asc); entry n_pass-1 is the cell the reference drops
 * (CORE/PCTSEA.java:421-424). Returns n_pass. */
int64_t orc_rank(int64_t n_cells, const double* score, const uint8_t* kept, double min_score,
                 int32_t* order_out);

/* ---- stage 3: weighted KS (a8), null (a9), p / norm / FDR (a10-a12) ------------------------ */
typedef struct {
    float   ews;        /* observed, compensated (a8 step 5) */
    float   dstat;      /* from the raw supremum */
    int32_t supX;       /* 1-based; -1 when NaN */
    double  norm_supX;  /* supX / Nu; -1 when NaN */
    int32_t sizeA, sizeB;
    float   raw_sup;    /* supremum before compensation */
    float   norm_ews;   /* a11 */
    double  p_emp;      /* a10 */
    double  fdr;        /* a12 */
    /* observed-only extras */
    float   sec_ews;    /* secondary supremum, ...Parallel.java:244-312 (0 when none) */
    int32_t sec_supX;   /* 1-based; 0 when none */
    double  ks_d;       /* two-sample KS D of the type's scores vs the others (:211-212), NaN when invalid */
} orc_type_result;

/* commons-math3 3.x KolmogorovSmirnovTest.kolmogorovSmirnovStatistic(x, y) (restated: sort both, merge walk
 * with ties processed together, integer running difference). */
double orc_ks_statistic(const double* x, int64_t n, const double* y, int64_t m);

/* s[Nu] ranked scores, lab[Nu] ranked labels in [-1, T). */
void orc_ks_observed(int64_t Nu, const double* s, const int32_t* lab, int32_t T, orc_type_result* out);
/* One permutation: null_ews[T], null_d[T] (raw supremum and dStat; NaN as in the reference). */
void orc_ks_null_one(int64_t Nu, const double* s, const int32_t* lab_perm, int32_t T,
                     float* null_ews, float* null_d);
/* perm_mode 0: JDK shuffle with one Random(seed) stream across all permutations (writes the index
 * arrays to replay_out[P][Nu] if non-NULL); 1: Feistel/Philox permutation (perm index = perm_begin+p).
 * null_ews/null_d are [T][P] row-major. nthreads>1 mirrors the reference's threads-over-types,
 * re-created per permutation (CORE/PCTSEA.java:2501-2519). */
void orc_null(int64_t Nu, const double* s, const int32_t* lab, int32_t T, int32_t P, int perm_mode,
              uint64_t seed, int32_t perm_begin, int rounds, int nthreads, int32_t* replay_out,
              float* null_ews, float* null_d);
/* Null from supplied index arrays replay[P][Nu] (label at rank i = lab[replay[p][i]]). */
void orc_null_replay(int64_t Nu, const double* s, const int32_t* lab, int32_t T, int32_t P,
                     const int32_t* replay, float* null_ews, float* null_d);
/* a10-a12 on a supplied null. mean_policy 0 = fp32 sequential sum / count (default), 1 = fp64 sum. */
void orc_reduce(int32_t T, int32_t P, const float* null_ews, int mean_policy, orc_type_result* inout,
                float* norm_null_out /* [T][P] or NULL */);

/* The FAST mode's definition: int32 fixed-point scores (scale 2^frac_bits, orc_fast_frac_bits: the largest
 * scale that keeps the sum of |fx| over the ranking below 2^31), exact integer running sums, supremum at hit /
 * pre-hit positions, one fp32 fused multiply-add per candidate. */
void orc_ks_null_fast_one(int64_t Nu, const double* s, const int32_t* lab_perm, int32_t T,
                          int frac_bits, float* null_ews);
int  orc_fast_frac_bits(int64_t Nu, const double* s);

#ifdef __cplusplus
}
#endif
#endif
