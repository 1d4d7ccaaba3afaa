/*
 * pctsea_oracle.c — CPU restatement of the PCTSEA hot path. TEST INFRASTRUCTURE ONLY; see the header.
 * PARITY UNPINNED (no reference golden vectors exist; reference is not runnable here).
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off -fPIC -shared -pthread (oracle/Makefile). No -ffast-math:
 * every float/double rounding below is part of the restated semantics.
 *
 * CORE/ = /root/reference/pctsea-core/src/main/java/edu/scripps/yates/pctsea/
 */
#define _POSIX_C_SOURCE 200809L
#include "pctsea_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ============================================================================================
 * JDK pieces (bit-exact restatements of OpenJDK 8 library code; used by the reference at
 * CORE/PCTSEA.java:1405 (Collections.shuffle), :1523-1537 (Arrays.sort / binarySearch),
 * CORE/utils/PCTSEAUtils.java:44-53 (Double.compare)).
 * ========================================================================================== */

#define JR_MULT 0x5DEECE66DULL
#define JR_ADD  0xBULL
#define JR_MASK ((1ULL << 48) - 1)

void orc_jrandom_init(orc_jrandom* r, int64_t seed) { r->seed = ((uint64_t)seed ^ JR_MULT) & JR_MASK; }

int32_t orc_jrandom_next(orc_jrandom* r, int bits) {
    r->seed = (r->seed * JR_MULT + JR_ADD) & JR_MASK;
    /* (int)(seed >>> (48 - bits)) */
    return (int32_t)(uint32_t)(r->seed >> (48 - bits));
}

int32_t orc_jrandom_next_int(orc_jrandom* r) { return orc_jrandom_next(r, 32); }

int32_t orc_jrandom_next_int_bound(orc_jrandom* r, int32_t bound) {
    /* JDK 8 Random.nextInt(int bound) */
    int32_t rr = orc_jrandom_next(r, 31);
    int32_t m = bound - 1;
    if ((bound & m) == 0) {
        rr = (int32_t)(((int64_t)bound * (int64_t)rr) >> 31);
    } else {
        int32_t u = rr;
        for (;;) {
            rr = u % bound;
            /* u - rr + m < 0 with int32 wrap-around */
            int32_t t = (int32_t)((uint32_t)u - (uint32_t)rr + (uint32_t)m);
            if (t >= 0) break;
            u = orc_jrandom_next(r, 31);
        }
    }
    return rr;
}

double orc_jrandom_next_double(orc_jrandom* r) {
    int64_t hi = (int64_t)orc_jrandom_next(r, 26);
    int64_t lo = (int64_t)orc_jrandom_next(r, 27);
    return (double)((hi << 27) + lo) * (1.0 / (double)(1LL << 53));
}

void orc_java_shuffle_i32(int32_t* a, int64_t n, orc_jrandom* r) {
    /* Collections.shuffle: for (int i=size; i>1; i--) swap(list, i-1, rnd.nextInt(i)); */
    for (int64_t i = n; i > 1; i--) {
        int32_t j = orc_jrandom_next_int_bound(r, (int32_t)i);
        int32_t t = a[i - 1];
        a[i - 1] = a[j];
        a[j] = t;
    }
}

static inline int32_t f32_bits(float f) { int32_t b; memcpy(&b, &f, 4); return b; }
static inline float bits_f32(int32_t b) { float f; memcpy(&f, &b, 4); return f; }
static inline int64_t f64_bits(double d) { int64_t b; memcpy(&b, &d, 8); return b; }

/* Float.floatToIntBits: canonical NaN */
static inline int32_t java_float_to_int_bits(float f) { return isnan(f) ? 0x7fc00000 : f32_bits(f); }

int32_t orc_java_binary_search_f32(const float* a, int32_t n, float key) {
    /* Arrays.binarySearch0(float[] a, 0, n, key) */
    int32_t low = 0, high = n - 1;
    while (low <= high) {
        int32_t mid = (int32_t)(((uint32_t)low + (uint32_t)high) >> 1);
        float midVal = a[mid];
        if (midVal < key) low = mid + 1;
        else if (midVal > key) high = mid - 1;
        else {
            int32_t midBits = java_float_to_int_bits(midVal);
            int32_t keyBits = java_float_to_int_bits(key);
            if (midBits == keyBits) return mid;
            else if (midBits < keyBits) low = mid + 1;
            else high = mid - 1;
        }
    }
    return -(low + 1);
}

/* Float.compare total order: -0.0 < 0.0, NaN greater than everything (all NaNs equal). */
static int java_float_compare(float a, float b) {
    if (a < b) return -1;
    if (a > b) return 1;
    int32_t ab = java_float_to_int_bits(a), bb = java_float_to_int_bits(b);
    return ab == bb ? 0 : (ab < bb ? -1 : 1);
}
static int cmp_f32_java(const void* pa, const void* pb) {
    return java_float_compare(*(const float*)pa, *(const float*)pb);
}
void orc_java_sort_f32(float* a, int64_t n) { qsort(a, (size_t)n, sizeof(float), cmp_f32_java); }

int orc_java_double_compare(double a, double b) {
    if (a < b) return -1;
    if (a > b) return 1;
    int64_t ab = isnan(a) ? 0x7ff8000000000000LL : f64_bits(a);
    int64_t bb = isnan(b) ? 0x7ff8000000000000LL : f64_bits(b);
    return ab == bb ? 0 : (ab < bb ? -1 : 1);
}

/* ============================================================================================
 * Philox4x32-10 (Salmon et al. 2011) — the repo's own counter-based generator: deterministic
 * jitter stream and the PHILOX-mode label permutation. Not part of the reference.
 * ========================================================================================== */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double orc_jitter_uniform(uint64_t seed, uint64_t cell, uint32_t vec, uint32_t i) {
    uint32_t ctr[4] = { i, vec | 0x4A000000u, (uint32_t)cell, (uint32_t)(cell >> 32) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t o[4];
    orc_philox4x32_10(ctr, key, o);
    uint64_t u = ((uint64_t)(o[0] >> 6) << 27) | (uint64_t)(o[1] >> 5);
    return (double)u * (1.0 / 9007199254740992.0);
}

static void feistel_dims(int64_t n, int64_t* a_out, int64_t* b_out) {
    /* a0 = ceil(sqrt n); b(a) = ceil(n/a) rounded up to a multiple of 4; the a in [a0, a0 + max(a0/16, 8)] with the
     * fewest out-of-range points a*b(a) - n (first on ties) */
    int64_t a0 = (int64_t)floor(sqrt((double)n));
    while (a0 * a0 < n) a0++;
    while (a0 > 1 && (a0 - 1) * (a0 - 1) >= n) a0--;
    if (a0 < 1) a0 = 1;
#define ORC_FEISTEL_B(a) (((n + (a) - 1) / (a) + 3) / 4 * 4)
    int64_t best_a = a0, best_w = a0 * ORC_FEISTEL_B(a0) - n;
    int64_t span = a0 / 16 > 8 ? a0 / 16 : 8;
    for (int64_t a = a0 + 1; a <= a0 + span && best_w > 0; a++) {
        int64_t w = a * ORC_FEISTEL_B(a) - n;
        if (w < best_w) { best_w = w; best_a = a; }
    }
    *a_out = best_a;
    *b_out = ORC_FEISTEL_B(best_a);
#undef ORC_FEISTEL_B
}

/* PHILOX-mode permutation, round 2 definition: a keyed Feistel bijection of [0,n) on Z_a x Z_b (a near sqrt n,
 * b = ceil(n / a), see feistel_dims; cycle-walking for the few out-of-range points) whose round functions are ARITHMETIC:
 *   F(v, key, mod) = floor(mix(v, key) * mod / 2^32),  mix = multiply - shift - multiply hash of v and the round key,
 * with one 32-bit key per round from Philox4x32-10(ctr = {j / 4, 0, p, "FEIS"}, key = seed)[j % 4]. No tables:
 * the GPU evaluates it in registers inside the null kernel. Round j modifies the high digit L iff (rounds-1-j) is
 * even, i.e. the LAST round always does: the class of a position is a range test on pi(x) (orc_perm_labels), decided
 * by its high digit, and a trailing low-digit round would be wasted. Measured against Collections.shuffle on the null
 * of the enrichment score (profiles/feistel_rounds_check.txt): 3 and 4 rounds inflate its standard deviation by 2 %
 * (the value histogram of a random round FUNCTION is uneven, and with the sorted arrangement that unevenness reaches
 * the class counts of long prefixes of the ranking); 5 and more rounds are indistinguishable. Default 5. */
/* mix: t = v * M + key spreads the digit over the word, t * (t >> 16) folds its high half back in non-linearly (three
 * GPU instructions). g_mix_variant selects other hashes for profiles/feistel_mix_check.py only (1: the multiply -
 * xorshift - multiply hash this replaced, statistically equivalent; 3: t * t, which fails on short rankings). */
static int g_mix_variant = 0;
void orc_set_mix_variant(int v) { g_mix_variant = v; }
static inline uint32_t feistel_mix(uint32_t v, uint32_t key) {
    uint32_t t = v * 0x9E3779B1u + key;
    if (g_mix_variant == 1) { t ^= t >> 15; return t * 0x2C1B3C6Du; }
    if (g_mix_variant == 3) return t * t;
    return t * (t >> 16);
}

void orc_feistel_keys(uint64_t seed, uint32_t p, int rounds, uint32_t* keys_out) {
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    for (int j0 = 0; j0 < rounds; j0 += 4) {
        uint32_t ctr[4] = { (uint32_t)(j0 >> 2), 0u, p, 0x46454953u };
        uint32_t o[4];
        orc_philox4x32_10(ctr, key, o);
        for (int q = 0; q < 4 && j0 + q < rounds; q++) keys_out[j0 + q] = o[q];
    }
}

void orc_feistel_perm(uint64_t seed, uint32_t p, int64_t n, int rounds, int32_t* perm_out) {
    if (n <= 0) return;
    int64_t a64, b64;
    feistel_dims(n, &a64, &b64);
    const uint32_t a = (uint32_t)a64, b = (uint32_t)b64;
    uint32_t keys[64];
    orc_feistel_keys(seed, p, rounds, keys);
    for (int64_t x = 0; x < n; x++) {
        uint32_t L = (uint32_t)(x / b), R = (uint32_t)(x % b);
        int64_t y;
        do {
            for (int j = 0; j < rounds; j++) {
                if (!((rounds - 1 - j) & 1)) { L += (uint32_t)(((uint64_t)feistel_mix(R, keys[j]) * a) >> 32); if (L >= a) L -= a; }
                else          { R += (uint32_t)(((uint64_t)feistel_mix(L, keys[j]) * b) >> 32); if (R >= b) R -= b; }
            }
            y = (int64_t)L * b + R;
        } while (y >= n);
        perm_out[x] = (int32_t)y;
    }
}

/* Permuted labels of PHILOX mode: position x carries the class at index pi_p(x) of the SIZE-SORTED arrangement of
 * the ranked labels: the T+1 classes (types 0..T-1 and the unknown-type class T) in ascending order of their number
 * of ranked cells, ties by class id, every class one contiguous run. A uniform permutation applied to any fixed
 * arrangement of the same multiset is a uniform arrangement, so which arrangement is permuted does not matter for
 * the null; a sorted one turns "class of index j" into a search in the T+1 cumulative class counts, and sorting by
 * size puts all the small classes (the ones that need a fine-grained look-up table) into one short prefix. */
void orc_size_sorted_classes(const int64_t* cnt, int32_t Tc, int32_t* order_out) {
    /* order_out[s] = class at sorted position s: rank by (count, id), O(Tc^2) like the device table builder */
    for (int32_t c = 0; c < Tc; c++) {
        int32_t r = 0;
        for (int32_t d = 0; d < Tc; d++) r += (cnt[d] < cnt[c] || (cnt[d] == cnt[c] && d < c)) ? 1 : 0;
        order_out[r] = c;
    }
}

void orc_perm_labels(uint64_t seed, uint32_t p, int64_t Nu, const int32_t* lab, int32_t T, int rounds,
                     int32_t* lab_perm_out) {
    if (Nu <= 0) return;
    int64_t* cnt = (int64_t*)calloc((size_t)T + 1, sizeof(int64_t));
    for (int64_t i = 0; i < Nu; i++) { int32_t c = (lab[i] < 0 || lab[i] >= T) ? T : lab[i]; cnt[c]++; }
    int32_t* order = (int32_t*)malloc(((size_t)T + 1) * sizeof(int32_t));
    orc_size_sorted_classes(cnt, T + 1, order);
    int32_t* sorted = (int32_t*)malloc((size_t)Nu * sizeof(int32_t));
    int64_t j = 0;
    for (int32_t s = 0; s <= T; s++) {
        const int32_t c = order[s];
        for (int64_t i = 0; i < cnt[c]; i++) sorted[j++] = (c == T) ? -1 : c;
    }
    int32_t* idx = (int32_t*)malloc((size_t)Nu * sizeof(int32_t));
    orc_feistel_perm(seed, p, Nu, rounds, idx);
    for (int64_t x = 0; x < Nu; x++) lab_perm_out[x] = sorted[idx[x]];
    free(cnt); free(order); free(sorted); free(idx);
}

/* ============================================================================================
 * Stage 1 — scoring. CORE/model/SingleCell.java:344-369 (a2), :380-463 (a3).
 * ========================================================================================== */

/* commons-math3 3.x: PearsonsCorrelation.correlation(xArray, yArray) feeds a SimpleRegression
 * (with intercept) pair by pair and returns getR(). Restated from the published algorithm
 * (SimpleRegression.addData / getSlope / getSumSquaredErrors / getRSquare / getR). */
double orc_pearson(const double* x, const double* y, int32_t n) {
    if (n < 2) return NAN; /* reference: MathIllegalArgumentException (run aborts) */
    double sumXX = 0.0, sumYY = 0.0, sumXY = 0.0, xbar = 0.0, ybar = 0.0;
    int64_t cnt = 0;
    for (int32_t i = 0; i < n; i++) {
        if (cnt == 0) {
            xbar = x[i];
            ybar = y[i];
        } else {
            double fact1 = 1.0 + (double)cnt;
            double fact2 = (double)cnt / (1.0 + (double)cnt);
            double dx = x[i] - xbar;
            double dy = y[i] - ybar;
            sumXX += dx * dx * fact2;
            sumYY += dy * dy * fact2;
            sumXY += dx * dy * fact2;
            xbar += dx / fact1;
            ybar += dy / fact1;
        }
        cnt++;
    }
    /* getSlope(): n<2 -> NaN; |sumXX| < 10*Double.MIN_VALUE -> NaN; else sumXY/sumXX */
    double b1;
    if (fabs(sumXX) < 10 * 4.9e-324) b1 = NAN; else b1 = sumXY / sumXX;
    /* getSumSquaredErrors(): FastMath.max(0d, sumYY - sumXY*sumXY/sumXX)  (NaN-propagating max) */
    double sse_arg = sumYY - sumXY * sumXY / sumXX;
    double sse;
    if (isnan(sse_arg)) sse = NAN; else sse = (sse_arg > 0.0) ? sse_arg : 0.0;
    /* getRSquare(): (ssto - sse)/ssto with ssto = sumYY */
    double rsq = (sumYY - sse) / sumYY;
    double result = sqrt(rsq);
    if (b1 < 0) result = -result;
    return result;
}

/* The accelerated path's canonical Pearson (this repo's choice, not the reference's): the textbook
 * two-pass centred sums, reduced the way a warp does it. Pair k of the cell (CSR order) goes to
 * sub-stream k % 32; pass 1 sums x and y per sub-stream (in k order) and adds the 32 partials pairwise
 * (offsets 16, 8, 4, 2, 1) to get the means; pass 2 does the same for (x-xbar)^2, (y-ybar)^2 and
 * (x-xbar)(y-ybar); the result goes through the same getR() formula as SimpleRegression. The reference's own
 * pair order is a Trove hash-set order (GeneExpressionsRetriever.java:142,266), so no summation order is
 * "the" reference order; this function agrees with orc_pearson to ~1e-14 relative (tested) and the GPU
 * kernel is bit-identical to it. */
static double tree32(double* v) {
    for (int o = 16; o >= 1; o >>= 1)
        for (int l = 0; l < o; l++) v[l] = v[l] + v[l + o];
    return v[0];
}

static double pearson_finish(double sumXX, double sumYY, double sumXY) {
    /* SimpleRegression.getSlope / getSumSquaredErrors / getRSquare / getR */
    double b1;
    if (fabs(sumXX) < 10 * 4.9e-324) b1 = NAN; else b1 = sumXY / sumXX;
    double sse_arg = sumYY - sumXY * sumXY / sumXX;
    double sse;
    if (isnan(sse_arg)) sse = NAN; else sse = (sse_arg > 0.0) ? sse_arg : 0.0;
    double rsq = (sumYY - sse) / sumYY;
    double result = sqrt(rsq);
    if (b1 < 0) result = -result;
    return result;
}

double orc_pearson_strided32(const double* x, const double* y, int32_t n) {
    if (n < 2) return NAN;
    double sx[32], sy[32], sxx[32], syy[32], sxy[32];
    for (int l = 0; l < 32; l++) { sx[l] = 0.0; sy[l] = 0.0; sxx[l] = 0.0; syy[l] = 0.0; sxy[l] = 0.0; }
    for (int32_t k = 0; k < n; k++) { sx[k & 31] += x[k]; sy[k & 31] += y[k]; }
    double xbar = tree32(sx) / (double)n, ybar = tree32(sy) / (double)n;
    for (int32_t k = 0; k < n; k++) {
        double dx = x[k] - xbar, dy = y[k] - ybar;
        sxx[k & 31] += dx * dx;
        syy[k & 31] += dy * dy;
        sxy[k & 31] += dx * dy;
    }
    double SXX = tree32(sxx), SYY = tree32(syy), SXY = tree32(sxy);
    return pearson_finish(SXX, SYY, SXY);
}

/* The gene-major scoring path's canonical Pearson (this repo's choice, see orc_pearson_strided32 for why a
 * choice is needed): pairs in ASCENDING GENE (matrix column) order, one sequential pass of sums shifted by the
 * first pair (dx = x - x[0], dy = y - y[0]: the textbook shifted-data algorithm, well conditioned because the
 * pivot is a sample), then Sxx = sum(dx^2) - sum(dx)^2/n (same for yy, xy) through the same getR() formula.
 * What the GPU's lane-per-cell gene-major kernel computes, bit for bit. */
double orc_pearson_shifted(const double* x, const double* y, int32_t n) {
    if (n < 2) return NAN;
    const double x0 = x[0], y0 = y[0];
    double sdx = 0.0, sdy = 0.0, sxx = 0.0, syy = 0.0, sxy = 0.0;
    for (int32_t i = 0; i < n; i++) {
        double dx = x[i] - x0, dy = y[i] - y0;
        sdx += dx; sdy += dy;
        sxx += dx * dx; syy += dy * dy; sxy += dx * dy;
    }
    const double nn = (double)n;
    double SXX = sxx - sdx * sdx / nn, SYY = syy - sdy * sdy / nn, SXY = sxy - sdx * sdy / nn;
    return pearson_finish(SXX, SYY, SXY);
}

typedef struct { int32_t g; double x, y; } gene_pair;
static int cmp_gene_pair(const void* a, const void* b) {
    int32_t ga = ((const gene_pair*)a)->g, gb = ((const gene_pair*)b)->g;
    return (ga > gb) - (ga < gb);
}

static int all_equal(const double* v, int32_t n) {
    /* edu.scripps.yates Maths.var(double[]) is un-vendored; only "== 0" is used
     * (CORE/model/SingleCell.java:430-441): policy = zero iff every element is equal (n>=1). */
    if (n < 1) return 0;
    for (int32_t i = 1; i < n; i++) if (v[i] != v[0]) return 0;
    return 1;
}

int orc_score2(int64_t n_cells, int32_t n_genes, const int64_t* row_ptr, const int32_t* col,
               const float* val, int32_t L, const int32_t* gene_idx, const float* abundance,
               const orc_score_params* prm, int canonical, double* score, int32_t* n_genes_used,
               int32_t* n_nonzero, uint8_t* kept) {
    if (prm->min_genes_cells > L) return -5; /* SingleCell.java:347-352 */
    int32_t* slot = (int32_t*)malloc((size_t)n_genes * sizeof(int32_t));
    for (int32_t g = 0; g < n_genes; g++) slot[g] = -1;
    for (int32_t l = 0; l < L; l++) if (gene_idx[l] >= 0 && gene_idx[l] < n_genes) slot[gene_idx[l]] = l;
    double* xs = (double*)malloc((size_t)(L > 0 ? L : 1) * sizeof(double)); /* interactors (abundance) */
    double* ys = (double*)malloc((size_t)(L > 0 ? L : 1) * sizeof(double)); /* gene expression in cell */
    gene_pair* gp = (gene_pair*)malloc((size_t)(L > 0 ? L : 1) * sizeof(gene_pair)); /* canonical == 2: by gene */
    for (int64_t c = 0; c < n_cells; c++) {
        int32_t np = 0, nz = 0;
        for (int64_t k = row_ptr[c]; k < row_ptr[c + 1]; k++) {
            int32_t g = col[k];
            if (g < 0 || g >= n_genes) continue;
            int32_t l = slot[g];
            if (l < 0) continue;
            float x = val[k];
            float y = abundance[l];
            if (x > 0.0f && y > 0.0f) { /* SingleCell.java:403 */
                if (np < L) { xs[np] = (double)y; ys[np] = (double)x; gp[np].g = g; gp[np].x = (double)y; gp[np].y = (double)x; }
                np++;
            }
            if (x != 0.0f) nz++; /* true for NaN, SingleCell.java:409 */
        }
        n_genes_used[c] = np;
        n_nonzero[c] = nz;
        /* a1: cells with no non-zero doc for any list gene never enter the list
         * (CORE/GeneExpressionsRetriever.java:213-222); a2: SingleCell.java:344-369 */
        int k = (nz > 0) && (np >= prm->min_genes_cells);
        kept[c] = (uint8_t)k;
        if (!k) { score[c] = NAN; continue; }
        if (np < 2) { score[c] = NAN; continue; } /* reference would throw inside commons-math */
        if (canonical == 2) { /* pair order of the gene-major path: ascending matrix column, whatever the row's storage order */
            qsort(gp, (size_t)np, sizeof(gene_pair), cmp_gene_pair);
            for (int32_t i = 0; i < np; i++) { xs[i] = gp[i].x; ys[i] = gp[i].y; }
        }
        int zx = all_equal(ys, np); /* geneExpressionVariance == 0, checked first (:433) */
        int zy = all_equal(xs, np); /* interactorsExpressionVariance == 0 (:441) */
        if ((zx || zy) && prm->jitter_mode == 0) { score[c] = NAN; continue; }
        if (zx) for (int32_t i = 0; i < np; i++) ys[i] = ys[i] + orc_jitter_uniform(prm->seed, (uint64_t)c, 0u, (uint32_t)i) / 1000000.0;
        if (zy) for (int32_t i = 0; i < np; i++) xs[i] = xs[i] + orc_jitter_uniform(prm->seed, (uint64_t)c, 1u, (uint32_t)i) / 1000000.0;
        /* correlation(interactors, genes) :449-450 */
        double r = canonical == 2 ? orc_pearson_shifted(xs, ys, np)
                 : canonical ? orc_pearson_strided32(xs, ys, np) : orc_pearson(xs, ys, np);
        if (prm->has_min_corr && r < prm->min_corr) r = NAN;     /* :455-456 */
        else if (prm->method == 1) r += (double)nz;              /* :457-458 */
        score[c] = r;
    }
    free(slot); free(xs); free(ys); free(gp);
    return 0;
}

int orc_score(int64_t n_cells, int32_t n_genes, const int64_t* row_ptr, const int32_t* col,
              const float* val, int32_t L, const int32_t* gene_idx, const float* abundance,
              const orc_score_params* prm, double* score, int32_t* n_genes_used,
              int32_t* n_nonzero, uint8_t* kept) {
    return orc_score2(n_cells, n_genes, row_ptr, col, val, L, gene_idx, abundance, prm, 0, score,
                      n_genes_used, n_nonzero, kept);
}

/* ============================================================================================
 * Stage 2 — threshold + rank. CORE/scoring/ScoreThreshold.java:57-75,98-106,
 * CORE/utils/PCTSEAUtils.java:44-53, CORE/PCTSEA.java:421-424,2497.
 * ========================================================================================== */

typedef struct { double s; int32_t idx; } rank_item;
static int cmp_desc(const void* pa, const void* pb) {
    const rank_item* a = (const rank_item*)pa; const rank_item* b = (const rank_item*)pb;
    int c = orc_java_double_compare(b->s, a->s);
    if (c) return c;
    return (a->idx > b->idx) - (a->idx < b->idx); /* stable merge sort from DB order */
}
static int cmp_asc(const void* pa, const void* pb) {
    const rank_item* a = (const rank_item*)pa; const rank_item* b = (const rank_item*)pb;
    int c = orc_java_double_compare(a->s, b->s);
    if (c) return c;
    return (a->idx > b->idx) - (a->idx < b->idx);
}

int64_t orc_rank(int64_t n_cells, const double* score, const uint8_t* kept, double min_score,
                 int32_t* order_out) {
    rank_item* it = (rank_item*)malloc((size_t)(n_cells > 0 ? n_cells : 1) * sizeof(rank_item));
    int64_t np = 0;
    for (int64_t c = 0; c < n_cells; c++) {
        if (kept && !kept[c]) continue;
        double s = score[c];
        int pass = (min_score >= 0.0) ? (s >= min_score) : (s < min_score); /* NaN fails both */
        if (pass) { it[np].s = s; it[np].idx = (int32_t)c; np++; }
    }
    if (min_score >= 0.0) {
        qsort(it, (size_t)np, sizeof(rank_item), cmp_desc);
        for (int64_t i = 0; i < np; i++) order_out[i] = it[i].idx;
    } else {
        /* ascending stable sort, subList drops the last, then PCTSEA.java:2497 re-sorts descending */
        qsort(it, (size_t)np, sizeof(rank_item), cmp_asc);
        if (np > 0) {
            int32_t dropped = it[np - 1].idx;
            qsort(it, (size_t)(np - 1), sizeof(rank_item), cmp_desc);
            for (int64_t i = 0; i < np - 1; i++) order_out[i] = it[i].idx;
            order_out[np - 1] = dropped;
        }
    }
    free(it);
    return np;
}

/* ============================================================================================
 * Stage 3 — weighted KS. CORE/utils/parallel/EnrichmentWeigthedScoreParallel.java:104-227,315-323,
 * 492-512.
 * ========================================================================================== */

static float dstat_of(float sup, int32_t sizea, int32_t sizeb) {
    /* :203-207 — sizea*sizeb is an int32 multiply (wraps), then widened */
    int32_t prod = (int32_t)((uint32_t)sizea * (uint32_t)sizeb);
    int32_t sum = (int32_t)((uint32_t)sizea + (uint32_t)sizeb);
    float sq = (float)sqrt(((double)prod * 1.0) / (1.0 * (double)sum));
    return sup * sq;
}

typedef struct {
    float sup; int32_t supX; float last_neg_at_sup; int has_neg_at_sup; int32_t nk; int valid;
} ks_walk;

/* One type, one label vector. observed!=0 also tracks the "last negative difference at 0-based
 * index < supX" needed by lookForPriorNegativeSupremum (:492-512). */
static ks_walk ks_one_type(int64_t Nu, const double* s, const int32_t* lab, int32_t t, int observed) {
    ks_walk w; memset(&w, 0, sizeof w);
    float denomA = 0.0f, denomB = 0.0f;
    int32_t nk = 0;
    for (int64_t i = 0; i < Nu; i++) { /* :110-117 */
        if (lab[i] == t) { nk++; denomA = (float)((double)denomA + s[i]); }
        else denomB = (float)((double)denomB + s[i]);
    }
    w.nk = nk;
    if (nk <= 1) { w.valid = 0; return w; } /* :119-127 */
    float sup = 0.0f, prevA = 0.0f, prevB = 0.0f, numA = 0.0f, numB = 0.0f;
    int32_t supX = 0;
    float last_neg = 0.0f; int has_neg = 0;
    float neg_at_sup = 0.0f; int has_neg_at_sup = 0;
    for (int64_t i = 0; i < Nu; i++) { /* :140-194 */
        float a, b;
        if (lab[i] == t) { numA = (float)((double)numA + s[i]); a = numA / denomA; b = prevB; }
        else             { numB = (float)((double)numB + s[i]); a = prevA; b = numB / denomB; }
        float diff = a - b;
        if (observed && diff < 0) { last_neg = diff; has_neg = 1; }
        if (fabsf(diff) > fabsf(sup)) {
            sup = diff; supX = (int32_t)(i + 1);
            neg_at_sup = last_neg; has_neg_at_sup = has_neg;
        }
        prevA = a; prevB = b;
    }
    int32_t sizea = nk, sizeb = (int32_t)(Nu - nk);
    w.sup = sup; w.supX = supX; w.last_neg_at_sup = neg_at_sup; w.has_neg_at_sup = has_neg_at_sup;
    w.valid = (sizea > 1 && sizeb > 1); /* :195-196 */
    return w;
}

static int cmp_f64_java(const void* pa, const void* pb) {
    return orc_java_double_compare(*(const double*)pa, *(const double*)pb);
}

double orc_ks_statistic(const double* x, int64_t n, const double* y, int64_t m) {
    /* KolmogorovSmirnovTest.integralKolmogorovSmirnovStatistic (commons-math3 3.6.x), un-vendored */
    double* sx = (double*)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    double* sy = (double*)malloc((size_t)(m > 0 ? m : 1) * sizeof(double));
    memcpy(sx, x, (size_t)n * sizeof(double));
    memcpy(sy, y, (size_t)m * sizeof(double));
    qsort(sx, (size_t)n, sizeof(double), cmp_f64_java);
    qsort(sy, (size_t)m, sizeof(double), cmp_f64_java);
    int64_t rankX = 0, rankY = 0, curD = 0, supD = 0;
    do {
        double z = orc_java_double_compare(sx[rankX], sy[rankY]) <= 0 ? sx[rankX] : sy[rankY];
        while (rankX < n && orc_java_double_compare(sx[rankX], z) == 0) { rankX += 1; curD += m; }
        while (rankY < m && orc_java_double_compare(sy[rankY], z) == 0) { rankY += 1; curD -= n; }
        if (curD > supD) supD = curD;
        else if (-curD > supD) supD = -curD;
    } while (rankX < n && rankY < m);
    free(sx); free(sy);
    return (double)supD / ((double)(n * m));
}

void orc_ks_observed(int64_t Nu, const double* s, const int32_t* lab, int32_t T, orc_type_result* out) {
    float* diffs = (float*)malloc((size_t)(Nu > 0 ? Nu : 1) * sizeof(float));
    float* as = (float*)malloc((size_t)(Nu > 0 ? Nu : 1) * sizeof(float));
    float* bs = (float*)malloc((size_t)(Nu > 0 ? Nu : 1) * sizeof(float));
    double* xa = (double*)malloc((size_t)(Nu > 0 ? Nu : 1) * sizeof(double));
    double* xb = (double*)malloc((size_t)(Nu > 0 ? Nu : 1) * sizeof(double));
    for (int32_t t = 0; t < T; t++) {
        ks_walk w = ks_one_type(Nu, s, lab, t, 1);
        orc_type_result* r = &out[t];
        r->sizeA = w.nk; r->sizeB = (int32_t)(Nu - w.nk);
        r->norm_ews = NAN; r->p_emp = NAN; r->fdr = NAN;
        r->sec_ews = 0.0f; r->sec_supX = 0; r->ks_d = NAN;
        if (!w.valid) { /* :120-127, :317-323 */
            r->ews = NAN; r->dstat = NAN; r->supX = -1; r->norm_supX = -1.0; r->raw_sup = NAN;
            continue;
        }
        float sup = w.sup;
        r->raw_sup = sup;
        r->dstat = dstat_of(sup, r->sizeA, r->sizeB);
        r->supX = w.supX;
        r->norm_supX = 1.0 * (double)w.supX / (double)Nu; /* :215 */
        if (w.has_neg_at_sup) sup = sup - fabsf(w.last_neg_at_sup); /* :216-227 */
        r->ews = sup;

        /* ---- observed-only extras ---- */
        /* KS D of the two score samples, :211-212 (NaN scores are not collected, :152,171) */
        int64_t na = 0, nb = 0;
        float denomA = 0.0f, denomB = 0.0f;
        for (int64_t i = 0; i < Nu; i++) {
            if (lab[i] == t) { denomA = (float)((double)denomA + s[i]); if (!isnan(s[i])) xa[na++] = s[i]; }
            else { denomB = (float)((double)denomB + s[i]); if (!isnan(s[i])) xb[nb++] = s[i]; }
        }
        r->ks_d = (na > 0 && nb > 0) ? orc_ks_statistic(xa, na, xb, nb) : NAN;
        /* differences again, then lookForSecondaryEnrichmentIndex :514-533 and the second walk :244-312 */
        float prevA = 0.0f, prevB = 0.0f, numA = 0.0f, numB = 0.0f;
        for (int64_t i = 0; i < Nu; i++) {
            float a, b;
            if (lab[i] == t) { numA = (float)((double)numA + s[i]); a = numA / denomA; b = prevB; }
            else             { numB = (float)((double)numB + s[i]); a = prevA; b = numB / denomB; }
            diffs[i] = a - b; as[i] = a; bs[i] = b;
            prevA = a; prevB = b;
        }
        int64_t secIdx = 0;
        double previousDifference = 0;
        for (int64_t i = (int64_t)w.supX - 1; i >= 0; i--) {
            double difference = diffs[i];
            if (difference > 0 && previousDifference < 0.0) { secIdx = i; break; }
            previousDifference = difference;
        }
        if (secIdx != 0) {
            float secSup = 0.0f; int32_t secX = 0;
            prevA = 0.0f; prevB = 0.0f; numA = 0.0f; numB = 0.0f;
            for (int64_t i = 0; i <= secIdx; i++) {
                float a, b;
                if (lab[i] == t) { numA = (float)((double)numA + fabs(s[i])); a = numA / denomA; b = prevB; }
                else             { numB = (float)((double)numB + fabs(s[i])); a = prevA; b = numB / denomB; }
                if (a > b) {
                    float difference = a - b;
                    if (fabsf(difference) > fabsf(secSup)) { secSup = difference; secX = (int32_t)(i + 1); }
                }
                prevA = a; prevB = b;
            }
            if (secX != 0) { r->sec_ews = secSup; r->sec_supX = secX; }
        }
    }
    free(diffs); free(as); free(bs); free(xa); free(xb);
}

void orc_ks_null_one(int64_t Nu, const double* s, const int32_t* lab_perm, int32_t T,
                     float* null_ews, float* null_d) {
    for (int32_t t = 0; t < T; t++) {
        ks_walk w = ks_one_type(Nu, s, lab_perm, t, 0);
        if (!w.valid) { null_ews[t] = NAN; null_d[t] = NAN; continue; }
        null_ews[t] = w.sup; /* :315 uncompensated */
        null_d[t] = dstat_of(w.sup, w.nk, (int32_t)(Nu - w.nk));
    }
}

/* ---- threads over types, re-created per permutation (CORE/PCTSEA.java:2498-2519) ----------- */
typedef struct {
    int64_t Nu; const double* s; const int32_t* lab; int32_t T; int32_t P; int32_t p;
    float* null_ews; float* null_d; volatile int32_t* next;
} null_job;

static void* null_worker(void* arg) {
    null_job* j = (null_job*)arg;
    for (;;) {
        int32_t t = __sync_fetch_and_add(j->next, 1);
        if (t >= j->T) break;
        ks_walk w = ks_one_type(j->Nu, j->s, j->lab, t, 0);
        size_t o = (size_t)t * (size_t)j->P + (size_t)j->p;
        if (!w.valid) { j->null_ews[o] = NAN; j->null_d[o] = NAN; }
        else { j->null_ews[o] = w.sup; j->null_d[o] = dstat_of(w.sup, w.nk, (int32_t)(j->Nu - w.nk)); }
    }
    return NULL;
}

static void null_one_perm(int64_t Nu, const double* s, const int32_t* labp, int32_t T, int32_t P,
                          int32_t p, int nthreads, float* null_ews, float* null_d) {
    volatile int32_t next = 0;
    null_job job = { Nu, s, labp, T, P, p, null_ews, null_d, &next };
    if (nthreads <= 1) { null_worker(&job); return; }
    pthread_t* th = (pthread_t*)malloc((size_t)nthreads * sizeof(pthread_t));
    for (int k = 0; k < nthreads; k++) pthread_create(&th[k], NULL, null_worker, &job);
    for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    free(th);
}

void orc_null(int64_t Nu, const double* s, const int32_t* lab, int32_t T, int32_t P, int perm_mode,
              uint64_t seed, int32_t perm_begin, int rounds, int nthreads, int32_t* replay_out,
              float* null_ews, float* null_d) {
    int32_t* idx = (int32_t*)malloc((size_t)(Nu > 0 ? Nu : 1) * sizeof(int32_t));
    int32_t* labp = (int32_t*)malloc((size_t)(Nu > 0 ? Nu : 1) * sizeof(int32_t));
    orc_jrandom jr; orc_jrandom_init(&jr, (int64_t)seed);
    for (int32_t p = 0; p < P; p++) {
        if (perm_mode == 0) {
            /* CORE/PCTSEA.java:1403-1413: copy, Collections.shuffle, label at rank i = shuffled[i] */
            for (int64_t i = 0; i < Nu; i++) idx[i] = (int32_t)i;
            orc_java_shuffle_i32(idx, Nu, &jr);
            for (int64_t i = 0; i < Nu; i++) labp[i] = lab[idx[i]];
            if (replay_out) memcpy(replay_out + (size_t)p * (size_t)Nu, idx, (size_t)Nu * sizeof(int32_t));
        } else {
            /* PHILOX mode permutes the size-sorted arrangement (orc_perm_labels); there is no index array to replay */
            orc_perm_labels(seed, (uint32_t)(perm_begin + p), Nu, lab, T, rounds, labp);
        }
        null_one_perm(Nu, s, labp, T, P, p, nthreads, null_ews, null_d);
    }
    free(idx); free(labp);
}

void orc_null_replay(int64_t Nu, const double* s, const int32_t* lab, int32_t T, int32_t P,
                     const int32_t* replay, float* null_ews, float* null_d) {
    int32_t* labp = (int32_t*)malloc((size_t)(Nu > 0 ? Nu : 1) * sizeof(int32_t));
    for (int32_t p = 0; p < P; p++) {
        const int32_t* idx = replay + (size_t)p * (size_t)Nu;
        for (int64_t i = 0; i < Nu; i++) labp[i] = lab[idx[i]];
        null_one_perm(Nu, s, labp, T, P, p, 1, null_ews, null_d);
    }
    free(labp);
}

/* ============================================================================================
 * a10 (CORE/PCTSEA.java:1449-1477,1799-1832), a11 (CORE/model/CellTypeClassification.java:407-441),
 * a12 (CORE/PCTSEA.java:1479-1549).
 * ========================================================================================== */

void orc_reduce(int32_t T, int32_t P, const float* null_ews, int mean_policy, orc_type_result* io,
                float* norm_null_out) {
    float* norm_null = norm_null_out ? norm_null_out : (float*)malloc((size_t)T * (size_t)(P > 0 ? P : 1) * sizeof(float));
    /* a10 */
    for (int32_t t = 0; t < T; t++) {
        float real = io[t].ews;
        if (isnan(real)) { io[t].p_emp = NAN; continue; }
        const float* nl = null_ews + (size_t)t * P;
        int32_t total = 0, cnt = 0;
        if (real >= 0.0f) {
            for (int32_t p = 0; p < P; p++) { float v = nl[p]; if (isnan(v)) continue; if (v <= 0.0f) continue; total++; if (v > real) cnt++; }
        } else {
            for (int32_t p = 0; p < P; p++) { float v = nl[p]; if (isnan(v)) continue; if (v >= 0.0f) continue; total++; if (v < real) cnt++; }
        }
        io[t].p_emp = 1.0 * (double)cnt / (double)total; /* 0/0 -> NaN */
    }
    /* a11 */
    for (int32_t t = 0; t < T; t++) {
        float real = io[t].ews;
        float* nn = norm_null + (size_t)t * P;
        if (isnan(real)) { io[t].norm_ews = NAN; for (int32_t p = 0; p < P; p++) nn[p] = NAN; continue; }
        const float* nl = null_ews + (size_t)t * P;
        /* Maths.mean(TFloatList) is un-vendored: policy 0 = float running sum in list order / size */
        float fsum_pos = 0.0f, fsum_neg = 0.0f; double dsum_pos = 0.0, dsum_neg = 0.0;
        int32_t npos = 0, nneg = 0;
        for (int32_t p = 0; p < P; p++) {
            float v = nl[p];
            if (v >= 0.0f) { fsum_pos += v; dsum_pos += (double)v; npos++; }
            else if (v < 0.0f) { fsum_neg += v; dsum_neg += (double)v; nneg++; }
        }
        float mean_pos = mean_policy ? (float)(dsum_pos / (double)npos) : fsum_pos / (float)npos;
        float mean_neg = mean_policy ? (float)(dsum_neg / (double)nneg) : fsum_neg / (float)nneg;
        float expct = (real >= 0.0f) ? fabsf(mean_pos) : fabsf(mean_neg);
        io[t].norm_ews = real / expct;
        for (int32_t p = 0; p < P; p++) nn[p] = nl[p] / expct;
    }
    /* a12 */
    int32_t nobs = 0; int64_t nnull = 0;
    float* R = (float*)malloc((size_t)(T > 0 ? T : 1) * sizeof(float));
    float* Z = (float*)malloc((size_t)T * (size_t)(P > 0 ? P : 1) * sizeof(float) + sizeof(float));
    for (int32_t t = 0; t < T; t++) {
        float ne = io[t].norm_ews;
        if (isnan(ne)) continue;
        if (!(ne < 0.0f)) R[nobs++] = ne;
        const float* nn = norm_null + (size_t)t * P;
        for (int32_t p = 0; p < P; p++) { float v = nn[p]; if (!(v < 0.0f)) Z[nnull++] = v; } /* NaN kept like Java's iterator.remove only on <0 */
    }
    orc_java_sort_f32(R, nobs);
    orc_java_sort_f32(Z, nnull);
    for (int32_t t = 0; t < T; t++) {
        float ne = io[t].norm_ews;
        if (isnan(ne) || ne < 0.0f) { io[t].fdr = NAN; continue; }
        int32_t i1 = orc_java_binary_search_f32(R, nobs, ne);
        int32_t i2 = orc_java_binary_search_f32(Z, (int32_t)nnull, ne);
        int32_t sobs = nobs - (i1 >= 0 ? i1 + 1 : -i1 - 1);
        int32_t snull = (int32_t)nnull - (i2 >= 0 ? i2 + 1 : -i2 - 1);
        double fdr;
        if (snull == 0) fdr = 0.0;
        else fdr = (1.0 * (double)snull / (double)sobs) * (1.0 * (double)nobs / (double)nnull);
        io[t].fdr = fdr;
    }
    free(R); free(Z);
    if (!norm_null_out) free(norm_null);
}

/* ============================================================================================
 * FAST-mode definition (this repo's own arithmetic, not the reference's; round 2 form): scores quantised to
 * int32 fixed point fx = llrint(s * 2^F), F chosen so that the sum of |fx| over the ranking fits an int32;
 * every running sum is then an exact int32, the supremum is evaluated only at hit and pre-hit positions, and
 * each candidate difference is ONE fp32 fused multiply-add of the converted sums:
 *     diff = numA/dA - (S - numA)/dB = fmaf((float)numA, c12, -((float)S * c2)),
 *     c2 = (float)(1/dB), c12 = (float)(1/dA + 1/dB)   (reciprocals and their sum in fp64).
 * Exact integer sums make the result independent of how the GPU splits the ranking into segments.
 * Differences from the reference beyond rounding (documented in include/pctsea_b200.h, ks_mode): a zero
 * denominator gives ES 0 (reference: NaN/Inf from the float division); equal |diff| resolves to the positive
 * one (reference: first position, strict '>'); hit / pre-hit evaluation assumes all ranked scores share a sign
 * (true on the reference's path: ScoreThreshold passes one side of the threshold).
 * ========================================================================================== */

int orc_fast_frac_bits(int64_t Nu, const double* s) {
    /* F0 = 25 - e with maxabs < 2^e  =>  |llrint(s * 2^F0)| <= 2^25; then the smallest right shift sh with
     * ceil(sum|fx0| / 2^sh) + Nu <= INT32_MAX  (=> sum |llrint(s * 2^(F0-sh))| <= INT32_MAX) */
    double maxabs = 0.0;
    for (int64_t i = 0; i < Nu; i++) { double v = fabs(s[i]); if (v > maxabs) maxabs = v; }
    int e = 0;
    if (maxabs > 0.0) e = ilogb(maxabs) + 1;
    int F0 = 25 - e;
    if (F0 > 1000) F0 = 1000;
    const double scale0 = ldexp(1.0, F0);
    int64_t A = 0;
    for (int64_t i = 0; i < Nu; i++) { int64_t v = llrint(s[i] * scale0); A += v < 0 ? -v : v; }
    int sh = 0;
    while (((A + (((int64_t)1 << sh) - 1)) >> sh) + Nu > (int64_t)INT32_MAX) sh++;
    return F0 - sh;
}

static inline int fast_better(float f, float sup) {
    /* larger |diff| wins; on equal |diff| the positive one wins (order-independent rule) */
    float af = fabsf(f), as = fabsf(sup);
    if (af > as) return 1;
    if (af == as && af > 0.0f && f > 0.0f && sup < 0.0f) return 1;
    return 0;
}

void orc_ks_null_fast_one(int64_t Nu, const double* s, const int32_t* lab, int32_t T,
                          int frac_bits, float* null_ews) {
    int32_t* fx = (int32_t*)malloc((size_t)(Nu > 0 ? Nu : 1) * sizeof(int32_t));
    int32_t* S = (int32_t*)malloc((size_t)(Nu > 0 ? Nu : 1) * sizeof(int32_t));
    int64_t acc = 0;
    double scale = ldexp(1.0, frac_bits);
    for (int64_t i = 0; i < Nu; i++) { fx[i] = (int32_t)llrint(s[i] * scale); acc += fx[i]; S[i] = (int32_t)acc; }
    int32_t Stot = (int32_t)acc;
    for (int32_t t = 0; t < T; t++) {
        int32_t dA = 0; int32_t nk = 0;
        for (int64_t i = 0; i < Nu; i++) if (lab[i] == t) { dA += fx[i]; nk++; }
        int32_t dB = Stot - dA;
        if (!(nk > 1 && Nu - nk > 1)) { null_ews[t] = NAN; continue; }
        if (dA == 0 || dB == 0) { null_ews[t] = 0.0f; continue; }
        const double r2 = 1.0 / (double)dB;
        const float c2 = (float)r2;
        const float c12 = (float)(1.0 / (double)dA + r2);
        float sup = 0.0f; int32_t numA = 0;
        for (int64_t i = 0; i < Nu; i++) {
            if (lab[i] != t) continue;
            if (i > 0) {
                float f0 = fmaf((float)numA, c12, -((float)S[i - 1] * c2));
                if (fast_better(f0, sup)) sup = f0;
            }
            numA += fx[i];
            float f1 = fmaf((float)numA, c12, -((float)S[i] * c2));
            if (fast_better(f1, sup)) sup = f1;
        }
        null_ews[t] = sup;
    }
    free(fx); free(S);
}
