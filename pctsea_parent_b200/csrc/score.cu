// score.cu — stage 1: per-cell scoring of the device-resident CSR against the input list.
//
// Replaces (reference paths under pctsea-core/src/main/java/edu/scripps/yates/pctsea/):
//   a2  model/SingleCell.java:344-369  hasMinNumberOfExpressedGenes   (driver PCTSEA.java:544-572)
//   a3  model/SingleCell.java:380-463  calculateCorrelation / calculateSimpleScore (driver PCTSEA.java:2305-2343)
//       with commons-math3 PearsonsCorrelation -> SimpleRegression restated (un-vendored dependency).
//
// Layout: one warp per cell. The warp streams the CSR row with 128-bit coalesced loads (each lane takes 4
// consecutive non-zeros), tests list membership against a shared-memory bitmap and compacts the matched
// (abundance, expression) pairs in CSR order into a per-warp shared-memory buffer (ballot + popc). Pair k then
// belongs to lane k % 32. Pearson is the two-pass centred form: pass 1 sums x and y per lane and adds the 32
// partials with an xor-shuffle tree to get the means, pass 2 does the same for the centred squares and cross
// products, and the result goes through commons-math's getR() formula. All fp64 arithmetic uses explicit
// round-to-nearest intrinsics (no FMA contraction), so the score is bit-identical to the CPU oracle's
// orc_pearson_strided32 (same sub-streams, same tree) and within ~1e-12 relative of the reference's
// sequential SimpleRegression update on non-degenerate cells (its own pair order is a hash-set order).
//
// Roofline: HBM. Algorithmic bytes = 8 B * nnz (col + val) + 16 B * N (row_ptr) + 17 B * N outputs.
#include "common.cuh"
#include <cstdlib>

namespace pctsea {

__global__ void k_list_lookup(const int32_t* __restrict__ gene_idx, const float* __restrict__ abund, int L, int G,
                              float* __restrict__ ybg, uint32_t* __restrict__ bitmap) {
    // two bits per gene, 16 genes per word: bit 0 = gene is in the list (counts towards "non-zero list
    // genes"), bit 1 = in the list with abundance > 0 (can form a pair)
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    int g = gene_idx[l];
    if (g < 0 || g >= G) return;
    const float y = abund[l];
    ybg[g] = y;
    atomicOr(&bitmap[g >> 4], (y > 0.0f ? 3u : 1u) << ((g & 15) * 2));
}

// Rank directory of the list genes in gene order, for the on-chip abundance gather: per 32 genes one entry
// {membership bits, number of list genes below the word}; rank(g) = entry.y + popc(entry.x & below(g)), and
// ab_sorted[rank] is the abundance. One CTA; words are scanned in tiles of blockDim.x with a running carry.
__global__ void k_list_rank(const uint32_t* __restrict__ bitmap, const float* __restrict__ ybg, int G, int nwords,
                            int2* __restrict__ rankw, float* __restrict__ ab_sorted) {
    __shared__ int s_scan[1024];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int w0 = 0; w0 < nwords; w0 += blockDim.x) {
        const int w = w0 + threadIdx.x;
        uint32_t bits = 0u;
        if (w < nwords) {
            // genes 32w .. 32w+31 live in bitmap words 2w and 2w+1 (two bits each); keep bit 0 of every entry
            const int bm_words = (G + 1 + 15) / 16;
#pragma unroll
            for (int hlf = 0; hlf < 2; hlf++) {
                const int bw = 2 * w + hlf;
                const uint32_t v = bw < bm_words ? bitmap[bw] : 0u;
#pragma unroll
                for (int j = 0; j < 16; j++) bits |= ((v >> (2 * j)) & 1u) << (16 * hlf + j);
            }
        }
        const int c = __popc(bits);
        s_scan[threadIdx.x] = c;
        __syncthreads();
        for (int o = 1; o < (int)blockDim.x; o <<= 1) {
            const int v = threadIdx.x >= (unsigned)o ? s_scan[threadIdx.x - o] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        const int excl = s_carry + s_scan[threadIdx.x] - c;
        if (w < nwords) {
            rankw[w] = make_int2((int)bits, excl);
            uint32_t b = bits;
            int r = excl;
            while (b) {
                const int j = __ffs(b) - 1;
                b &= b - 1;
                ab_sorted[r++] = ybg[32 * w + j];
            }
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry += s_scan[threadIdx.x];
        __syncthreads();
    }
}

__device__ __forceinline__ double jitter_uniform(uint64_t seed, uint64_t cell, uint32_t vec, uint32_t i) {
    uint32_t o[4];
    philox4x32_10(i, vec | 0x4A000000u, (uint32_t)cell, (uint32_t)(cell >> 32), (uint32_t)seed,
                  (uint32_t)(seed >> 32), o);
    uint64_t u = ((uint64_t)(o[0] >> 6) << 27) | (uint64_t)(o[1] >> 5);
    return __dmul_rn((double)u, 1.0 / 9007199254740992.0);
}

__device__ __forceinline__ double warp_sum_tree(double v) {
    // offsets 16, 8, 4, 2, 1: lane l ends with the same value as the oracle's tree32 slot 0 (fp add commutes)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// The same two (three) trees as warp_sum_tree, but packed: after the offset-16 exchange the lower half-warp
// carries the first value and the upper half the second, so every level costs one shuffle + add for both (and
// the means need one division per lane instead of two). Each partial sum pairs exactly the operands the plain
// butterfly pairs (fp add commutes), so the results are bit-identical to warp_sum_tree on each value.
__device__ __forceinline__ void warp_means_packed(double sx, double sy, double n, int lane, double& xbar, double& ybar) {
    const bool hi = (lane & 16) != 0;
    double v = __dadd_rn(hi ? sy : sx, __shfl_xor_sync(0xffffffffu, hi ? sx : sy, 16));
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    const double m = __ddiv_rn(v, n);   // lanes 0-15: mean of x, lanes 16-31: mean of y
    xbar = __shfl_sync(0xffffffffu, m, 0);
    ybar = __shfl_sync(0xffffffffu, m, 16);
}

// Three sums; lane 0 ends with all three (the other lanes' outputs are unspecified).
__device__ __forceinline__ void warp_sum3_packed(double& a, double& b, double& c, int lane) {
    const bool hi = (lane & 16) != 0, q8 = (lane & 8) != 0;
    // offset 16: lower half keeps (a, b), upper half keeps (c, -); c travels down, a and b travel up
    const double r0 = __shfl_xor_sync(0xffffffffu, hi ? a : c, 16);
    const double r1 = __shfl_xor_sync(0xffffffffu, b, 16);
    double u = __dadd_rn(hi ? c : a, r0);   // lower: a_l + a_{l+16}; upper: c_l + c_{l-16}
    double w = __dadd_rn(b, r1);            // b_l + b_{l^16} (used by the lower half only)
    // offset 8: inside the lower half, lanes with bit 3 clear keep a, the others keep b; the upper half keeps c
    const double r2 = __shfl_xor_sync(0xffffffffu, (hi || q8) ? u : w, 8);
    double v = __dadd_rn((hi || !q8) ? u : w, r2);
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    a = v;                                      // lane 0
    b = __shfl_sync(0xffffffffu, v, 8);
    c = __shfl_sync(0xffffffffu, v, 16);
}

// SimpleRegression.getR(): sign(slope) * sqrt(RSquare)
__device__ __forceinline__ double pearson_finish(double sumXX, double sumYY, double sumXY) {
    const double ten_min = __longlong_as_double(10LL);  // 10 * Double.MIN_VALUE
    const double b1 = (fabs(sumXX) < ten_min) ? __longlong_as_double(0x7ff8000000000000LL) : __ddiv_rn(sumXY, sumXX);
    const double sse_arg = __dsub_rn(sumYY, __ddiv_rn(__dmul_rn(sumXY, sumXY), sumXX));
    const double sse = (sse_arg != sse_arg) ? sse_arg : (sse_arg > 0.0 ? sse_arg : 0.0);
    const double rsq = __ddiv_rn(__dsub_rn(sumYY, sse), sumYY);
    double r = __dsqrt_rn(rsq);
    if (b1 < 0.0) r = -r;
    return r;
}

// Zero-variance cells (all matched abundances or all matched expressions equal): the constant vector gets the
// Philox jitter (model/SingleCell.java:430-441 adds Math.random()/1e6) and the two-pass sums are redone on the
// jittered values. Rare and bulky (two Philox streams), so it stays out of line to keep the scoring kernel's
// instruction footprint small. Returns the warp-reduced centred sums.
struct CentredSums { double sxx, syy, sxy; };
__device__ __noinline__ CentredSums pearson_sums_jittered(const float* pair_a, const float* pair_e, int np, int lane,
                                                          bool eq_list, bool eq_cell, uint64_t seed, uint64_t cell) {
    double sx = 0.0, sy = 0.0;
    for (int k = lane; k < np; k += 32) {
        double x = (double)pair_a[k], y = (double)pair_e[k];
        if (eq_cell) y = __dadd_rn(y, __ddiv_rn(jitter_uniform(seed, cell, 0u, (uint32_t)k), 1000000.0));
        if (eq_list) x = __dadd_rn(x, __ddiv_rn(jitter_uniform(seed, cell, 1u, (uint32_t)k), 1000000.0));
        sx = __dadd_rn(sx, x);
        sy = __dadd_rn(sy, y);
    }
    const double xbar = __ddiv_rn(warp_sum_tree(sx), (double)np);
    const double ybar = __ddiv_rn(warp_sum_tree(sy), (double)np);
    double sxx = 0.0, syy = 0.0, sxy = 0.0;
    for (int k = lane; k < np; k += 32) {
        double x = (double)pair_a[k], y = (double)pair_e[k];
        if (eq_cell) y = __dadd_rn(y, __ddiv_rn(jitter_uniform(seed, cell, 0u, (uint32_t)k), 1000000.0));
        if (eq_list) x = __dadd_rn(x, __ddiv_rn(jitter_uniform(seed, cell, 1u, (uint32_t)k), 1000000.0));
        const double dx = __dsub_rn(x, xbar), dy = __dsub_rn(y, ybar);
        sxx = __dadd_rn(sxx, __dmul_rn(dx, dx));
        syy = __dadd_rn(syy, __dmul_rn(dy, dy));
        sxy = __dadd_rn(sxy, __dmul_rn(dx, dy));
    }
    CentredSums r;
    r.sxx = warp_sum_tree(sxx);
    r.syy = warp_sum_tree(syy);
    r.sxy = warp_sum_tree(sxy);
    return r;
}

struct ScoreKernelArgs {
    const int64_t* row_ptr;
    const int32_t* col;
    const float* val;
    const int32_t* col_end;  // col + nnz: 128-bit loads stop here
    int64_t N;
    int32_t G;
    const float* ybg;
    const uint32_t* bitmap;
    int32_t bm_words;
    const int2* rankw;        // rank directory (k_list_rank), copied to shared memory when rank_words > 0
    const float* ab_sorted;   // abundances of the list genes in gene order
    int32_t rank_words, rank_bytes, n_list;  // rank_bytes: shared-memory bytes of both arrays (multiple of 16)
    int32_t tab_bytes;  // shared-memory bytes of the membership table (multiple of 16)
    int32_t cap;        // pairs per warp buffer, >= L
    int32_t method, min_genes, has_min_corr, jitter_mode;
    double min_corr;
    uint64_t seed;
    double* sums;       // [N][3] centred sums sxx, syy, sxy (NaN where no correlation is defined)
    int32_t* nused;
    int32_t* nnzc;
    uint8_t* kept;
};

constexpr int kScoreMaxWarps = 24;    // one CTA per SM; 24 warps * 85 registers fill the register file
constexpr int kScoreCellBatch = 32;   // consecutive cells a CTA takes at a time (interleaved over the grid)

// 128 consecutive non-zeros of one row as the warp holds them: lane l has elements 4l..4l+3 of the step.
struct RowChunk {
    int32_t g[4];
    float x[4];
};

// Membership entry of gene g: 0x11 = in the list with abundance > 0 (pairs with a positive expression),
// 0x10 = in the list only (counts towards the non-zero list genes), 0 otherwise. The low and high nibbles
// of the sum over a lane's four elements are its pair and non-zero counts.
template <bool BYTE_TAB>
__device__ __forceinline__ uint32_t tab_lookup(const unsigned char* tab, uint32_t g) {
    if (BYTE_TAB) return tab[g];
    const uint32_t bits = (reinterpret_cast<const uint32_t*>(tab)[g >> 4] >> ((g & 15u) * 2u)) & 3u;
    return ((bits & 1u) << 4) | (bits >> 1);
}

// One warp per cell, one CTA per SM. Two decoupled roles run in each warp:
//   * the cursor walks the warp's cells step by step (128 non-zeros per step) and issues the 128-bit loads
//     DEPTH steps ahead of their use, across cell boundaries; cells are handed out from a CTA-local counter so
//     the warps of an SM finish together. Each cell the cursor enters is announced to the consumer through a
//     small per-warp shared-memory queue (cell id, row length from the first step's base, row-begin offset);
//   * the consumer tests list membership (one shared-memory table read per element), compacts the matching
//     (column, expression) pairs in CSR order into the warp's pair buffer, and on the last step of a row runs
//     the two-pass Pearson regression over the buffer.
// The in-flight steps live in DEPTH named register sets that the unrolled loop visits in turn, so no register
// moves are needed to rotate the pipeline.
template <bool BYTE_TAB, int MAXW, int DEPTH, bool ALL_POSITIVE>
__global__ void __launch_bounds__(MAXW * 32, 1) k_score(const ScoreKernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_next;
    __shared__ int4 s_queue[MAXW][8];
    unsigned char* tab = smem_raw;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // pair buffer, two planes: matrix column (int) until the abundances are gathered, then the abundance;
    // and the expression value
    int2* rankw = reinterpret_cast<int2*>(smem_raw + a.tab_bytes);
    float* ab_sorted = reinterpret_cast<float*>(rankw + a.rank_words);
    float* pair_a = reinterpret_cast<float*>(smem_raw + a.tab_bytes + a.rank_bytes) + (size_t)warp * 2 * a.cap;
    float* pair_e = pair_a + a.cap;

    if (BYTE_TAB) {
        for (int g = threadIdx.x; g <= a.G; g += blockDim.x) {
            const uint32_t bits = (a.bitmap[g >> 4] >> ((g & 15) * 2)) & 3u;
            tab[g] = (unsigned char)(((bits & 1u) << 4) | (bits >> 1));
        }
    } else {
        for (int i = threadIdx.x; i < a.bm_words; i += blockDim.x) reinterpret_cast<uint32_t*>(tab)[i] = a.bitmap[i];
    }
    if (a.rank_words > 0) {
        for (int i = threadIdx.x; i < a.rank_words; i += blockDim.x) rankw[i] = a.rankw[i];
        for (int i = threadIdx.x; i < a.n_list; i += blockDim.x) ab_sorted[i] = a.ab_sorted[i];
    }
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();

    const unsigned FULL = 0xffffffffu;
    const unsigned lt_mask = (1u << lane) - 1u;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const int32_t Ncl = (int32_t)a.N;
    const int lane4 = 4 * lane;
    const bool on_chip = a.rank_words > 0;
    // abundance of list gene g: rank directory + sorted abundances in shared memory (short lists), else the
    // by-gene table in L2
    auto abundance_of = [&](int g) -> float {
        if (!on_chip) return __ldg(a.ybg + g);
        const int2 e = rankw[g >> 5];
        return ab_sorted[e.y + __popc((uint32_t)e.x & ((1u << (g & 31)) - 1u))];
    };

    // ---- cursor -------------------------------------------------------------------------------------------
    int32_t f_cell, f_ncell;         // cell being fetched, and the warp's following cell
    int64_t f_nbeg = 0, f_nend = 0;  // row bounds of f_ncell
    const int32_t* pc = a.col;       // this lane's position in the row
    const float* pv = a.val;
    int remw = 0;                    // row end minus the warp's step base
    int f_n = 0;                     // cells announced so far
    auto grab = [&]() -> int32_t {   // next cell of this CTA (clamped to N when none is left)
        int j = 0;
        if (lane == 0) j = atomicAdd(&s_next, 1);
        j = __shfl_sync(FULL, j, 0);
        const int64_t id = ((int64_t)(j / kScoreCellBatch) * gridDim.x + blockIdx.x) * kScoreCellBatch + (j % kScoreCellBatch);
        return (int32_t)(id < a.N ? id : a.N);
    };
    auto bounds = [&](int32_t cl, int64_t& b, int64_t& e) {
        b = 0; e = 0;
        if (cl < Ncl) { b = a.row_ptr[cl]; e = a.row_ptr[cl + 1]; }
    };
    auto enter = [&](int32_t cl, int64_t b, int64_t e) {
        const int64_t ks = b & ~(int64_t)3;
        remw = (int)(e - ks);
        pc = a.col + ks + lane4;
        pv = a.val + ks + lane4;
        if (lane == 0) s_queue[warp][f_n & 7] = make_int4(cl, remw, (int)(b & 3), 0);
        f_n++;
    };
    auto fetch = [&]() -> RowChunk {
        RowChunk c;
        if (remw > 128) {   // not the last step of the row: every lane's quad lies inside the arrays
            const int4 gv = __ldcs(reinterpret_cast<const int4*>(pc));
            const float4 xv = __ldcs(reinterpret_cast<const float4*>(pv));
            c.g[0] = gv.x; c.g[1] = gv.y; c.g[2] = gv.z; c.g[3] = gv.w;
            c.x[0] = xv.x; c.x[1] = xv.y; c.x[2] = xv.z; c.x[3] = xv.w;
            pc += 128; pv += 128; remw -= 128;
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) { c.g[q] = 0; c.x[q] = 0.0f; }
            if (lane4 < remw) {
                if (pc + 4 <= a.col_end) {
                    const int4 gv = __ldcs(reinterpret_cast<const int4*>(pc));
                    const float4 xv = __ldcs(reinterpret_cast<const float4*>(pv));
                    c.g[0] = gv.x; c.g[1] = gv.y; c.g[2] = gv.z; c.g[3] = gv.w;
                    c.x[0] = xv.x; c.x[1] = xv.y; c.x[2] = xv.z; c.x[3] = xv.w;
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (lane4 + q < remw) { c.g[q] = __ldcs(pc + q); c.x[q] = __ldcs(pv + q); }
                }
            }
            f_cell = f_ncell;
            enter(f_cell, f_nbeg, f_nend);
            f_ncell = grab();
            bounds(f_ncell, f_nbeg, f_nend);
        }
        return c;
    };

    // ---- consumer -----------------------------------------------------------------------------------------
    int cnt = 0, my_nz = 0;
    int32_t c_cell; int c_remw, c_boff; int c_n = 0;
    auto next_cell = [&]() {
        __syncwarp();
        const int4 q = s_queue[warp][c_n & 7];
        c_n++;
        c_cell = q.x; c_remw = q.y; c_boff = q.z;
    };
    auto consume = [&](RowChunk& c) {
        if ((c_boff != 0) | (c_remw <= 128)) {   // head or tail step: drop the elements outside the row
            const int rem = c_remw - lane4, off = lane4 - c_boff;
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (!(q < rem && off + q >= 0)) { c.g[q] = a.G; c.x[q] = 1.0f; }   // the table's zero entry
        }
        uint32_t t[4];
#pragma unroll
        for (int q = 0; q < 4; q++) t[q] = tab_lookup<BYTE_TAB>(tab, (uint32_t)c.g[q]);
        // ALL_POSITIVE: the load-time check found every stored value > 0 (the usual CSR of counts), nothing to test
        const bool pos4 = ALL_POSITIVE || ((c.x[0] > 0.0f) && (c.x[1] > 0.0f) && (c.x[2] > 0.0f) && (c.x[3] > 0.0f));
        if (!ALL_POSITIVE && !__all_sync(FULL, pos4)) {   // zeros, negatives or NaN stored in the matrix: rare
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (!(c.x[q] > 0.0f)) t[q] &= 0xF0u;    // cannot form a pair
                if (!(c.x[q] != 0.0f)) t[q] = 0u;       // an explicit zero is not an expressed gene (NaN is)
            }
        }
        const uint32_t S = t[0] + t[1] + t[2] + t[3];
        my_nz += (int)(S >> 4);
        // exclusive prefix over lanes of the lane's pair count (S & 15, 0..4) from the ballots of its three bit
        // planes; CSR order inside the step is (lane, q): lower lanes first, then my own earlier elements
        const unsigned p0 = __ballot_sync(FULL, (S & 1u) != 0u), p1 = __ballot_sync(FULL, (S & 2u) != 0u),
                       p2 = __ballot_sync(FULL, (S & 4u) != 0u);
        int pos = cnt + __popc(p0 & lt_mask) + 2 * __popc(p1 & lt_mask) + 4 * __popc(p2 & lt_mask);
        cnt += __reduce_add_sync(FULL, (int)(S & 15u));
        if (S & 15u) {
            if (cnt <= a.cap) {   // warp-uniform: the whole step fits (always, unless a row repeats columns)
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (t[q] & 1u) { pair_a[pos] = __int_as_float(c.g[q]); pair_e[pos] = c.x[q]; pos++; }
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (t[q] & 1u) {
                        if (pos < a.cap) { pair_a[pos] = __int_as_float(c.g[q]); pair_e[pos] = c.x[q]; }
                        pos++;
                    }
            }
        }
    };
    // Pearson over the warp's pair buffer; writes the cell's outputs and resets the per-cell counters
    auto finish = [&](int32_t cell) {
        const int cnt_nz = __reduce_add_sync(FULL, my_nz);
        __syncwarp();
        const int np = min(cnt, a.cap);  // cnt > cap only with duplicate columns in a row (unsupported input)
        const bool keep = (cnt_nz > 0) && (cnt >= a.min_genes);
        double o_xx = qnan, o_yy = qnan, o_xy = qnan;
        if (keep && np >= 2) {
            // one sweep: gather the abundances of the matched genes (L1/L2-resident table), test both vectors
            // for zero variance <=> all elements equal (Maths.var == 0, model/SingleCell.java:430-441), and
            // take the plain sums for the means
            const float a0 = abundance_of(__float_as_int(pair_a[0])), e0 = pair_e[0];
            __syncwarp();
            bool eq_list = true, eq_cell = true;
            double sx = 0.0, sy = 0.0;
#pragma unroll 4
            for (int k = lane; k < np; k += 32) {
                const float ab = abundance_of(__float_as_int(pair_a[k])), ex = pair_e[k];
                pair_a[k] = ab;
                eq_list = eq_list && (ab == a0);
                eq_cell = eq_cell && (ex == e0);
                sx = __dadd_rn(sx, (double)ab);
                sy = __dadd_rn(sy, (double)ex);
            }
            eq_list = __all_sync(FULL, eq_list);
            eq_cell = __all_sync(FULL, eq_cell);
            const bool degenerate = eq_cell || eq_list;
            if (!(degenerate && a.jitter_mode == PCTSEA_JITTER_NAN)) {
                double sxx = 0.0, syy = 0.0, sxy = 0.0;
                if (!degenerate) {
                    double xbar, ybar;
                    warp_means_packed(sx, sy, (double)np, lane, xbar, ybar);
#pragma unroll 2
                    for (int k = lane; k < np; k += 32) {
                        const double dx = __dsub_rn((double)pair_a[k], xbar), dy = __dsub_rn((double)pair_e[k], ybar);
                        sxx = __dadd_rn(sxx, __dmul_rn(dx, dx));
                        syy = __dadd_rn(syy, __dmul_rn(dy, dy));
                        sxy = __dadd_rn(sxy, __dmul_rn(dx, dy));
                    }
                    warp_sum3_packed(sxx, syy, sxy, lane);   // complete in lane 0, which writes them
                } else {
                    const CentredSums cs = pearson_sums_jittered(pair_a, pair_e, np, lane, eq_list, eq_cell, a.seed, (uint64_t)cell);
                    sxx = cs.sxx; syy = cs.syy; sxy = cs.sxy;
                }
                o_xx = sxx; o_yy = syy; o_xy = sxy;
            }
        }
        if (lane == 0) {
            double* o = a.sums + 3 * (size_t)cell;
            o[0] = o_xx; o[1] = o_yy; o[2] = o_xy;
            a.nused[cell] = cnt;
            a.nnzc[cell] = cnt_nz;
            a.kept[cell] = keep ? 1 : 0;
        }
        __syncwarp();
        cnt = 0; my_nz = 0;
    };

    {   // prime the cursor with the warp's first two cells
        f_cell = grab();
        int64_t b, e;
        bounds(f_cell, b, e);
        enter(f_cell, b, e);
        f_ncell = grab();
        bounds(f_ncell, f_nbeg, f_nend);
    }
    static_assert(DEPTH == 3 || DEPTH == 4, "the register rotation below is written for 3 or 4 steps in flight");
    RowChunk c0 = fetch(), c1 = fetch(), c2 = fetch(), c3;
    if (DEPTH == 4) c3 = fetch();
    next_cell();
    if (c_cell >= Ncl) return;
    for (;;) {   // one cell per iteration; c0 is always the oldest outstanding step here
        int used;   // register sets consumed (and refilled) since c0 was the oldest, mod DEPTH
        for (;;) {
            consume(c0); c0 = fetch(); used = 1;
            if (c_remw <= 128) break;
            c_remw -= 128; c_boff = 0;
            consume(c1); c1 = fetch(); used = 2;
            if (c_remw <= 128) break;
            c_remw -= 128; c_boff = 0;
            consume(c2); c2 = fetch(); used = (DEPTH == 3) ? 0 : 3;
            if (c_remw <= 128) break;
            c_remw -= 128; c_boff = 0;
            if (DEPTH == 4) {
                consume(c3); c3 = fetch(); used = 0;
                if (c_remw <= 128) break;
                c_remw -= 128; c_boff = 0;
            }
        }
        finish(c_cell);
        // rotate the register sets (after the regression, when the loads have landed) so that the oldest step
        // is c0 again: once per cell instead of a move chain per step
        if (DEPTH == 3) {
            if (used == 1) { const RowChunk t = c0; c0 = c1; c1 = c2; c2 = t; }
            else if (used == 2) { const RowChunk t = c0; c0 = c2; c2 = c1; c1 = t; }
        } else {
            if (used == 1) { const RowChunk t = c0; c0 = c1; c1 = c2; c2 = c3; c3 = t; }
            else if (used == 2) { RowChunk t = c0; c0 = c2; c2 = t; t = c1; c1 = c3; c3 = t; }
            else if (used == 3) { const RowChunk t = c3; c3 = c2; c2 = c1; c1 = c0; c0 = t; }
        }
        next_cell();
        if (c_cell >= Ncl) return;   // the cursor ran past the CTA's last cell
    }
}

// Thread per cell: the centred sums -> SimpleRegression.getR(), then the score thresholds. Split from k_score so
// that its warps do not spend ~200 instructions per cell on scalar divisions and a square root.
__global__ void k_score_final(const double* __restrict__ sums, const int32_t* __restrict__ nnzc, int64_t N, int method,
                              int has_min_corr, double min_corr, double* __restrict__ score) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double sxx = sums[3 * i], syy = sums[3 * i + 1], sxy = sums[3 * i + 2];
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    double r = qnan;
    if (!(sxx != sxx && syy != syy && sxy != sxy)) {   // all three NaN: no correlation defined for this cell
        r = pearson_finish(sxx, syy, sxy);
        if (has_min_corr && r < min_corr) r = qnan;
        else if (method == PCTSEA_METHOD_SIMPLE_SCORE) r = __dadd_rn(r, (double)nnzc[i]);
    }
    score[i] = r;
}

__global__ void k_validate_csr(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                               const float* __restrict__ val, int64_t N, int32_t G, int64_t nnz, int* __restrict__ bad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int b = 0;
    for (int64_t i = t0; i < N; i += stride)
        if (row_ptr[i + 1] < row_ptr[i]) b |= 1;
    if (t0 == 0 && (row_ptr[0] != 0 || row_ptr[N] != nnz)) b |= 1;
    for (int64_t k = t0; k < nnz; k += stride) {
        if ((unsigned)col[k] >= (unsigned)G) b |= 2;
        if (!(val[k] > 0.0f)) b |= 4;   // a stored zero, negative or NaN: not an error, but the scorer must test for it
    }
    if (b) atomicOr(bad, b);
}

// Bit 3: some row stores a column twice. The reference keeps ONE value per (cell, gene) (model/SingleCell.java:80, the
// last one written), the two scoring kernels would each do something else with a repeated column (the CSR scan pairs
// it twice, the gene-major build collapses it into one mask bit and races on the value), so such a matrix is
// rejected. One warp per row, any storage order: the row's columns are marked in the warp's bitmap of G bits in
// shared memory (atomicOr returns the previous word), then the same walk clears them again.
__global__ void __launch_bounds__(256) k_validate_unique_cols(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                                              int64_t N, int32_t G, int words, int* __restrict__ bad) {
    extern __shared__ unsigned s_bits[];   // [warps][words]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned* bits = s_bits + (size_t)warp * words;
    for (int w = lane; w < words; w += 32) bits[w] = 0u;
    __syncwarp();
    bool dup = false;
    for (int64_t i = (int64_t)blockIdx.x * nwarps + warp; i < N; i += (int64_t)gridDim.x * nwarps) {
        const int64_t b = row_ptr[i], e = row_ptr[i + 1];
        for (int64_t k = b + lane; k < e; k += 32) {
            const unsigned g = (unsigned)col[k];
            if (g < (unsigned)G) dup |= (atomicOr(&bits[g >> 5], 1u << (g & 31)) >> (g & 31)) & 1u;
        }
        __syncwarp();
        for (int64_t k = b + lane; k < e; k += 32) {
            const unsigned g = (unsigned)col[k];
            if (g < (unsigned)G) bits[g >> 5] = 0u;
        }
        __syncwarp();
    }
    if (dup) atomicOr(bad, 8);
}

// Checks the device CSR once when a matrix is loaded or wrapped: the scoring kernel indexes its shared-memory
// membership table with the stored columns and trusts the row bounds. Bit 2 of the result is informational: some
// stored value is not > 0 (the scoring kernel then keeps its per-step test for such values).
int validate_csr(pctsea_ctx* c, const pctsea_matrix* m) {
    c->counters.reserve(64);
    int* bad = c->counters.as<int>();
    PCT_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), c->stream));
    k_validate_csr<<<c->num_sms * 8, 256, 0, c->stream>>>(m->row_ptr, m->col, m->val, m->n_cells, m->n_genes, m->nnz, bad);
    PCT_CUDA(cudaGetLastError());
    {
        int h0 = 0;   // the uniqueness pass walks the rows: only behind valid row bounds
        PCT_CUDA(cudaMemcpyAsync(&h0, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        PCT_CUDA(cudaStreamSynchronize(c->stream));
        if (!(h0 & 1) && m->n_cells > 0 && m->nnz > 0) {
            const int words = (int)div_up(std::max(m->n_genes, 1), 32);
            int warps = 8;
            while (warps > 1 && (size_t)warps * words * 4 > (size_t)c->smem_per_block_optin) warps >>= 1;
            const size_t smem = (size_t)warps * words * 4;
            PCT_REQUIRE(smem <= (size_t)c->smem_per_block_optin, "n_genes too large for the duplicate-column check");
            ensure_dyn_smem(c, k_validate_unique_cols, smem);
            const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)c->smem_per_sm / (smem + 1024)));
            k_validate_unique_cols<<<c->num_sms * per_sm, warps * 32, smem, c->stream>>>(m->row_ptr, m->col, m->n_cells,
                                                                                         m->n_genes, words, bad);
            PCT_CUDA(cudaGetLastError());
        }
    }
    int h = 0;
    PCT_CUDA(cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PCT_CUDA(cudaStreamSynchronize(c->stream));
    return h;
}

void launch_list_lookup(pctsea_ctx* c, const ScoreArgs& sa) {
    k_list_lookup<<<(sa.L + 255) / 256, 256, 0, c->stream>>>(sa.gene_idx_dev, sa.abund_dev, sa.L, sa.m->n_genes,
                                                              c->ybg.as<float>(), c->bitmap.as<uint32_t>());
    count_launch(c);
}

void launch_score_final(pctsea_ctx* c, const ScoreArgs& sa) {
    const int64_t N = sa.m->n_cells;
    k_score_final<<<(unsigned)div_up(N, 256), 256, 0, c->stream>>>(c->score_sums.as<double>(), c->nnzc.as<int32_t>(), N,
                                                                    sa.prm.method, sa.prm.has_min_corr, sa.prm.min_corr,
                                                                    c->score.as<double>());
    count_launch(c);
    PCT_CUDA(cudaGetLastError());
}

void launch_score(pctsea_ctx* c, const ScoreArgs& sa) {
    const pctsea_matrix* m = sa.m;
    if (m->gm_mo && m->n_cells > 0) {   // the matrix carries its gene-major copy: lane-per-cell kernel (score_gm.cu)
        if (!c->opt_score_path) { launch_score_gm(c, sa); return; }   // PCTSEA_OPT_SCORE_PATH = 1 forces the CSR scan
    }
    const int64_t N = m->n_cells;
    const int32_t G = m->n_genes;
    const int32_t bm_words = (G + 1 + 15) / 16;  // 2 bits per gene, plus the always-zero entry G
    // rank directory + sorted abundances for the on-chip gather (lists up to 4096 genes, <= 16 KB + G/4 bytes)
    const int rank_words = (sa.L > 0 && sa.L <= 4096) ? (G + 1 + 31) / 32 : 0;
    const size_t rank_bytes = rank_words ? (size_t)round_up((int64_t)rank_words * 8 + (int64_t)sa.L * 4, 16) : 0;
    c->ybg.reserve((size_t)G * 4 + 16);
    c->bitmap.reserve((size_t)bm_words * 4 + 16);
    c->list_rank.reserve((size_t)std::max<size_t>(rank_bytes, 16));
    c->score.reserve((size_t)N * 8 + 16);
    c->score_sums.reserve((size_t)N * 24 + 16);
    c->nused.reserve((size_t)N * 4 + 16);
    c->nnzc.reserve((size_t)N * 4 + 16);
    c->kept.reserve((size_t)N + 16);
    PCT_CUDA(cudaMemsetAsync(c->ybg.p, 0, (size_t)G * 4, c->stream));
    PCT_CUDA(cudaMemsetAsync(c->bitmap.p, 0, (size_t)bm_words * 4, c->stream));
    if (rank_bytes) PCT_CUDA(cudaMemsetAsync(c->list_rank.p, 0, rank_bytes, c->stream));
    if (sa.L > 0) {
        launch_list_lookup(c, sa);
        if (rank_words) {
            k_list_rank<<<1, 1024, 0, c->stream>>>(c->bitmap.as<uint32_t>(), c->ybg.as<float>(), G, rank_words,
                                                   c->list_rank.as<int2>(), c->list_rank.as<float>() + 2 * (size_t)rank_words);
            count_launch(c);
        }
    }
    if (N == 0) return;

    ScoreKernelArgs ka;
    ka.row_ptr = m->row_ptr; ka.col = m->col; ka.val = m->val; ka.col_end = m->col + m->nnz; ka.N = N; ka.G = G;
    ka.ybg = c->ybg.as<float>(); ka.bitmap = c->bitmap.as<uint32_t>(); ka.bm_words = bm_words;
    ka.rankw = c->list_rank.as<int2>(); ka.ab_sorted = c->list_rank.as<float>() + 2 * (size_t)rank_words;
    ka.rank_words = rank_words; ka.rank_bytes = (int32_t)rank_bytes; ka.n_list = sa.L;
    ka.method = sa.prm.method; ka.min_genes = sa.prm.min_genes_cells; ka.has_min_corr = sa.prm.has_min_corr;
    ka.jitter_mode = sa.prm.jitter_mode; ka.min_corr = sa.prm.min_corr; ka.seed = sa.prm.seed;
    ka.sums = c->score_sums.as<double>(); ka.nused = c->nused.as<int32_t>(); ka.nnzc = c->nnzc.as<int32_t>();
    ka.kept = c->kept.as<uint8_t>();

    // Shared memory: the membership table (one byte per gene when that leaves room for >= 16 warps, else two
    // bits per gene), the rank directory with the sorted abundances, and one pair buffer of >= L pairs per warp so that any cell fits.
    const int cap = (int)round_up(std::max(sa.L, 32), 4);
    // 227 KB per CTA minus the kernel's static shared memory (cell counter and per-warp queues, ~4 KB)
    const size_t smem_limit = 227 * 1024 - (sizeof(int4) * 8 * kScoreMaxWarps + 64);
    const size_t pair_bytes = (size_t)cap * 8, rank_smem = rank_bytes;
    const size_t byte_tab = (size_t)round_up((int64_t)G + 1, 16), bit_tab = (size_t)round_up((int64_t)bm_words * 4, 16);
    const bool byte_mode = byte_tab + rank_smem + 16 * pair_bytes <= smem_limit;
    const size_t tab_bytes = byte_mode ? byte_tab : bit_tab;
    if (tab_bytes + rank_smem + pair_bytes > smem_limit)
        throw Error(PCTSEA_EINVAL, "input list too long for the on-chip pair buffer (L=" + std::to_string(sa.L) + ")");
    int maxw = kScoreMaxWarps;
    if (c->opt_score_warps > 0) maxw = c->opt_score_warps;   // PCTSEA_OPT_SCORE_WARPS (tuning aid)
    const int warps = (int)std::max<size_t>(1, std::min<size_t>(maxw, (smem_limit - tab_bytes - rank_smem) / pair_bytes));
    const size_t smem = tab_bytes + rank_smem + (size_t)warps * pair_bytes;
    ka.cap = cap;
    ka.tab_bytes = (int32_t)tab_bytes;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(c->num_sms, div_up(N, kScoreCellBatch)));
#define PCT_LAUNCH_SCORE2(B, W, AP)                                                                           \
    do {                                                                                                      \
        ensure_dyn_smem(c, k_score<B, W, 3, AP>, (size_t)(smem)); \
        k_score<B, W, 3, AP><<<grid, warps * 32, smem, c->stream>>>(ka);                                      \
    } while (0)
#define PCT_LAUNCH_SCORE(B, W)                                                                                \
    do {                                                                                                      \
        if (m->all_positive) PCT_LAUNCH_SCORE2(B, W, true); else PCT_LAUNCH_SCORE2(B, W, false);              \
    } while (0)
    // the launch bound (registers per thread) follows the warp count that shared memory allows
    if (byte_mode) {
        if (warps > 20) PCT_LAUNCH_SCORE(true, 24);
        else if (warps > 16) PCT_LAUNCH_SCORE(true, 20);
        else PCT_LAUNCH_SCORE(true, 16);
    } else {
        if (warps > 20) PCT_LAUNCH_SCORE(false, 24);
        else if (warps > 16) PCT_LAUNCH_SCORE(false, 20);
        else PCT_LAUNCH_SCORE(false, 16);
    }
#undef PCT_LAUNCH_SCORE2
#undef PCT_LAUNCH_SCORE
    count_launch(c);
    PCT_CUDA(cudaGetLastError());
    launch_score_final(c, sa);
}

}  // namespace pctsea
