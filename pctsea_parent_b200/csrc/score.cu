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

struct ScoreKernelArgs {
    const int64_t* row_ptr;
    const int32_t* col;
    const float* val;
    int64_t N;
    int32_t G;
    const float* ybg;
    const uint32_t* bitmap;
    int32_t bm_words;
    int32_t cap;  // pairs per warp buffer, >= L
    int32_t method, min_genes, has_min_corr, jitter_mode;
    double min_corr;
    uint64_t seed;
    double* score;
    int32_t* nused;
    int32_t* nnzc;
    uint8_t* kept;
};

struct RowChunk {
    int32_t g[4];
    float x[4];
};

// 4 consecutive non-zeros of the row per lane, starting at row offset j (a multiple of 4 relative to the
// 16-byte-aligned position at or below the row start): one 128-bit load each for col and val when the warp's
// whole 128-element step lies inside the row (`interior`, warp-uniform), element-wise at the row's head and
// tail. Elements outside the row come back as g = -1.
__device__ __forceinline__ RowChunk load_chunk(const int32_t* __restrict__ col, const float* __restrict__ val,
                                               int64_t kk, int64_t beg, int64_t end, bool interior) {
    RowChunk c;
    if (interior) {
        const int4 gv = __ldcs(reinterpret_cast<const int4*>(col + kk));
        const float4 xv = __ldcs(reinterpret_cast<const float4*>(val + kk));
        c.g[0] = gv.x; c.g[1] = gv.y; c.g[2] = gv.z; c.g[3] = gv.w;
        c.x[0] = xv.x; c.x[1] = xv.y; c.x[2] = xv.z; c.x[3] = xv.w;
        return c;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) { c.g[q] = -1; c.x[q] = 0.0f; }
    if (kk < end) {
        if (kk >= beg && kk + 4 <= end) {
            const int4 gv = __ldcs(reinterpret_cast<const int4*>(col + kk));
            const float4 xv = __ldcs(reinterpret_cast<const float4*>(val + kk));
            c.g[0] = gv.x; c.g[1] = gv.y; c.g[2] = gv.z; c.g[3] = gv.w;
            c.x[0] = xv.x; c.x[1] = xv.y; c.x[2] = xv.z; c.x[3] = xv.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int64_t k = kk + q;
                if (k >= beg && k < end) { c.g[q] = __ldcs(col + k); c.x[q] = __ldcs(val + k); }
            }
        }
    }
    return c;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_score(const ScoreKernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* bm = reinterpret_cast<uint32_t*>(smem_raw);          // 2 bits per gene (in list, pairable)
    const int bm_bytes = ((a.bm_words * 4 + 15) / 16) * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // pair buffer: .x = matrix column (int bits) until the abundances are gathered, then the abundance
    float2* pairs = reinterpret_cast<float2*>(smem_raw + bm_bytes) + (size_t)warp * a.cap;

    for (int i = threadIdx.x; i < a.bm_words; i += WARPS * 32) bm[i] = a.bitmap[i];
    __syncthreads();

    const unsigned lt_mask = (1u << lane) - 1u;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const int64_t warp_global = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t warp_stride = (int64_t)gridDim.x * WARPS;
    const unsigned Gu = (unsigned)a.G;  // bm has a zero entry for index G: out-of-range / padding slots

    int64_t cell = warp_global;
    if (cell >= a.N) return;
    int64_t beg = a.row_ptr[cell], end = a.row_ptr[cell + 1];
    // chunk `idx` (128 non-zeros per warp) of the row [b, e); past the row's last chunk it comes back empty
    auto fetch = [&](int64_t b, int64_t e, int64_t idx) {
        const int64_t k0 = (b & ~(int64_t)3) + idx * 128;
        return load_chunk(a.col, a.val, k0 + 4 * lane, b, e, k0 >= b && k0 + 128 <= e);
    };
    // software pipeline two steps deep: while step i is compacted, steps i+1 and i+2 (or the first steps of the
    // warp's next cell) are in flight
    RowChunk cur = fetch(beg, end, 0), n1 = fetch(beg, end, 1);
    while (cell < a.N) {
        // row bounds of the warp's next cell, fetched while this one is scanned
        const int64_t ncell = cell + warp_stride;
        int64_t nbeg = 0, nend = 0;
        if (ncell < a.N) { nbeg = a.row_ptr[ncell]; nend = a.row_ptr[ncell + 1]; }
        const int64_t kstart = beg & ~(int64_t)3;
        const int64_t m = (end > beg) ? (end - kstart + 127) / 128 : 0;   // steps of this row
        int cnt = 0, my_nz = 0;
        for (int64_t i = 0; i < m; i++) {
            const RowChunk n2 = (i + 2 < m) ? fetch(beg, end, i + 2) : fetch(nbeg, nend, i + 2 - m);
            unsigned pmask = 0u, zmask = 0u;   // bit q: element q is a pair / a non-zero list gene
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned g = min((unsigned)cur.g[q], Gu);   // -1 (padding) and bad columns -> the zero entry
                const unsigned bits = (bm[g >> 4] >> ((g & 15u) * 2u)) & 3u;
                pmask |= ((bits >> 1) & (cur.x[q] > 0.0f ? 1u : 0u)) << q;
                zmask |= ((bits & 1u) & (cur.x[q] != 0.0f ? 1u : 0u)) << q;  // true for NaN
            }
            my_nz += __popc(zmask);
            const int mine = __popc(pmask);
            // exclusive prefix over lanes of `mine` (0..4) from the ballots of its three bit planes;
            // CSR order inside the step is (lane, q): lower lanes first, then my own earlier elements
            const unsigned p0 = __ballot_sync(0xffffffffu, mine & 1), p1 = __ballot_sync(0xffffffffu, mine & 2),
                           p2 = __ballot_sync(0xffffffffu, mine & 4);
            int pos = cnt + __popc(p0 & lt_mask) + 2 * __popc(p1 & lt_mask) + 4 * __popc(p2 & lt_mask);
            cnt += __popc(p0) + 2 * __popc(p1) + 4 * __popc(p2);
            if (pmask) {
                if (cnt <= a.cap) {   // warp-uniform: the whole step fits (always, unless a row repeats columns)
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (pmask & (1u << q)) pairs[pos++] = make_float2(__int_as_float(cur.g[q]), cur.x[q]);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (pmask & (1u << q)) { if (pos < a.cap) pairs[pos] = make_float2(__int_as_float(cur.g[q]), cur.x[q]); pos++; }
                }
            }
            cur = n1;
            n1 = n2;
        }
        // invariant for the next cell: cur = its step 0, n1 = its step 1. Rows with fewer than two steps were
        // primed with their own (empty) steps instead, so refill those slots here.
        if (m == 0) { cur = fetch(nbeg, nend, 0); n1 = fetch(nbeg, nend, 1); }
        else if (m == 1) cur = fetch(nbeg, nend, 0);
        int cnt_nz = my_nz;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) cnt_nz += __shfl_xor_sync(0xffffffffu, cnt_nz, o);
        __syncwarp();
        const int np = min(cnt, a.cap);  // cnt > cap only with duplicate columns in a row (unsupported input)
        const bool keep = (cnt_nz > 0) && (cnt >= a.min_genes);
        double sc = qnan;
        if (keep && np >= 2) {
            // gather the abundances of the matched genes (L1/L2-resident table)
            for (int k = lane; k < np; k += 32) pairs[k].x = __ldg(a.ybg + __float_as_int(pairs[k].x));
            __syncwarp();
            // zero variance <=> all elements equal (Maths.var == 0, model/SingleCell.java:430-441)
            const float2 v0 = pairs[0];
            bool eq_list = true, eq_cell = true;
            for (int k = lane; k < np; k += 32) {
                const float2 v = pairs[k];
                eq_list = eq_list && (v.x == v0.x);
                eq_cell = eq_cell && (v.y == v0.y);
            }
            eq_list = __all_sync(0xffffffffu, eq_list);
            eq_cell = __all_sync(0xffffffffu, eq_cell);
            if (!((eq_cell || eq_list) && a.jitter_mode == PCTSEA_JITTER_NAN)) {
                double sx = 0.0, sy = 0.0;
                for (int k = lane; k < np; k += 32) {
                    const float2 v = pairs[k];
                    double x = (double)v.x, y = (double)v.y;
                    if (eq_cell) y = __dadd_rn(y, __ddiv_rn(jitter_uniform(a.seed, (uint64_t)cell, 0u, (uint32_t)k), 1000000.0));
                    if (eq_list) x = __dadd_rn(x, __ddiv_rn(jitter_uniform(a.seed, (uint64_t)cell, 1u, (uint32_t)k), 1000000.0));
                    sx = __dadd_rn(sx, x);
                    sy = __dadd_rn(sy, y);
                }
                const double xbar = __ddiv_rn(warp_sum_tree(sx), (double)np);
                const double ybar = __ddiv_rn(warp_sum_tree(sy), (double)np);
                double sxx = 0.0, syy = 0.0, sxy = 0.0;
                for (int k = lane; k < np; k += 32) {
                    const float2 v = pairs[k];
                    double x = (double)v.x, y = (double)v.y;
                    if (eq_cell) y = __dadd_rn(y, __ddiv_rn(jitter_uniform(a.seed, (uint64_t)cell, 0u, (uint32_t)k), 1000000.0));
                    if (eq_list) x = __dadd_rn(x, __ddiv_rn(jitter_uniform(a.seed, (uint64_t)cell, 1u, (uint32_t)k), 1000000.0));
                    const double dx = __dsub_rn(x, xbar), dy = __dsub_rn(y, ybar);
                    sxx = __dadd_rn(sxx, __dmul_rn(dx, dx));
                    syy = __dadd_rn(syy, __dmul_rn(dy, dy));
                    sxy = __dadd_rn(sxy, __dmul_rn(dx, dy));
                }
                sxx = warp_sum_tree(sxx);
                syy = warp_sum_tree(syy);
                sxy = warp_sum_tree(sxy);
                double r = pearson_finish(sxx, syy, sxy);
                if (a.has_min_corr && r < a.min_corr) r = qnan;
                else if (a.method == PCTSEA_METHOD_SIMPLE_SCORE) r = __dadd_rn(r, (double)cnt_nz);
                sc = r;
            }
        }
        if (lane == 0) {
            a.score[cell] = sc;
            a.nused[cell] = cnt;
            a.nnzc[cell] = cnt_nz;
            a.kept[cell] = keep ? 1 : 0;
        }
        __syncwarp();
        cell = ncell; beg = nbeg; end = nend;
    }
}

void launch_score(pctsea_ctx* c, const ScoreArgs& sa) {
    const pctsea_matrix* m = sa.m;
    const int64_t N = m->n_cells;
    const int32_t G = m->n_genes;
    const int32_t bm_words = (G + 1 + 15) / 16;  // 2 bits per gene, plus the always-zero entry G
    c->ybg.reserve((size_t)G * 4 + 16);
    c->bitmap.reserve((size_t)bm_words * 4 + 16);
    c->score.reserve((size_t)N * 8 + 16);
    c->nused.reserve((size_t)N * 4 + 16);
    c->nnzc.reserve((size_t)N * 4 + 16);
    c->kept.reserve((size_t)N + 16);
    PCT_CUDA(cudaMemsetAsync(c->ybg.p, 0, (size_t)G * 4, c->stream));
    PCT_CUDA(cudaMemsetAsync(c->bitmap.p, 0, (size_t)bm_words * 4, c->stream));
    if (sa.L > 0) {
        k_list_lookup<<<(sa.L + 255) / 256, 256, 0, c->stream>>>(sa.gene_idx_dev, sa.abund_dev, sa.L, G,
                                                                  c->ybg.as<float>(), c->bitmap.as<uint32_t>());
        count_launch(c);
    }
    if (N == 0) return;

    ScoreKernelArgs ka;
    ka.row_ptr = m->row_ptr; ka.col = m->col; ka.val = m->val; ka.N = N; ka.G = G;
    ka.ybg = c->ybg.as<float>(); ka.bitmap = c->bitmap.as<uint32_t>(); ka.bm_words = bm_words;
    ka.method = sa.prm.method; ka.min_genes = sa.prm.min_genes_cells; ka.has_min_corr = sa.prm.has_min_corr;
    ka.jitter_mode = sa.prm.jitter_mode; ka.min_corr = sa.prm.min_corr; ka.seed = sa.prm.seed;
    ka.score = c->score.as<double>(); ka.nused = c->nused.as<int32_t>(); ka.nnzc = c->nnzc.as<int32_t>();
    ka.kept = c->kept.as<uint8_t>();

    // Pair buffer: >= L pairs per warp so any cell fits; as many warps per SM as shared memory allows.
    const int bm_bytes = ((bm_words * 4 + 15) / 16) * 16;
    int cap = (int)round_up(std::max(sa.L, 32), 32);
    const size_t smem_limit = 220 * 1024;
    int warps = 8;
    while (warps > 1 && (size_t)bm_bytes + (size_t)warps * cap * 8 > smem_limit) warps >>= 1;
    size_t smem = (size_t)bm_bytes + (size_t)warps * cap * 8;
    if (smem > 227 * 1024)
        throw Error(PCTSEA_EINVAL, "input list too long for the on-chip pair buffer (L=" + std::to_string(sa.L) + ")");
    ka.cap = cap;
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(64 / warps, (224 * 1024) / (smem + 1024)));
    int grid = (int)std::min<int64_t>((int64_t)kNumSMs * ctas_per_sm, div_up(N, warps));
    if (grid < 1) grid = 1;
#define PCT_LAUNCH_SCORE(W)                                                                                   \
    do {                                                                                                      \
        PCT_CUDA(cudaFuncSetAttribute(k_score<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        k_score<W><<<grid, W * 32, smem, c->stream>>>(ka);                                                    \
    } while (0)
    switch (warps) {
        case 8: PCT_LAUNCH_SCORE(8); break;
        case 4: PCT_LAUNCH_SCORE(4); break;
        case 2: PCT_LAUNCH_SCORE(2); break;
        default: PCT_LAUNCH_SCORE(1); break;
    }
#undef PCT_LAUNCH_SCORE
    count_launch(c);
    PCT_CUDA(cudaGetLastError());
}

}  // namespace pctsea
