// score.cu — stage 1: per-cell scoring of the device-resident CSR against the input list.
//
// Replaces (reference paths under pctsea-core/src/main/java/edu/scripps/yates/pctsea/):
//   a2  model/SingleCell.java:344-369  hasMinNumberOfExpressedGenes   (driver PCTSEA.java:544-572)
//   a3  model/SingleCell.java:380-463  calculateCorrelation / calculateSimpleScore (driver PCTSEA.java:2305-2343)
//       with commons-math3 PearsonsCorrelation -> SimpleRegression restated (un-vendored dependency).
//
// Layout: one warp owns a batch of 32 consecutive cells. For each cell the warp streams the CSR row with
// coalesced loads, tests list membership against a shared-memory bitmap, and compacts the matched
// (abundance, expression) pairs in CSR order into a per-warp shared-memory buffer (ballot + popc).
// When the buffer is full or the batch is done, lane c runs cell c's SimpleRegression update sequentially
// in fp64 with explicit round-to-nearest intrinsics (no FMA contraction), so the score is bit-identical to
// the CPU oracle's, which walks the same pairs in the same order.
//
// Roofline: HBM. Algorithmic bytes = 8 B * nnz (col + val) + 16 B * N (row_ptr) + 17 B * N outputs.
#include "common.cuh"

namespace pctsea {

__global__ void k_list_lookup(const int32_t* __restrict__ gene_idx, const float* __restrict__ abund, int L, int G,
                              float* __restrict__ ybg, uint32_t* __restrict__ bitmap) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    int g = gene_idx[l];
    if (g < 0 || g >= G) return;
    ybg[g] = abund[l];
    atomicOr(&bitmap[g >> 5], 1u << (g & 31));
}

__device__ __forceinline__ double jitter_uniform(uint64_t seed, uint64_t cell, uint32_t vec, uint32_t i) {
    uint32_t o[4];
    philox4x32_10(i, vec | 0x4A000000u, (uint32_t)cell, (uint32_t)(cell >> 32), (uint32_t)seed,
                  (uint32_t)(seed >> 32), o);
    uint64_t u = ((uint64_t)(o[0] >> 6) << 27) | (uint64_t)(o[1] >> 5);
    return __dmul_rn((double)u, 1.0 / 9007199254740992.0);
}

// commons-math3 SimpleRegression.addData + getR, restated; pairs are (x = abundance, y = expression).
__device__ double pearson_sequential(const float2* __restrict__ pr, int np, bool jit_cell, bool jit_list,
                                     uint64_t seed, uint64_t cell) {
    double sumXX = 0.0, sumYY = 0.0, sumXY = 0.0, xbar = 0.0, ybar = 0.0;
    for (int i = 0; i < np; i++) {
        float2 v = pr[i];
        double x = (double)v.x, y = (double)v.y;
        if (jit_cell) y = __dadd_rn(y, __ddiv_rn(jitter_uniform(seed, cell, 0u, (uint32_t)i), 1000000.0));
        if (jit_list) x = __dadd_rn(x, __ddiv_rn(jitter_uniform(seed, cell, 1u, (uint32_t)i), 1000000.0));
        if (i == 0) {
            xbar = x; ybar = y;
        } else {
            double nd = (double)i;
            double fact1 = __dadd_rn(1.0, nd);
            double fact2 = __ddiv_rn(nd, __dadd_rn(1.0, nd));
            double dx = __dsub_rn(x, xbar), dy = __dsub_rn(y, ybar);
            sumXX = __dadd_rn(sumXX, __dmul_rn(__dmul_rn(dx, dx), fact2));
            sumYY = __dadd_rn(sumYY, __dmul_rn(__dmul_rn(dy, dy), fact2));
            sumXY = __dadd_rn(sumXY, __dmul_rn(__dmul_rn(dx, dy), fact2));
            xbar = __dadd_rn(xbar, __ddiv_rn(dx, fact1));
            ybar = __dadd_rn(ybar, __ddiv_rn(dy, fact1));
        }
    }
    const double ten_min = __longlong_as_double(10LL);  // 10 * Double.MIN_VALUE
    double b1 = (fabs(sumXX) < ten_min) ? __longlong_as_double(0x7ff8000000000000LL) : __ddiv_rn(sumXY, sumXX);
    double sse_arg = __dsub_rn(sumYY, __ddiv_rn(__dmul_rn(sumXY, sumXY), sumXX));
    double sse = (sse_arg != sse_arg) ? sse_arg : (sse_arg > 0.0 ? sse_arg : 0.0);
    double rsq = __ddiv_rn(__dsub_rn(sumYY, sse), sumYY);
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

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_score(const ScoreKernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* bm = reinterpret_cast<uint32_t*>(smem_raw);
    const int bm_bytes = ((a.bm_words * 4 + 15) / 16) * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* pairs = reinterpret_cast<float2*>(smem_raw + bm_bytes) + (size_t)warp * a.cap;

    for (int i = threadIdx.x; i < a.bm_words; i += WARPS * 32) bm[i] = a.bitmap[i];
    __syncthreads();

    const unsigned lt_mask = (1u << lane) - 1u;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const int64_t n_batches = (a.N + 31) / 32;
    const int64_t warp_global = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t warp_stride = (int64_t)gridDim.x * WARPS;

    for (int64_t batch = warp_global; batch < n_batches; batch += warp_stride) {
        const int64_t cell0 = batch * 32;
        const int64_t my_cell = cell0 + lane;
        int64_t rp_lo = 0, rp_hi = 0;
        if (my_cell < a.N) { rp_lo = a.row_ptr[my_cell]; rp_hi = a.row_ptr[my_cell + 1]; }
        const int ncell = (int)min((int64_t)32, a.N - cell0);

        int my_off = 0, my_np = 0, my_nz = 0;
        int used = 0, sub_start = 0, c = 0;

        // Flush: lanes sub_start..last run their cell's regression from the compacted pairs.
        auto flush = [&](int first, int last) {
            __syncwarp();
            if (lane >= first && lane <= last) {
                const int np = my_np, nz = my_nz;
                const bool keep = (nz > 0) && (np >= a.min_genes);
                double sc = qnan;
                if (keep && np >= 2) {
                    const float2* pr = pairs + my_off;
                    bool eq_cell = true, eq_list = true;  // zero variance <=> all elements equal
                    const float2 v0 = pr[0];
                    for (int i = 1; i < np; i++) {
                        float2 v = pr[i];
                        eq_list = eq_list && (v.x == v0.x);
                        eq_cell = eq_cell && (v.y == v0.y);
                    }
                    if (!((eq_cell || eq_list) && a.jitter_mode == PCTSEA_JITTER_NAN)) {
                        double r = pearson_sequential(pr, np, eq_cell, eq_list, a.seed, (uint64_t)my_cell);
                        if (a.has_min_corr && r < a.min_corr) r = qnan;
                        else if (a.method == PCTSEA_METHOD_SIMPLE_SCORE) r = __dadd_rn(r, (double)nz);
                        sc = r;
                    }
                }
                a.score[my_cell] = sc;
                a.nused[my_cell] = np;
                a.nnzc[my_cell] = nz;
                a.kept[my_cell] = keep ? 1 : 0;
            }
            __syncwarp();
        };

        while (c < ncell) {
            const int64_t beg = __shfl_sync(0xffffffffu, rp_lo, c);
            const int64_t end = __shfl_sync(0xffffffffu, rp_hi, c);
            int cnt = 0, cnt_nz = 0;
            for (int64_t k0 = beg; k0 < end; k0 += 128) {
                int32_t g[4];
                float x[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    int64_t k = k0 + u * 32 + lane;
                    bool ok = k < end;
                    g[u] = ok ? __ldcs(a.col + k) : -1;
                    x[u] = ok ? __ldcs(a.val + k) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    bool inlist = false;
                    if (g[u] >= 0 && g[u] < a.G) inlist = (bm[g[u] >> 5] >> (g[u] & 31)) & 1u;
                    float y = inlist ? __ldg(a.ybg + g[u]) : 0.0f;
                    bool ispair = inlist && (x[u] > 0.0f) && (y > 0.0f);
                    bool isnz = inlist && (x[u] != 0.0f);  // true for NaN
                    unsigned pm = __ballot_sync(0xffffffffu, ispair);
                    unsigned zm = __ballot_sync(0xffffffffu, isnz);
                    if (ispair) {
                        int pos = used + cnt + __popc(pm & lt_mask);
                        if (pos < a.cap) pairs[pos] = make_float2(y, x[u]);
                    }
                    cnt += __popc(pm);
                    cnt_nz += __popc(zm);
                }
            }
            if (used == 0 && cnt > a.cap) cnt = a.cap;  // only with duplicate columns in a row (unsupported input)
            if (used + cnt > a.cap) {
                // does not fit behind the cells already buffered: score those, then rescan this cell
                // (cnt <= L <= cap, so after the flush it always fits; the row is still in L2)
                flush(sub_start, c - 1);
                used = 0;
                sub_start = c;
                continue;
            }
            if (lane == c) { my_off = used; my_np = cnt; my_nz = cnt_nz; }
            used += cnt;
            c++;
        }
        flush(sub_start, ncell - 1);
    }
}

void launch_score(pctsea_ctx* c, const ScoreArgs& sa) {
    const pctsea_matrix* m = sa.m;
    const int64_t N = m->n_cells;
    const int32_t G = m->n_genes;
    const int32_t bm_words = (G + 31) / 32;
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

    // Pair buffer: >= L pairs per warp so any single cell fits; sized so that 2-3 CTAs share an SM.
    const int bm_bytes = ((bm_words * 4 + 15) / 16) * 16;
    int cap = (int)round_up(std::max(sa.L, 1024), 32);
    const size_t smem_limit = 200 * 1024;
    int warps = 8;
    while (warps > 1 && (size_t)bm_bytes + (size_t)warps * cap * 8 > smem_limit) warps >>= 1;
    size_t smem = (size_t)bm_bytes + (size_t)warps * cap * 8;
    if (smem > 227 * 1024)
        throw Error(PCTSEA_EINVAL, "input list too long for the on-chip pair buffer (L=" + std::to_string(sa.L) + ")");
    ka.cap = cap;
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / (smem + 1024)));
    int64_t n_batches = (N + 31) / 32;
    int grid = (int)std::min<int64_t>((int64_t)kNumSMs * ctas_per_sm, div_up(n_batches, warps));
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
