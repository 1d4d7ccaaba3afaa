// score.cu — stage 1: per-cell scoring of the device-resident CSR against the input list.
//
// Replaces (reference paths under pctsea-core/src/main/java/edu/scripps/yates/pctsea/):
//   a2  model/SingleCell.java:344-369  hasMinNumberOfExpressedGenes   (driver PCTSEA.java:544-572)
//   a3  model/SingleCell.java:380-463  calculateCorrelation / calculateSimpleScore (driver PCTSEA.java:2305-2343)
//       with commons-math3 PearsonsCorrelation -> SimpleRegression restated (un-vendored dependency).
//
// Layout: one warp per cell. The warp streams the CSR row with coalesced loads, tests list membership
// against a shared-memory bitmap and compacts the matched (abundance, expression) pairs in CSR order into a
// per-warp shared-memory buffer (ballot + popc). Pair k then belongs to lane k % 32: every lane runs the
// SimpleRegression recurrence over its strided sub-stream in fp64, and the 32 partial states are merged with
// a fixed xor-shuffle tree (Chan et al. parallel update; lower lane = left operand). All arithmetic uses
// explicit round-to-nearest intrinsics (no FMA contraction), so the score is bit-identical to the CPU
// oracle's orc_pearson_strided32, which walks the same sub-streams and the same tree, and within ~1e-13
// relative of the reference's sequential update (whose own pair order is a hash-set order).
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

struct RegState {
    int n;
    double xbar, ybar, sxx, syy, sxy;
};

// commons-math3 SimpleRegression.addData, restated (x = abundance, y = expression).
__device__ __forceinline__ void reg_add(RegState& s, double x, double y) {
    if (s.n == 0) {
        s.xbar = x; s.ybar = y;
    } else {
        const double nd = (double)s.n;
        const double fact1 = __dadd_rn(1.0, nd);
        const double fact2 = __ddiv_rn(nd, __dadd_rn(1.0, nd));
        const double dx = __dsub_rn(x, s.xbar), dy = __dsub_rn(y, s.ybar);
        s.sxx = __dadd_rn(s.sxx, __dmul_rn(__dmul_rn(dx, dx), fact2));
        s.syy = __dadd_rn(s.syy, __dmul_rn(__dmul_rn(dy, dy), fact2));
        s.sxy = __dadd_rn(s.sxy, __dmul_rn(__dmul_rn(dx, dy), fact2));
        s.xbar = __dadd_rn(s.xbar, __ddiv_rn(dx, fact1));
        s.ybar = __dadd_rn(s.ybar, __ddiv_rn(dy, fact1));
    }
    s.n++;
}

// Parallel update of two partial states (a = left operand).
__device__ __forceinline__ RegState reg_merge(const RegState& a, const RegState& b) {
    if (b.n == 0) return a;
    if (a.n == 0) return b;
    RegState r;
    const double na = (double)a.n, nb = (double)b.n;
    const double n = __dadd_rn(na, nb);
    const double dx = __dsub_rn(b.xbar, a.xbar), dy = __dsub_rn(b.ybar, a.ybar);
    const double w = __ddiv_rn(__dmul_rn(na, nb), n);
    r.sxx = __dadd_rn(__dadd_rn(a.sxx, b.sxx), __dmul_rn(__dmul_rn(dx, dx), w));
    r.syy = __dadd_rn(__dadd_rn(a.syy, b.syy), __dmul_rn(__dmul_rn(dy, dy), w));
    r.sxy = __dadd_rn(__dadd_rn(a.sxy, b.sxy), __dmul_rn(__dmul_rn(dx, dy), w));
    const double f = __ddiv_rn(nb, n);
    r.xbar = __dadd_rn(a.xbar, __dmul_rn(dx, f));
    r.ybar = __dadd_rn(a.ybar, __dmul_rn(dy, f));
    r.n = a.n + b.n;
    return r;
}

// SimpleRegression.getR(): sign(slope) * sqrt(RSquare)
__device__ __forceinline__ double reg_finish(const RegState& s) {
    const double sumXX = s.sxx, sumYY = s.syy, sumXY = s.sxy;
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
    const int64_t warp_global = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t warp_stride = (int64_t)gridDim.x * WARPS;

    for (int64_t cell = warp_global; cell < a.N; cell += warp_stride) {
        const int64_t beg = a.row_ptr[cell], end = a.row_ptr[cell + 1];
        const int32_t* colp = a.col + beg;
        const float* valp = a.val + beg;
        const int len = (int)(end - beg);
        int cnt = 0, cnt_nz = 0;
        for (int j0 = 0; j0 < len; j0 += 128) {
            int32_t g[4];
            float x[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = j0 + u * 32 + lane;
                const bool ok = j < len;
                g[u] = ok ? __ldcs(colp + j) : -1;
                x[u] = ok ? __ldcs(valp + j) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                bool inlist = false;
                if ((unsigned)g[u] < (unsigned)a.G) inlist = (bm[g[u] >> 5] >> (g[u] & 31)) & 1u;
                const float y = inlist ? __ldg(a.ybg + g[u]) : 0.0f;
                const bool ispair = inlist && (x[u] > 0.0f) && (y > 0.0f);
                const bool isnz = inlist && (x[u] != 0.0f);  // true for NaN
                const unsigned pm = __ballot_sync(0xffffffffu, ispair);
                const unsigned zm = __ballot_sync(0xffffffffu, isnz);
                if (ispair) {
                    const int pos = cnt + __popc(pm & lt_mask);
                    if (pos < a.cap) pairs[pos] = make_float2(y, x[u]);
                }
                cnt += __popc(pm);
                cnt_nz += __popc(zm);
            }
        }
        __syncwarp();
        const int np = min(cnt, a.cap);  // cnt > cap only with duplicate columns in a row (unsupported input)
        const bool keep = (cnt_nz > 0) && (cnt >= a.min_genes);
        double sc = qnan;
        if (keep && np >= 2) {
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
                RegState st;
                st.n = 0; st.xbar = 0.0; st.ybar = 0.0; st.sxx = 0.0; st.syy = 0.0; st.sxy = 0.0;
                for (int k = lane; k < np; k += 32) {
                    const float2 v = pairs[k];
                    double x = (double)v.x, y = (double)v.y;
                    if (eq_cell) y = __dadd_rn(y, __ddiv_rn(jitter_uniform(a.seed, (uint64_t)cell, 0u, (uint32_t)k), 1000000.0));
                    if (eq_list) x = __dadd_rn(x, __ddiv_rn(jitter_uniform(a.seed, (uint64_t)cell, 1u, (uint32_t)k), 1000000.0));
                    reg_add(st, x, y);
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    RegState pt;
                    pt.n = __shfl_xor_sync(0xffffffffu, st.n, o);
                    pt.xbar = __shfl_xor_sync(0xffffffffu, st.xbar, o);
                    pt.ybar = __shfl_xor_sync(0xffffffffu, st.ybar, o);
                    pt.sxx = __shfl_xor_sync(0xffffffffu, st.sxx, o);
                    pt.syy = __shfl_xor_sync(0xffffffffu, st.syy, o);
                    pt.sxy = __shfl_xor_sync(0xffffffffu, st.sxy, o);
                    st = (lane & o) ? reg_merge(pt, st) : reg_merge(st, pt);
                }
                double r = reg_finish(st);
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
