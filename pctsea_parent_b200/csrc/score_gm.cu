// score_gm.cu — stage 1 over a gene-major copy of the matrix (built once per dataset, opt-in).
//
// Replaces the same reference code as score.cu (a2: model/SingleCell.java:344-369, a3: :380-463, drivers
// PCTSEA.java:544-572, 2305-2343; paths under pctsea-core/src/main/java/edu/scripps/yates/pctsea/).
//
// Why a second layout: the CSR scan of score.cu reads all 8 B * nnz of the matrix although only the list's genes
// (~4 % of the columns, ~13 % of the non-zeros of config B) can form pairs, and it is bound by the instructions of
// the membership test per non-zero, not by HBM. The reference holds, per cell, exactly the list-gene values
// (SingleCell.expressionIndexes, model/SingleCell.java:49-104, filled per input list from MongoDB); the layout
// here lets the GPU fetch just those without a per-list rebuild:
//
//   cells are grouped by 32 (group = warp, lane = cell). For every (group, gene):
//     mo(group, gene) = { mask: bit l set iff the group's cell l has a non-zero of this gene,
//                         off : position of the group's first value of this gene among the values of its TILE }
//   stored in tiles of 8 groups (= the warps of one scoring CTA), gene-major inside a tile (see gm_row_offset),
//   gval[tile_base + off + popc(mask & lanes below l)] = that cell's value: the tile's non-zeros in
//   (gene, group, cell) order, so that the values of one gene for the 256 cells of a CTA are ONE contiguous run
//   (~13 values on config B) instead of eight runs of ~1.6 values in eight different sectors.
//   Size: 8 B * G * N/32 (config B 3.75 GB, D 15 GB) + 4 B * nnz.
//
// Scoring a list: one warp per group walks the list's genes, sorted by gene, in chunks of 32; the chunk's 32 masks
// are bit-transposed so that each lane holds the set of the chunk's genes its own cell expresses and pops them in
// ascending order (the popped gene's 8-byte entry comes from the lane that fetched it, by shuffle); each lane keeps
// its cell's running sums in registers. No shared memory, no atomics, no cross-lane reduction; the pair order of a
// cell is the gene order whatever the storage order of its CSR row. Cells are assigned to groups by cell type and row
// length (a chunk costs a group as many steps as its busiest cell has genes in it). The Pearson sums are the shifted-data form (deviations from
// the cell's first pair), bit-identical to the oracle's orc_pearson_shifted and within 1e-9 relative of the
// reference's sequential SimpleRegression update (measured <= 1e-11).
//
// Roofline: the bytes this path must read are 8 B per (list gene, group) entry + the matched values (config B:
// 0.15 GB + 0.7 GB useful, 1.31 GB of DRAM sectors measured — 2.06 GB while every group had its own value array),
// not the 6.5 GB CSR: it sits between instruction issue (61 %) and the latency of the dependent entry -> value loads,
// not at the HBM roofline (profiles/README.md).
#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace pctsea {

namespace {

constexpr unsigned FULL = 0xffffffffu;

// Layout of the {mask, offset} entries: tiles of GM_TILE consecutive groups (= the warps of one scoring CTA), and
// inside a tile gene-major: entry(group, gene) = mo[((group / GM_TILE) * G + gene) * GM_TILE + group % GM_TILE].
// The eight entries of a gene for the eight groups of a CTA share one 64-byte line, and a CTA's whole tile is one
// contiguous 64 B * G region.
constexpr int GM_TILE = 8;
__host__ __device__ __forceinline__ size_t gm_row_offset(int64_t grp, int32_t G) {
    return (size_t)(grp / GM_TILE) * (size_t)G * GM_TILE + (size_t)(grp % GM_TILE);
}

// ---- build, three kernels. k_gm_masks: one CTA per group of 32 cells; the genes go through shared memory in ranges of
// Gs (one pass for matrices of up to ~51 k genes, more passes over the group's rows for larger gene sets) and the
// {mask, -} entries are written. k_gm_offsets: one CTA per tile turns the eight groups' value counts of every gene into
// positions inside the tile's value array. k_gm_values: one CTA per group copies the values to their positions. ------
__global__ void __launch_bounds__(256) k_gm_masks(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                                  const int32_t* __restrict__ slot_cell, int32_t G, int32_t Gs,
                                                  uint2* __restrict__ mo) {
    extern __shared__ unsigned s_mask[];   // [Gs]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int64_t grp = blockIdx.x;
    const int32_t* cells = slot_cell + grp * 32;   // the group's cells (-1 = empty slot)
    uint2* row = mo + gm_row_offset(grp, G);       // entry of gene g: row[g * GM_TILE]
    for (int32_t r0 = 0; r0 < G; r0 += Gs) {
        const int32_t r1 = min(r0 + Gs, G), ng = r1 - r0;
        for (int g = tid; g < ng; g += blockDim.x) s_mask[g] = 0u;
        __syncthreads();
        for (int ci = warp; ci < 32; ci += nwarps) {
            const int32_t cl = cells[ci];
            if (cl < 0) continue;
            const int64_t b = row_ptr[cl], e = row_ptr[cl + 1];
            for (int64_t k = b + lane; k < e; k += 32) {
                const int32_t g = col[k];
                if (g >= r0 && g < r1) atomicOr(&s_mask[g - r0], 1u << ci);
            }
        }
        __syncthreads();
        for (int g = tid; g < ng; g += blockDim.x) row[(size_t)(r0 + g) * GM_TILE] = make_uint2(s_mask[g], 0u);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_gm_offsets(int64_t n_groups, int32_t G, uint2* __restrict__ mo) {
    __shared__ unsigned s_scan[256];
    __shared__ unsigned s_carry;
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    uint2* base = mo + (size_t)tile * (size_t)G * GM_TILE;
    const int live = (int)min((long long)GM_TILE, (long long)(n_groups - tile * GM_TILE));   // groups of the last tile that exist
    if (tid == 0) s_carry = 0u;
    __syncthreads();
    for (int32_t g0 = 0; g0 < G; g0 += 256) {
        const int32_t g = g0 + tid;
        unsigned cnt[GM_TILE], tot = 0;
#pragma unroll
        for (int w = 0; w < GM_TILE; w++) {
            unsigned m = 0u;
            if (g < G && w < live) m = base[(size_t)g * GM_TILE + w].x;
            cnt[w] = __popc(m);
            tot += cnt[w];
        }
        s_scan[tid] = tot;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const unsigned v = tid >= o ? s_scan[tid - o] : 0u;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        unsigned run = s_carry + s_scan[tid] - tot;   // values of the tile in lower genes
        if (g < G) {
#pragma unroll
            for (int w = 0; w < GM_TILE; w++) {
                if (w < live) base[(size_t)g * GM_TILE + w].y = run;
                run += cnt[w];
            }
        }
        __syncthreads();
        if (tid == 255) s_carry += s_scan[255];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_gm_values(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                                   const float* __restrict__ val, const int32_t* __restrict__ slot_cell,
                                                   const int64_t* __restrict__ grp_base, int32_t G,
                                                   const uint2* __restrict__ mo, float* __restrict__ gval) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int64_t grp = blockIdx.x;
    const int32_t* cells = slot_cell + grp * 32;
    const uint2* row = mo + gm_row_offset(grp, G);
    const int64_t base = grp_base[grp];   // first value of the group's TILE
    for (int ci = warp; ci < 32; ci += nwarps) {
        const int32_t cl = cells[ci];
        if (cl < 0) continue;
        const int64_t b = row_ptr[cl], e = row_ptr[cl + 1];
        const unsigned below = (1u << ci) - 1u;
        for (int64_t k = b + lane; k < e; k += 32) {
            const uint2 x = row[(size_t)col[k] * GM_TILE];
            gval[base + x.y + __popc(x.x & below)] = val[k];
        }
    }
}

// ---- list in ascending gene order (from the 2-bit membership bitmap and the by-gene abundance table) -----------
// One CTA; same tile scan as k_list_rank of score.cu, but it emits (gene, abundance) of every list gene.
__global__ void k_gm_sorted_list(const uint32_t* __restrict__ bitmap, const float* __restrict__ ybg, int G,
                                 int32_t* __restrict__ gs_gene, float* __restrict__ gs_ab, int* __restrict__ n_out) {
    __shared__ int s_scan[1024];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int bm_words = (G + 1 + 15) / 16;   // 16 genes per word, bit 0 of each 2-bit entry = in the list
    for (int w0 = 0; w0 < bm_words; w0 += blockDim.x) {
        const int w = w0 + threadIdx.x;
        uint32_t bits = 0u;
        if (w < bm_words) {
            const uint32_t v = bitmap[w];
#pragma unroll
            for (int j = 0; j < 16; j++) bits |= ((v >> (2 * j)) & 1u) << j;
        }
        const int cnt = __popc(bits);
        s_scan[threadIdx.x] = cnt;
        __syncthreads();
        for (int o = 1; o < (int)blockDim.x; o <<= 1) {
            const int v = threadIdx.x >= (unsigned)o ? s_scan[threadIdx.x - o] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        int r = s_carry + s_scan[threadIdx.x] - cnt;
        while (bits) {
            const int j = __ffs(bits) - 1;
            bits &= bits - 1;
            const int g = 16 * w + j;
            if (g < G) { gs_gene[r] = g; gs_ab[r] = ybg[g]; r++; }
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry += s_scan[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = s_carry;
}

struct GmArgs {
    const uint2* mo;
    const float* gval;
    const int32_t* slot_cell;   // [n_groups * 32] cell of every (group, lane) slot, -1 = empty
    const int64_t* grp_base;    // [n_groups] first value of the group's tile in gval
    int64_t N, n_groups;
    int32_t G;
    const int32_t* gs_gene;
    const float* gs_ab;
    int32_t Ls;
    int32_t min_genes, jitter_mode;
    uint64_t seed;
    double* sums;
    int32_t* nused;
    int32_t* nnzc;
    uint8_t* kept;
    int64_t tile_begin, tile_stride;   // CTA b scores tile tile_begin + b * tile_stride (a tile = the 8 groups of one CTA):
                                       // all tiles on one GPU; rank, rank + nranks, ... when stage 1 is sharded
    int2* fix_list;   // slots whose pairs have zero variance (jitter mode PHILOX): {slot, bit 0 = list vector constant, bit 1 = cell vector constant}
    int* fix_count;
};

// Running sums of one cell (one lane): deviations from the first pair.
struct CellSums {
    double x0 = 0.0, y0 = 0.0, sdx = 0.0, sdy = 0.0, sxx = 0.0, syy = 0.0, sxy = 0.0;
    int n = 0;
    // FIRST_POSSIBLE = false: the caller knows that this cell already has its first pair (no pivot select)
    // COUNT = false: the caller has already added this pair to n
    template <bool FIRST_POSSIBLE = true, bool COUNT = true>
    __device__ __forceinline__ void add(double x, double y) {
        if (FIRST_POSSIBLE && n == 0) { x0 = x; y0 = y; }
        const double dx = __dsub_rn(x, x0), dy = __dsub_rn(y, y0);
        sdx = __dadd_rn(sdx, dx);
        sdy = __dadd_rn(sdy, dy);
        sxx = __dadd_rn(sxx, __dmul_rn(dx, dx));
        syy = __dadd_rn(syy, __dmul_rn(dy, dy));
        sxy = __dadd_rn(sxy, __dmul_rn(dx, dy));
        if (COUNT) n++;
    }
    // "All elements equal" (Maths.var == 0, model/SingleCell.java:430-441) without a flag per pair: a sum of squares
    // of deviations from the first element is zero iff every deviation is (x and y are floats widened to double,
    // so a non-zero deviation is at least 2^-149 and its square cannot underflow).
    __device__ __forceinline__ bool varx() const { return sxx != 0.0; }
    __device__ __forceinline__ bool vary() const { return syy != 0.0; }
    // centred sums: Sxx = sum(dx^2) - sum(dx)^2 / n, ...
    __device__ __forceinline__ void centred(double& SXX, double& SYY, double& SXY) const {
        const double nn = (double)n;
        SXX = __dsub_rn(sxx, __ddiv_rn(__dmul_rn(sdx, sdx), nn));
        SYY = __dsub_rn(syy, __ddiv_rn(__dmul_rn(sdy, sdy), nn));
        SXY = __dsub_rn(sxy, __ddiv_rn(__dmul_rn(sdx, sdy), nn));
    }
};

__device__ __forceinline__ double gm_jitter(uint64_t seed, uint64_t cell, uint32_t vec, uint32_t i) {
    uint32_t o[4];
    philox4x32_10(i, vec | 0x4A000000u, (uint32_t)cell, (uint32_t)(cell >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const uint64_t u = ((uint64_t)(o[0] >> 6) << 27) | (uint64_t)(o[1] >> 5);
    return __ddiv_rn(__dmul_rn((double)u, 1.0 / 9007199254740992.0), 1000000.0);
}

// 32 x 32 bit-matrix transpose across a warp: lane j gives row j (bit l = column l), lane l gets column l
// (bit j = row j's bit l). Five butterfly stages.
__device__ __forceinline__ unsigned warp_bit_transpose(unsigned x, int lane) {
    unsigned m = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const unsigned y = __shfl_xor_sync(FULL, x, j);
        if ((lane & j) == 0) { const unsigned t = ((x >> j) ^ y) & m; x ^= (t << j); }
        else { const unsigned t = ((y >> j) ^ x) & m; x ^= t; }
        m ^= (m << (j >> 1));
    }
    return x;
}

// One warp per group of 32 cells, lane = cell. The list is walked in chunks of 32 genes: lane j fetches gene j's
// {mask, offset} entry, the 32 masks are bit-transposed so that every lane holds the set of the chunk's genes ITS
// cell expresses, and each lane then pops its own genes in ascending order (the entry of the popped gene comes
// from the lane that fetched it, by a variable-source shuffle). A chunk costs as many steps as its busiest cell
// has genes (not 32), and the lanes of a step all do useful work as long as they have genes left. U steps are
// fetched together so that several value loads of a lane are in flight.
template <bool ALL_POSITIVE, int U>
__global__ void __launch_bounds__(256, 4) k_score_gm(const GmArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t grp = (a.tile_begin + (int64_t)blockIdx.x * a.tile_stride) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (grp >= a.n_groups) return;
    const int64_t cell = a.slot_cell[grp * 32 + lane];
    const float* gv = a.gval + a.grp_base[grp];
    asm volatile("" : "+l"(gv));   // one pointer for the group's values: a value address is then base + 32-bit index
    // (Fully gene-major entries, [gene][group], were measured 10-35 % slower: every load of a chunk then lands on a
    // different page. The tiled layout keeps a CTA inside one contiguous region.)
    const uint2* mo_row = a.mo + gm_row_offset(grp, a.G);
    const unsigned below = (1u << lane) - 1u;
    CellSums cs;
    int nz = 0;

    auto fetch = [&](int j0, uint2& mo, float& ab) {
        const int j = j0 + lane;
        mo = make_uint2(0u, 0u);
        ab = 0.0f;
        if (j < a.Ls) {
            ab = __ldg(a.gs_ab + j);
            mo = __ldg(mo_row + (size_t)__ldg(a.gs_gene + j) * GM_TILE);
        }
    };
    uint2 mo_n; float ab_n;
    fetch(0, mo_n, ab_n);
    // One chunk of 32 list genes. `first_tag`: some cell of the group may still be waiting for its first pair (the
    // pivot of its shifted sums); once every cell has one, the chunks run without the pivot select.
    auto chunk = [&](auto first_tag, const uint2 mo, const float ab) {
        constexpr bool FIRST = decltype(first_tag)::value;
        unsigned mine = warp_bit_transpose(mo.x, lane);   // bit k: my cell expresses the chunk's gene k
        if (ALL_POSITIVE) {
            // every stored value is > 0: an expressed list gene always counts, and it pairs iff its abundance is
            // > 0 — both are known for the whole chunk from the bit sets, nothing to test per gene
            nz += __popc(mine);
            mine &= __ballot_sync(FULL, ab > 0.0f);
            if (!FIRST) cs.n += __popc(mine);
        }
        int iters = __reduce_max_sync(FULL, (unsigned)__popc(mine));
        for (; iters > 0; iters -= U) {   // warp-uniform trip count
            unsigned m[U]; float abk[U], e[U]; bool has[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                has[u] = mine != 0u;
                const int k = has[u] ? (__ffs(mine) - 1) : 0;
                mine &= mine - 1u;   // 0 stays 0
                m[u] = __shfl_sync(FULL, mo.x, k);
                const unsigned off = __shfl_sync(FULL, mo.y, k);
                abk[u] = __shfl_sync(FULL, ab, k);
                e[u] = 0.0f;
                if (has[u]) e[u] = __ldg(gv + (uint32_t)(off + __popc(m[u] & below)));   // a group holds < 2^32 values
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (ALL_POSITIVE) {
                    if (has[u]) cs.template add<FIRST, FIRST>((double)abk[u], (double)e[u]);   // model/SingleCell.java:403
                } else if (has[u]) {
                    if (e[u] != 0.0f) nz++;                                      // NaN counts (:409)
                    if (abk[u] > 0.0f && e[u] > 0.0f) cs.template add<FIRST>((double)abk[u], (double)e[u]);
                }
            }
        }
    };
    bool all_started = false;   // warp-uniform
    for (int j0 = 0; j0 < a.Ls; j0 += 32) {
        const uint2 mo = mo_n;
        const float ab = ab_n;
        fetch(j0 + 32, mo_n, ab_n);   // the next chunk's entries travel while this chunk is processed
        if (all_started) chunk(std::false_type{}, mo, ab);
        else {
            chunk(std::true_type{}, mo, ab);
            all_started = __all_sync(FULL, cs.n > 0 || cell < 0);
        }
    }
    if (cell < 0) return;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const bool keep = (nz > 0) && (cs.n >= a.min_genes);
    double o_xx = qnan, o_yy = qnan, o_xy = qnan;
    if (keep && cs.n >= 2) {
        const bool degenerate = !cs.varx() || !cs.vary();   // Maths.var == 0 <=> all elements equal (:430-441)
        if (!degenerate) cs.centred(o_xx, o_yy, o_xy);
        else if (a.jitter_mode == PCTSEA_JITTER_PHILOX) {
            const int slot = atomicAdd(a.fix_count, 1);
            a.fix_list[slot] = make_int2((int)(grp * 32 + lane), (cs.varx() ? 0 : 1) | (cs.vary() ? 0 : 2));
        }
    }
    double* o = a.sums + 3 * (size_t)cell;
    o[0] = o_xx; o[1] = o_yy; o[2] = o_xy;
    a.nused[cell] = cs.n;
    a.nnzc[cell] = nz;
    a.kept[cell] = keep ? 1 : 0;
}

// Zero-variance cells (rare): one warp per listed cell re-walks the list's genes, the lane holding a pair adds
// the Philox jitter of its pair index to the constant vector(s), and the pairs are folded in gene order.
__global__ void __launch_bounds__(256) k_score_gm_jitter(const GmArgs a) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n_fix = *a.fix_count;
    const unsigned below = (1u << lane) - 1u;
    for (int i = w; i < n_fix; i += gridDim.x * (blockDim.x >> 5)) {
        const int2 f = a.fix_list[i];
        const int64_t grp = f.x >> 5, cell = a.slot_cell[f.x];
        const int cl = f.x & 31;
        const bool eq_list = (f.y & 1) != 0, eq_cell = (f.y & 2) != 0;
        const float* gv = a.gval + a.grp_base[grp];
        const uint2* mo_row = a.mo + gm_row_offset(grp, a.G);
        CellSums cs;
        int kbase = 0;
        for (int j0 = 0; j0 < a.Ls; j0 += 32) {
            const int j = j0 + lane;
            bool has = false;
            double x = 0.0, y = 0.0;
            if (j < a.Ls) {
                const float ab = a.gs_ab[j];
                const uint2 mo = mo_row[(size_t)a.gs_gene[j] * GM_TILE];
                if ((mo.x >> cl) & 1u) {
                    const float e = gv[mo.y + __popc(mo.x & ((1u << cl) - 1u))];
                    if (ab > 0.0f && e > 0.0f) { has = true; x = (double)ab; y = (double)e; }
                }
            }
            unsigned bal = __ballot_sync(FULL, has);
            if (has) {
                const uint32_t k = (uint32_t)(kbase + __popc(bal & below));
                if (eq_cell) y = __dadd_rn(y, gm_jitter(a.seed, (uint64_t)cell, 0u, k));
                if (eq_list) x = __dadd_rn(x, gm_jitter(a.seed, (uint64_t)cell, 1u, k));
            }
            kbase += __popc(bal);
            while (bal) {
                const int l = __ffs(bal) - 1;
                bal &= bal - 1u;
                cs.add(__shfl_sync(FULL, x, l), __shfl_sync(FULL, y, l));
            }
        }
        if (lane == 0) {
            double* o = a.sums + 3 * (size_t)cell;
            cs.centred(o[0], o[1], o[2]);
        }
    }
}

// ---- stage 1 sharded over the ranks of a communicator (multi-GPU pctsea_analyze) --------------------------------
// Rank r scores the tiles r, r + R, ... (interleaved: the tiles are ordered by cell type and row length, so any
// contiguous range would be a biased sample of the work). Its results travel in tile order: pack[(r * tpr + j) * 256 + s]
// is slot s of tile j * R + r; one in-place all-gather; every rank scatters all records back to cell order.
struct __align__(16) ScoreRec { double score; int32_t nused; int32_t kept; };

__global__ void k_score_pack(const double* __restrict__ score, const int32_t* __restrict__ nused, const uint8_t* __restrict__ kept,
                             const int32_t* __restrict__ slot_cell, int64_t n_slots, int64_t tiles_per_rank, int32_t R, int32_t r,
                             ScoreRec* __restrict__ pack) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (j, s)
    if (id >= tiles_per_rank * 256) return;
    const int64_t j = id >> 8, s = id & 255, slot = (j * R + r) * 256 + s;
    ScoreRec rec;
    rec.score = __longlong_as_double(0x7ff8000000000000LL); rec.nused = 0; rec.kept = -1;   // kept -1: no cell here
    if (slot < n_slots) {
        const int32_t cell = slot_cell[slot];
        if (cell >= 0) { rec.score = score[cell]; rec.nused = nused[cell]; rec.kept = kept[cell]; }
    }
    pack[(int64_t)r * tiles_per_rank * 256 + id] = rec;
}

__global__ void k_score_unpack(const ScoreRec* __restrict__ pack, const int32_t* __restrict__ slot_cell, int64_t n_slots,
                               int64_t tiles_per_rank, int32_t R, double* __restrict__ score, int32_t* __restrict__ nused,
                               uint8_t* __restrict__ kept) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (rr, j, s)
    if (id >= (int64_t)R * tiles_per_rank * 256) return;
    const int64_t per = tiles_per_rank * 256, rr = id / per, rem = id - rr * per, j = rem >> 8, s = rem & 255;
    const int64_t slot = (j * R + rr) * 256 + s;
    if (slot >= n_slots) return;
    const int32_t cell = slot_cell[slot];
    if (cell < 0) return;
    const ScoreRec rec = pack[id];
    score[cell] = rec.score; nused[cell] = rec.nused; kept[cell] = (uint8_t)rec.kept;
}

}  // namespace

// Which cells share a group is free (cells are independent), and it matters: a chunk of the list costs a group as
// many steps as its busiest cell has genes in it. Cells are therefore grouped by cell type (cells of a type have
// similar expression profiles; unknown types first) and, inside a type, by their number of non-zeros. The order is a
// layout decision only: every result is written to the cell's own index.
void build_gene_index(pctsea_ctx* c, pctsea_matrix* m) {
    if (m->gm_mo) return;
    const int64_t N = m->n_cells;
    const int32_t G = m->n_genes;
    PCT_REQUIRE(N >= 1 && G >= 1, "empty matrix");
    const int32_t Gs = std::min<int32_t>(G, 50 * 1024);   // genes per pass of the build kernel (one mask word each)
    const size_t smem = (size_t)Gs * 4;
    const int64_t n_groups = div_up(N, 32);
    PCT_REQUIRE(N < ((int64_t)1 << 31), "too many cells");
    std::vector<int64_t> rp((size_t)N + 1);
    std::vector<int32_t> lab;
    PCT_CUDA(cudaMemcpyAsync(rp.data(), m->row_ptr, ((size_t)N + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    if (m->label) {
        lab.resize((size_t)N);
        PCT_CUDA(cudaMemcpyAsync(lab.data(), m->label, (size_t)N * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    PCT_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<int32_t> slot((size_t)n_groups * 32, -1);
    for (int64_t i = 0; i < N; i++) slot[(size_t)i] = (int32_t)i;
    std::stable_sort(slot.begin(), slot.begin() + N, [&](int32_t x, int32_t y) {
        if (!lab.empty() && lab[x] != lab[y]) return lab[x] < lab[y];
        return rp[(size_t)x + 1] - rp[x] < rp[(size_t)y + 1] - rp[y];
    });
    // (Ordering the groups by decreasing size, so that the short ones fill the tail of the last wave, was measured
    // 12 % slower: the long groups then all run at the same time.)
    // base[g] = first value of group g's tile (GM_TILE groups, the warps of one scoring CTA, share a value array)
    std::vector<int64_t> base((size_t)n_groups + 1, 0);
    {
        int64_t run = 0, tile_start = 0;
        for (int64_t g = 0; g < n_groups; g++) {
            if (g % GM_TILE == 0) tile_start = run;
            int64_t cnt = 0;
            for (int l = 0; l < 32; l++) {
                const int32_t cl = slot[(size_t)g * 32 + l];
                if (cl >= 0) cnt += rp[(size_t)cl + 1] - rp[cl];
            }
            PCT_REQUIRE(run + cnt - tile_start < ((int64_t)1 << 32), "a tile of 256 cells holds 2^32 non-zeros or more");
            base[(size_t)g] = tile_start;
            run += cnt;
        }
        base[(size_t)n_groups] = run;
    }
    uint2* mo = nullptr;
    float* gval = nullptr;
    int32_t* slot_dev = nullptr;
    int64_t* base_dev = nullptr;
    const size_t mo_bytes = (size_t)round_up(n_groups, GM_TILE) * (size_t)G * sizeof(uint2);
    auto release = [&]() { if (mo) cudaFree(mo); if (gval) cudaFree(gval); if (slot_dev) cudaFree(slot_dev); if (base_dev) cudaFree(base_dev); };
    cudaError_t e = cudaMalloc((void**)&mo, mo_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&gval, std::max<size_t>((size_t)m->nnz * 4, 16));
    if (e == cudaSuccess) e = cudaMalloc((void**)&slot_dev, slot.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&base_dev, base.size() * 8);
    if (e != cudaSuccess) {
        cudaGetLastError();
        release();
        throw Error(PCTSEA_ENOMEM, "gene index: cudaMalloc of " + std::to_string(mo_bytes + (size_t)m->nnz * 4) + " bytes failed");
    }
    cudaError_t le = cudaMemcpyAsync(slot_dev, slot.data(), slot.size() * 4, cudaMemcpyHostToDevice, c->stream);
    if (le == cudaSuccess) le = cudaMemcpyAsync(base_dev, base.data(), base.size() * 8, cudaMemcpyHostToDevice, c->stream);
    if (le == cudaSuccess) le = cudaFuncSetAttribute(k_gm_masks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (le == cudaSuccess) {
        k_gm_masks<<<(unsigned)n_groups, 256, smem, c->stream>>>(m->row_ptr, m->col, slot_dev, G, Gs, mo);
        k_gm_offsets<<<(unsigned)div_up(n_groups, GM_TILE), 256, 0, c->stream>>>(n_groups, G, mo);
        k_gm_values<<<(unsigned)n_groups, 256, 0, c->stream>>>(m->row_ptr, m->col, m->val, slot_dev, base_dev, G, mo, gval);
        count_launch(c, 3);
        le = cudaGetLastError();
    }
    if (le == cudaSuccess) le = cudaStreamSynchronize(c->stream);   // the host vectors above are the copy sources
    if (le != cudaSuccess) {
        release();
        throw Error(PCTSEA_ECUDA, std::string("gene index build failed: ") + cudaGetErrorString(le));
    }
    m->gm_mo = mo;
    m->gm_val = gval;
    m->gm_slot_cell = slot_dev;
    m->gm_base = base_dev;
    m->n_groups = n_groups;
}

// Stage 1 over the gene-major copy. List tables (ybg, bitmap) as in launch_score; then the sorted list, the
// scoring kernel, the jitter fix-up and the shared k_score_final.
void launch_score_gm(pctsea_ctx* c, const ScoreArgs& sa) {
    const pctsea_matrix* m = sa.m;
    const int64_t N = m->n_cells;
    const int32_t G = m->n_genes;
    const int32_t bm_words = (G + 1 + 15) / 16;
    c->ybg.reserve((size_t)G * 4 + 16);
    c->bitmap.reserve((size_t)bm_words * 4 + 16);
    c->list_rank.reserve((size_t)std::max(sa.L, 1) * 8 + 64);
    c->score.reserve((size_t)N * 8 + 16);
    c->score_sums.reserve((size_t)N * 24 + 16);
    c->nused.reserve((size_t)N * 4 + 16);
    c->nnzc.reserve((size_t)N * 4 + 16);
    c->kept.reserve((size_t)N + 16);
    c->gm_fix.reserve((size_t)N * 8 + 128);   // [0] sorted list length, [1] cells to fix up; the cell list from byte 64
    PCT_CUDA(cudaMemsetAsync(c->ybg.p, 0, (size_t)G * 4, c->stream));
    PCT_CUDA(cudaMemsetAsync(c->bitmap.p, 0, (size_t)bm_words * 4, c->stream));
    PCT_CUDA(cudaMemsetAsync(c->gm_fix.p, 0, 64, c->stream));
    int32_t* gs_gene = c->list_rank.as<int32_t>();
    float* gs_ab = c->list_rank.as<float>() + std::max(sa.L, 1);
    int* counters = c->gm_fix.as<int>();
    if (sa.L > 0) {
        launch_list_lookup(c, sa);
        k_gm_sorted_list<<<1, 1024, 0, c->stream>>>(c->bitmap.as<uint32_t>(), c->ybg.as<float>(), G, gs_gene, gs_ab, counters);
        count_launch(c);
    }
    GmArgs a;
    a.mo = m->gm_mo; a.gval = m->gm_val; a.slot_cell = m->gm_slot_cell; a.grp_base = m->gm_base; a.N = N;
    a.n_groups = m->n_groups; a.G = G;
    a.gs_gene = gs_gene; a.gs_ab = gs_ab; a.Ls = c->list_mapped;
    a.min_genes = sa.prm.min_genes_cells; a.jitter_mode = sa.prm.jitter_mode; a.seed = sa.prm.seed;
    a.sums = c->score_sums.as<double>(); a.nused = c->nused.as<int32_t>(); a.nnzc = c->nnzc.as<int32_t>();
    a.kept = c->kept.as<uint8_t>();
    a.fix_list = c->gm_fix.as<int2>() + 8; a.fix_count = counters + 1;
    // multi-GPU analysis: this rank scores every R-th tile, the results are all-gathered below
    const int32_t R = c->shard_score ? c->comm_nranks : 1, r = c->shard_score ? c->comm_rank : 0;
    const int64_t n_tiles = div_up(m->n_groups, 8);
    a.tile_begin = r; a.tile_stride = R;
    const unsigned grid = (unsigned)std::max<int64_t>((n_tiles - r + R - 1) / R, 0);
    // (a rank can be left without tiles when the matrix has fewer tiles than the communicator has ranks)
    int unroll = 8;
    if (c->opt_gm_unroll > 0) unroll = c->opt_gm_unroll;   // PCTSEA_OPT_GM_UNROLL (tuning aid)
    if (grid == 0) {
    } else if (m->all_positive) {
        if (unroll >= 8) k_score_gm<true, 8><<<grid, 256, 0, c->stream>>>(a);
        else if (unroll >= 4) k_score_gm<true, 4><<<grid, 256, 0, c->stream>>>(a);
        else if (unroll <= 1) k_score_gm<true, 1><<<grid, 256, 0, c->stream>>>(a);
        else k_score_gm<true, 2><<<grid, 256, 0, c->stream>>>(a);
    } else k_score_gm<false, 8><<<grid, 256, 0, c->stream>>>(a);
    count_launch(c);
    PCT_CUDA(cudaGetLastError());
    if (sa.prm.jitter_mode == PCTSEA_JITTER_PHILOX) {
        k_score_gm_jitter<<<c->num_sms * 2, 256, 0, c->stream>>>(a);
        count_launch(c);
        PCT_CUDA(cudaGetLastError());
    }
    launch_score_final(c, sa);   // over all cells: the scores of the tiles of other ranks are overwritten below
    if (R > 1) {
        const int64_t tpr = div_up(n_tiles, R), n_slots = m->n_groups * 32;
        const size_t rank_bytes = (size_t)tpr * 256 * sizeof(ScoreRec);
        c->score_pack.reserve(rank_bytes * R + 64);
        ScoreRec* pack = c->score_pack.as<ScoreRec>();
        k_score_pack<<<(unsigned)div_up(tpr * 256, 256), 256, 0, c->stream>>>(c->score.as<double>(), c->nused.as<int32_t>(),
                                                                             c->kept.as<uint8_t>(), m->gm_slot_cell, n_slots, tpr, R, r, pack);
        comm_all_gather_in_place(c, pack, rank_bytes);
        k_score_unpack<<<(unsigned)div_up((int64_t)R * tpr * 256, 256), 256, 0, c->stream>>>(pack, m->gm_slot_cell, n_slots, tpr, R,
                                                                                            c->score.as<double>(), c->nused.as<int32_t>(),
                                                                                            c->kept.as<uint8_t>());
        count_launch(c, 2);
        PCT_CUDA(cudaGetLastError());
    }
}

}  // namespace pctsea
