// rank.cu — stage 2: score threshold + ranking, all on device.
//
// Replaces (reference paths under pctsea-core/src/main/java/edu/scripps/yates/pctsea/):
//   a4  scoring/ScoreThreshold.java:98-106          passThreshold (NaN never passes)
//   a5  scoring/ScoreThreshold.java:57-75,120-144, utils/PCTSEAUtils.java:44-53
//       stable sorts by Double.compare => total order (score desc, cell index asc)
//   a7  PCTSEA.java:421-424                          subList(0, size-1): the last cell is dropped
//
// Keys are the 64-bit Double.compare-orderable image of the score, inverted for descending order; the sort
// is a hand-written stable LSD radix sort (8-bit digits) over (key, cell index) pairs that starts from
// index order, so ties keep DB order exactly like the reference's stable merge sorts. Digit positions on
// which every key agrees are skipped (decided from one up-front histogram of all 8 digits).
//
// Roofline: HBM, 2 * 12 B * Np per executed pass; Np <= 2M so the stage is launch-latency sized.
#include "common.cuh"
#include <cstring>

namespace pctsea {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 2048 keys per tile

__device__ __forceinline__ uint64_t orderable_desc(double s) {
    // Double.compare order: -0.0 < 0.0, ascending in the two's-complement image below; inverted => descending.
    uint64_t b = (uint64_t)__double_as_longlong(s);
    uint64_t u = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
    return ~u;
}

__device__ __forceinline__ bool passes(double s, double thr) { return thr >= 0.0 ? (s >= thr) : (s < thr); }

// ---- compaction of the passing cells (index order) ----------------------------------------------
__global__ void __launch_bounds__(RS_THREADS) k_pass_count(const double* __restrict__ score,
                                                          const uint8_t* __restrict__ kept, int64_t N, double thr,
                                                          uint32_t* __restrict__ tile_cnt) {
    __shared__ uint32_t wsum[RS_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * RS_TILE;
    uint32_t cnt = 0;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        int64_t i = base + r * RS_THREADS + threadIdx.x;
        bool p = false;
        if (i < N) p = (kept == nullptr || kept[i]) && passes(score[i], thr);
        cnt += __popc(__ballot_sync(0xffffffffu, p));
    }
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < RS_THREADS / 32; w++) t += wsum[w];
        tile_cnt[blockIdx.x] = t;
    }
}

// Exclusive scan of n uint32 values by one block; total written to *total (u64).
__global__ void __launch_bounds__(1024) k_scan_u32(uint32_t* __restrict__ data, int64_t n,
                                                  unsigned long long* __restrict__ total) {
    __shared__ uint64_t part[1024];
    int64_t per = (n + 1023) / 1024;
    int64_t lo = (int64_t)threadIdx.x * per, hi = min(lo + per, n);
    uint64_t s = 0;
    for (int64_t i = lo; i < hi; i++) s += data[i];
    part[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (int off = 1; off < 1024; off <<= 1) {
        uint64_t v = (threadIdx.x >= off) ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint64_t run = (threadIdx.x == 0) ? 0 : part[threadIdx.x - 1];
    for (int64_t i = lo; i < hi; i++) { uint32_t v = data[i]; data[i] = (uint32_t)run; run += v; }
    if (threadIdx.x == 1023 && total) *total = part[1023];
}

__global__ void __launch_bounds__(RS_THREADS) k_pass_scatter(const double* __restrict__ score,
                                                            const uint8_t* __restrict__ kept, int64_t N, double thr,
                                                            const uint32_t* __restrict__ tile_off,
                                                            uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                            uint32_t* __restrict__ hist /*[8][256]*/) {
    __shared__ uint32_t wcnt[RS_THREADS / 32];
    __shared__ uint32_t shist[8 * 256];
    __shared__ uint32_t running;
    for (int i = threadIdx.x; i < 8 * 256; i += RS_THREADS) shist[i] = 0;
    if (threadIdx.x == 0) running = tile_off[blockIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t base = (int64_t)blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ITEMS; r++) {
        int64_t i = base + r * RS_THREADS + threadIdx.x;
        bool p = false;
        double s = 0.0;
        if (i < N) { s = score[i]; p = (kept == nullptr || kept[i]) && passes(s, thr); }
        unsigned m = __ballot_sync(0xffffffffu, p);
        if (lane == 0) wcnt[warp] = __popc(m);
        __syncthreads();
        uint32_t off = running;
        for (int w = 0; w < warp; w++) off += wcnt[w];
        if (p) {
            uint32_t dst = off + __popc(m & ((1u << lane) - 1u));
            uint64_t k = orderable_desc(s);
            keys[dst] = k;
            vals[dst] = (uint32_t)i;
#pragma unroll
            for (int d = 0; d < 8; d++) atomicAdd(&shist[d * 256 + (int)((k >> (8 * d)) & 0xff)], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < RS_THREADS / 32; w++) t += wcnt[w];
            running += t;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 8 * 256; i += RS_THREADS)
        if (shist[i]) atomicAdd(&hist[i], shist[i]);
}

// ---- radix sort passes ----------------------------------------------------------------------------
__global__ void __launch_bounds__(RS_THREADS) k_digit_hist(const uint64_t* __restrict__ keys, int64_t n,
                                                          uint32_t* __restrict__ hist, int ndig) {
    __shared__ uint32_t shist[8 * 256];
    for (int i = threadIdx.x; i < ndig * 256; i += RS_THREADS) shist[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RS_THREADS) {
        uint64_t k = keys[i];
        for (int d = 0; d < ndig; d++) atomicAdd(&shist[d * 256 + (int)((k >> (8 * d)) & 0xff)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ndig * 256; i += RS_THREADS)
        if (shist[i]) atomicAdd(&hist[i], shist[i]);
}

__global__ void __launch_bounds__(RS_THREADS) k_tile_count(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                          uint32_t* __restrict__ tile_hist, int n_tiles) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        int64_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&sh[(int)((keys[i] >> shift) & 0xff)], 1u);
    }
    __syncthreads();
    tile_hist[(size_t)blockIdx.x * 256 + threadIdx.x] = sh[threadIdx.x];   // [tile][digit]
}

// Long rankings (more than RS_FUSED_TILES tiles): the per-tile counts are turned into exclusive prefixes over the tiles
// by a pass of their own — one warp per digit column, 32 tiles per step — instead of every scatter CTA summing every
// column (quadratic in the number of tiles: 2.8 ms of a 2 M-cell ranking's 7 passes). tile_hist[tile][d] becomes the
// number of keys with digit d in earlier tiles, digit_total[d] the column sum.
constexpr int RS_FUSED_TILES = 256;

__global__ void __launch_bounds__(256) k_tile_prefix(uint32_t* __restrict__ tile_hist, int n_tiles,
                                                     uint32_t* __restrict__ digit_total) {
    const int lane = threadIdx.x & 31, d = blockIdx.x * 8 + (threadIdx.x >> 5);
    uint32_t run = 0;
    for (int j0 = 0; j0 < n_tiles; j0 += 32) {
        const int j = j0 + lane;
        const uint32_t v = j < n_tiles ? tile_hist[(size_t)j * 256 + d] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (j < n_tiles) tile_hist[(size_t)j * 256 + d] = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) digit_total[d] = run;
}

__global__ void __launch_bounds__(RS_THREADS) k_tile_scatter(const uint64_t* __restrict__ keys_in,
                                                            const uint32_t* __restrict__ vals_in, int64_t n, int shift,
                                                            const uint32_t* __restrict__ tile_hist, int n_tiles,
                                                            const uint32_t* __restrict__ digit_total,
                                                            uint64_t* __restrict__ keys_out,
                                                            uint32_t* __restrict__ vals_out) {
    // Element order inside the tile: (warp, round, lane) <-> index warp*256 + round*32 + lane.
    __shared__ uint32_t cnt[RS_THREADS / 32][256];
    __shared__ uint32_t toff[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < (RS_THREADS / 32) * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    // Output offset of (digit d, this tile) = keys with a smaller digit anywhere + keys with digit d in earlier
    // tiles, straight from the per-tile counts ([tile][digit], L2-resident): thread d sums its column, then the
    // block scans the 256 digit totals. This replaces a separate single-CTA scan launch per pass.
    {
        __shared__ uint32_t wtot[RS_THREADS / 32];
        const uint32_t* col = tile_hist + threadIdx.x;
        uint32_t before = 0, total = 0;
        int j = 0;
        if (digit_total) {   // prefixes over the tiles were computed by k_tile_prefix
            before = col[(size_t)blockIdx.x * 256];
            total = digit_total[threadIdx.x];
            j = n_tiles;
        }
        for (; j + 8 <= n_tiles; j += 8) {
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = col[(size_t)(j + u) * 256];
#pragma unroll
            for (int u = 0; u < 8; u++) { total += v[u]; before += (j + u < (int)blockIdx.x) ? v[u] : 0u; }
        }
        for (; j < n_tiles; j++) { const uint32_t v = col[(size_t)j * 256]; total += v; before += (j < (int)blockIdx.x) ? v : 0u; }
        uint32_t incl = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < warp; w++) wbase += wtot[w];
        toff[threadIdx.x] = wbase + incl - total + before;
    }
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * RS_TILE + warp * (RS_ITEMS * 32);
    uint64_t key[RS_ITEMS];
    uint32_t val[RS_ITEMS];
    int dig[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        int64_t i = base + r * 32 + lane;
        if (i < n) { key[r] = keys_in[i]; val[r] = vals_in[i]; dig[r] = (int)((key[r] >> shift) & 0xff); }
        else { key[r] = 0; val[r] = 0; dig[r] = 256 + lane; }  // unique, never counted
    }
    // phase A: per-warp digit counts
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        unsigned m = __match_any_sync(0xffffffffu, dig[r]);
        if (dig[r] < 256 && (m & lt) == 0) cnt[warp][dig[r]] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
    // phase B: exclusive prefix over warps, per digit
    {
        uint32_t run = 0;
        for (int w = 0; w < RS_THREADS / 32; w++) { uint32_t v = cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = run; run += v; }
    }
    __syncthreads();
    // phase C: stable ranks and scatter
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        unsigned m = __match_any_sync(0xffffffffu, dig[r]);
        uint32_t rank = 0;
        if (dig[r] < 256) rank = cnt[warp][dig[r]] + __popc(m & lt);
        __syncwarp();
        if (dig[r] < 256 && (m & lt) == 0) cnt[warp][dig[r]] += __popc(m);
        __syncwarp();
        if (dig[r] < 256) {
            uint32_t dst = toff[dig[r]] + rank;
            keys_out[dst] = key[r];
            vals_out[dst] = val[r];
        }
    }
}

// thr < 0 only: position of the dropped cell = (#keys equal to keys[0]) - 1
__global__ void k_count_top_ties(const uint64_t* __restrict__ keys, int64_t n, unsigned long long* __restrict__ out) {
    uint64_t k0 = keys[0];
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (keys[i] == k0) c++;
    if (c) atomicAdd(out, c);
}

__global__ void k_write_order(const uint32_t* __restrict__ vals, int64_t n, const unsigned long long* __restrict__ ties,
                              int neg_thr, int32_t* __restrict__ order) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!neg_thr) { order[i] = (int32_t)vals[i]; return; }
    int64_t pos = (int64_t)(*ties) - 1;  // dropped element
    int32_t v;
    if (i < pos) v = (int32_t)vals[i];
    else if (i < n - 1) v = (int32_t)vals[i + 1];
    else v = (int32_t)vals[pos];
    order[i] = v;
}

// Sorts n (key,value) pairs by key ascending, stable. hist_dev must hold the [8][256] digit histogram of
// the keys (filled by the caller or by k_digit_hist) already copied to hist_host.
static void radix_passes(pctsea_ctx* c, int64_t n, const uint32_t* hist_host, int ndig, uint64_t** keys,
                         uint64_t** keys_alt, uint32_t** vals, uint32_t** vals_alt) {
    const int n_tiles = (int)div_up(n, RS_TILE);
    c->tile_hist.reserve((size_t)256 * (n_tiles + 1) * 4 + 16);
    uint32_t* digit_total = n_tiles > RS_FUSED_TILES ? c->tile_hist.as<uint32_t>() + (size_t)256 * n_tiles : nullptr;
    for (int d = 0; d < ndig; d++) {
        bool trivial = false;
        for (int b = 0; b < 256; b++)
            if ((int64_t)hist_host[d * 256 + b] == n) { trivial = true; break; }
        if (trivial) continue;
        const int shift = 8 * d;
        k_tile_count<<<n_tiles, RS_THREADS, 0, c->stream>>>(*keys, n, shift, c->tile_hist.as<uint32_t>(), n_tiles);
        if (digit_total) {
            k_tile_prefix<<<32, 256, 0, c->stream>>>(c->tile_hist.as<uint32_t>(), n_tiles, digit_total);
            count_launch(c);
        }
        k_tile_scatter<<<n_tiles, RS_THREADS, 0, c->stream>>>(*keys, *vals, n, shift, c->tile_hist.as<uint32_t>(),
                                                              n_tiles, digit_total, *keys_alt, *vals_alt);
        count_launch(c, 2);
        std::swap(*keys, *keys_alt);
        std::swap(*vals, *vals_alt);
    }
    PCT_CUDA(cudaGetLastError());
}

void radix_sort_pairs(pctsea_ctx* c, int64_t n, uint64_t** keys, uint64_t** keys_alt, uint32_t** vals,
                      uint32_t** vals_alt, int key_bits) {
    if (n <= 1) return;
    const int ndig = (key_bits + 7) / 8;
    c->hist.reserve(8 * 256 * 4);
    c->pin_small.reserve(8 * 256 * 4 + 64);
    PCT_CUDA(cudaMemsetAsync(c->hist.p, 0, 8 * 256 * 4, c->stream));
    int grid = (int)std::min<int64_t>(div_up(n, RS_THREADS), (int64_t)c->num_sms * 8);
    k_digit_hist<<<grid, RS_THREADS, 0, c->stream>>>(*keys, n, c->hist.as<uint32_t>(), ndig);
    count_launch(c);
    PCT_CUDA(cudaMemcpyAsync(c->pin_small.p, c->hist.p, 8 * 256 * 4, cudaMemcpyDeviceToHost, c->stream));
    PCT_CUDA(cudaStreamSynchronize(c->stream));
    radix_passes(c, n, c->pin_small.as<uint32_t>(), ndig, keys, keys_alt, vals, vals_alt);
}

// Per-type cell counts for the host-side hypergeometric statistics (PCTSEA.java:2169-2264): cells of the type
// among the cells that survive the min_genes_cells filter, and among those the ones passing the score threshold.
// counts[0..T) = kept, counts[T..2T) = kept and passing, counts[2T] = all kept cells (unknown types included).
__global__ void __launch_bounds__(256) k_type_counts(const double* __restrict__ score, const uint8_t* __restrict__ kept,
                                                     const int32_t* __restrict__ label, int64_t N, double thr, int32_t T,
                                                     int32_t* __restrict__ counts, int32_t hist_n) {
    extern __shared__ int32_t sh[];
    for (int j = threadIdx.x; j < hist_n; j += 256) sh[j] = 0;
    __syncthreads();
    int32_t* dst = hist_n ? sh : counts;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < N; i += (int64_t)gridDim.x * 256) {
        if (kept != nullptr && !kept[i]) continue;
        atomicAdd(&dst[2 * T], 1);
        const int32_t lab = label[i];
        if (lab < 0 || lab >= T) continue;
        atomicAdd(&dst[lab], 1);
        if (passes(score[i], thr)) atomicAdd(&dst[T + lab], 1);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < hist_n; j += 256)
        if (sh[j]) atomicAdd(&counts[j], sh[j]);
}

void launch_type_counts(pctsea_ctx* c, int64_t N, const double* score_dev, const uint8_t* kept_dev,
                        const int32_t* label_dev, double min_score, int32_t T) {
    const int32_t n = 2 * T + 1;
    c->type_counts.reserve((size_t)n * 4 + 64);
    PCT_CUDA(cudaMemsetAsync(c->type_counts.p, 0, (size_t)n * 4, c->stream));
    c->type_counts_T = T;
    if (N == 0) return;
    const int32_t hist_n = (n <= 8192) ? n : 0;
    int grid = (int)std::min<int64_t>(div_up(N, 256), (int64_t)c->num_sms * 4);
    k_type_counts<<<grid, 256, (size_t)hist_n * 4, c->stream>>>(score_dev, kept_dev, label_dev, N, min_score, T,
                                                                c->type_counts.as<int32_t>(), hist_n);
    count_launch(c);
    PCT_CUDA(cudaGetLastError());
}

int64_t launch_rank(pctsea_ctx* c, int64_t N, const double* score_dev, const uint8_t* kept_dev, double min_score) {
    c->order.reserve((size_t)std::max<int64_t>(N, 1) * 4 + 16);
    if (N == 0) return 0;
    const int n_tiles = (int)div_up(N, RS_TILE);
    c->keysA.reserve((size_t)N * 8 + 16);
    c->keysB.reserve((size_t)N * 8 + 16);
    c->valsA.reserve((size_t)N * 4 + 16);
    c->valsB.reserve((size_t)N * 4 + 16);
    c->scan_tmp.reserve((size_t)n_tiles * 4 + 16);
    c->hist.reserve(8 * 256 * 4);
    c->counters.reserve(64);
    c->pin_small.reserve(8 * 256 * 4 + 64);
    PCT_CUDA(cudaMemsetAsync(c->hist.p, 0, 8 * 256 * 4, c->stream));
    PCT_CUDA(cudaMemsetAsync(c->counters.p, 0, 64, c->stream));
    k_pass_count<<<n_tiles, RS_THREADS, 0, c->stream>>>(score_dev, kept_dev, N, min_score, c->scan_tmp.as<uint32_t>());
    k_scan_u32<<<1, 1024, 0, c->stream>>>(c->scan_tmp.as<uint32_t>(), n_tiles, c->counters.as<unsigned long long>());
    k_pass_scatter<<<n_tiles, RS_THREADS, 0, c->stream>>>(score_dev, kept_dev, N, min_score, c->scan_tmp.as<uint32_t>(),
                                                          c->keysA.as<uint64_t>(), c->valsA.as<uint32_t>(),
                                                          c->hist.as<uint32_t>());
    count_launch(c, 3);
    // one sync: Np and the 8 digit histograms
    uint32_t* hh = c->pin_small.as<uint32_t>();
    PCT_CUDA(cudaMemcpyAsync(hh, c->hist.p, 8 * 256 * 4, cudaMemcpyDeviceToHost, c->stream));
    PCT_CUDA(cudaMemcpyAsync(hh + 8 * 256, c->counters.p, 8, cudaMemcpyDeviceToHost, c->stream));
    PCT_CUDA(cudaStreamSynchronize(c->stream));
    unsigned long long np_ull;
    memcpy(&np_ull, hh + 8 * 256, 8);
    const int64_t Np = (int64_t)np_ull;
    if (Np == 0) return 0;
    uint64_t *k = c->keysA.as<uint64_t>(), *k2 = c->keysB.as<uint64_t>();
    uint32_t *v = c->valsA.as<uint32_t>(), *v2 = c->valsB.as<uint32_t>();
    if (Np > 1) radix_passes(c, Np, hh, 8, &k, &k2, &v, &v2);
    const int neg = (min_score < 0.0) ? 1 : 0;
    unsigned long long* ties = c->counters.as<unsigned long long>() + 1;
    if (neg) {
        int grid = (int)std::min<int64_t>(div_up(Np, 256), (int64_t)c->num_sms * 4);
        k_count_top_ties<<<grid, 256, 0, c->stream>>>(k, Np, ties);
        count_launch(c);
    }
    k_write_order<<<(unsigned)div_up(Np, 256), 256, 0, c->stream>>>(v, Np, ties, neg, c->order.as<int32_t>());
    count_launch(c);
    PCT_CUDA(cudaGetLastError());
    return Np;
}

}  // namespace pctsea
