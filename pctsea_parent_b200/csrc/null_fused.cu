// null_fused.cu — the label-permutation null of stage 3 in FAST arithmetic: ONE kernel per call that generates the
// permuted labels and evaluates the weighted-KS supremum of every cell type, permutation by permutation.
//
// Replaces (reference paths under pctsea-core/src/main/java/edu/scripps/yates/pctsea/):
//   a9  PCTSEA.java:1390-1443 (permutation driver: Collections.shuffle of the ranked labels, P times)
//   a8  utils/parallel/EnrichmentWeigthedScoreParallel.java:104-227 evaluated once per (type, permutation)
//
// Permutation (PCTSEA_PERM_PHILOX): position x of permutation p carries the class at index pi_p(x) of the
// SIZE-SORTED arrangement of the ranked labels (classes in ascending order of their cell counts; a uniform
// permutation of any fixed arrangement of the same multiset is a uniform arrangement). pi_p is a keyed Feistel
// bijection of [0,Nu) on Z_a x Z_b (a near sqrt Nu, b = ceil(Nu/a) rounded up to a multiple of 4, cycle-walking for
// the few out-of-range points) with ARITHMETIC round functions
//   F(v, key, mod) = umulhi(mix(v * M + key), mod), mix(t) = t * (t >> 16), one Philox4x32-10 key per round and permutation,
// so it lives entirely in registers: no tables, and "class of index j" is a look-up in a two-region bucket table of the
// T+1 cumulative class counts in shared memory. The round that is applied LAST modifies the high digit (the one
// that decides the class); with the sorted arrangement the class of a position is a range test on that digit, which
// exposes the unevenness of a random round FUNCTION after 3-4 rounds as a 2 % inflation of the null's standard
// deviation (profiles/feistel_rounds_check.txt) — 5 rounds are the default, measured indistinguishable from
// Collections.shuffle.
//
// Arithmetic (PCTSEA_KS_FAST, round 2 form): scores quantised to int32 fixed point with the sum of |fx| over the
// ranking below 2^31, so every running sum is an exact int32 and the result cannot depend on the segmentation.
// Every ranked position is a hit for exactly one type, so sup|a-b| of a type can only sit at one of its hits or just
// before one, and there the two running sums are closed-form: numA from the per-(type, segment) state, the total
// prefix S from a register. One candidate = one fp32 FMA of the converted sums (the oracle restates the same definition).
//
// Kernel: persistent CTAs take permutations from a ticket counter; thread k of a CTA owns the segment
// [k*seg, (k+1)*seg) of the ranking and keeps its own per-type running sums in shared memory (numA[t][k], k fastest:
// conflict-free), 4 bytes per (type, thread):
//   phase 1  generate the labels of the segment (registers only), accumulate the per-type segment sums, park the
//            labels (1 B per position) in a per-CTA scratch row that stays in L2
//   phase 2  exclusive prefix over the segments per type; totals -> 1/denomA, 1/denomB
//   phase 3  second walk: running sums continue from the prefix, two candidates per position, per-type maxima by
//            shared-memory atomicMax behind a stale-read filter
//   phase 4  decode -> null ES and dStat
#include "common.cuh"

namespace pctsea {

// ================================================================================================
// fixed-point scores
// ================================================================================================
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

// F0 = 25 - e with maxabs < 2^e, so |llrint(s * 2^F0)| <= 2^25 (same rule as the oracle)
__device__ __forceinline__ int fast_frac_bits0(unsigned long long maxabs_bits) {
    const double maxabs = __longlong_as_double((long long)maxabs_bits);
    int e = 0;
    if (maxabs > 0.0) e = ilogb(maxabs) + 1;
    const int F = 25 - e;
    return F > 1000 ? 1000 : F;
}

// ... minus the smallest shift that makes ceil(sum|fx0| / 2^sh) + Nu fit an int32
__device__ __forceinline__ int fast_frac_bits(unsigned long long maxabs_bits, unsigned long long sumabs0, int64_t Nu) {
    int sh = 0;
    while ((long long)((sumabs0 + ((1ull << sh) - 1ull)) >> sh) + (long long)Nu > 2147483647LL) sh++;
    return fast_frac_bits0(maxabs_bits) - sh;
}

__device__ __forceinline__ double pow2_double(int F) {   // exact for the exponents that can occur (|F| <= 1000)
    return __longlong_as_double((long long)(1023 + F) << 52);
}

// sum over the ranking of |llrint(s * 2^F0)|: exact integer, order-independent
__global__ void __launch_bounds__(SC_THREADS) k_fx_abs_sum(const double* __restrict__ s_rank, int64_t Nu,
                                                          const unsigned long long* __restrict__ maxabs_bits,
                                                          unsigned long long* __restrict__ sumabs) {
    const double scale = pow2_double(fast_frac_bits0(*maxabs_bits));
    const int64_t base = (int64_t)blockIdx.x * SC_TILE;
    unsigned long long sum = 0;
#pragma unroll
    for (int r = 0; r < SC_ITEMS; r++) {
        const int64_t i = base + r * SC_THREADS + threadIdx.x;
        if (i < Nu) {
            const long long v = __double2ll_rn(__dmul_rn(s_rank[i], scale));
            sum += (unsigned long long)(v < 0 ? -v : v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(sumabs, sum);
}

// fx = llrint(s * 2^F) per ranked cell, plus the sum of every SC_TILE-sized tile for the prefix scan.
__global__ void __launch_bounds__(SC_THREADS) k_fixed_point(const double* __restrict__ s_rank, int64_t Nu,
                                                           const unsigned long long* __restrict__ maxabs_bits,
                                                           const unsigned long long* __restrict__ sumabs,
                                                           int* __restrict__ fx, int32_t* __restrict__ frac_out,
                                                           int* __restrict__ tile_sum) {
    __shared__ int ws[SC_THREADS / 32];
    const int F = fast_frac_bits(*maxabs_bits, *sumabs, Nu);
    if (blockIdx.x == 0 && threadIdx.x == 0) *frac_out = F;
    const double scale = pow2_double(F);
    const int64_t base = (int64_t)blockIdx.x * SC_TILE;
    int sum = 0;
#pragma unroll
    for (int r = 0; r < SC_ITEMS; r++) {
        const int64_t i = base + r * SC_THREADS + threadIdx.x;
        if (i < Nu) {
            const int v = (int)__double2ll_rn(__dmul_rn(s_rank[i], scale));
            fx[i] = v;
            sum += v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < SC_THREADS / 32; w++) t += ws[w];
        tile_sum[blockIdx.x] = t;
    }
}

__device__ __forceinline__ int warp_incl_scan_i32(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// Exact int32 prefix sums of the fixed-point scores: every tile first adds up the sums of the tiles before it (a few
// hundred values written by k_fixed_point), then scans its own items.
__global__ void __launch_bounds__(SC_THREADS) k_tile_scan_i32(const int* __restrict__ in, int64_t n,
                                                             const int* __restrict__ tile_sum, int* __restrict__ out) {
    // thread t owns items [t*8, t*8+8) of the tile => contiguous, so the scan is in index order
    __shared__ int ws[SC_THREADS / 32];
    __shared__ int s_before;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        int b = 0;
        for (int j = threadIdx.x; j < (int)blockIdx.x; j += SC_THREADS) b += tile_sum[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b += __shfl_down_sync(0xffffffffu, b, o);
        if (lane == 0) ws[warp] = b;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < SC_THREADS / 32; w++) t += ws[w];
            s_before = t;
        }
        __syncthreads();
    }
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    int v[SC_ITEMS];
    int s = 0;
#pragma unroll
    for (int r = 0; r < SC_ITEMS; r++) { const int64_t i = base + r; v[r] = (i < n) ? in[i] : 0; s += v[r]; v[r] = s; }
    const int incl = warp_incl_scan_i32(s, lane);
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; w++) woff += ws[w];
    const int off = s_before + woff + (incl - s);
#pragma unroll
    for (int r = 0; r < SC_ITEMS; r++) { const int64_t i = base + r; if (i < n) out[i] = v[r] + off; }
}

// Lane layout of the fixed-point scores: thread k of the null kernel owns [k*seg, (k+1)*seg), and fxT4[i4][k] holds
// its positions 4*i4 .. 4*i4+3, so that every thread's stream is a coalesced load.
__global__ void k_fx_transpose(const int* __restrict__ fx, int64_t Nu, int32_t seg, int32_t Kp, int4* __restrict__ fxT4) {
    const int seg4 = seg >> 2;
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (int64_t)seg4 * Kp) return;
    const int k = (int)(id % Kp), i4 = (int)(id / Kp);
    const int64_t x = (int64_t)k * seg + 4 * i4;
    int4 v;
    v.x = (x + 0 < Nu) ? fx[x + 0] : 0;
    v.y = (x + 1 < Nu) ? fx[x + 1] : 0;
    v.z = (x + 2 < Nu) ? fx[x + 2] : 0;
    v.w = (x + 3 < Nu) ? fx[x + 3] : 0;
    fxT4[id] = v;
}

// ================================================================================================
// class of an index of the size-sorted arrangement
// ================================================================================================
// The arrangement that is permuted lists the Tc = T + 1 classes (the last id = "unknown type") in ascending order of
// their number of ranked cells, ties by id; sorder[s] = id of the class at SORTED POSITION s, and
// cum[s] = number of ranked cells at sorted positions below s (s = 0 .. Tc). The null kernel works in sorted positions
// throughout (its per-type columns, the parked labels) and maps back to ids when it writes its results.
//   class(j) = the s with cum[s] <= j < cum[s+1].
// Look-up: one 8-byte shared-memory entry per bucket of indices,
//   .x = index of the bucket's only class boundary (0xffffffff: none inside), .y = class of the bucket's first index + 1,
// so that class(j) = .y - (j < .x): one read and a subtract with borrow. Real cell-type size distributions are
// long-tailed (config B's ranking: 100 types with a median of 87 cells next to buckets of 256 indices), and a bucket
// that holds several boundaries needs more than that. Because the arrangement is sorted by size, all classes narrower
// than a bucket sit in one short prefix [0, X) of the indices: that prefix gets its own, finer buckets (region A,
// 2^shiftA indices each), everything behind it the coarse ones (region B, 2^shiftB indices, at most one boundary each
// by construction). The host only fixes the table's capacity (what the CTA's shared memory has left); the split is
// chosen on the device from the class sizes: the (shiftB, shiftA) that leaves the fewest indices in classes narrower
// than a fine bucket. Which region an index belongs to is one compare and
// two selects, no divergence. (Round 2's first form resolved the several-boundary buckets through a second-level
// table in global memory: only 2 % of the indices took that path, but some lane of almost every warp did — 12 % of
// the kernel's warp instructions.) A region-A bucket that still holds several boundaries (classes of a handful of
// cells) has bit 31 of .y set: walk cum[] from the class.
constexpr uint32_t CT_MULTI = 0x80000000u;

struct ClassHdr {     // written by k_class_tables, read by the kernels that use the tables
    uint32_t X;       // indices below X use region A
    uint32_t shiftA;  // log2 of region A's bucket width
    uint32_t su;      // sorted position of the unknown-type class
    uint32_t nA;      // region A entries in use
    uint32_t shiftB;  // log2 of region B's bucket width
    uint32_t offB;    // first region B entry (region B is indexed by j >> shiftB over the whole ranking)
    uint32_t pad0, pad1;
};

// Global loads with a cache policy. The kernel streams its fixed-point scores and parked labels through L1 (every
// byte is used once per walk): the streams do not allocate there.
// The fixed-point scores (16 bytes per 4 positions, shared by every CTA and re-read in every walk of every
// permutation) are asked to stay in L2 (evict-last policy): on a long ranking the per-CTA scratch streams are many
// times the L2 and would otherwise push them out between two uses.
__device__ __forceinline__ uint64_t l2_keep_policy() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_stream_policy() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ int4 ld_keep(const int4* p, uint64_t pol) {
    int4 v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;"
        : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) {   // the CTA's own scratch: ordinary (coherent) load
    uint32_t v;
    asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint2 ld_stream(const uint2* p) {
    uint2 v;
    asm volatile("ld.global.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t class_search(const uint32_t* __restrict__ cum, uint32_t Tc, uint32_t j) {
    uint32_t lo = 0, hi = Tc;   // invariant: cum[lo] <= j < cum[hi]; empty classes (equal counts) are skipped
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cum[mid] <= j) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint2 class_entry(const uint32_t* __restrict__ cum, uint32_t Tc, long long start, long long width,
                                             long long Nu) {
    uint2 e = make_uint2(0xffffffffu, 1u);
    if (start < Nu) {
        const uint32_t last = (uint32_t)(min(start + width, Nu) - 1);
        const uint32_t ca = class_search(cum, Tc, (uint32_t)start), cb = class_search(cum, Tc, last);
        e.y = ca + 1u;   // indices below .x borrow: class ca; the others: ca + 1
        if (cb == ca + 1u) e.x = cum[cb];
        else if (cb > ca + 1u) e.y = (ca + 1u) | CT_MULTI;
    }
    return e;
}

// One CTA: sorted order of the classes, their cumulative counts, the two-region bucket table of `cap` entries
// (region A first, region B behind it).
__global__ void __launch_bounds__(1024) k_class_tables(const int32_t* __restrict__ class_cnt, int32_t Tc, int64_t Nu,
                                                      int shift_min, int32_t cap, uint32_t* __restrict__ cum,
                                                      int32_t* __restrict__ sorder, ClassHdr* __restrict__ hdr,
                                                      uint2* __restrict__ ctab) {
    __shared__ uint32_t part[1024];
    __shared__ uint32_t s_buf[2048];   // class counts by id, later the cumulative counts, when they fit
    __shared__ uint32_t s_shiftA, s_nA, s_shiftB, s_offB;
    const bool fits = Tc + 1 <= 2048;
    if (fits) for (int c = threadIdx.x; c < Tc; c += 1024) s_buf[c] = (uint32_t)class_cnt[c];
    __syncthreads();
    // sorted position of a class = number of classes that come before it in (count, id) order; the unsorted counts
    // land in cum[1 ..] at their sorted positions
    for (int c = threadIdx.x; c < Tc; c += 1024) {
        const uint32_t mine = fits ? s_buf[c] : (uint32_t)class_cnt[c];
        uint32_t r = 0;
        for (int d = 0; d < Tc; d++) {
            const uint32_t o = fits ? s_buf[d] : (uint32_t)class_cnt[d];
            r += (o < mine || (o == mine && d < c)) ? 1u : 0u;
        }
        sorder[r] = c;
        cum[r + 1] = mine;
        if (c == Tc - 1) hdr->su = r;
    }
    __syncthreads();
    // inclusive scan of cum[1 .. Tc] in place (thread t owns a contiguous chunk)
    const int per = (Tc + 1023) / 1024;
    const int c0 = 1 + threadIdx.x * per, c1 = min(c0 + per, (int)Tc + 1);
    uint32_t sum = 0;
    for (int c = c0; c < c1; c++) sum += cum[c];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 1024; i++) { const uint32_t v = part[i]; part[i] = run; run += v; }
        cum[0] = 0u;
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (int c = c0; c < c1; c++) { run += cum[c]; cum[c] = run; }
    __syncthreads();
    // the binary searches below run on a shared-memory copy of the cumulative counts when it fits: ~40 dependent
    // look-ups per thread, and this single CTA sits on the critical path of every analysis
    if (fits) {
        for (int c = threadIdx.x; c <= Tc; c += 1024) s_buf[c] = cum[c];
        __syncthreads();
        cum = s_buf;
    }
    if (threadIdx.x == 0) {
        // number of ranked cells in classes of fewer than w cells = cum[first sorted position whose class has >= w]
        auto cells_below = [&](unsigned long long w) -> unsigned long long {
            int lo = 0, hi = Tc;   // sorted by size: the first position with count >= w
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((unsigned long long)(cum[mid + 1] - cum[mid]) < w) lo = mid + 1; else hi = mid;
            }
            return cum[lo];
        };
        // Candidates: region B with buckets of 2^sB indices for the few sB that fit the capacity; what is left goes
        // to region A = [0, X) (the classes narrower than a region-B bucket, up to the next bucket boundary), as
        // fine as it can be. The cost of a candidate = the indices in classes narrower than a fine bucket (they
        // are resolved by walking cum[]); fewest wins, ties to the smaller sB.
        unsigned long long best_cost = ~0ull;
        uint32_t bX = 0, bsA = 0, bnA = 0, bsB = (uint32_t)shift_min, boff = 0;
        for (int sB = shift_min; sB <= shift_min + 5 && sB < 32; sB++) {
            const unsigned long long W = 1ull << sB;
            const long long nbB = (long long)(((unsigned long long)Nu + W - 1) >> sB);
            const long long capA = (long long)cap - nbB;
            if (capA < 0) continue;
            const unsigned long long raw = cells_below(W);
            unsigned long long X = (raw + W - 1) / W * W;
            const unsigned long long cover = min(X, (unsigned long long)Nu);
            uint32_t sA = 0;
            while (capA > 0 && ((cover + (1ull << sA) - 1) >> sA) > (unsigned long long)capA) sA++;
            unsigned long long cost;
            uint32_t nA = 0;
            if (capA <= 0 || (int)sA >= sB || raw == 0) { X = 0; sA = 0; cost = raw; }   // no finer than region B: no region A
            else { nA = (uint32_t)((cover + (1ull << sA) - 1) >> sA); cost = cells_below(1ull << sA); }
            if (cost < best_cost) {
                best_cost = cost; bX = (uint32_t)min(X, 0xffffffffull); bsA = sA; bnA = nA; bsB = (uint32_t)sB;
                boff = (uint32_t)(cap - nbB);
            }
            if (cost == 0) break;
        }
        s_shiftA = bsA; s_nA = bnA; s_shiftB = bsB; s_offB = boff;
        hdr->X = bX; hdr->shiftA = bsA; hdr->nA = bnA; hdr->shiftB = bsB; hdr->offB = boff;
    }
    __syncthreads();
    const uint32_t sA = s_shiftA, nA = s_nA, sB = s_shiftB, offB = s_offB;
    for (int i = threadIdx.x; i < (int)offB; i += 1024)
        ctab[i] = (uint32_t)i < nA ? class_entry(cum, (uint32_t)Tc, (long long)i << sA, 1ll << sA, Nu) : make_uint2(0xffffffffu, 1u);
    for (int i = threadIdx.x; i < cap - (int)offB; i += 1024)
        ctab[offB + i] = class_entry(cum, (uint32_t)Tc, (long long)i << sB, 1ll << sB, Nu);
}

// Sorted positions (classes) of four indices. saddrA / saddrB = addresses of the two regions of the bucket table in
// the shared window.
__device__ __forceinline__ void class_of4(const uint32_t (&j)[4], uint32_t saddrA, uint32_t saddrB, uint32_t X, uint32_t shiftA,
                                          uint32_t shiftB, const uint32_t* __restrict__ cum, uint32_t (&c)[4]) {
    uint32_t ex[4], ey[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const bool inA = j[q] < X;
        const uint32_t base = inA ? saddrA : saddrB, sh = inA ? shiftA : shiftB;
        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ex[q]), "=r"(ey[q]) : "r"(base + ((j[q] >> sh) << 3)));
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        // c = ey - (j < ex): .y holds the class of the bucket's first index PLUS ONE; the borrow of j - ex takes it back
        asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\tsubc.u32 %0, %3, 0;\n\t}" : "=r"(c[q]) : "r"(j[q]), "r"(ex[q]), "r"(ey[q]));
    }
    if ((c[0] | c[1] | c[2] | c[3]) & CT_MULTI) {   // a bucket of region A with several boundaries: rare
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (c[q] & CT_MULTI) {
                uint32_t cc = (ey[q] & ~CT_MULTI) - 1u;   // class of the bucket's first index
                while (j[q] >= cum[cc + 1]) cc++;          // ends: cum[Tc] = Nu > j
                c[q] = cc;
            }
    }
}

// ================================================================================================
// the keyed Feistel bijection
// ================================================================================================
struct FeistelParams {
    uint32_t a, b;
    int rounds;
    uint64_t seed;
};

// multiply - shift - multiply: t = v * M + key spreads the digit over the word, t * (t >> 16) folds its high half back
// in non-linearly (without that the round function is affine in the digit and neighbouring positions correlate).
// Three instructions, one of them on the ALU pipe; the multiply - xorshift - multiply hash it replaces took four
// (two on the ALU pipe, the kernel's busiest) and gives the same null (profiles/feistel_mix_check.txt).
__host__ __device__ __forceinline__ uint32_t feistel_mix(uint32_t v, uint32_t key) {
    const uint32_t t = v * 0x9E3779B1u + key;
    return t * (t >> 16);
}

// round keys of permutation p: Philox4x32-10(ctr = {j / 4, 0, p, "FEIS"}, key = seed)[j % 4]
template <int NR>
__device__ __forceinline__ void feistel_keys(uint64_t seed, uint32_t p, uint32_t (&keys)[NR]) {
#pragma unroll
    for (int j0 = 0; j0 < NR; j0 += 4) {
        uint32_t o[4];
        philox4x32_10((uint32_t)(j0 >> 2), 0u, p, 0x46454953u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (j0 + q < NR) keys[j0 + q] = o[q];
    }
}

__device__ __forceinline__ uint32_t mad_hi_u32(uint32_t x, uint32_t y, uint32_t z) {   // hi32(x * y) + z, one IMAD.HI
    uint32_t d;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(y), "r"(z));
    return d;
}

// One pass of the rounds over the digits (L, R). NR > 0: unrolled, keys in registers; NR == 0: `rounds` at run time,
// keys from memory.
template <int NR, typename Keys>
__device__ __forceinline__ void feistel_rounds(const Keys& keys, uint32_t a, uint32_t b, int rounds_rt, uint32_t& L, uint32_t& R) {
    const int rounds = NR > 0 ? NR : rounds_rt;
#pragma unroll
    for (int j = 0; j < rounds; j++) {
        if (((rounds - 1 - j) & 1) == 0) {   // the last round modifies the high digit
            L = mad_hi_u32(feistel_mix(R, keys[j]), a, L);
            L = min(L, L - a);   // unsigned: L - a wraps above L exactly when L < a
        } else {
            R = mad_hi_u32(feistel_mix(L, keys[j]), b, R);
            R = min(R, R - b);
        }
    }
}

template <int NR, typename Keys>
__device__ __forceinline__ uint32_t feistel_apply(const Keys& keys, uint32_t a, uint32_t b, int rounds_rt, uint32_t L,
                                                  uint32_t R, uint32_t n) {
    uint32_t y;
    do {
        feistel_rounds<NR>(keys, a, b, rounds_rt, L, R);
        y = L * b + R;
    } while (y >= n);  // cycle-walk: < a of the a*b points are out of range
    return y;
}

// Four labels of one thread's segment = one word of the lane layout.
template <typename LabT> struct LaneWord;
template <> struct LaneWord<uint8_t> {
    typedef uint32_t type;
    static __device__ __forceinline__ uint32_t get(uint32_t w, int q) { return __byte_perm(w, 0u, 0x4440u + (uint32_t)q); }   // one PRMT
    static __device__ __forceinline__ uint32_t pack(const uint32_t c[4]) {
        return __byte_perm(__byte_perm(c[0], c[1], 0x0040u), __byte_perm(c[2], c[3], 0x0040u), 0x5410u);
    }
};
template <> struct LaneWord<uint16_t> {
    typedef uint2 type;
    static __device__ __forceinline__ uint32_t get(uint2 w, int q) { return ((q < 2 ? w.x : w.y) >> (16 * (q & 1))) & 0xffffu; }
    static __device__ __forceinline__ uint2 pack(const uint32_t c[4]) { return make_uint2(c[0] | (c[1] << 16), c[2] | (c[3] << 16)); }
};

// Row-major permuted labels for the EXACT (validation) null: labels[p][x], one thread per position.
template <typename LabT>
__global__ void __launch_bounds__(256) k_labels_philox_rm(int64_t Nu, int64_t Nu_pad, int32_t T, FeistelParams fp,
                                                         int32_t perm_begin, const uint32_t* __restrict__ cum,
                                                         const int32_t* __restrict__ sorder, LabT* __restrict__ labels) {
    __shared__ uint32_t keys[32];
    const int p = blockIdx.y;
    if (threadIdx.x < 8 && (int)threadIdx.x * 4 < fp.rounds) {
        uint32_t o[4];
        philox4x32_10(threadIdx.x, 0u, (uint32_t)(perm_begin + p), 0x46454953u, (uint32_t)fp.seed, (uint32_t)(fp.seed >> 32), o);
#pragma unroll
        for (int q = 0; q < 4; q++) keys[threadIdx.x * 4 + q] = o[q];
    }
    __syncthreads();
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= Nu_pad) return;
    uint32_t cls = (uint32_t)T;
    if (x < Nu) {
        const uint32_t L = (uint32_t)(x / fp.b), R = (uint32_t)(x - (int64_t)L * fp.b);
        const uint32_t j = feistel_apply<0>(keys, fp.a, fp.b, fp.rounds, L, R, (uint32_t)Nu);
        cls = (uint32_t)sorder[class_search(cum, (uint32_t)T + 1u, j)];   // sorted position -> class id
    }
    labels[(size_t)p * Nu_pad + x] = (LabT)cls;
}

// ================================================================================================
// the fused null kernel
// ================================================================================================
__device__ __forceinline__ float dstat_of_f(float sup, int32_t sizea, int32_t sizeb) {
    // EnrichmentWeigthedScoreParallel.java:203-207: sizea*sizeb is an int32 multiply (wraps)
    const int32_t prod = (int32_t)((uint32_t)sizea * (uint32_t)sizeb);
    const int32_t sum = (int32_t)((uint32_t)sizea + (uint32_t)sizeb);
    const float sq = (float)__dsqrt_rn(__ddiv_rn((double)prod, (double)sum));
    return __fmul_rn(sup, sq);
}

__device__ __forceinline__ uint32_t sup_key_fast(float f) {
    // larger |diff| wins; on equal |diff| the positive one wins (an order-independent rule). +0 maps to 1 and -0 to
    // 0: both lose to every non-zero difference and decode to +0.
    const uint32_t b = (uint32_t)__float_as_int(f);
    return __funnelshift_l(~b, b, 1);   // (b << 1) | (~b >> 31)
}

constexpr int NF_PRE = 4;   // label / score words in flight per thread in the second walk

struct __align__(16) TypeConst {   // per cell type and permutation
    float c12, c2;     // the two constants of a candidate difference
    uint32_t best;     // ordering key of the running supremum (sup_key_fast)
    uint32_t top;      // float bits of its magnitude; +inf for a type that cannot have a supremum (zero constants)
};

// Shared-memory accesses by address in the shared window (32-bit address arithmetic, no generic-pointer conversion in
// the loops).
__device__ __forceinline__ int lds_i32(uint32_t saddr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_i32(uint32_t saddr, int v) {
    asm volatile("st.shared.s32 [%0], %1;" :: "r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 lds_u64(uint32_t saddr) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}

// Four read-modify-writes col[t[q]] += f[q] in program order with the four reads issued together: a repeated type
// takes the running value of its earlier occurrence from a register instead of waiting for the store (one
// shared-memory latency per four positions instead of four). bq / aq = value before / after each update.
__device__ __forceinline__ void rmw4(uint32_t col_saddr, uint32_t stride_bytes, const uint32_t (&t)[4], const int (&f)[4],
                                     int (&bq)[4], int (&aq)[4]) {
    uint32_t ad[4];
    int ld[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { ad[q] = col_saddr + t[q] * stride_bytes; ld[q] = lds_i32(ad[q]); }
    // All four reads precede the stores, so a repeated type reads the same old value at each occurrence: the value
    // before position q = what was read + the scores of the earlier positions of the same type. Those sums depend
    // on the types and scores only and are ready before the shared-memory reads return (two dependent adds behind
    // the read instead of a chain of selects and adds through all four positions).
    const int s10 = (t[1] == t[0]) ? f[0] : 0;
    const int s20 = (t[2] == t[0]) ? f[0] : 0, s21 = (t[2] == t[1]) ? f[1] : 0;
    const int s30 = (t[3] == t[0]) ? f[0] : 0, s31 = (t[3] == t[1]) ? f[1] : 0, s32 = (t[3] == t[2]) ? f[2] : 0;
    bq[0] = ld[0];
    bq[1] = ld[1] + s10;
    bq[2] = ld[2] + s20 + s21;
    bq[3] = ld[3] + (s30 + s31) + s32;
#pragma unroll
    for (int q = 0; q < 4; q++) aq[q] = bq[q] + f[q];
#pragma unroll
    for (int q = 0; q < 4; q++) sts_i32(ad[q], aq[q]);   // same address: the later store carries the later value
}

// 16-byte shared-memory read that is re-issued every time (the maxima change under the reader); saddr = address in
// the shared window
__device__ __forceinline__ uint4 lds128_volatile(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(saddr));
    return v;
}

struct NullArgs {
    const int4* fxT4;            // [seg4][Kp] fixed-point scores, lane layout
    const int* Sfx;              // [Nu] inclusive int32 prefix of the fixed-point scores
    const int32_t* class_cnt;    // [Tc]
    const uint32_t* cum;         // [Tc + 1] cumulative counts of the size-sorted classes
    const int32_t* sorder;       // [Tc] class id at every sorted position
    const ClassHdr* hdr;
    const uint2* ctab;           // [tab_cap]
    void* scratch;               // [grid][seg4][Kp] label words
    uint4* cstream;              // SLICED: [grid][seg + 4][Kp] compact records {class, score, prefix before, -} of the sparse slices
    int* ticket;
    float* null_ews;             // [T][P_stride]
    float* null_d;
    int64_t Nu;
    int32_t T_all, Ts, seg, Pc, perm_begin, P_stride, p_off, tab_cap, cum_smem;
    FeistelParams fp;
};

// Classes are sorted positions of the size-sorted arrangement everywhere inside the kernel (columns of the per-thread
// state, parked labels); sorder[] maps back to type ids when the results are written. The unknown-type class (sorted
// position su) has a column like every other class but zero constants, so it can never set a maximum; padding
// positions carry it with zero scores.
// SLICED: more classes than the per-thread shared-memory state can hold at once. The labels are then generated
// in a pass of their own, and phases 1-4 run once per slice of Ts sorted positions, with every other class mapped to
// the slice's dummy column (Ts), which takes the same read-modify-write but has zero constants. The classes are
// sorted by size, so only the LAST slice is dense (it holds the largest classes, i.e. most positions: config D's
// upper 150 of 301 classes own 89 % of the cells); it walks the parked labels of all positions. Every earlier slice
// is sparse: the label pass appends a {class, score, prefix of all scores before the position} record for each of
// their positions to a per-thread compact stream, and their phases 1 and 3 walk those records only — the work per
// permutation is ~1.2 walks instead of one per slice. The parked labels serve the dense slice alone, so they are
// parked as its LOCAL ids (members: offset in the slice; everything else: the dummy column), which always fit a
// byte.
template <typename LabT, int NR, bool SLICED>
__global__ void __launch_bounds__(512, 1) k_null_fused(const NullArgs A) {
    typedef typename LaneWord<LabT>::type Word;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_p;
    __shared__ uint32_t s_keys[32];
    const int Ts = A.Ts, T_all = A.T_all, Tc_all = T_all + 1;
    const int Tc = SLICED ? Ts + 1 : Tc_all;   // columns: the slice + its dummy, or every class
    const int Kp = blockDim.x;
    const int k = threadIdx.x, lane = k & 31, warp = k >> 5, nwarp = Kp >> 5;
    int* numA = reinterpret_cast<int*>(smem_raw);                               // [Tc][Kp]
    TypeConst* cb = reinterpret_cast<TypeConst*>(numA + (size_t)Tc * Kp);       // [Tc]
    uint2* ctab_s = reinterpret_cast<uint2*>(cb + Tc);                          // [tab_cap]
    uint32_t* cum_s = reinterpret_cast<uint32_t*>(ctab_s + A.tab_cap);          // [T_all + 2] when it fits
    const uint32_t* cum = A.cum_smem ? cum_s : A.cum;
    const uint32_t cb_saddr = (uint32_t)__cvta_generic_to_shared(cb);
    const uint32_t ctab_saddr = (uint32_t)__cvta_generic_to_shared(ctab_s);
    const uint32_t col_saddr = (uint32_t)__cvta_generic_to_shared(numA) + 4u * (uint32_t)threadIdx.x;
    const uint32_t Kp4 = 4u * (uint32_t)blockDim.x;
    for (int i = k; i < A.tab_cap; i += Kp) ctab_s[i] = A.ctab[i];
    const uint32_t ctabB_saddr = ctab_saddr + 8u * A.hdr->offB;
    const uint32_t X = A.hdr->X, shiftA = A.hdr->shiftA, shiftB = A.hdr->shiftB;
    const int su = (int)A.hdr->su;
    if (A.cum_smem) for (int i = k; i < T_all + 2; i += Kp) cum_s[i] = A.cum[i];
    const int seg = A.seg;
    const uint32_t n = (uint32_t)A.Nu, a = A.fp.a, b = A.fp.b;
    Word* lab = reinterpret_cast<Word*>(A.scratch) + (size_t)blockIdx.x * (seg >> 2) * Kp + k;
    uint4* cs = SLICED ? A.cstream + (size_t)blockIdx.x * (seg + 4) * Kp + k : nullptr;
    const uint32_t dense0 = SLICED ? (uint32_t)((Tc_all - 1) / Ts * Ts) : 0u;   // first sorted position of the last (dense) slice
    const int4* fxp = A.fxT4 + k;
    const uint64_t keep = l2_keep_policy();
    int* col = numA + k;   // this thread's private column: type t lives at col[t * Kp] (bank = k % 32 for every t)
    const uint32_t x0 = (uint32_t)k * (uint32_t)seg;
    const uint32_t L00 = x0 / b, R00 = x0 - L00 * b;
    // my positions: nv valid ones = nfull whole words + possibly one partial word; nothing behind them is visited
    // (the fixed-point scores are zero-padded, the labels of a partial word's tail are the unknown class)
    const int nv = x0 < n ? (int)min((uint32_t)seg, n - x0) : 0;
    const int nfull = nv >> 2, nw = (nv + 3) >> 2, nwp = (nw + NF_PRE - 1) / NF_PRE * NF_PRE;   // seg / 4 is a multiple of NF_PRE
    const int Stot = A.Sfx[A.Nu - 1];
    const int S0 = (x0 > 0u && x0 <= n) ? A.Sfx[x0 - 1u] : 0;   // exact prefix before my segment

    for (;;) {
        __syncthreads();   // the previous permutation is done with the shared state (and the tables are loaded)
        if (k == 0) s_p = atomicAdd(A.ticket, 1);
        __syncthreads();
        const int p = s_p;
        if (p >= A.Pc) break;
        uint32_t keys[NR > 0 ? NR : 1];
        if constexpr (NR > 0) feistel_keys<NR>(A.fp.seed, (uint32_t)(A.perm_begin + p), keys);
        else {
            if (k < 8 && k * 4 < A.fp.rounds) {
                uint32_t o[4];
                philox4x32_10((uint32_t)k, 0u, (uint32_t)(A.perm_begin + p), 0x46454953u, (uint32_t)A.fp.seed,
                              (uint32_t)(A.fp.seed >> 32), o);
#pragma unroll
                for (int q = 0; q < 4; q++) s_keys[k * 4 + q] = o[q];
            }
            __syncthreads();
        }
        auto rounds = [&](uint32_t& L, uint32_t& R) {
            if constexpr (NR > 0) feistel_rounds<NR>(keys, a, b, NR, L, R);
            else feistel_rounds<0>(s_keys, a, b, A.fp.rounds, L, R);
        };
        // (SLICED: the dense slice comes first and takes its segment sums from the label pass, like the one-pass form)
        for (int t = 0; t < Tc; t++) col[t * Kp] = 0;
        for (int t = k; t < Tc; t += Kp) cb[t].best = 0u;

        // ---- phase 1: labels of my segment + per-type segment sums (one shared-memory read-modify-write per
        // position, in program order). Four positions at a time: four independent round chains, no per-position tests.
        uint32_t nrec = 0;   // SLICED: records in my compact stream
        {
            int Srun = S0;       // SLICED: exact prefix of all scores before the current word
            uint4* csw = cs;     // SLICED: next free record
            uint32_t L = L00, R = R00;
            uint32_t idx = 0;   // i4 * Kp: the same index addresses the label word and the score quad of (i4, k)
            int4 fn = nw > 0 ? ld_keep(fxp, keep) : make_int4(0, 0, 0, 0);
            for (int i4 = 0; i4 < nw; i4++) {
                const int4 fc = fn;
                if (i4 + 1 < nw) fn = ld_keep(fxp + (idx + (uint32_t)Kp), keep);
                const int fq[4] = { fc.x, fc.y, fc.z, fc.w };
                uint32_t y[4], cls[4];
                if (i4 < nfull) {
                    uint32_t Lq[4], Rq[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        // digits of position x + q: rows are a multiple of 4 positions long and x is a multiple of 4, so
                        // the four positions of a word share their row
                        Lq[q] = L;
                        Rq[q] = R + (uint32_t)q;
                        rounds(Lq[q], Rq[q]);
                        y[q] = Lq[q] * b + Rq[q];
                    }
                    R += 4u;
                    if (R == b) { R = 0u; L++; }
                    if (max(max(y[0], y[1]), max(y[2], y[3])) >= n) {   // cycle-walk: rare
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            while (y[q] >= n) { rounds(Lq[q], Rq[q]); y[q] = Lq[q] * b + Rq[q]; }
                    }
                    class_of4(y, ctab_saddr, ctabB_saddr, X, shiftA, shiftB, cum, cls);
                } else {   // the partial last word of the ranking
                    y[0] = y[1] = y[2] = y[3] = 0xffffffffu;
#pragma unroll 1
                    for (int q = 0; q < 4; q++) {
                        if (i4 * 4 + q < nv) {
                            uint32_t Lc = L, Rc = R, yy;
                            do { rounds(Lc, Rc); yy = Lc * b + Rc; } while (yy >= n);
                            y[0] = q == 0 ? yy : y[0]; y[1] = q == 1 ? yy : y[1];   // (no dynamic indexing: y[] stays in registers)
                            y[2] = q == 2 ? yy : y[2]; y[3] = q == 3 ? yy : y[3];
                        }
                        R++;
                        if (R >= b) { R = 0u; L++; }
                    }
                    uint32_t jj[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) jj[q] = min(y[q], n - 1u);
                    class_of4(jj, ctab_saddr, ctabB_saddr, X, shiftA, shiftB, cum, cls);
#pragma unroll
                    for (int q = 0; q < 4; q++) cls[q] = (y[q] == 0xffffffffu) ? (uint32_t)su : cls[q];
                }
                if (!SLICED) {
                    int bq[4], aq[4];
                    rmw4(col_saddr, Kp4, cls, fq, bq, aq);
                } else {
                    int sb = Srun;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (cls[q] + 1u <= dense0 && i4 * 4 + q < nv) {   // a position of a sparse slice
                            *csw = make_uint4(cls[q], (uint32_t)fq[q], (uint32_t)sb, 0u);
                            csw += Kp;
                            nrec++;
                        }
                        sb += fq[q];
                        // parked for the dense slice: its local id, or the dummy column
                        const uint32_t tl = cls[q] - dense0;
                        cls[q] = tl < (uint32_t)(Tc_all - (int)dense0) ? tl : (uint32_t)Ts;
                    }
                    Srun = sb;
                    int bq[4], aq[4];
                    rmw4(col_saddr, Kp4, cls, fq, bq, aq);   // the dense slice's segment sums
                }
                lab[idx] = LaneWord<LabT>::pack(cls);
                idx += (uint32_t)Kp;
            }
            // pad my words to a whole number of NF_PRE groups (unknown class, zero scores): the second walk has no tail
            const uint32_t pad = SLICED ? (uint32_t)Ts : (uint32_t)su;   // zero scores on a column that cannot set a maximum
            const uint32_t dummy[4] = { pad, pad, pad, pad };
            for (int i4 = nw; i4 < nwp; i4++) { lab[idx] = LaneWord<LabT>::pack(dummy); idx += (uint32_t)Kp; }
            if (SLICED)   // whole groups of four records: the padding belongs to no slice (dummy column, zero score)
                for (; nrec & 3u; nrec++) { *csw = make_uint4(0xffffffffu, 0u, 0u, 0u); csw += Kp; }
        }

      // one iteration unless SLICED; then the dense (last) slice first, the sparse ones behind it
      for (int sl = 0, t0 = SLICED ? (int)dense0 : 0; t0 >= 0 && t0 < Tc_all;
           sl++, t0 = SLICED ? ((sl - 1) * Ts < (int)dense0 ? (sl - 1) * Ts : -1) : Tc_all) {
        const int T = SLICED ? min(Ts, Tc_all - t0) : Tc_all;   // classes of this slice: local ids [0, T) (dummy = Ts when SLICED)
        // (parked labels are column ids already: the sorted positions themselves, or the dense slice's local ids)
        const bool sparse = SLICED && (uint32_t)t0 < dense0;
        // four records of my compact stream -> local classes (members of this slice keep their offset, everything else
        // goes to the dummy column), score, and the exact prefix before / after the position
        auto load_records = [&](const uint4* p, uint4 (&rec)[4]) {
#pragma unroll
            for (int u = 0; u < 4; u++) rec[u] = ld_stream(p + (size_t)u * Kp);
        };
        auto decode_records = [&](const uint4 (&rec)[4], uint32_t (&t)[4], int (&f)[4], int (&sb)[4], int (&sa)[4]) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t tl = rec[u].x - (uint32_t)t0;
                t[u] = tl < (uint32_t)T ? tl : (uint32_t)Ts;
                f[u] = (int)rec[u].y;
                sb[u] = (int)rec[u].z;
                sa[u] = sb[u] + f[u];
            }
        };
        if (sparse) {
            for (int t = 0; t < Tc; t++) col[t * Kp] = 0;
            for (int t = k; t < Tc; t += Kp) cb[t].best = 0u;
            // phase 1 on the compact records (each thread re-reads what it wrote itself: no barrier needed)
            {
                uint4 rec[4], nxt[4];   // the next group's loads are in flight while this one is processed
                if (nrec) load_records(cs, rec);
                for (uint32_t r = 0; r < nrec; r += 4) {
                    if (r + 4 < nrec) load_records(cs + (size_t)(r + 4) * Kp, nxt);
                    uint32_t t[4];
                    int f[4], sb[4], sa[4], bq[4], aq[4];
                    decode_records(rec, t, f, sb, sa);
                    rmw4(col_saddr, Kp4, t, f, bq, aq);
#pragma unroll
                    for (int u = 0; u < 4; u++) rec[u] = nxt[u];
                }
            }
        }
        __syncthreads();

        // ---- phase 2: exclusive prefix over the segments (warp w scans types w, w + nwarp, ...: 128 segments per step,
        // four per lane), then the reciprocals, one type per lane
        for (int base = 0; base < Tc; base += 32 * nwarp) {   // one pass unless Tc > 32 * nwarp
            int mytot = 0;
            for (int i = 0; i < 32; i++) {
                const int t = base + warp + i * nwarp;
                if (t >= Tc) break;
                int* row = numA + t * Kp;
                int carry = 0;
                for (int k0 = 0; k0 < Kp; k0 += 128) {
                    const int kk = k0 + 4 * lane;
                    const bool in = kk < Kp;   // Kp is a multiple of 32
                    int4 v = make_int4(0, 0, 0, 0);
                    if (in) v = *reinterpret_cast<int4*>(row + kk);
                    const int s1 = v.x, s2 = s1 + v.y, s3 = s2 + v.z, s4 = s3 + v.w;
                    const int incl = warp_incl_scan_i32(s4, lane);
                    const int ex = carry + (incl - s4);
                    if (in) *reinterpret_cast<int4*>(row + kk) = make_int4(ex, ex + s1, ex + s2, ex + s3);
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == i) mytot = carry;
            }
            const int t = base + warp + lane * nwarp;
            if (t < Tc) {
                const int dA = mytot, dB = Stot - mytot;
                float c12 = 0.0f, c2 = 0.0f;
                if (t < T && t0 + t != su && dA != 0 && dB != 0) {
                    const double r2 = __ddiv_rn(1.0, (double)dB);
                    c2 = __double2float_rn(r2);
                    c12 = __double2float_rn(__dadd_rn(__ddiv_rn(1.0, (double)dA), r2));
                }
                cb[t].c12 = c12;
                cb[t].c2 = c2;
                cb[t].top = (c12 != 0.0f || c2 != 0.0f) ? 0u : 0x7f800000u;
            }
        }
        __syncthreads();

        // ---- phase 3: suprema at hit / pre-hit positions. At the global position 0 the pre-hit candidate is
        // exactly 0 and cannot win.
        if (sparse) {
            uint4 rec[4], nxt[4];
            if (nrec) load_records(cs, rec);
            for (uint32_t r = 0; r < nrec; r += 4) {
                if (r + 4 < nrec) load_records(cs + (size_t)(r + 4) * Kp, nxt);
                uint32_t t[4];
                int f[4], sb[4], sa[4], bq[4], aq[4];
                decode_records(rec, t, f, sb, sa);
#pragma unroll
                for (int u = 0; u < 4; u++) rec[u] = nxt[u];
                uint4 cq[4];
#pragma unroll
                for (int q = 0; q < 4; q++) cq[q] = lds128_volatile(cb_saddr + t[q] * 16u);
                rmw4(col_saddr, Kp4, t, f, bq, aq);
                float d0[4], d1[4];
                bool rec = false;
#pragma unroll
                for (int q = 0; q < 4; q++) {   // the same expressions as the dense walk below
                    const float c12 = __uint_as_float(cq[q].x), c2 = __uint_as_float(cq[q].y), top = __uint_as_float(cq[q].w);
                    d1[q] = __fmaf_rn(__int2float_rn(aq[q]), c12, -__fmul_rn(__int2float_rn(sa[q]), c2));
                    d0[q] = __fmaf_rn(__int2float_rn(bq[q]), c12, -__fmul_rn(__int2float_rn(sb[q]), c2));
                    rec = rec || (fabsf(d1[q]) >= top) || (fabsf(d0[q]) >= top);
                }
                if (rec) {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t k1 = sup_key_fast(d1[q]), k0 = sup_key_fast(d0[q]);
                        const uint32_t key = k0 > k1 ? k0 : k1;
                        if (key > cq[q].z) {
                            atomicMax(&cb[t[q]].best, key);
                            atomicMax(&cb[t[q]].top, key >> 1);
                        }
                    }
                }
            }
        } else {
            int S = S0;
            float Sf = __int2float_rn(S);
            Word wc[NF_PRE];   // register sets refilled in place: the loads run NF_PRE - 1 words ahead of their use
            int4 fc[NF_PRE];
            const uint32_t Kpu = (uint32_t)Kp;
            uint32_t idx = 0;   // index of the group's first word
            if (nwp > 0) {
#pragma unroll
                for (int u = 0; u < NF_PRE; u++) { wc[u] = ld_stream(lab + (idx + u * Kpu)); fc[u] = ld_keep(fxp + (idx + u * Kpu), keep); }
            }
            for (int i4 = 0; i4 < nwp; i4 += NF_PRE) {
                const bool more = i4 + NF_PRE < nwp;
                idx += NF_PRE * Kpu;
#pragma unroll
                for (int u = 0; u < NF_PRE; u++) {
                    const int fq[4] = { fc[u].x, fc[u].y, fc[u].z, fc[u].w };
                    const Word wcu = wc[u];
                    if (more) { wc[u] = ld_stream(lab + (idx + u * Kpu)); fc[u] = ld_keep(fxp + (idx + u * Kpu), keep); }
                    uint32_t t[4];
                    int bq[4], aq[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) t[q] = LaneWord<LabT>::get(wcu, q);
                    // the constants and the current maximum of each type in one 16-byte read, issued together with
                    // the running-sum reads so that their latency hides behind each other
                    uint4 cq[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) cq[q] = lds128_volatile(cb_saddr + t[q] * 16u);
                    rmw4(col_saddr, Kp4, t, fq, bq, aq);
                    // Two candidates per position, compared as floats against the largest |diff| of the type so far
                    // (.w; it only grows, so a candidate that does not reach the possibly stale value read here
                    // cannot beat the current one either). Record-setting candidates are rare: the exact ordering
                    // keys and the shared-memory atomics sit behind one branch per group of four positions.
                    float d0[4], d1[4];
                    bool rec = false;
                    // the prefixes of all scores after each of the four positions, each one add away from S (exact:
                    // |S| <= sum |fx| < 2^31), so that the four conversions do not wait for each other
                    const int p2 = fq[0] + fq[1];
                    const int Sq[4] = { S + fq[0], S + p2, S + (p2 + fq[2]), S + (p2 + (fq[2] + fq[3])) };
                    float Sfq[5];
                    Sfq[0] = Sf;
#pragma unroll
                    for (int q = 0; q < 4; q++) Sfq[q + 1] = __int2float_rn(Sq[q]);
                    S = Sq[3];
                    Sf = Sfq[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const float c12 = __uint_as_float(cq[q].x), c2 = __uint_as_float(cq[q].y), top = __uint_as_float(cq[q].w);
                        d1[q] = __fmaf_rn(__int2float_rn(aq[q]), c12, -__fmul_rn(Sfq[q + 1], c2));
                        d0[q] = __fmaf_rn(__int2float_rn(bq[q]), c12, -__fmul_rn(Sfq[q], c2));
                        rec = rec || (fabsf(d1[q]) >= top) || (fabsf(d0[q]) >= top);
                    }
                    if (rec) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint32_t k1 = sup_key_fast(d1[q]), k0 = sup_key_fast(d0[q]);
                            const uint32_t key = k0 > k1 ? k0 : k1;
                            if (key > cq[q].z) {
                                atomicMax(&cb[t[q]].best, key);
                                atomicMax(&cb[t[q]].top, key >> 1);   // the bits of |diff|: non-negative floats order as integers
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase 4: decode
        for (int ts = k; ts < T; ts += Kp) {
            if (t0 + ts == su) continue;                 // the unknown-type class has no result
            const int t = A.sorder[t0 + ts] - t0;        // (t0 + t) = type id of this sorted position
            const int32_t nk = A.class_cnt[t0 + t];
            const int32_t sizeb = (int32_t)(A.Nu - nk);
            float v, d;
            if (!(nk > 1 && sizeb > 1)) { v = __int_as_float(0x7fc00000); d = v; }
            else {
                const uint32_t key = cb[ts].best;
                v = (key <= 1u) ? 0.0f : __int_as_float((int)((key >> 1) | (((key & 1u) ^ 1u) << 31)));
                d = dstat_of_f(v, nk, sizeb);
            }
            A.null_ews[(size_t)(t0 + t) * A.P_stride + A.p_off + p] = v;
            A.null_d[(size_t)(t0 + t) * A.P_stride + A.p_off + p] = d;
        }
        if (SLICED) __syncthreads();   // the next slice reuses the state
      }
    }
}

// ================================================================================================
// host side
// ================================================================================================
static void feistel_params(const EnrichGeom& g, uint64_t seed, int rounds, FeistelParams* fp) {
    feistel_dims(g.Nu, &fp->a, &fp->b);
    fp->rounds = rounds;
    fp->seed = seed;
}

// Bucket table geometry: at most 1024 buckets.
static void class_table_geom(int64_t Nu, int* shift, int32_t* nb) {
    int s = 0;
    while (((Nu - 1) >> s) + 1 > 1024) s++;
    *shift = s;
    *nb = (int32_t)(((Nu - 1) >> s) + 1);
}

NullPlanFast plan_null_fast(const pctsea_ctx* c, const EnrichGeom& g) {
    NullPlanFast pl;
    class_table_geom(std::max<int64_t>(g.Nu, 1), &pl.shift, &pl.nb);
    pl.cum_smem = (g.Tc + 1 <= 1024) ? 1 : 0;
    const size_t smem_sm = (size_t)c->smem_per_sm, smem_cta_max = (size_t)c->smem_per_block_optin;
    // columns of the per-thread state: every class (T + 1) in one pass, or a slice of ts classes + the dummy column
    auto fixed = [&](int cols) { return (size_t)cols * 16 + (size_t)pl.nb * 8 + (pl.cum_smem ? (size_t)(g.Tc + 1) * 4 : 0) + 64; };
    const int max_threads_sm = 65536 / 128;   // the kernel is compiled for up to 128 registers per thread
    // threads per CTA when `ctas` CTAs share an SM (the per-CTA budget includes the 1 KB the driver reserves)
    // per-CTA budget: the SM's shared memory over the co-resident CTAs, minus the 1 KB the driver reserves per CTA and
    // the kernel's few static variables
    auto budget_for = [&](int ctas) { return std::min(smem_cta_max, smem_sm / (size_t)ctas - 1024) - 256; };
    auto threads_for_n = [&](int cols, int ctas) {
        const size_t f = fixed(cols);
        const size_t budget = budget_for(ctas);
        if (f + (size_t)cols * 4 * 32 > budget) return 0;
        const int kp = (int)std::min<size_t>((size_t)(max_threads_sm / ctas), (budget - f) / ((size_t)cols * 4));
        return (kp / 32) * 32;
    };
    // Two co-resident CTAs of half the size beat one large CTA at the same warp count (their serial phases 2 and 4
    // hide behind the other CTA's walks: 0.97 -> 0.87 ms on config B), so among the geometries with the most threads
    // per SM the one with more CTAs wins, down to 128 threads per CTA.
    int ctas_chosen = 1;
    auto threads_for = [&](int cols) {
        int best = threads_for_n(cols, 1), best_total = best;
        ctas_chosen = 1;
        const int max_ctas = c->opt_null_ctas_per_sm > 0 ? c->opt_null_ctas_per_sm : 2;
        for (int n = 2; n <= max_ctas; n++) {
            const int kp = threads_for_n(cols, n);
            if (kp >= 128 && kp * n >= best_total) { best = kp; best_total = kp * n; ctas_chosen = n; }
        }
        return best;
    };
    const int n_classes = g.T + 1;
    int ts = g.T, cols = n_classes, kp = threads_for(cols);   // ts == T: not sliced
    const int min_kp = c->opt_null_min_threads > 0 ? c->opt_null_min_threads : 192;
    if (kp < min_kp) {
        // slices of equal size over the T + 1 classes, as few as give every slice at least 256 threads (or the
        // maximum that fits); a slice has one more column, the dummy
        // Only the last slice is dense (the classes are sorted by size; the earlier slices walk their compact record
        // streams), so more slices cost little and occupancy decides: the fewest slices that reach 480 threads per
        // SM, else the geometry with the most.
        int best_total = -1, best_ts = 1, best_cols = 2, best_kp = 0, best_ctas = 1;
        for (int n_slices = 2;; n_slices++) {
            ts = (int)div_up(n_classes, n_slices);
            cols = ts + 1;
            kp = threads_for(cols);
            if (kp * ctas_chosen > best_total) {
                best_total = kp * ctas_chosen; best_ts = ts; best_cols = cols; best_kp = kp; best_ctas = ctas_chosen;
            }
            if (best_total >= 480 || ts == 1 || n_slices >= 16) break;
        }
        ts = best_ts; cols = best_cols; kp = best_kp; ctas_chosen = best_ctas;
        PCT_REQUIRE(kp >= 32, "shared memory too small for the null kernel's per-type state");
    }
    if (c->opt_null_threads > 0) kp = std::min(kp, (int)round_up(c->opt_null_threads, 32));
    // short rankings: at least ~8 positions per thread
    const int64_t want = std::max<int64_t>(32, round_up(div_up(std::max<int64_t>(g.Nu, 1), 8), 32));
    kp = (int)std::min<int64_t>(kp, want);
    pl.Kp = kp;
    pl.Ts = ts;
    pl.seg = (int32_t)round_up(div_up(std::max<int64_t>(g.Nu, 1), kp), 4 * NF_PRE);
    // The class table takes what is left of the CTA's share of shared memory on top of the nb coarse entries the
    // geometry was sized with, up to 1024 more; how the capacity is split between fine and coarse buckets is decided
    // on the device from the class sizes (k_class_tables).
    {
        const size_t used = (size_t)cols * kp * 4 + fixed(cols);
        const size_t budget = budget_for(ctas_chosen);
        int32_t extra = (int32_t)std::min<size_t>(1024, budget > used ? (budget - used) / 8 : 0);
        if (c->opt_null_fine_entries >= 0) extra = std::min(extra, c->opt_null_fine_entries);
        pl.nA_cap = extra;
    }
    pl.tab_cap = pl.nb + pl.nA_cap;
    pl.smem = (size_t)cols * kp * 4 + fixed(cols) + (size_t)pl.nA_cap * 8;
    int per_sm = (int)std::max<size_t>(1, smem_sm / (pl.smem + 1024));
    per_sm = std::min(per_sm, std::max(1, 2048 / kp));
    per_sm = std::min(per_sm, 4);
    if (c->opt_null_ctas_per_sm > 0) per_sm = std::min(per_sm, c->opt_null_ctas_per_sm);
    pl.ctas_per_sm = per_sm;
    // parked labels: sorted positions (one pass, at most T) or the dense slice's local ids (at most ts)
    pl.word_bytes = (ts < g.T ? ts : g.T) <= 255 ? 4 : 8;
    return pl;
}

// Once per enrich call (FAST arithmetic): fixed-point scores, their prefix and lane layout, class tables.
void launch_null_fast_prepare(pctsea_ctx* c, const EnrichGeom& g, const NullPlanFast& pl) {
    if (g.Nu == 0) return;
    unsigned long long* maxabs = c->counters.as<unsigned long long>() + 2;
    unsigned long long* sumabs = c->counters.as<unsigned long long>() + 4;
    int32_t* frac_dev = reinterpret_cast<int32_t*>(c->counters.as<unsigned long long>() + 3);
    c->fx.reserve((size_t)g.Nu * 4 + 64);
    c->Sfx.reserve((size_t)g.Nu * 4 + 64);
    const int n_tiles = (int)div_up(g.Nu, SC_TILE);
    c->scan_tmp.reserve((size_t)n_tiles * 4 + 64);
    k_fx_abs_sum<<<n_tiles, SC_THREADS, 0, c->stream>>>(c->s_rank.as<double>(), g.Nu, maxabs, sumabs);
    k_fixed_point<<<n_tiles, SC_THREADS, 0, c->stream>>>(c->s_rank.as<double>(), g.Nu, maxabs, sumabs, c->fx.as<int>(),
                                                         frac_dev, c->scan_tmp.as<int>());
    k_tile_scan_i32<<<n_tiles, SC_THREADS, 0, c->stream>>>(c->fx.as<int>(), g.Nu, c->scan_tmp.as<int>(), c->Sfx.as<int>());
    const int seg4 = pl.seg >> 2;
    c->fxT.reserve((size_t)seg4 * pl.Kp * 16 + 64);
    k_fx_transpose<<<(unsigned)div_up((int64_t)seg4 * pl.Kp, 256), 256, 0, c->stream>>>(c->fx.as<int>(), g.Nu, pl.seg, pl.Kp,
                                                                                      c->fxT.as<int4>());
    count_launch(c, 4);
    PCT_CUDA(cudaGetLastError());
}

// Layout of c->cls_tab: cum[Tc + 1] | sorder[Tc] | ClassHdr | ctab[nA_cap + nb]; needs class_cnt (k_gather_ranked).
struct ClassTabLayout { size_t sorder, hdr, ctab, bytes; };
static ClassTabLayout class_tab_layout(const EnrichGeom& g, int32_t cap) {
    ClassTabLayout l;
    l.sorder = (size_t)round_up((int64_t)(g.Tc + 1) * 4, 16);
    l.hdr = l.sorder + (size_t)round_up((int64_t)g.Tc * 4, 16);
    l.ctab = l.hdr + 32;
    l.bytes = l.ctab + (size_t)std::max(cap, 0) * 8 + 64;
    return l;
}

void launch_class_tables(pctsea_ctx* c, const EnrichGeom& g, int shift_min, int32_t cap) {
    const ClassTabLayout l = class_tab_layout(g, cap);
    c->cls_tab.reserve(l.bytes);
    unsigned char* base = c->cls_tab.as<unsigned char>();
    c->cls_tab_cap = cap;
    k_class_tables<<<1, 1024, 0, c->stream>>>(c->class_cnt.as<int32_t>(), g.Tc, g.Nu, shift_min, cap,
                                              reinterpret_cast<uint32_t*>(base), reinterpret_cast<int32_t*>(base + l.sorder),
                                              reinterpret_cast<ClassHdr*>(base + l.hdr), reinterpret_cast<uint2*>(base + l.ctab));
    count_launch(c);
    PCT_CUDA(cudaGetLastError());
}

void launch_labels_philox_rm(pctsea_ctx* c, const EnrichGeom& g, uint64_t seed, int32_t perm_begin, int32_t Pc, int rounds) {
    if (Pc == 0 || g.Nu == 0) return;
    FeistelParams fp;
    feistel_params(g, seed, rounds, &fp);
    const uint32_t* cum = c->cls_tab.as<uint32_t>();
    const int32_t* sorder = reinterpret_cast<const int32_t*>(c->cls_tab.as<unsigned char>() +
                                                             class_tab_layout(g, c->cls_tab_cap).sorder);
    for (int32_t p0 = 0; p0 < Pc; p0 += 32768) {
        const int32_t pc = std::min<int32_t>(32768, Pc - p0);
        dim3 grid((unsigned)div_up(g.Nu_pad, 256), (unsigned)pc);
        if (g.lab_bytes == 1)
            k_labels_philox_rm<uint8_t><<<grid, 256, 0, c->stream>>>(g.Nu, g.Nu_pad, g.T, fp, perm_begin + p0, cum, sorder,
                                                                     c->labels.as<uint8_t>() + (size_t)p0 * g.Nu_pad);
        else
            k_labels_philox_rm<uint16_t><<<grid, 256, 0, c->stream>>>(g.Nu, g.Nu_pad, g.T, fp, perm_begin + p0, cum, sorder,
                                                                      c->labels.as<uint16_t>() + (size_t)p0 * g.Nu_pad);
        count_launch(c);
    }
    PCT_CUDA(cudaGetLastError());
}

void launch_null_fused(pctsea_ctx* c, const EnrichGeom& g, const NullPlanFast& pl, uint64_t seed, int32_t perm_begin,
                       int32_t Pc, int rounds, int32_t P_stride, int32_t p_off) {
    if (Pc == 0 || g.T == 0 || g.Nu == 0) return;
    NullArgs A;
    A.fxT4 = c->fxT.as<int4>();
    A.Sfx = c->Sfx.as<int>();
    A.class_cnt = c->class_cnt.as<int32_t>();
    const ClassTabLayout l = class_tab_layout(g, pl.tab_cap);
    A.cum = c->cls_tab.as<uint32_t>();
    A.sorder = reinterpret_cast<const int32_t*>(c->cls_tab.as<unsigned char>() + l.sorder);
    A.hdr = reinterpret_cast<const ClassHdr*>(c->cls_tab.as<unsigned char>() + l.hdr);
    A.ctab = reinterpret_cast<const uint2*>(c->cls_tab.as<unsigned char>() + l.ctab);
    const int seg4 = pl.seg >> 2;
    const bool sliced = pl.Ts < g.T;
    // One scratch row per resident CTA: the parked labels and, when the classes go in slices, the compact records
    // (16 B per position at worst). The persistent grid shrinks when that would pass 24 GiB (a 2 M-cell, 300-type
    // ranking takes 10 GB with all 296 CTAs; a 50 M-cell one would take 250 GB).
    const size_t row_bytes = (size_t)seg4 * pl.Kp * pl.word_bytes + (sliced ? (size_t)(pl.seg + 4) * pl.Kp * sizeof(uint4) : 0);
    const int64_t grid_cap = std::max<int64_t>(1, (int64_t)(((size_t)24 << 30) / std::max<size_t>(row_bytes, 1)));
    const unsigned grid = (unsigned)std::min<int64_t>(std::min<int64_t>(Pc, (int64_t)pl.ctas_per_sm * c->num_sms), grid_cap);
    const size_t lab_bytes = (size_t)round_up((int64_t)((size_t)grid * seg4 * pl.Kp * pl.word_bytes), 256);
    const size_t cs_bytes = sliced ? (size_t)grid * (pl.seg + 4) * pl.Kp * sizeof(uint4) : 0;
    c->labels.reserve(lab_bytes + cs_bytes + 64);
    A.scratch = c->labels.p;
    A.cstream = sliced ? reinterpret_cast<uint4*>(c->labels.as<unsigned char>() + lab_bytes) : nullptr;
    c->lab_ticket.reserve(64);
    PCT_CUDA(cudaMemsetAsync(c->lab_ticket.p, 0, 4, c->stream));
    A.ticket = c->lab_ticket.as<int>();
    A.null_ews = c->null_ews.as<float>();
    A.null_d = c->null_d.as<float>();
    A.Nu = g.Nu;
    A.T_all = g.T; A.Ts = pl.Ts; A.seg = pl.seg; A.Pc = Pc; A.perm_begin = perm_begin;
    A.P_stride = P_stride; A.p_off = p_off; A.tab_cap = pl.tab_cap; A.cum_smem = pl.cum_smem;
    feistel_params(g, seed, rounds, &A.fp);
#define PCT_LAUNCH_NULL(LT, NR, SL)                                                          \
    do {                                                                                     \
        ensure_dyn_smem(c, k_null_fused<LT, NR, SL>, pl.smem);                               \
        k_null_fused<LT, NR, SL><<<grid, pl.Kp, pl.smem, c->stream>>>(A);                    \
    } while (0)
#define PCT_LAUNCH_NULL_R(LT, SL)                                                            \
    do {                                                                                     \
        if (rounds == PCTSEA_DEFAULT_FEISTEL_ROUNDS) PCT_LAUNCH_NULL(LT, PCTSEA_DEFAULT_FEISTEL_ROUNDS, SL); \
        else PCT_LAUNCH_NULL(LT, 0, SL);                                                     \
    } while (0)
    if (pl.word_bytes == 8) { if (sliced) PCT_LAUNCH_NULL_R(uint16_t, true); else PCT_LAUNCH_NULL_R(uint16_t, false); }
    else                  { if (sliced) PCT_LAUNCH_NULL_R(uint8_t, true); else PCT_LAUNCH_NULL_R(uint8_t, false); }
#undef PCT_LAUNCH_NULL_R
#undef PCT_LAUNCH_NULL
    count_launch(c);
    PCT_CUDA(cudaGetLastError());
}

}  // namespace pctsea
