// null_fused.cu — the label-permutation null of stage 3 in FAST arithmetic: ONE kernel per call that generates the
// permuted labels and evaluates the weighted-KS supremum of every cell type, permutation by permutation.
//
// Replaces (reference paths under pctsea-core/src/main/java/edu/scripps/yates/pctsea/):
//   a9  PCTSEA.java:1390-1443 (permutation driver: Collections.shuffle of the ranked labels, P times)
//   a8  utils/parallel/EnrichmentWeigthedScoreParallel.java:104-227 evaluated once per (type, permutation)
//
// Permutation (PCTSEA_PERM_PHILOX): position x of permutation p carries the class at index pi_p(x) of the
// TYPE-SORTED arrangement of the ranked labels (a uniform permutation of any fixed arrangement of the same multiset
// is a uniform arrangement). pi_p is a keyed Feistel bijection of [0,Nu) on Z_a x Z_b (a = ceil(sqrt Nu),
// b = ceil(Nu/a), cycle-walking for the < a out-of-range points) with ARITHMETIC round functions
//   F(v, key, mod) = umulhi(mix(v * M1 + key), mod),   one 32-bit Philox4x32-10 key per round and permutation,
// so it lives entirely in registers: no tables, no shared memory, and "class of index j" is a look-up in a coarse
// bucket table of the T+1 cumulative class counts. The round that is applied LAST modifies the high digit (the one
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
// class of an index of the type-sorted arrangement
// ================================================================================================
// cum[c] = number of ranked cells of a class below c (c = 0 .. Tc, Tc = T + 1 classes, the last one "unknown type");
// class(j) = the c with cum[c] <= j < cum[c+1]. ctab[j >> shift] holds, per bucket of 2^shift indices,
//   .x = index of the bucket's only class boundary (0xffffffff: none inside), .y = class of the bucket's first index + 1,
// so that class(j) = .y - (j < .x): one 8-byte shared-memory read and a subtract with borrow. A bucket with several
// boundaries (or one that skips empty classes) points to a second-level table instead: .y = CT_SUB | slot, and
// sub[slot][(j >> shift2) & (CT_NSUB - 1)] is an entry of the same form for one of the bucket's CT_NSUB sub-buckets
// (real cell-type size distributions are long-tailed: config B's ranking has 100 types with a median of 87 cells
// next to buckets of 256 indices, so most words of four labels touch such a bucket; walking cum[] for them cost 18 %
// more warp instructions than a ranking without small types). A sub-bucket that still holds several boundaries
// (types of a dozen cells, empty types) has bit 31 of .y set: walk cum[] from the class.
constexpr uint32_t CT_MULTI = 0x80000000u;
constexpr uint32_t CT_SUB = 0x40000000u;
constexpr int CT_NSUB_LOG2 = 5, CT_NSUB = 1 << CT_NSUB_LOG2;

// Global loads with a cache policy. The kernel streams its fixed-point scores and parked labels through L1 (every
// byte is used once per walk), while the few KB of second-level class-table entries are touched by some lane of
// almost every word: the streams do not allocate in L1, the table entries are kept there (before this the table
// loads missed L1 — 3.6 % hit rate — and the first use of a class entry was the hottest stall of the kernel).
__device__ __forceinline__ int4 ld_stream(const int4* p) {
    int4 v;
    asm("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) {   // the CTA's own scratch: ordinary (coherent) load
    uint32_t v;
    asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint2 ld_stream(const uint2* p) {
    uint2 v;
    asm volatile("ld.global.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint2 ld_keep(const uint2* p) {
    uint2 v;
    asm("ld.global.nc.L1::evict_last.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ uint32_t class_search(const uint32_t* __restrict__ cum, uint32_t Tc, uint32_t j) {
    uint32_t lo = 0, hi = Tc;   // invariant: cum[lo] <= j < cum[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cum[mid] <= j) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint2 class_entry(const uint32_t* __restrict__ cum, uint32_t Tc, long long start, long long width,
                                             long long Nu, bool* several) {
    uint2 e = make_uint2(0xffffffffu, 1u);
    *several = false;
    if (start < Nu) {
        const uint32_t last = (uint32_t)(min(start + width, Nu) - 1);
        const uint32_t ca = class_search(cum, Tc, (uint32_t)start), cb = class_search(cum, Tc, last);
        e.y = ca + 1u;   // indices below .x borrow: class ca; the others: ca + 1
        if (cb == ca + 1u) e.x = cum[cb];
        else if (cb > ca + 1u) { e.y = (ca + 1u) | CT_MULTI; *several = true; }
    }
    return e;
}

__global__ void __launch_bounds__(1024) k_class_tables(const int32_t* __restrict__ class_cnt, int32_t Tc, int64_t Nu,
                                                      int shift, int32_t nb, uint32_t* __restrict__ cum,
                                                      uint2* __restrict__ ctab, uint2* __restrict__ sub) {
    // one CTA: exclusive scan of the class counts (thread t owns a contiguous chunk), then the bucket table
    __shared__ uint32_t part[1024];
    const int per = (Tc + 1023) / 1024;
    const int c0 = threadIdx.x * per, c1 = min(c0 + per, (int)Tc);
    uint32_t s = 0;
    for (int c = c0; c < c1; c++) s += (uint32_t)class_cnt[c];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 1024; i++) { const uint32_t v = part[i]; part[i] = run; run += v; }
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (int c = c0; c < c1; c++) { cum[c] = run; run += (uint32_t)class_cnt[c]; }
    if (threadIdx.x == 0) cum[Tc] = (uint32_t)Nu;
    __syncthreads();
    // the binary searches below run on a shared-memory copy of the cumulative counts when it fits (it does up to
    // 2047 classes): ~70 dependent look-ups per thread, and this single CTA sits on the critical path of every analysis
    __shared__ uint32_t s_cum[2048];
    if (Tc + 1 <= 2048) {
        for (int c = threadIdx.x; c <= Tc; c += 1024) s_cum[c] = cum[c];
        __syncthreads();
        cum = s_cum;
    }
    // first level: nb <= 1024 buckets, one per thread; the buckets with several boundaries get a slot of the second level
    const int i = threadIdx.x;
    bool several = false;
    uint2 e = make_uint2(0xffffffffu, 1u);
    if (i < nb) e = class_entry(cum, (uint32_t)Tc, (long long)i << shift, 1ll << shift, Nu, &several);
    part[i] = several ? 1u : 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t r2 = 0;
        for (int k = 0; k < 1024; k++) { const uint32_t v = part[k]; part[k] = r2; r2 += v; }
    }
    __syncthreads();
    __shared__ int s_multi[1024];   // bucket of every second-level slot
    __shared__ int s_nmulti;
    if (i < nb) {
        if (several) {
            s_multi[part[i]] = i;
            e = make_uint2(0xffffffffu, CT_SUB | part[i]);
        }
        ctab[i] = e;
    }
    if (i == 1023) s_nmulti = (int)part[1023] + (several ? 1 : 0);
    __syncthreads();
    const int sub_log2 = min(CT_NSUB_LOG2, shift);   // a bucket of fewer than CT_NSUB indices: one index per sub-bucket
    const int shift2 = shift - sub_log2, nsub = 1 << sub_log2;
    for (int w = threadIdx.x; w < s_nmulti * nsub; w += 1024) {   // one sub-bucket per thread
        const int slot = w >> sub_log2, k = w & (nsub - 1), bkt = s_multi[slot];
        bool sev2;
        sub[(size_t)slot * CT_NSUB + k] = class_entry(cum, (uint32_t)Tc, ((long long)bkt << shift) + ((long long)k << shift2),
                                                      1ll << shift2, Nu, &sev2);
    }
}

// Classes of four indices; the several-boundaries case is checked once for the four. ctab_saddr = address of the
// bucket table in the shared window.
__device__ __forceinline__ void class_of4(const uint32_t (&j)[4], uint32_t ctab_saddr, const uint2* __restrict__ sub,
                                          const uint32_t* __restrict__ cum, int shift, int shift2, uint32_t sub_mask, uint32_t (&c)[4]) {
    uint32_t ex[4], ey[4];
#pragma unroll
    for (int q = 0; q < 4; q++)
        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ex[q]), "=r"(ey[q]) : "r"(ctab_saddr + ((j[q] >> shift) << 3)));
    if ((ey[0] | ey[1] | ey[2] | ey[3]) & CT_SUB) {   // some index of the four sits in a bucket with several boundaries
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (ey[q] & CT_SUB) {
                const uint2 e2 = ld_keep(sub + (((ey[q] & (CT_SUB - 1u)) << CT_NSUB_LOG2) + ((j[q] >> shift2) & sub_mask)));
                ex[q] = e2.x; ey[q] = e2.y;
            }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        // c = ey - (j < ex): .y holds the class of the bucket's first index PLUS ONE; the borrow of j - ex takes it back
        asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\tsubc.u32 %0, %3, 0;\n\t}" : "=r"(c[q]) : "r"(j[q]), "r"(ex[q]), "r"(ey[q]));
    }
    if ((c[0] | c[1] | c[2] | c[3]) & CT_MULTI) {
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (c[q] & CT_MULTI) {
                uint32_t cc = c[q] & ~CT_MULTI;
                while (j[q] >= cum[cc + 1]) cc++;   // ends: cum[Tc] = Nu > j
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

__host__ __device__ __forceinline__ uint32_t feistel_mix(uint32_t v, uint32_t key) {
    uint32_t t = v * 0x9E3779B1u + key;
    t ^= t >> 15;
    t *= 0x2C1B3C6Du;
    return t;
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
                                                         LabT* __restrict__ labels) {
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
        cls = class_search(cum, (uint32_t)T + 1u, j);
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
#pragma unroll
    for (int q = 0; q < 4; q++) { ad[q] = col_saddr + t[q] * stride_bytes; bq[q] = lds_i32(ad[q]); }
    aq[0] = bq[0] + f[0];
    bq[1] = (t[1] == t[0]) ? aq[0] : bq[1];
    aq[1] = bq[1] + f[1];
    bq[2] = (t[2] == t[1]) ? aq[1] : ((t[2] == t[0]) ? aq[0] : bq[2]);
    aq[2] = bq[2] + f[2];
    bq[3] = (t[3] == t[2]) ? aq[2] : ((t[3] == t[1]) ? aq[1] : ((t[3] == t[0]) ? aq[0] : bq[3]));
    aq[3] = bq[3] + f[3];
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
    const uint32_t* cum;         // [Tc + 1]
    const uint2* ctab;           // [nb]
    const uint2* sub;            // [<= nb][CT_NSUB] second level of the buckets with several class boundaries
    void* scratch;               // [grid][seg4][Kp] label words
    int* ticket;
    float* null_ews;             // [T][P_stride]
    float* null_d;
    int64_t Nu;
    int32_t T_all, Ts, seg, Pc, perm_begin, P_stride, p_off, nb, shift, cum_smem;
    FeistelParams fp;
};

// SLICED: more cell types than the per-thread shared-memory state can hold at once. The labels are then generated
// in a pass of their own, and phases 1-4 run once per slice of Ts types on the parked labels, with every other type
// mapped to the slice's dummy column (Ts), which takes the same read-modify-write but has zero constants, so it can
// never set a maximum.
template <typename LabT, int NR, bool SLICED>
__global__ void __launch_bounds__(512, 1) k_null_fused(const NullArgs A) {
    typedef typename LaneWord<LabT>::type Word;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_p;
    __shared__ uint32_t s_keys[32];
    const int Ts = A.Ts, Tc = Ts + 1, T_all = A.T_all;
    const int Kp = blockDim.x;
    const int k = threadIdx.x, lane = k & 31, warp = k >> 5, nwarp = Kp >> 5;
    int* numA = reinterpret_cast<int*>(smem_raw);                               // [Tc][Kp]
    TypeConst* cb = reinterpret_cast<TypeConst*>(numA + (size_t)Tc * Kp);       // [Tc]
    uint2* ctab_s = reinterpret_cast<uint2*>(cb + Tc);                          // [nb]
    uint32_t* cum_s = reinterpret_cast<uint32_t*>(ctab_s + A.nb);               // [T_all + 2] when it fits
    const uint32_t* cum = A.cum_smem ? cum_s : A.cum;
    const uint32_t cb_saddr = (uint32_t)__cvta_generic_to_shared(cb);
    const uint32_t ctab_saddr = (uint32_t)__cvta_generic_to_shared(ctab_s);
    const uint32_t col_saddr = (uint32_t)__cvta_generic_to_shared(numA) + 4u * (uint32_t)threadIdx.x;
    const uint32_t Kp4 = 4u * (uint32_t)blockDim.x;
    for (int i = k; i < A.nb; i += Kp) ctab_s[i] = A.ctab[i];
    if (A.cum_smem) for (int i = k; i < T_all + 2; i += Kp) cum_s[i] = A.cum[i];
    const int seg = A.seg;
    const uint32_t n = (uint32_t)A.Nu, a = A.fp.a, b = A.fp.b;
    const int shift = A.shift, shift2 = A.shift - min(CT_NSUB_LOG2, A.shift);
    const uint32_t sub_mask = (1u << min(CT_NSUB_LOG2, A.shift)) - 1u;   // a bucket narrower than CT_NSUB indices has one index per sub-bucket
    Word* lab = reinterpret_cast<Word*>(A.scratch) + (size_t)blockIdx.x * (seg >> 2) * Kp + k;
    const int4* fxp = A.fxT4 + k;
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
        if (!SLICED) {
            for (int t = 0; t < Tc; t++) col[t * Kp] = 0;
            for (int t = k; t < Tc; t += Kp) cb[t].best = 0u;
        }

        // ---- phase 1: labels of my segment + per-type segment sums (one shared-memory read-modify-write per
        // position, in program order). Four positions at a time: four independent round chains, no per-position tests.
        {
            uint32_t L = L00, R = R00;
            uint32_t idx = 0;   // i4 * Kp: the same index addresses the label word and the score quad of (i4, k)
            int4 fn = nw > 0 ? ld_stream(fxp) : make_int4(0, 0, 0, 0);
            for (int i4 = 0; i4 < nw; i4++) {
                const int4 fc = fn;
                if (i4 + 1 < nw) fn = ld_stream(fxp + (idx + (uint32_t)Kp));
                const int fq[4] = { fc.x, fc.y, fc.z, fc.w };
                uint32_t y[4], cls[4];
                if (i4 < nfull) {
                    uint32_t Lq[4], Rq[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        // digits of position x + q, branch-free across a row boundary (lanes cross rows at different words)
                        const uint32_t r = R + (uint32_t)q;
                        Lq[q] = L + (r >= b ? 1u : 0u);
                        Rq[q] = min(r, r - b);
                        rounds(Lq[q], Rq[q]);
                        y[q] = Lq[q] * b + Rq[q];
                    }
                    R += 4u;
                    if (R >= b) { R -= b; L++; }
                    if (max(max(y[0], y[1]), max(y[2], y[3])) >= n) {   // cycle-walk: rare
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            while (y[q] >= n) { rounds(Lq[q], Rq[q]); y[q] = Lq[q] * b + Rq[q]; }
                    }
                    class_of4(y, ctab_saddr, A.sub, cum, shift, shift2, sub_mask, cls);
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
                    class_of4(jj, ctab_saddr, A.sub, cum, shift, shift2, sub_mask, cls);
#pragma unroll
                    for (int q = 0; q < 4; q++) cls[q] = (y[q] == 0xffffffffu) ? (uint32_t)T_all : cls[q];
                }
                if (!SLICED) {
                    int bq[4], aq[4];
                    rmw4(col_saddr, Kp4, cls, fq, bq, aq);
                }
                lab[idx] = LaneWord<LabT>::pack(cls);
                idx += (uint32_t)Kp;
            }
            // pad my words to a whole number of NF_PRE groups (unknown class, zero scores): the second walk has no tail
            const uint32_t dummy[4] = { (uint32_t)T_all, (uint32_t)T_all, (uint32_t)T_all, (uint32_t)T_all };
            for (int i4 = nw; i4 < nwp; i4++) { lab[idx] = LaneWord<LabT>::pack(dummy); idx += (uint32_t)Kp; }
        }

      for (int t0 = 0; t0 < T_all; t0 += Ts) {   // one iteration unless SLICED
        const int T = min(Ts, T_all - t0);        // real types of this slice: local ids [0, T), dummy = Ts
        // local type id of a label: slice members keep their offset, everything else goes to the dummy column
        auto local_type = [&](uint32_t t) -> uint32_t {
            if (!SLICED) return t;   // labels are already in [0, T_all] with T_all = Ts the padding / unknown class
            const uint32_t tl = t - (uint32_t)t0;
            return tl < (uint32_t)T ? tl : (uint32_t)Ts;
        };
        if (SLICED) {
            for (int t = 0; t < Tc; t++) col[t * Kp] = 0;
            for (int t = k; t < Tc; t += Kp) cb[t].best = 0u;
            // phase 1 on the parked labels (each thread re-reads what it wrote itself: no barrier needed)
            for (int i4 = 0; i4 < nw; i4++) {
                const Word w = ld_stream(lab + (uint32_t)i4 * (uint32_t)Kp);
                const int4 fc = ld_stream(fxp + (uint32_t)i4 * (uint32_t)Kp);
                const int fq[4] = { fc.x, fc.y, fc.z, fc.w };
                uint32_t t[4];
                int bq[4], aq[4];
#pragma unroll
                for (int q = 0; q < 4; q++) t[q] = local_type(LaneWord<LabT>::get(w, q));
                rmw4(col_saddr, Kp4, t, fq, bq, aq);
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
                if (t < T && dA != 0 && dB != 0) {
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
        {
            int S = S0;
            float Sf = __int2float_rn(S);
            Word wc[NF_PRE];   // register sets refilled in place: the loads run NF_PRE - 1 words ahead of their use
            int4 fc[NF_PRE];
            const uint32_t Kpu = (uint32_t)Kp;
            uint32_t idx = 0;   // index of the group's first word
            if (nwp > 0) {
#pragma unroll
                for (int u = 0; u < NF_PRE; u++) { wc[u] = ld_stream(lab + (idx + u * Kpu)); fc[u] = ld_stream(fxp + (idx + u * Kpu)); }
            }
            for (int i4 = 0; i4 < nwp; i4 += NF_PRE) {
                const bool more = i4 + NF_PRE < nwp;
                idx += NF_PRE * Kpu;
#pragma unroll
                for (int u = 0; u < NF_PRE; u++) {
                    const int fq[4] = { fc[u].x, fc[u].y, fc[u].z, fc[u].w };
                    const Word wcu = wc[u];
                    if (more) { wc[u] = ld_stream(lab + (idx + u * Kpu)); fc[u] = ld_stream(fxp + (idx + u * Kpu)); }
                    uint32_t t[4];
                    int bq[4], aq[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) t[q] = local_type(LaneWord<LabT>::get(wcu, q));
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
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const float c12 = __uint_as_float(cq[q].x), c2 = __uint_as_float(cq[q].y), top = __uint_as_float(cq[q].w);
                        const float Spf = Sf;
                        S += fq[q];   // exact: |S| <= sum |fx| < 2^31
                        Sf = __int2float_rn(S);
                        d1[q] = __fmaf_rn(__int2float_rn(aq[q]), c12, -__fmul_rn(Sf, c2));
                        d0[q] = __fmaf_rn(__int2float_rn(bq[q]), c12, -__fmul_rn(Spf, c2));
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
        for (int t = k; t < T; t += Kp) {
            const int32_t nk = A.class_cnt[t0 + t];
            const int32_t sizeb = (int32_t)(A.Nu - nk);
            float v, d;
            if (!(nk > 1 && sizeb > 1)) { v = __int_as_float(0x7fc00000); d = v; }
            else {
                const uint32_t key = cb[t].best;
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
    auto fixed = [&](int ts) { return (size_t)(ts + 1) * 16 + (size_t)pl.nb * 8 + (pl.cum_smem ? (size_t)(g.Tc + 1) * 4 : 0) + 64; };
    // threads per CTA when `ctas` CTAs share an SM (the per-CTA budget includes the 1 KB the driver reserves)
    auto threads_for_n = [&](int ts, int ctas) {
        const size_t f = fixed(ts);
        const size_t budget = std::min(smem_cta_max, smem_sm / (size_t)ctas - 1024);
        if (f + (size_t)(ts + 1) * 4 * 32 > budget) return 0;
        const int kp = (int)std::min<size_t>(512, (budget - f) / ((size_t)(ts + 1) * 4));
        return (kp / 32) * 32;
    };
    // Two co-resident CTAs of half the size beat one large CTA at the same warp count (their serial phases 2 and 4
    // hide behind the other CTA's walks: 0.97 -> 0.87 ms on config B), so among the geometries with the most threads
    // per SM the one with more CTAs wins, down to 128 threads per CTA.
    auto threads_for = [&](int ts) {
        int best = threads_for_n(ts, 1), best_total = best;
        const int max_ctas = c->opt_null_ctas_per_sm > 0 ? c->opt_null_ctas_per_sm : 2;
        for (int n = 2; n <= max_ctas; n++) {
            const int kp = threads_for_n(ts, n);
            if (kp >= 128 && kp * n >= best_total) { best = kp; best_total = kp * n; }
        }
        return best;
    };
    int ts = std::max(g.T, 1), kp = threads_for(ts);
    const int min_kp = c->opt_null_min_threads > 0 ? c->opt_null_min_threads : 192;
    if (kp < min_kp) {
        // slices of equal size, as few as give every slice at least 256 threads (or the maximum that fits)
        int n_slices = 2;
        for (;; n_slices++) {
            ts = (int)div_up(g.T, n_slices);
            kp = threads_for(ts);
            if (kp >= 256 || ts == 1) break;
        }
        PCT_REQUIRE(kp >= 32, "shared memory too small for the null kernel's per-type state");
    }
    if (c->opt_null_threads > 0) kp = std::min(kp, (int)round_up(c->opt_null_threads, 32));
    // short rankings: at least ~8 positions per thread
    const int64_t want = std::max<int64_t>(32, round_up(div_up(std::max<int64_t>(g.Nu, 1), 8), 32));
    kp = (int)std::min<int64_t>(kp, want);
    pl.Kp = kp;
    pl.Ts = ts;
    pl.seg = (int32_t)round_up(div_up(std::max<int64_t>(g.Nu, 1), kp), 4 * NF_PRE);
    pl.smem = (size_t)(ts + 1) * kp * 4 + fixed(ts);
    int per_sm = (int)std::max<size_t>(1, smem_sm / (pl.smem + 1024));
    per_sm = std::min(per_sm, std::max(1, 2048 / kp));
    per_sm = std::min(per_sm, 4);
    if (c->opt_null_ctas_per_sm > 0) per_sm = std::min(per_sm, c->opt_null_ctas_per_sm);
    pl.ctas_per_sm = per_sm;
    pl.word_bytes = 4 * g.lab_bytes;
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

// cum[] (+ the bucket table when nb > 0) of the type-sorted arrangement; needs class_cnt (k_gather_ranked).
static size_t class_tab_cum_bytes(const EnrichGeom& g) { return (size_t)round_up((int64_t)(g.Tc + 1) * 4, 16); }
static size_t class_tab_sub_offset(const EnrichGeom& g, int32_t nb) { return class_tab_cum_bytes(g) + (size_t)round_up((int64_t)std::max(nb, 1) * 8, 16); }

void launch_class_tables(pctsea_ctx* c, const EnrichGeom& g, int shift, int32_t nb) {
    const size_t cum_bytes = class_tab_cum_bytes(g), sub_off = class_tab_sub_offset(g, nb);
    c->cls_tab.reserve(sub_off + (size_t)std::max(nb, 1) * CT_NSUB * 8 + 64);
    uint32_t* cum = c->cls_tab.as<uint32_t>();
    uint2* ctab = reinterpret_cast<uint2*>(c->cls_tab.as<unsigned char>() + cum_bytes);
    uint2* sub = reinterpret_cast<uint2*>(c->cls_tab.as<unsigned char>() + sub_off);
    k_class_tables<<<1, 1024, 0, c->stream>>>(c->class_cnt.as<int32_t>(), g.Tc, g.Nu, shift, nb, cum, ctab, sub);
    count_launch(c);
    PCT_CUDA(cudaGetLastError());
}

void launch_labels_philox_rm(pctsea_ctx* c, const EnrichGeom& g, uint64_t seed, int32_t perm_begin, int32_t Pc, int rounds) {
    if (Pc == 0 || g.Nu == 0) return;
    FeistelParams fp;
    feistel_params(g, seed, rounds, &fp);
    const uint32_t* cum = c->cls_tab.as<uint32_t>();
    for (int32_t p0 = 0; p0 < Pc; p0 += 32768) {
        const int32_t pc = std::min<int32_t>(32768, Pc - p0);
        dim3 grid((unsigned)div_up(g.Nu_pad, 256), (unsigned)pc);
        if (g.lab_bytes == 1)
            k_labels_philox_rm<uint8_t><<<grid, 256, 0, c->stream>>>(g.Nu, g.Nu_pad, g.T, fp, perm_begin + p0, cum,
                                                                     c->labels.as<uint8_t>() + (size_t)p0 * g.Nu_pad);
        else
            k_labels_philox_rm<uint16_t><<<grid, 256, 0, c->stream>>>(g.Nu, g.Nu_pad, g.T, fp, perm_begin + p0, cum,
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
    A.cum = c->cls_tab.as<uint32_t>();
    A.ctab = reinterpret_cast<const uint2*>(c->cls_tab.as<unsigned char>() + class_tab_cum_bytes(g));
    A.sub = reinterpret_cast<const uint2*>(c->cls_tab.as<unsigned char>() + class_tab_sub_offset(g, pl.nb));
    const unsigned grid = (unsigned)std::min<int64_t>(Pc, (int64_t)pl.ctas_per_sm * c->num_sms);
    const int seg4 = pl.seg >> 2;
    c->labels.reserve((size_t)grid * seg4 * pl.Kp * pl.word_bytes + 64);
    A.scratch = c->labels.p;
    c->lab_ticket.reserve(64);
    PCT_CUDA(cudaMemsetAsync(c->lab_ticket.p, 0, 4, c->stream));
    A.ticket = c->lab_ticket.as<int>();
    A.null_ews = c->null_ews.as<float>();
    A.null_d = c->null_d.as<float>();
    A.Nu = g.Nu;
    A.T_all = g.T; A.Ts = pl.Ts; A.seg = pl.seg; A.Pc = Pc; A.perm_begin = perm_begin;
    A.P_stride = P_stride; A.p_off = p_off; A.nb = pl.nb; A.shift = pl.shift; A.cum_smem = pl.cum_smem;
    feistel_params(g, seed, rounds, &A.fp);
    const bool sliced = pl.Ts < g.T;
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
    if (g.lab_bytes == 2) { if (sliced) PCT_LAUNCH_NULL_R(uint16_t, true); else PCT_LAUNCH_NULL_R(uint16_t, false); }
    else                  { if (sliced) PCT_LAUNCH_NULL_R(uint8_t, true); else PCT_LAUNCH_NULL_R(uint8_t, false); }
#undef PCT_LAUNCH_NULL_R
#undef PCT_LAUNCH_NULL
    count_launch(c);
    PCT_CUDA(cudaGetLastError());
}

}  // namespace pctsea
