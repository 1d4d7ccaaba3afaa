// enrich.cu — stage 3: the ranked arrays, the observed weighted Kolmogorov-Smirnov enrichment score per cell type
// (reference arithmetic, bit-exact) and the validation-mode (EXACT) null. The production null (FAST arithmetic,
// generated permutations) is null_fused.cu.
//
// Replaces (reference paths under pctsea-core/src/main/java/edu/scripps/yates/pctsea/):
//   a8  utils/parallel/EnrichmentWeigthedScoreParallel.java:72-353 (run), :492-512 (lookForPriorNegativeSupremum)
//   a9  PCTSEA.java:1390-1443 (permutation driver) in validation mode: replayed index arrays, sequential arithmetic
//
// Data layout in HBM (all in rank order, Nu = Np - 1 cells):
//   s_rank  fp64 [Nu]            ranked scores
//   lab_rank LabT [Nu_pad]       observed class per rank (class T = "unknown type", label -1)
//   labels  LabT [Pc][Nu_pad]    permuted classes of the EXACT null, one row per permutation (u8 if T < 256 else u16)
//
// EXACT null: one thread per (type, permutation) replays the reference's sequential float/double recurrence bit for
// bit (lanes = types; scores and labels are warp-broadcast loads) — O(T * Nu * P), validation only.
#include "common.cuh"

namespace pctsea {

// ================================================================================================
// ranked arrays
// ================================================================================================
template <typename LabT>
__global__ void k_gather_ranked(const double* __restrict__ score, const int32_t* __restrict__ label,
                                const int32_t* __restrict__ order, int64_t Nu, int64_t Nu_pad, int32_t T,
                                double* __restrict__ s_rank, LabT* __restrict__ lab_rank,
                                int32_t* __restrict__ class_cnt, unsigned long long* __restrict__ maxabs_bits,
                                int32_t hist_smem_n) {
    // shared histogram of classes when it fits (hist_n = T+1), else global atomics (hist_n = 0)
    extern __shared__ int32_t shist[];
    const int hist_n = (int)hist_smem_n;
    for (int j = threadIdx.x; j < hist_n; j += blockDim.x) shist[j] = 0;
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long ab = 0ull;
    if (i < Nu) {
        int32_t idx = order[i];
        double s = score[idx];
        if (s != s) s = 0.0;  // PCTSEA.java:2461-2465 (NaN -> 0; a no-op after the threshold)
        int32_t lab = label[idx];
        int32_t cls = (lab < 0 || lab >= T) ? T : lab;
        s_rank[i] = s;
        lab_rank[i] = (LabT)cls;
        if (hist_n) atomicAdd(&shist[cls], 1); else atomicAdd(&class_cnt[cls], 1);
        ab = (unsigned long long)__double_as_longlong(fabs(s));
    } else if (i < Nu_pad) {
        lab_rank[i] = (LabT)T;
    }
    // warp-reduce the max before touching the global atomic (all lanes participate)
    unsigned long long m = ab;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long u = __shfl_xor_sync(0xffffffffu, m, o); m = u > m ? u : m; }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(maxabs_bits, m);
    __syncthreads();
    for (int j = threadIdx.x; j < hist_n; j += blockDim.x)
        if (shist[j]) atomicAdd(&class_cnt[j], shist[j]);
}

void launch_prepare_ranked(pctsea_ctx* c, int64_t N, const double* score_dev, const int32_t* label_dev,
                           const int32_t* order_dev, EnrichGeom& g) {
    (void)N;
    g.Tc = g.T + 1;
    g.lab_bytes = (g.Tc <= 256) ? 1 : 2;
    c->last_lab_bytes = g.lab_bytes;
    g.Nu_pad = round_up(std::max<int64_t>(g.Nu, 1), 128);
    c->s_rank.reserve((size_t)g.Nu_pad * 8 + 64);
    c->lab_rank.reserve((size_t)g.Nu_pad * g.lab_bytes + 64);
    c->class_cnt.reserve((size_t)g.Tc * 4 + 64);
    c->counters.reserve(64);
    PCT_CUDA(cudaMemsetAsync(c->class_cnt.p, 0, (size_t)g.Tc * 4, c->stream));
    PCT_CUDA(cudaMemsetAsync(c->counters.p, 0, 64, c->stream));
    unsigned long long* maxabs = c->counters.as<unsigned long long>() + 2;
    int32_t* frac_dev = reinterpret_cast<int32_t*>(c->counters.as<unsigned long long>() + 3);
    unsigned grid = (unsigned)div_up(g.Nu_pad, 256);
    const int32_t hist_n = (g.Tc <= 8192) ? g.Tc : 0;
    const size_t hist_smem = (size_t)hist_n * 4;
    if (g.lab_bytes == 1)
        k_gather_ranked<uint8_t><<<grid, 256, hist_smem, c->stream>>>(score_dev, label_dev, order_dev, g.Nu, g.Nu_pad,
                                                                      g.T, c->s_rank.as<double>(),
                                                                      c->lab_rank.as<uint8_t>(),
                                                                      c->class_cnt.as<int32_t>(), maxabs, hist_n);
    else
        k_gather_ranked<uint16_t><<<grid, 256, hist_smem, c->stream>>>(score_dev, label_dev, order_dev, g.Nu, g.Nu_pad,
                                                                       g.T, c->s_rank.as<double>(),
                                                                       c->lab_rank.as<uint16_t>(),
                                                                       c->class_cnt.as<int32_t>(), maxabs, hist_n);
    count_launch(c);
    PCT_CUDA(cudaGetLastError());
}

// ================================================================================================
// EXACT KS: lanes = types
// ================================================================================================
__device__ __forceinline__ float dstat_of(float sup, int32_t sizea, int32_t sizeb) {
    // EnrichmentWeigthedScoreParallel.java:203-207: sizea*sizeb is an int32 multiply (wraps)
    int32_t prod = (int32_t)((uint32_t)sizea * (uint32_t)sizeb);
    int32_t sum = (int32_t)((uint32_t)sizea + (uint32_t)sizeb);
    float sq = (float)__dsqrt_rn(__ddiv_rn((double)prod, (double)sum));
    return __fmul_rn(sup, sq);
}

template <typename LabT> struct LabVec;
template <> struct LabVec<uint8_t> {
    static constexpr int N = 16;
    __device__ static void load(const uint8_t* p, int out[16]) {
        uint4 v = *reinterpret_cast<const uint4*>(p);
        uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int i = 0; i < 16; i++) out[i] = (int)((w[i >> 2] >> (8 * (i & 3))) & 0xffu);
    }
};
template <> struct LabVec<uint16_t> {
    static constexpr int N = 8;
    __device__ static void load(const uint16_t* p, int out[8]) {
        uint4 v = *reinterpret_cast<const uint4*>(p);
        uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int i = 0; i < 8; i++) out[i] = (int)((w[i >> 1] >> (16 * (i & 1))) & 0xffffu);
    }
};

// Null, EXACT arithmetic: one thread = one (type, permutation); replays the reference's two passes
// (denominators :104-117, running sums :140-194) with its float accumulators and double adds.
template <typename LabT>
__global__ void __launch_bounds__(128) k_ks_exact(const double* __restrict__ s_rank, const LabT* __restrict__ labels,
                                                  int64_t Nu, int64_t Nu_pad, int32_t T,
                                                  float* __restrict__ null_ews, float* __restrict__ null_d,
                                                  int32_t P_stride, int32_t p_off) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    const int p = blockIdx.y;
    if (t >= T) return;
    const LabT* lab = labels + (size_t)p * Nu_pad;
    constexpr int V = LabVec<LabT>::N;
    const float qnan = __int_as_float(0x7fc00000);

    float denomA = 0.0f, denomB = 0.0f;
    int32_t nk = 0;
    for (int64_t i0 = 0; i0 < Nu; i0 += V) {
        int lv[V];
        LabVec<LabT>::load(lab + i0, lv);
#pragma unroll
        for (int j = 0; j < V; j++) {
            int64_t i = i0 + j;
            if (i < Nu) {
                const double s = s_rank[i];
                const bool hit = lv[j] == t;
                const float acc = hit ? denomA : denomB;
                const float r = __double2float_rn(__dadd_rn((double)acc, s));
                denomA = hit ? r : denomA;
                denomB = hit ? denomB : r;
                nk += hit ? 1 : 0;
            }
        }
    }
    const int32_t sizea = nk, sizeb = (int32_t)(Nu - nk);
    if (!((nk > 1) && (sizeb > 1))) {  // :119-127, :195-196
        null_ews[(size_t)t * P_stride + p_off + p] = qnan;
        null_d[(size_t)t * P_stride + p_off + p] = qnan;
        return;
    }
    float sup = 0.0f, a = 0.0f, b = 0.0f, numA = 0.0f, numB = 0.0f;
    for (int64_t i0 = 0; i0 < Nu; i0 += V) {
        int lv[V];
        LabVec<LabT>::load(lab + i0, lv);
#pragma unroll
        for (int j = 0; j < V; j++) {
            int64_t i = i0 + j;
            if (i < Nu) {
                const double s = s_rank[i];
                const bool hit = lv[j] == t;
                const float acc = hit ? numA : numB;
                const float r = __double2float_rn(__dadd_rn((double)acc, s));
                const float q = __fdiv_rn(r, hit ? denomA : denomB);
                numA = hit ? r : numA;
                numB = hit ? numB : r;
                a = hit ? q : a;
                b = hit ? b : q;
                const float diff = __fsub_rn(a, b);
                if (fabsf(diff) > fabsf(sup)) sup = diff;
            }
        }
    }
    null_ews[(size_t)t * P_stride + p_off + p] = sup;  // :315 uncompensated
    null_d[(size_t)t * P_stride + p_off + p] = dstat_of(sup, sizea, sizeb);
}

// ------------------------------------------------------------------------------------------------
// Observed scores, EXACT arithmetic, evaluated in parallel.
//
// The reference's running sums are float accumulators updated with a double add,
//     acc' = (float)((double)acc + s),
// one per position and type (numeratorA over the hits, numeratorB over the misses; the denominators of
// :104-117 are the final values of the same two recurrences). That recurrence looks sequential, but inside
// one float binade it is not: write acc = m * U with U = ulp(acc) = 2^(e-23) and m an integer in
// [2^23, 2^24), and s = q * U. While the sum stays in the binade, (double)acc + s is rounded to the double
// grid U * 2^-29 (independently of m, which is a multiple of that grid), and the float rounding then adds the
// integer r = RN(RN_{2^-29}(q)) to m — again independently of m unless the rounded q sits exactly on a
// half-integer (a tie, resolved by the parity of m). So a run of positions that causes no binade crossing and
// no tie is a plain integer prefix sum of r. k_obs_chain walks the ranked list in blocks of 1024 positions:
// every thread quantises its score under the current binade of its accumulator (A if its cell is a hit, B
// otherwise), a block scan gives the prefix sums, the first "event" in the block (accumulator still zero or
// subnormal, sign change, q >= 2^24, tie, or m + prefix >= 2^24) is executed with the real float/double
// instructions by one thread, and the walk resumes behind it. Events are rare (about two dozen binade
// crossings per accumulator), so the cost is ~Nu/1024 block steps per type instead of Nu dependent steps.
// The result is bit-identical to the sequential recurrence (tests compare against the oracle's loop).
//
//   k_obs_chain  one CTA per type: both accumulators after every position, type-major [T][Nu] float2
//   k_obs_sup    all (position, type) pairs in parallel: a = numA/denomA, b = numB/denomB, diff = a - b,
//                first position of max |diff| (strict '>' of :182-184)
//   k_obs_neg    last negative diff at 0-based index < supX (lookForPriorNegativeSupremum :492-512)
//   k_obs_finish compensation, dStat, normSupX
// "not touched yet" is recorded as -0.0f, which the recurrences can never produce.
// ------------------------------------------------------------------------------------------------
constexpr int OBS_THREADS = 512;
constexpr uint32_t OBS_SAT = 1u << 30;

__device__ __forceinline__ uint32_t sat_add(uint32_t a, uint32_t b) {
    uint32_t c = a + b;
    return c > OBS_SAT ? OBS_SAT : c;
}

// Inclusive block scan (saturating) of two values; `scratch` holds 2 x 32 warp totals.
__device__ __forceinline__ void block_scan2_sat(uint32_t& va, uint32_t& vb, uint32_t* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ua = __shfl_up_sync(0xffffffffu, va, o), ub = __shfl_up_sync(0xffffffffu, vb, o);
        if (lane >= o) { va = sat_add(va, ua); vb = sat_add(vb, ub); }
    }
    if (lane == 31) { scratch[warp] = va; scratch[32 + warp] = vb; }
    __syncthreads();
    if (warp == 0) {
        uint32_t ta = scratch[lane], tb = scratch[32 + lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ua = __shfl_up_sync(0xffffffffu, ta, o), ub = __shfl_up_sync(0xffffffffu, tb, o);
            if (lane >= o) { ta = sat_add(ta, ua); tb = sat_add(tb, ub); }
        }
        scratch[lane] = ta; scratch[32 + lane] = tb;
    }
    __syncthreads();
    if (warp > 0) { va = sat_add(va, scratch[warp - 1]); vb = sat_add(vb, scratch[32 + warp - 1]); }
    // no trailing barrier: the caller passes at least one barrier before `scratch` is written again
}

constexpr int OBS_V = 4;                          // consecutive positions per thread and block step
constexpr int OBS_STEP = OBS_THREADS * OBS_V;     // positions per block step

template <typename LabT>
__global__ void __launch_bounds__(OBS_THREADS) k_obs_chain(const double* __restrict__ s_rank, const LabT* __restrict__ lab,
                                                           int64_t Nu, int32_t Tpad, int32_t t0, float2* __restrict__ chainT,
                                                           float* __restrict__ denom /*[2][Tpad]*/) {
    __shared__ uint32_t scratch[64];
    __shared__ float st_acc[2];   // accumulators A, B
    __shared__ int st_any[2];     // touched yet?
    __shared__ int st_first;
    const int t = t0 + blockIdx.x;   // the history buffer holds the types of one slice: row = t - t0
    const int tid = threadIdx.x;
    float2* out = chainT + (size_t)blockIdx.x * Nu;
    const float neg0 = __int_as_float((int)0x80000000);
    if (tid == 0) { st_acc[0] = 0.0f; st_acc[1] = 0.0f; st_any[0] = 0; st_any[1] = 0; }
    __syncthreads();

    int64_t pos = 0;
    while (pos < Nu) {
        const float accA = st_acc[0], accB = st_acc[1];
        const int anyA = st_any[0], anyB = st_any[1];
        if (tid == 0) st_first = OBS_STEP;
        const uint32_t bitsA = (uint32_t)__float_as_int(accA), bitsB = (uint32_t)__float_as_int(accB);
        const uint32_t mA = (bitsA & 0x7fffffu) | 0x800000u, mB = (bitsB & 0x7fffffu) | 0x800000u;
        bool valid[OBS_V], hit[OBS_V], ev[OBS_V];
        double sv[OBS_V];
        uint32_t ra[OBS_V], rb[OBS_V];   // inclusive in-thread prefixes of the quantised increments
        uint32_t ta = 0, tb = 0;
#pragma unroll
        for (int v = 0; v < OBS_V; v++) {
            const int64_t i = pos + (int64_t)tid * OBS_V + v;
            valid[v] = i < Nu;
            hit[v] = valid[v] && ((int)lab[i] == t);
            sv[v] = valid[v] ? s_rank[i] : 0.0;
            const uint32_t ab = hit[v] ? bitsA : bitsB;
            const uint32_t ex = (ab >> 23) & 0xffu;
            ev[v] = false;
            uint32_t r = 0;
            if (valid[v]) {
                const bool neg_acc = (ab >> 31) != 0u;
                const bool neg_s = __double_as_longlong(sv[v]) < 0;  // sign bit (also -0.0)
                const double as = fabs(sv[v]);
                if (ex == 0u || ex == 255u) ev[v] = true;                   // zero / subnormal / inf / NaN accumulator
                else if (as != 0.0 && (neg_s != neg_acc)) ev[v] = true;     // opposite signs: not a pure accumulation
                else {
                    // q = |s| / ulp(acc) = |s| * 2^(23 - e), e = ex - 127  (exact power-of-two scaling)
                    const double scale = __longlong_as_double((long long)(1023 + 23 - ((int)ex - 127)) << 52);
                    const double q = __dmul_rn(as, scale);
                    if (!(q < 16777216.0)) ev[v] = true;                    // would leave the binade on its own (or NaN)
                    else {
                        const long long Q = __double2ll_rn(__dmul_rn(q, 536870912.0));  // q on the 2^-29 grid, RNE
                        if ((Q & 0x1fffffffLL) == 0x10000000LL) ev[v] = true;           // exactly half-way: parity of m decides
                        else r = (uint32_t)((Q + 0x10000000LL) >> 29);
                    }
                }
            }
            ta = sat_add(ta, (hit[v] && !ev[v]) ? r : 0u);
            tb = sat_add(tb, (valid[v] && !hit[v] && !ev[v]) ? r : 0u);
            ra[v] = ta; rb[v] = tb;
        }
        uint32_t pa = ta, pb = tb;
        block_scan2_sat(pa, pb, scratch);   // inclusive over threads
        const uint32_t ea = pa - ta, eb = pb - tb;   // exclusive (no saturation can have happened below 2^30 when it matters)
        int my_first = OBS_STEP;
#pragma unroll
        for (int v = OBS_V - 1; v >= 0; v--) {
            bool e = ev[v];
            if (valid[v] && !e) e = hit[v] ? (mA + sat_add(ea, ra[v]) >= 0x1000000u) : (mB + sat_add(eb, rb[v]) >= 0x1000000u);
            if (e) my_first = tid * OBS_V + v;
        }
        if (my_first < OBS_STEP) atomicMin(&st_first, my_first);
        __syncthreads();
        const int first = st_first;
        const int64_t remaining = Nu - pos;
        const int n_ok = (int)min((int64_t)first, min((int64_t)OBS_STEP, remaining));
#pragma unroll
        for (int v = 0; v < OBS_V; v++) {
            const int e_idx = tid * OBS_V + v;
            if (e_idx < n_ok) {
                // still inside both binades: accumulator = (m + prefix) * ulp, written straight as float bits
                const uint32_t ba = (bitsA & 0xff800000u) | ((mA + ea + ra[v]) & 0x7fffffu);
                const uint32_t bb = (bitsB & 0xff800000u) | ((mB + eb + rb[v]) & 0x7fffffu);
                const float fa = anyA ? __int_as_float((int)ba) : neg0;
                const float fb = anyB ? __int_as_float((int)bb) : neg0;
                out[pos + e_idx] = make_float2(fa, fb);
                if (e_idx == n_ok - 1) {
                    if (anyA) st_acc[0] = fa;
                    if (anyB) st_acc[1] = fb;
                }
            }
        }
        if (first < OBS_STEP) __syncthreads();   // (uniform) the event below reads the accumulators written just above
        if (first < OBS_STEP && (first / OBS_V) == tid) {
            // the event position: the reference's own instruction sequence on the up-to-date accumulator
            const int v = first % OBS_V;
            double s = sv[0];
            bool h = hit[0];
#pragma unroll
            for (int u = 1; u < OBS_V; u++) if (u == v) { s = sv[u]; h = hit[u]; }
            const int ch = h ? 0 : 1;
            const float cur = st_acc[ch];
            const float nxt = __double2float_rn(__dadd_rn((double)cur, s));
            st_acc[ch] = nxt;
            st_any[ch] = 1;
            const float fa = (ch == 0) ? nxt : (st_any[0] ? st_acc[0] : neg0);
            const float fb = (ch == 1) ? nxt : (st_any[1] ? st_acc[1] : neg0);
            out[pos + first] = make_float2(fa, fb);
        }
        pos += n_ok + ((first < OBS_STEP) ? 1 : 0);
        __syncthreads();
    }
    if (tid == 0) {
        denom[t] = st_acc[0];
        denom[Tpad + t] = st_acc[1];
    }
}

__device__ __forceinline__ float obs_diff(float2 v, float dA, float dB) {
    // a (b) keeps its initial 0.0f until the first hit (miss), :144-169
    const float a = (__float_as_int(v.x) == (int)0x80000000) ? 0.0f : __fdiv_rn(v.x, dA);
    const float b = (__float_as_int(v.y) == (int)0x80000000) ? 0.0f : __fdiv_rn(v.y, dB);
    return __fsub_rn(a, b);
}

constexpr int OBS_TILE = 2048;  // positions per CTA (256 threads x 8)

__global__ void __launch_bounds__(256) k_obs_sup(const float2* __restrict__ chainT, const float* __restrict__ denom,
                                                 const int32_t* __restrict__ class_cnt, int64_t Nu, int32_t Tpad, int32_t t0,
                                                 unsigned long long* __restrict__ supkey) {
    __shared__ unsigned long long wbest[8];
    const int t = t0 + blockIdx.y;
    const float dA = denom[t], dB = denom[Tpad + t];
    const float2* row = chainT + (size_t)blockIdx.y * Nu;
    unsigned long long best = 0ull;
    const int64_t i_beg = (int64_t)blockIdx.x * OBS_TILE, i_end = min(i_beg + OBS_TILE, Nu);
    for (int64_t i = i_beg + threadIdx.x; i < i_end; i += 256) {
        const float d = obs_diff(row[i], dA, dB);
        const uint32_t ab = (uint32_t)__float_as_int(d) & 0x7fffffffu;
        // NaN diffs never win in the reference (|NaN| > x is false)
        if (d == d && ab != 0u) {
            const unsigned long long key = ((unsigned long long)ab << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
            best = key > best ? key : best;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long u = __shfl_xor_sync(0xffffffffu, best, o);
        best = u > best ? u : best;
    }
    if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) best = wbest[w] > best ? wbest[w] : best;
        if (class_cnt[t] > 1 && best) atomicMax(&supkey[t], best);
    }
}

// One pass over [0, supX): the last negative diff at 0-based index < supX (lookForPriorNegativeSupremum
// :492-512) and, for the secondary supremum, lookForSecondaryEnrichmentIndex (:514-533): the largest
// i < supX-1 with diff[i] > 0 and diff[i+1] < 0 (0 = none, also when the only such i is 0).
__global__ void __launch_bounds__(256) k_obs_neg(const float2* __restrict__ chainT, const float* __restrict__ denom,
                                                 int64_t Nu, int32_t Tpad, int32_t t0, const unsigned long long* __restrict__ supkey,
                                                 unsigned int* __restrict__ lastneg /* index+1, 0 = none */,
                                                 unsigned int* __restrict__ secidx) {
    const int t = t0 + blockIdx.y;
    const unsigned long long key = supkey[t];
    if (key == 0ull) return;
    const int64_t supX = (int64_t)(0xffffffffu - (uint32_t)(key & 0xffffffffull)) + 1;  // 1-based
    const int64_t i_beg = (int64_t)blockIdx.x * OBS_TILE, i_end = min(min(i_beg + OBS_TILE, Nu), supX);
    if (i_beg >= i_end) return;
    const float dA = denom[t], dB = denom[Tpad + t];
    const float2* row = chainT + (size_t)blockIdx.y * Nu;
    unsigned int best = 0, bsec = 0;
    for (int64_t i = i_beg + threadIdx.x; i < i_end; i += 256) {
        const float d = obs_diff(row[i], dA, dB);
        if (d < 0.0f) best = (unsigned int)(i + 1);
        if (d > 0.0f && i + 1 < supX - 0 && i < supX - 1 && i + 1 < Nu) {
            const float d1 = obs_diff(row[i + 1], dA, dB);
            if (d1 < 0.0f) bsec = (unsigned int)i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned int u = __shfl_xor_sync(0xffffffffu, best, o), v = __shfl_xor_sync(0xffffffffu, bsec, o);
        best = u > best ? u : best;
        bsec = v > bsec ? v : bsec;
    }
    if ((threadIdx.x & 31) == 0) {
        if (best) atomicMax(&lastneg[t], best);
        if (bsec) atomicMax(&secidx[t], bsec);
    }
}

// ---- observed-only extras ---------------------------------------------------------
// KS D statistic of the type's scores against all other scores = what the reference stores as
// "KSTestPvalue" (:211-212,228) and, cast to float, as the unweighted enrichment score
// (PCTSEA.java:2606-2616). commons-math3 sorts both samples and walks them with ties processed together;
// on the ranked list that is: D = max over tie-group boundaries of |hits*m - misses*n| / (n*m), with
// integer counts (Double.compare equality, so -0.0 and 0.0 are different values).
// Three small kernels: per-tile hit counts of every type, their exclusive scan over the tiles, then every
// (tile, type) pair in parallel.
constexpr int KSD_V = 4;                           // consecutive positions per thread
constexpr int KSD_TILE = OBS_THREADS * KSD_V;      // positions per tile

template <typename LabT>
__global__ void __launch_bounds__(OBS_THREADS) k_obs_tilecnt(const LabT* __restrict__ lab, int64_t Nu, int32_t Tc,
                                                             int32_t* __restrict__ tilecnt /*[tiles][Tc]*/) {
    extern __shared__ int32_t sh_cnt[];
    for (int j = threadIdx.x; j < Tc; j += OBS_THREADS) sh_cnt[j] = 0;
    __syncthreads();
    const int64_t i0 = (int64_t)blockIdx.x * KSD_TILE + (int64_t)threadIdx.x * KSD_V;
#pragma unroll
    for (int v = 0; v < KSD_V; v++)
        if (i0 + v < Nu) atomicAdd(&sh_cnt[(int)lab[i0 + v]], 1);
    __syncthreads();
    for (int j = threadIdx.x; j < Tc; j += OBS_THREADS) tilecnt[(size_t)blockIdx.x * Tc + j] = sh_cnt[j];
}

// exclusive scan over the tiles, one warp per class (32 tiles per step)
__global__ void __launch_bounds__(128) k_obs_tilescan(int32_t* __restrict__ tilecnt, int32_t n_tiles, int32_t Tc) {
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= Tc) return;
    int32_t run = 0;
    for (int32_t k0 = 0; k0 < n_tiles; k0 += 32) {
        const int32_t k = k0 + lane;
        const int32_t v = k < n_tiles ? tilecnt[(size_t)k * Tc + t] : 0;
        int32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (k < n_tiles) tilecnt[(size_t)k * Tc + t] = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

template <typename LabT>
__global__ void __launch_bounds__(OBS_THREADS) k_obs_ksd(const double* __restrict__ s_rank, const LabT* __restrict__ lab,
                                                         int64_t Nu, int32_t Tc, const int32_t* __restrict__ class_cnt,
                                                         const int32_t* __restrict__ tilebase, int32_t t_first,
                                                         unsigned long long* __restrict__ ksd_sup) {
    __shared__ int wtot[32];
    __shared__ unsigned long long wbest[32];
    const int t = t_first + blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n = class_cnt[t], m = Nu - n;
    const int64_t i0 = (int64_t)blockIdx.x * KSD_TILE + (int64_t)tid * KSD_V;
    bool hit[KSD_V];
    long long sb[KSD_V + 1];
    int mine = 0;
#pragma unroll
    for (int v = 0; v < KSD_V; v++) {
        const int64_t i = i0 + v;
        hit[v] = (i < Nu) && ((int)lab[i] == t);
        sb[v] = (i < Nu) ? __double_as_longlong(s_rank[i]) : 0;
        mine += hit[v] ? 1 : 0;
    }
    sb[KSD_V] = (i0 + KSD_V < Nu) ? __double_as_longlong(s_rank[i0 + KSD_V]) : 0;
    // exclusive prefix of `mine` over the block
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int v = wtot[lane];
        int wi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += u; }
        wtot[lane] = wi - v;
    }
    __syncthreads();
    long long hx = (long long)tilebase[(size_t)blockIdx.x * Tc + t] + wtot[warp] + (incl - mine);
    unsigned long long best = 0ull;
#pragma unroll
    for (int v = 0; v < KSD_V; v++) {
        const int64_t i = i0 + v;
        hx += hit[v] ? 1 : 0;
        if (i + 1 < Nu && sb[v] != sb[v + 1]) {  // end of a tie group
            const long long hy = (i + 1) - hx;
            long long cur = hx * m - hy * n;
            cur = cur < 0 ? -cur : cur;
            best = (unsigned long long)cur > best ? (unsigned long long)cur : best;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long u = __shfl_xor_sync(0xffffffffu, best, o); best = u > best ? u : best; }
    if (lane == 0) wbest[warp] = best;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < OBS_THREADS / 32; w++) best = wbest[w] > best ? wbest[w] : best;
        if (best) atomicMax(&ksd_sup[t], best);
    }
}

// Secondary supremum (:244-312), second walk.
// Second walk over [0, secidx]: with |score| numerators and the same denominators a' = +-a, b' = +-b (all
// ranked scores share a sign), so the candidates are e = +-diff > 0; first position of the largest one.
__global__ void __launch_bounds__(256) k_obs_sec_sup(const float2* __restrict__ chainT, const float* __restrict__ denom,
                                                     const double* __restrict__ s_rank, int64_t Nu, int32_t Tpad, int32_t t0,
                                                     const unsigned int* __restrict__ secidx,
                                                     unsigned long long* __restrict__ seckey) {
    const int t = t0 + blockIdx.y;
    const unsigned int sidx = secidx[t];
    if (sidx == 0u) return;
    const int64_t i_beg = (int64_t)blockIdx.x * OBS_TILE, i_end = min(min(i_beg + OBS_TILE, Nu), (int64_t)sidx + 1);
    if (i_beg >= i_end) return;
    const bool neg_mode = (s_rank[Nu - 1] < 0.0) && (s_rank[0] <= 0.0);
    const float dA = denom[t], dB = denom[Tpad + t];
    const float2* row = chainT + (size_t)blockIdx.y * Nu;
    unsigned long long best = 0ull;
    for (int64_t i = i_beg + threadIdx.x; i < i_end; i += 256) {
        float e = obs_diff(row[i], dA, dB);
        e = neg_mode ? -e : e;
        if (e > 0.0f) {
            const unsigned long long key = ((unsigned long long)(uint32_t)__float_as_int(e) << 32) |
                                           (unsigned long long)(0xffffffffu - (uint32_t)i);
            best = key > best ? key : best;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long u = __shfl_xor_sync(0xffffffffu, best, o); best = u > best ? u : best; }
    if ((threadIdx.x & 31) == 0 && best) atomicMax(&seckey[t], best);
}

__global__ void k_obs_finish(const float2* __restrict__ chainT, const float* __restrict__ denom,
                             const int32_t* __restrict__ class_cnt, int64_t Nu, int32_t T, int32_t Tpad, int32_t t0, int32_t t1,
                             const unsigned long long* __restrict__ supkey, const unsigned int* __restrict__ lastneg,
                             const unsigned long long* __restrict__ seckey,
                             const unsigned long long* __restrict__ ksd_sup, pctsea_type_result* __restrict__ results) {
    const int t = t0 + blockIdx.x * blockDim.x + threadIdx.x;   // the types [t0, t1) of the resident history slice
    if (t >= t1 || t >= T) return;
    const float qnan = __int_as_float(0x7fc00000);
    const double dnan = __longlong_as_double(0x7ff8000000000000LL);
    pctsea_type_result r;
    const int32_t sizea = class_cnt[t], sizeb = (int32_t)(Nu - class_cnt[t]);
    r.sizeA = sizea; r.sizeB = sizeb;
    r.norm_ews = qnan; r.p_emp = dnan; r.fdr = dnan;
    r.sec_ews = 0.0f; r.sec_supX = 0; r.ks_d = dnan;
    if (!(sizea > 1 && sizeb > 1)) {  // :119-127, :317-323
        r.ews = qnan; r.dstat = qnan; r.supX = -1; r.norm_supX = -1.0; r.raw_sup = qnan;
        results[t] = r;
        return;
    }
    const float dA = denom[t], dB = denom[Tpad + t];
    const float2* row = chainT + (size_t)(t - t0) * Nu;
    const unsigned long long key = supkey[t];
    float sup = 0.0f;
    int32_t supX = 0;
    if (key) {
        supX = (int32_t)(0xffffffffu - (uint32_t)(key & 0xffffffffull)) + 1;
        sup = obs_diff(row[supX - 1], dA, dB);
    }
    r.raw_sup = sup;
    r.dstat = dstat_of(sup, sizea, sizeb);
    r.supX = supX;
    r.norm_supX = __ddiv_rn(__dmul_rn(1.0, (double)supX), (double)Nu);  // :215
    float es = sup;
    const unsigned int ln = lastneg[t];
    if (ln) {  // :216-227
        const float dn = obs_diff(row[ln - 1], dA, dB);
        es = __fsub_rn(sup, fabsf(dn));
    }
    r.ews = es;
    r.ks_d = __ddiv_rn((double)(long long)ksd_sup[t], (double)((long long)sizea * (long long)sizeb));
    const unsigned long long sk = seckey[t];
    if (sk) {
        r.sec_ews = __int_as_float((int)(uint32_t)(sk >> 32));
        r.sec_supX = (int32_t)(0xffffffffu - (uint32_t)(sk & 0xffffffffull)) + 1;
    }
    results[t] = r;
    reinterpret_cast<int32_t*>(&results[t])[3] = 0;   // the struct's padding word: records compare equal as bytes
}

// The accumulator history is kept for a slice of the cell types at a time: as many types as fit obs_chain_budget bytes
// (6 GiB unless PCTSEA_OPT_OBS_CHAIN_MB says otherwise). Config B's 100 types x 219 530 cells are 176 MB and go in one
// slice, and so does an atlas-scale ranking (2 M cells x 300 types = 4.8 GB): one launch of 300 chain CTAs keeps two or
// three of them resident per SM (the walk is latency-bound: 15 ms in five slices of 64 types under a 1 GiB budget).
static int32_t obs_types_per_slice(const pctsea_ctx* c, const EnrichGeom& g) {
    const size_t budget = c->opt_obs_chain_mb > 0 ? (size_t)c->opt_obs_chain_mb << 20 : (size_t)6 << 30;
    const size_t per_type = (size_t)std::max<int64_t>(g.Nu, 1) * sizeof(float2);
    return (int32_t)std::max<size_t>(1, std::min<size_t>((size_t)std::max(g.T, 1), budget / per_type));
}

// Workspaces of the observed stage, reserved up front (before the null takes its memory budget).
void reserve_ks_observed(pctsea_ctx* c, const EnrichGeom& g) {
    c->results.reserve((size_t)std::max(g.T, 1) * sizeof(pctsea_type_result) + 64);
    if (g.T == 0) return;
    const int32_t Tpad = (int32_t)round_up(g.T, 32);
    c->obs_chain.reserve((size_t)g.Nu * obs_types_per_slice(c, g) * sizeof(float2) + 64);
    c->obs_misc.reserve((size_t)Tpad * 40 + 64);
    c->obs_tiles.reserve((size_t)div_up(g.Nu, KSD_TILE) * g.Tc * 4 + 64);
}

template <typename LabT>
static void launch_obs_chain_slice(pctsea_ctx* c, const EnrichGeom& g, int32_t Tpad, int32_t t0, int32_t nt, cudaStream_t st) {
    k_obs_chain<LabT><<<nt, OBS_THREADS, 0, st>>>(c->s_rank.as<double>(), c->lab_rank.as<LabT>(), g.Nu, Tpad, t0,
                                                 c->obs_chain.as<float2>(), c->obs_misc.as<float>());
}

// Observed statistics of the cell types [tb, te) (all of them unless a multi-GPU analysis shards the types over its
// ranks: every type is an independent walk of the ranking).
void launch_ks_observed(pctsea_ctx* c, const EnrichGeom& g, cudaStream_t st, int32_t tb, int32_t te) {
    reserve_ks_observed(c, g);
    if (te < 0) te = g.T;
    c->obs_slice_t0 = 0;
    c->obs_slice_n = 0;
    if (g.T == 0 || te <= tb) return;
    const int32_t Tpad = (int32_t)round_up(g.T, 32);
    // [2][Tpad] denominators | [Tpad] u64 supremum keys | [Tpad] u64 secondary keys | [Tpad] f64 KS D |
    // [Tpad] u32 last-negative indices | [Tpad] u32 secondary indices
    const size_t off_key = (size_t)2 * Tpad * 4, off_sec = off_key + (size_t)Tpad * 8, off_ksd = off_sec + (size_t)Tpad * 8,
                 off_neg = off_ksd + (size_t)Tpad * 8, off_sidx = off_neg + (size_t)Tpad * 4, total = off_sidx + (size_t)Tpad * 4;
    c->obs_misc.reserve(total + 64);
    unsigned char* misc = c->obs_misc.as<unsigned char>();
    float* denom = reinterpret_cast<float*>(misc);
    unsigned long long* supkey = reinterpret_cast<unsigned long long*>(misc + off_key);
    unsigned long long* seckey = reinterpret_cast<unsigned long long*>(misc + off_sec);
    unsigned long long* ksd = reinterpret_cast<unsigned long long*>(misc + off_ksd);
    unsigned int* lastneg = reinterpret_cast<unsigned int*>(misc + off_neg);
    unsigned int* secidx = reinterpret_cast<unsigned int*>(misc + off_sidx);
    PCT_CUDA(cudaMemsetAsync(misc + off_key, 0, total - off_key, st));
    float2* chain = c->obs_chain.as<float2>();
    const int32_t* cnt = c->class_cnt.as<int32_t>();
    const int32_t n_tiles = (int32_t)div_up(g.Nu, KSD_TILE);
    c->obs_tiles.reserve((size_t)n_tiles * g.Tc * 4 + 64);
    int32_t* tiles = c->obs_tiles.as<int32_t>();
    dim3 kgrid((unsigned)n_tiles, (unsigned)(te - tb));
    if (g.lab_bytes == 1) {
        k_obs_tilecnt<uint8_t><<<n_tiles, OBS_THREADS, (size_t)g.Tc * 4, st>>>(c->lab_rank.as<uint8_t>(), g.Nu, g.Tc, tiles);
        k_obs_tilescan<<<(unsigned)div_up(g.Tc, 4), 128, 0, st>>>(tiles, n_tiles, g.Tc);
        k_obs_ksd<uint8_t><<<kgrid, OBS_THREADS, 0, st>>>(c->s_rank.as<double>(), c->lab_rank.as<uint8_t>(), g.Nu, g.Tc,
                                                                 cnt, tiles, tb, ksd);
    } else {
        ensure_dyn_smem(c, k_obs_tilecnt<uint16_t>, (size_t)(((size_t)g.Tc * 4)));
        k_obs_tilecnt<uint16_t><<<n_tiles, OBS_THREADS, (size_t)g.Tc * 4, st>>>(c->lab_rank.as<uint16_t>(), g.Nu, g.Tc, tiles);
        k_obs_tilescan<<<(unsigned)div_up(g.Tc, 4), 128, 0, st>>>(tiles, n_tiles, g.Tc);
        k_obs_ksd<uint16_t><<<kgrid, OBS_THREADS, 0, st>>>(c->s_rank.as<double>(), c->lab_rank.as<uint16_t>(), g.Nu, g.Tc,
                                                                  cnt, tiles, tb, ksd);
    }
    count_launch(c, 3);
    const int32_t Ts = obs_types_per_slice(c, g);
    for (int32_t t0 = tb; t0 < te; t0 += Ts) {
        const int32_t nt = std::min(Ts, te - t0);
        if (g.lab_bytes == 1) launch_obs_chain_slice<uint8_t>(c, g, Tpad, t0, nt, st);
        else launch_obs_chain_slice<uint16_t>(c, g, Tpad, t0, nt, st);
        dim3 grid((unsigned)div_up(g.Nu, OBS_TILE), (unsigned)nt);
        k_obs_sup<<<grid, 256, 0, st>>>(chain, denom, cnt, g.Nu, Tpad, t0, supkey);
        k_obs_neg<<<grid, 256, 0, st>>>(chain, denom, g.Nu, Tpad, t0, supkey, lastneg, secidx);
        k_obs_sec_sup<<<grid, 256, 0, st>>>(chain, denom, c->s_rank.as<double>(), g.Nu, Tpad, t0, secidx, seckey);
        k_obs_finish<<<(unsigned)div_up(nt, 128), 128, 0, st>>>(chain, denom, cnt, g.Nu, g.T, Tpad, t0, t0 + nt, supkey, lastneg,
                                                                seckey, ksd, c->results.as<pctsea_type_result>());
        count_launch(c, 5);
        c->obs_slice_t0 = t0;
        c->obs_slice_n = nt;
    }
    PCT_CUDA(cudaGetLastError());
}

__global__ void k_obs_curve(const float2* __restrict__ row, float dA, float dB, int64_t Nu, float* __restrict__ a_out,
                            float* __restrict__ b_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nu) return;
    const float2 v = row[i];
    // same expressions as obs_diff: a (b) keeps its initial 0.0f until the first hit (miss)
    a_out[i] = (__float_as_int(v.x) == (int)0x80000000) ? 0.0f : __fdiv_rn(v.x, dA);
    b_out[i] = (__float_as_int(v.y) == (int)0x80000000) ? 0.0f : __fdiv_rn(v.y, dB);
}

void launch_obs_curve(pctsea_ctx* c, int32_t type_id, int64_t Nu, int32_t T, float* a_dev, float* b_dev) {
    const int32_t Tpad = (int32_t)round_up(T, 32);
    float d[2];
    const float* denom = c->obs_misc.as<float>();
    if (type_id < c->obs_slice_t0 || type_id >= c->obs_slice_t0 + c->obs_slice_n) {
        // not in the resident history slice (rankings whose history went through several slices, or a type that
        // another rank of a multi-GPU analysis walked): walk this one type again (the ranked arrays of the last call
        // are still there; the walk writes the type's two denominators)
        EnrichGeom g;
        g.Nu = Nu; g.T = T;
        if (c->last_lab_bytes == 1) launch_obs_chain_slice<uint8_t>(c, g, Tpad, type_id, 1, c->stream);
        else launch_obs_chain_slice<uint16_t>(c, g, Tpad, type_id, 1, c->stream);
        count_launch(c);
        c->obs_slice_t0 = type_id;
        c->obs_slice_n = 1;
    }
    PCT_CUDA(cudaMemcpyAsync(&d[0], denom + type_id, 4, cudaMemcpyDeviceToHost, c->stream));
    PCT_CUDA(cudaMemcpyAsync(&d[1], denom + Tpad + type_id, 4, cudaMemcpyDeviceToHost, c->stream));
    PCT_CUDA(cudaStreamSynchronize(c->stream));
    k_obs_curve<<<(unsigned)div_up(Nu, 256), 256, 0, c->stream>>>(c->obs_chain.as<float2>() + (size_t)(type_id - c->obs_slice_t0) * Nu,
                                                                 d[0], d[1], Nu, a_dev, b_dev);
    count_launch(c);
    PCT_CUDA(cudaGetLastError());
}

void launch_ks_exact_null(pctsea_ctx* c, const EnrichGeom& g, int32_t Pc, int32_t P_stride, int32_t p_off,
                          size_t label_off) {
    if (g.T == 0 || Pc == 0) return;
    unsigned char* lbase = c->labels.as<unsigned char>() + label_off;
    // gridDim.y is limited to 65535
    for (int32_t p0 = 0; p0 < Pc; p0 += 32768) {
        int32_t pc = std::min<int32_t>(32768, Pc - p0);
        dim3 grid((unsigned)div_up(g.T, 128), (unsigned)pc);
        if (g.lab_bytes == 1)
            k_ks_exact<uint8_t><<<grid, 128, 0, c->stream>>>(
                c->s_rank.as<double>(), reinterpret_cast<uint8_t*>(lbase) + (size_t)p0 * g.Nu_pad, g.Nu, g.Nu_pad, g.T,
                c->null_ews.as<float>(), c->null_d.as<float>(), P_stride, p_off + p0);
        else
            k_ks_exact<uint16_t><<<grid, 128, 0, c->stream>>>(
                c->s_rank.as<double>(), reinterpret_cast<uint16_t*>(lbase) + (size_t)p0 * g.Nu_pad, g.Nu, g.Nu_pad, g.T,
                c->null_ews.as<float>(), c->null_d.as<float>(), P_stride, p_off + p0);
        count_launch(c);
    }
    PCT_CUDA(cudaGetLastError());
}

// ================================================================================================
// permuted labels
// ================================================================================================
template <typename LabT>
__global__ void k_labels_replay(const LabT* __restrict__ lab_rank, const int32_t* __restrict__ replay, int64_t Nu,
                                int64_t Nu_pad, int32_t T, LabT* __restrict__ labels) {
    // PCTSEA.java:1403-1413: label at rank i = shuffled[i] = original label of rank replay[p][i]
    const int p = blockIdx.y;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nu_pad) return;
    LabT v = (LabT)T;
    if (i < Nu) {
        int32_t src = replay[(size_t)p * Nu + i];
        v = lab_rank[src];
    }
    labels[(size_t)p * Nu_pad + i] = v;
}

void launch_labels_replay(pctsea_ctx* c, const EnrichGeom& g, const int32_t* replay_dev, int32_t Pc) {
    if (Pc == 0) return;
    for (int32_t p0 = 0; p0 < Pc; p0 += 32768) {
        int32_t pc = std::min<int32_t>(32768, Pc - p0);
        dim3 grid((unsigned)div_up(g.Nu_pad, 256), (unsigned)pc);
        if (g.lab_bytes == 1)
            k_labels_replay<uint8_t><<<grid, 256, 0, c->stream>>>(c->lab_rank.as<uint8_t>(), replay_dev + (size_t)p0 * g.Nu,
                                                                  g.Nu, g.Nu_pad, g.T,
                                                                  c->labels.as<uint8_t>() + (size_t)p0 * g.Nu_pad);
        else
            k_labels_replay<uint16_t><<<grid, 256, 0, c->stream>>>(c->lab_rank.as<uint16_t>(), replay_dev + (size_t)p0 * g.Nu,
                                                                   g.Nu, g.Nu_pad, g.T,
                                                                   c->labels.as<uint16_t>() + (size_t)p0 * g.Nu_pad);
        count_launch(c);
    }
    PCT_CUDA(cudaGetLastError());
}

}  // namespace pctsea
