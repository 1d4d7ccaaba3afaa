// reduce.cu — a10-a12: empirical p-value, normalised enrichment score, pooled FDR from the null matrix.
//
// Replaces (reference paths under pctsea-core/src/main/java/edu/scripps/yates/pctsea/):
//   a10 PCTSEA.java:1449-1477 + calculateProportionPositive/Negative :1799-1832
//   a11 model/CellTypeClassification.java:407-441 (getNormalizedEnrichmentScore; Maths.mean un-vendored:
//       policy 0 = fp32 running sum in permutation order / count, policy 1 = fp64 sum)
//   a12 PCTSEA.java:1479-1549 (pooled, sorted, Arrays.binarySearch). The reference sorts the pooled normalised
//       null only to binary-search the observed scores in it. The index that search returns depends on nothing
//       but the array length and, for the searched key, how many pooled values are smaller and how many are equal
//       (the probe at `mid` sees "< key" iff mid < #smaller and "> key" iff mid >= #smaller + #equal). So the pool
//       is never materialised or sorted: one pass counts the pooled values between and at the observed scores,
//       and the search is replayed probe for probe on those counts — the duplicate-key corner lands on the same
//       index as the JDK.
//
// Inputs: null fp32 [T][P] (device), results[t].ews (device). All of it is O(T*P) <= a few MB.
#include "common.cuh"
#include <cstring>

namespace pctsea {

__device__ __forceinline__ uint32_t java_orderable_f32(float f) {
    // Float.compare / Arrays.sort(float[]) total order: -0.0 < 0.0, every NaN = canonical NaN, last
    uint32_t b = (f != f) ? 0x7fc00000u : (uint32_t)__float_as_int(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// a10 / a11 in three small kernels.
//   k_red_counts     one warp per type: the four counts of a10 / a11 by ballots over coalesced 32-value chunks
//   k_red_transpose  null [T][P] -> nullT [T / 32][P][32] (32 x 32 tiles through shared memory)
//   k_red_chain      one THREAD per type: the fp32 / fp64 running sums of a11. They must keep the permutation order
//                    (the Maths.mean policy), so each is a dependent chain of P additions whatever the layout. In
//                    the transposed copy the 32 types of a warp form one contiguous stream, which lane 0 brings into
//                    a shared-memory ring with bulk asynchronous copies (TMA, completion on mbarriers): nothing
//                    stands between two additions but a conflict-free shared-memory read and the add itself. (A
//                    warp-per-type version that broadcast each value by shuffle spent 49 cycles per value: 202 us
//                    at P = 8000, 87 % of the whole reduction; thread-per-type walks with plain global loads wait one
//                    L2 round trip per batch of loads in flight: 0.55 - 0.6 ms.)
struct RedCounts { int32_t total, cnt, npos, nneg; };

__global__ void __launch_bounds__(32) k_red_counts(const float* __restrict__ null_ews, int32_t T, int32_t P,
                                                   const pctsea_type_result* __restrict__ res, RedCounts* __restrict__ counts) {
    const int t = blockIdx.x, lane = threadIdx.x;
    const float real = res[t].ews;
    const float fnan = __int_as_float(0x7fc00000);
    const float* nl = null_ews + (size_t)t * P;
    int32_t total = 0, cnt = 0, npos = 0, nneg = 0;
    const bool pos_side = real >= 0.0f;
    for (int32_t p0 = 0; p0 < P; p0 += 32) {
        const int32_t p = p0 + lane;
        const bool in = p < P;
        const float v = in ? nl[p] : fnan;
        // a10: NaNs dropped, zeros excluded, strict comparisons
        const bool side = in && (v == v) && (pos_side ? !(v <= 0.0f) : !(v >= 0.0f));
        total += __popc(__ballot_sync(0xffffffffu, side));
        cnt += __popc(__ballot_sync(0xffffffffu, side && (pos_side ? (v > real) : (v < real))));
        // a11: zeros go with the positives
        npos += __popc(__ballot_sync(0xffffffffu, in && v >= 0.0f));
        nneg += __popc(__ballot_sync(0xffffffffu, in && v < 0.0f));
    }
    if (lane == 0) counts[t] = RedCounts{ total, cnt, npos, nneg };
}

// null [T][P] -> nullT [T / 32][P][32]: the 32 types of one chain warp are one contiguous stream of 128-byte lines
// (padding types and nothing else hold NaN), 32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256) k_red_transpose(const float* __restrict__ in, int32_t T, int32_t P,
                                                       float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int p0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
    const float fnan = __int_as_float(0x7fc00000);
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int t = t0 + r, p = p0 + tx;
        tile[r][tx] = (t < T && p < P) ? in[(size_t)t * P + p] : fnan;
    }
    __syncthreads();
    float* dst = out + (size_t)blockIdx.y * P * 32;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int p = p0 + r;
        if (p < P) dst[(size_t)p * 32 + tx] = tile[tx][r];
    }
}

// ---- bulk asynchronous copies (TMA, 1-D) into a shared-memory ring, completion through mbarriers ----------------
__device__ __forceinline__ void mbar_init(uint32_t bar_saddr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar_saddr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_saddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar_saddr, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(bar_saddr), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint32_t bar_saddr) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_saddr), "l"(src), "r"(bytes), "r"(bar_saddr) : "memory");
}

constexpr int RC_STAGE_PERMS = 128;                       // permutations per ring stage: 128 x 128 B = 16 KB
constexpr int RC_STAGES = 6;                              // 96 KB in flight per chain warp (dynamic shared memory)

// One warp per 32 types, lane = type. The warp's stream nullT[tb][P][32] is contiguous, so lane 0 keeps RC_STAGES bulk
// copies of 16 KB in flight (cp.async.bulk + mbarrier complete_tx) and every lane walks its own column of the stage
// that has landed: conflict-free shared-memory reads, no global-load latency anywhere near the dependent adds.
__global__ void __launch_bounds__(32) k_red_chain(const float* __restrict__ nullT, int32_t T, int32_t P,
                                                  int mean_policy, const RedCounts* __restrict__ counts,
                                                  pctsea_type_result* __restrict__ res, float* __restrict__ expct_out) {
    extern __shared__ __align__(128) float ring_dyn[];   // [RC_STAGES][RC_STAGE_PERMS * 32]
    float (*ring)[RC_STAGE_PERMS * 32] = reinterpret_cast<float (*)[RC_STAGE_PERMS * 32]>(ring_dyn);
    __shared__ __align__(8) unsigned long long bars[RC_STAGES];
    const int lane = threadIdx.x;
    const int t = blockIdx.x * 32 + lane;
    const float* src = nullT + (size_t)blockIdx.x * P * 32;
    const uint32_t ring_saddr = (uint32_t)__cvta_generic_to_shared(&ring[0][0]);
    const uint32_t bar_saddr = (uint32_t)__cvta_generic_to_shared(&bars[0]);
    const int n_stage = (P + RC_STAGE_PERMS - 1) / RC_STAGE_PERMS;
    auto stage_perms = [&](int k) { return min(RC_STAGE_PERMS, P - k * RC_STAGE_PERMS); };
    if (lane == 0) {
        for (int k = 0; k < RC_STAGES; k++) mbar_init(bar_saddr + 8u * k, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        for (int k = 0; k < RC_STAGES && k < n_stage; k++) {
            const uint32_t bytes = (uint32_t)stage_perms(k) * 128u;
            mbar_expect_tx(bar_saddr + 8u * k, bytes);
            bulk_g2s(ring_saddr + (uint32_t)k * RC_STAGE_PERMS * 128u, src + (size_t)k * RC_STAGE_PERMS * 32, bytes, bar_saddr + 8u * k);
        }
    }
    float fsum_pos = 0.0f, fsum_neg = 0.0f;
    double dsum_pos = 0.0, dsum_neg = 0.0;
    for (int k = 0; k < n_stage; k++) {
        const int slot = k % RC_STAGES;
        mbar_wait(bar_saddr + 8u * slot, (uint32_t)((k / RC_STAGES) & 1));
        const float* col = &ring[slot][lane];
        const int np = stage_perms(k);
        if (mean_policy) {
            for (int j = 0; j < np; j++) {
                const float vj = col[j * 32];   // NaNs (real ones and the padding types) fail both tests
                if (vj >= 0.0f) dsum_pos = __dadd_rn(dsum_pos, (double)vj);
                else if (vj < 0.0f) dsum_neg = __dadd_rn(dsum_neg, (double)vj);
            }
        } else {
#pragma unroll 8
            for (int j = 0; j < np; j++) {
                const float vj = col[j * 32];
                // adding -0.0f leaves any fp32 accumulator bit-unchanged (also a -0.0f one): branch-free chains
                fsum_pos = __fadd_rn(fsum_pos, (vj >= 0.0f) ? vj : -0.0f);
                fsum_neg = __fadd_rn(fsum_neg, (vj < 0.0f) ? vj : -0.0f);
            }
        }
        __syncwarp();   // every lane is done with the slot before it is refilled
        if (lane == 0 && k + RC_STAGES < n_stage) {
            const int kn = k + RC_STAGES;
            const uint32_t bytes = (uint32_t)stage_perms(kn) * 128u;
            mbar_expect_tx(bar_saddr + 8u * slot, bytes);
            bulk_g2s(ring_saddr + (uint32_t)slot * RC_STAGE_PERMS * 128u, src + (size_t)kn * RC_STAGE_PERMS * 32, bytes, bar_saddr + 8u * slot);
        }
    }
    if (t >= T) return;
    const float real = res[t].ews;
    const double dnan = __longlong_as_double(0x7ff8000000000000LL);
    const float fnan = __int_as_float(0x7fc00000);
    if (real != real) {
        res[t].p_emp = dnan; res[t].norm_ews = fnan; res[t].fdr = dnan; expct_out[t] = fnan;
        return;
    }
    const RedCounts c = counts[t];
    const bool pos_side = real >= 0.0f;
    res[t].p_emp = __ddiv_rn(__dmul_rn(1.0, (double)c.cnt), (double)c.total);  // 0/0 -> NaN
    float mean_pos = mean_policy ? __double2float_rn(__ddiv_rn(dsum_pos, (double)c.npos)) : __fdiv_rn(fsum_pos, (float)c.npos);
    float mean_neg = mean_policy ? __double2float_rn(__ddiv_rn(dsum_neg, (double)c.nneg)) : __fdiv_rn(fsum_neg, (float)c.nneg);
    float expct = pos_side ? fabsf(mean_pos) : fabsf(mean_neg);
    res[t].norm_ews = __fdiv_rn(real, expct);
    expct_out[t] = expct;
}

// Arrays.binarySearch on a sorted array of length n that holds `less` elements below the key and `eq` equal to it.
__device__ __forceinline__ int32_t java_binary_search_counts(int32_t n, int32_t less, int32_t eq) {
    int32_t low = 0, high = n - 1;
    while (low <= high) {
        const int32_t mid = (int32_t)(((uint32_t)low + (uint32_t)high) >> 1);
        if (mid < less) low = mid + 1;
        else if (mid >= less + eq) high = mid - 1;
        else return mid;
    }
    return -(low + 1);
}

__device__ __forceinline__ int32_t java_binary_search(const uint32_t* a, int32_t n, uint32_t key) {
    int32_t low = 0, high = n - 1;
    while (low <= high) {
        const int32_t mid = (int32_t)(((uint32_t)low + (uint32_t)high) >> 1);
        const uint32_t mv = a[mid];
        if (mv < key) low = mid + 1;
        else if (mv > key) high = mid - 1;
        else return mid;
    }
    return -(low + 1);
}

__device__ __forceinline__ bool fdr_candidate(float ne) { return ne == ne && !(ne < 0.0f); }

// One block: R = the observed normalised scores that are not NaN and not < 0, sorted (rank sort, T is small).
// Also clears the pooled-null bins: bins[2j] counts pooled values strictly between R[j-1] and R[j] (bins[2 nobs]:
// above R[nobs-1]), bins[2j+1] those equal to R[j] (accumulated at the first index of a run of equal scores).
__global__ void __launch_bounds__(1024) k_red_rank(int32_t T, const pctsea_type_result* __restrict__ res,
                                                   uint32_t* __restrict__ Rsorted, int32_t* __restrict__ nobs_out,
                                                   unsigned int* __restrict__ bins) {
    for (int i = threadIdx.x; i < 2 * T + 2; i += 1024) bins[i] = 0u;
    int32_t mine = 0;
    for (int t = threadIdx.x; t < T; t += 1024) {
        const float ne = res[t].norm_ews;
        if (!fdr_candidate(ne)) continue;
        mine++;
        const uint32_t k = java_orderable_f32(ne);
        int32_t rank = 0;
        for (int u = 0; u < T; u++) {
            const float nu = res[u].norm_ews;
            if (!fdr_candidate(nu)) continue;
            const uint32_t ku = java_orderable_f32(nu);
            if (ku < k || (ku == k && u < t)) rank++;
        }
        Rsorted[rank] = k;
    }
    if (threadIdx.x == 0) *nobs_out = 0;
    __syncthreads();
    if (mine) atomicAdd(nobs_out, mine);
}

// The pooled normalised null = null[t][p] / expected[t] of every type whose normalised ES is not NaN, entries < 0
// dropped (PCTSEA.java:1479-1500). Each value is located among the observed scores and counted into its bin.
__global__ void __launch_bounds__(256) k_red_count(const float* __restrict__ null_ews, int32_t T, int32_t P,
                                                   const pctsea_type_result* __restrict__ res, const float* __restrict__ expct,
                                                   const uint32_t* __restrict__ Rsorted, const int32_t* __restrict__ nobs_p,
                                                   unsigned int* __restrict__ bins, int use_smem) {
    extern __shared__ unsigned int sbins[];
    const int32_t nobs = *nobs_p;
    const int nb = 2 * nobs + 1;
    if (use_smem) {
        for (int i = threadIdx.x; i < nb; i += 256) sbins[i] = 0u;
        __syncthreads();
    }
    unsigned int* dst = use_smem ? sbins : bins;
    const int64_t n = (int64_t)T * P;
    for (int64_t id = (int64_t)blockIdx.x * 256 + threadIdx.x; id < n; id += (int64_t)gridDim.x * 256) {
        const int t = (int)(id / P);
        const float ne = res[t].norm_ews;
        if (ne != ne) continue;
        const float v = __fdiv_rn(null_ews[id], expct[t]);
        if (v < 0.0f) continue;
        const uint32_t k = java_orderable_f32(v);
        int lo = 0, hi = nobs;   // lower bound: first j with R[j] >= k
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (Rsorted[mid] < k) lo = mid + 1; else hi = mid;
        }
        const bool eq = lo < nobs && Rsorted[lo] == k;
        atomicAdd(&dst[2 * lo + (eq ? 1 : 0)], 1u);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < nb; i += 256)
            if (sbins[i]) atomicAdd(&bins[i], sbins[i]);
    }
}

// One block: prefix over the bins, then per type the two binary searches of PCTSEA.java:1502-1549.
__global__ void __launch_bounds__(1024) k_red_fdr(int32_t T, pctsea_type_result* __restrict__ res,
                                                  const uint32_t* __restrict__ Rsorted, const int32_t* __restrict__ nobs_p,
                                                  unsigned int* __restrict__ bins) {
    __shared__ long long s_nnull;
    const int32_t nobs = *nobs_p;
    const int nb = 2 * nobs + 1;
    // exclusive prefix of the bins in place (nb is a few hundred to a few thousand): one warp, 32 bins per step
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        long long run = 0;
        for (int base = 0; base < nb; base += 32) {
            const int i = base + lane;
            const unsigned int v = i < nb ? bins[i] : 0u;
            unsigned int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            if (i < nb) bins[i] = (unsigned int)run + (incl - v);
            run += (long long)__shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) { bins[nb] = (unsigned int)run; s_nnull = run; }
    }
    __syncthreads();
    const long long nnull = s_nnull;
    const double dnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int t = threadIdx.x; t < T; t += 1024) {
        const float ne = res[t].norm_ews;
        if (ne != ne || ne < 0.0f) { res[t].fdr = dnan; continue; }
        const uint32_t k = java_orderable_f32(ne);
        const int32_t i1 = java_binary_search(Rsorted, nobs, k);
        // first index of this score's run in R: its bins hold the pooled values below / equal to the score
        int lo = 0, hi = nobs;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (Rsorted[mid] < k) lo = mid + 1; else hi = mid;
        }
        const int32_t less = (int32_t)bins[2 * lo + 1];                   // everything before the "equal" bin
        const int32_t eq = (int32_t)(bins[2 * lo + 2] - bins[2 * lo + 1]);
        const int32_t i2 = java_binary_search_counts((int32_t)nnull, less, eq);
        const int32_t sobs = nobs - (i1 >= 0 ? i1 + 1 : -i1 - 1);
        const int32_t snull = (int32_t)nnull - (i2 >= 0 ? i2 + 1 : -i2 - 1);
        double fdr;
        if (snull == 0) fdr = 0.0;
        else fdr = __dmul_rn(__ddiv_rn(__dmul_rn(1.0, (double)snull), (double)sobs),
                             __ddiv_rn(__dmul_rn(1.0, (double)nobs), (double)nnull));
        res[t].fdr = fdr;
    }
}

void launch_reduce(pctsea_ctx* c, int32_t T, int32_t P, const float* null_dev, int mean_policy,
                   pctsea_type_result* results_dev) {
    if (T == 0) return;
    const int64_t n = (int64_t)T * P;
    if (n > 0x7fffffffll) throw Error(PCTSEA_EINVAL, "pooled null larger than a Java int");
    // [T] expected values | [T] sorted observed keys | nobs | [2T + 2] bins | [T] counts | [Tpad / 32][P][32] transposed null
    const int32_t Tpad = (int32_t)round_up(T, 32);
    const size_t off_r = (size_t)round_up((int64_t)T * 4, 16), off_n = off_r + (size_t)round_up((int64_t)T * 4, 16),
                 off_b = off_n + 16, off_c = (size_t)round_up((int64_t)(off_b + ((size_t)2 * T + 2) * 4), 16),
                 off_t = (size_t)round_up((int64_t)(off_c + (size_t)T * sizeof(RedCounts)), 128), total = off_t + (size_t)std::max(P, 1) * Tpad * 4;
    c->red_misc.reserve(total + 64);
    unsigned char* base = c->red_misc.as<unsigned char>();
    float* expct = reinterpret_cast<float*>(base);
    uint32_t* Rsorted = reinterpret_cast<uint32_t*>(base + off_r);
    int32_t* nobs = reinterpret_cast<int32_t*>(base + off_n);
    unsigned int* bins = reinterpret_cast<unsigned int*>(base + off_b);
    RedCounts* counts = reinterpret_cast<RedCounts*>(base + off_c);
    float* nullT = reinterpret_cast<float*>(base + off_t);

    k_red_counts<<<(unsigned)T, 32, 0, c->stream>>>(null_dev, T, P, results_dev, counts);
    if (P > 0) {
        dim3 tg((unsigned)div_up(P, 32), (unsigned)(Tpad / 32));
        k_red_transpose<<<tg, 256, 0, c->stream>>>(null_dev, T, P, nullT);
    }
    const size_t ring_bytes = (size_t)RC_STAGES * RC_STAGE_PERMS * 128;
    ensure_dyn_smem(c, k_red_chain, ring_bytes);
    k_red_chain<<<(unsigned)(Tpad / 32), 32, ring_bytes, c->stream>>>(nullT, T, P, mean_policy, counts, results_dev, expct);
    count_launch(c, P > 0 ? 2 : 1);
    k_red_rank<<<1, 1024, 0, c->stream>>>(T, results_dev, Rsorted, nobs, bins);
    if (n > 0) {
        const int use_smem = T <= 4096;
        const int grid = (int)std::min<int64_t>(div_up(n, 256), (int64_t)c->num_sms * 8);
        k_red_count<<<grid, 256, use_smem ? ((size_t)2 * T + 2) * 4 : 0, c->stream>>>(null_dev, T, P, results_dev, expct, Rsorted,
                                                                                     nobs, bins, use_smem);
    }
    k_red_fdr<<<1, 1024, 0, c->stream>>>(T, results_dev, Rsorted, nobs, bins);
    count_launch(c, n > 0 ? 4 : 3);
    PCT_CUDA(cudaGetLastError());
}

}  // namespace pctsea
