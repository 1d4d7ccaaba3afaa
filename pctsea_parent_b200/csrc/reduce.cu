// reduce.cu — a10-a12: empirical p-value, normalised enrichment score, pooled FDR from the null matrix.
//
// Replaces (reference paths under pctsea-core/src/main/java/edu/scripps/yates/pctsea/):
//   a10 PCTSEA.java:1449-1477 + calculateProportionPositive/Negative :1799-1832
//   a11 model/CellTypeClassification.java:407-441 (getNormalizedEnrichmentScore; Maths.mean un-vendored:
//       policy 0 = fp32 running sum in permutation order / count, policy 1 = fp64 sum)
//   a12 PCTSEA.java:1479-1549 (pooled, sorted, Arrays.binarySearch — restated probe for probe so the
//       duplicate-key corner lands on the same index as the JDK)
//
// Inputs: null fp32 [T][P] (device), results[t].ews (device). All of it is O(T*P) <= a few MB; the
// only heavy step is sorting the pooled normalised null (<= T*P floats) with the stage-2 radix sort.
#include "common.cuh"
#include <cstring>

namespace pctsea {

void radix_passes_public(pctsea_ctx* c, int64_t n, const uint32_t* hist_host, int ndig, uint64_t** keys,
                         uint64_t** keys_alt, uint32_t** vals, uint32_t** vals_alt);

__device__ __forceinline__ uint32_t java_orderable_f32(float f) {
    // Float.compare / Arrays.sort(float[]) total order: -0.0 < 0.0, every NaN = canonical NaN, last
    uint32_t b = (f != f) ? 0x7fc00000u : (uint32_t)__float_as_int(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// One warp per type. Counts are warp-parallel (ballots); the fp32 / fp64 running sums of a11 keep the
// permutation order (the Maths.mean policy) by replaying each coalesced 32-value chunk lane by lane.
__global__ void __launch_bounds__(32) k_red_type(const float* __restrict__ null_ews, int32_t T, int32_t P,
                                                 int mean_policy, pctsea_type_result* __restrict__ res,
                                                 float* __restrict__ expct_out) {
    const int t = blockIdx.x, lane = threadIdx.x;
    const float real = res[t].ews;
    const double dnan = __longlong_as_double(0x7ff8000000000000LL);
    const float fnan = __int_as_float(0x7fc00000);
    if (real != real) {
        if (lane == 0) { res[t].p_emp = dnan; res[t].norm_ews = fnan; res[t].fdr = dnan; expct_out[t] = fnan; }
        return;
    }
    const float* nl = null_ews + (size_t)t * P;
    int32_t total = 0, cnt = 0, npos = 0, nneg = 0;
    float fsum_pos = 0.0f, fsum_neg = 0.0f;
    double dsum_pos = 0.0, dsum_neg = 0.0;
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
        const int n = min(32, P - p0);
        // only the accumulator of the selected mean policy: the two sequential chains (positives, negatives) are
        // what bounds this kernel
        if (mean_policy) {
            for (int j = 0; j < n; j++) {
                const float vj = __shfl_sync(0xffffffffu, v, j);
                if (vj >= 0.0f) dsum_pos = __dadd_rn(dsum_pos, (double)vj);
                else if (vj < 0.0f) dsum_neg = __dadd_rn(dsum_neg, (double)vj);
            }
        } else {
#pragma unroll 8
            for (int j = 0; j < n; j++) {
                const float vj = __shfl_sync(0xffffffffu, v, j);
                // adding -0.0f leaves any fp32 accumulator bit-unchanged (also a -0.0f one), so the two chains can
                // be branch-free
                fsum_pos = __fadd_rn(fsum_pos, (vj >= 0.0f) ? vj : -0.0f);
                fsum_neg = __fadd_rn(fsum_neg, (vj < 0.0f) ? vj : -0.0f);
            }
        }
    }
    if (lane != 0) return;
    res[t].p_emp = __ddiv_rn(__dmul_rn(1.0, (double)cnt), (double)total);  // 0/0 -> NaN
    float mean_pos = mean_policy ? __double2float_rn(__ddiv_rn(dsum_pos, (double)npos)) : __fdiv_rn(fsum_pos, (float)npos);
    float mean_neg = mean_policy ? __double2float_rn(__ddiv_rn(dsum_neg, (double)nneg)) : __fdiv_rn(fsum_neg, (float)nneg);
    float expct = pos_side ? fabsf(mean_pos) : fabsf(mean_neg);
    res[t].norm_ews = __fdiv_rn(real, expct);
    expct_out[t] = expct;
}

// Pool the normalised nulls of every type whose normalised ES is not NaN, dropping entries < 0.
__global__ void __launch_bounds__(256) k_red_pool(const float* __restrict__ null_ews, int32_t T, int32_t P,
                                                  const pctsea_type_result* __restrict__ res,
                                                  const float* __restrict__ expct, uint64_t* __restrict__ keys,
                                                  unsigned long long* __restrict__ count, uint32_t* __restrict__ hist) {
    __shared__ uint32_t shist[4 * 256];
    for (int i = threadIdx.x; i < 4 * 256; i += 256) shist[i] = 0;
    __syncthreads();
    const int64_t n = (int64_t)T * P;
    const int lane = threadIdx.x & 31;
    for (int64_t base = (int64_t)blockIdx.x * 256; base < n; base += (int64_t)gridDim.x * 256) {
        int64_t id = base + threadIdx.x;
        bool take = false;
        uint32_t k = 0;
        if (id < n) {
            int t = (int)(id / P);
            float ne = res[t].norm_ews;
            if (ne == ne) {
                float v = __fdiv_rn(null_ews[id], expct[t]);
                if (!(v < 0.0f)) { take = true; k = java_orderable_f32(v); }
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, take);
        unsigned long long wbase = 0;
        if (lane == 0 && m) wbase = atomicAdd(count, (unsigned long long)__popc(m));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (take) {
            keys[wbase + __popc(m & ((1u << lane) - 1u))] = (uint64_t)k;
#pragma unroll
            for (int d = 0; d < 4; d++) atomicAdd(&shist[d * 256 + ((k >> (8 * d)) & 0xff)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * 256; i += 256)
        if (shist[i]) atomicAdd(&hist[i], shist[i]);
}

__device__ __forceinline__ int32_t java_binary_search(const uint64_t* a, int32_t n, uint64_t key) {
    int32_t low = 0, high = n - 1;
    while (low <= high) {
        int32_t mid = (int32_t)(((uint32_t)low + (uint32_t)high) >> 1);
        uint64_t mv = a[mid];
        if (mv < key) low = mid + 1;
        else if (mv > key) high = mid - 1;
        else return mid;
    }
    return -(low + 1);
}

// One block: sorts the observed normalised scores (rank sort, T is small) and evaluates the FDR per type.
__global__ void __launch_bounds__(1024) k_red_fdr(int32_t T, pctsea_type_result* __restrict__ res,
                                                  const uint64_t* __restrict__ Z, const unsigned long long* __restrict__ nnull_p,
                                                  uint64_t* __restrict__ Rsorted) {
    __shared__ int32_t nobs_s;
    if (threadIdx.x == 0) nobs_s = 0;
    __syncthreads();
    // R = non-NaN normalised scores that are not < 0
    for (int t = threadIdx.x; t < T; t += 1024) {
        float ne = res[t].norm_ews;
        if (ne == ne && !(ne < 0.0f)) atomicAdd(&nobs_s, 1);
    }
    __syncthreads();
    const int32_t nobs = nobs_s;
    for (int t = threadIdx.x; t < T; t += 1024) {
        float ne = res[t].norm_ews;
        if (!(ne == ne && !(ne < 0.0f))) continue;
        uint32_t k = java_orderable_f32(ne);
        int32_t rank = 0;
        for (int u = 0; u < T; u++) {
            float nu = res[u].norm_ews;
            if (!(nu == nu && !(nu < 0.0f))) continue;
            uint32_t ku = java_orderable_f32(nu);
            if (ku < k || (ku == k && u < t)) rank++;
        }
        Rsorted[rank] = (uint64_t)k;
    }
    __syncthreads();
    const long long nnull = (long long)(*nnull_p);
    const double dnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int t = threadIdx.x; t < T; t += 1024) {
        float ne = res[t].norm_ews;
        if (ne != ne || ne < 0.0f) { res[t].fdr = dnan; continue; }
        uint64_t k = (uint64_t)java_orderable_f32(ne);
        int32_t i1 = java_binary_search(Rsorted, nobs, k);
        int32_t i2 = java_binary_search(Z, (int32_t)nnull, k);
        int32_t sobs = nobs - (i1 >= 0 ? i1 + 1 : -i1 - 1);
        int32_t snull = (int32_t)nnull - (i2 >= 0 ? i2 + 1 : -i2 - 1);
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
    c->red_misc.reserve((size_t)T * 4 + (size_t)T * 8 + 256);
    float* expct = c->red_misc.as<float>();
    uint64_t* Rsorted = reinterpret_cast<uint64_t*>(c->red_misc.as<unsigned char>() + round_up((int64_t)T * 4, 16));
    c->red_keys.reserve((size_t)std::max<int64_t>(n, 1) * 8 + 64);
    c->red_keys2.reserve((size_t)std::max<int64_t>(n, 1) * 8 + 64);
    c->red_vals.reserve((size_t)std::max<int64_t>(n, 1) * 4 + 64);
    c->red_vals2.reserve((size_t)std::max<int64_t>(n, 1) * 4 + 64);
    c->hist.reserve(8 * 256 * 4);
    c->counters.reserve(64);
    c->pin_small.reserve(8 * 256 * 4 + 64);
    PCT_CUDA(cudaMemsetAsync(c->hist.p, 0, 8 * 256 * 4, c->stream));
    PCT_CUDA(cudaMemsetAsync(c->counters.p, 0, 64, c->stream));
    unsigned long long* count = c->counters.as<unsigned long long>() + 4;

    k_red_type<<<(unsigned)T, 32, 0, c->stream>>>(null_dev, T, P, mean_policy, results_dev, expct);
    count_launch(c);
    uint64_t *k = c->red_keys.as<uint64_t>(), *k2 = c->red_keys2.as<uint64_t>();
    uint32_t *v = c->red_vals.as<uint32_t>(), *v2 = c->red_vals2.as<uint32_t>();
    if (n > 0) {
        int grid = (int)std::min<int64_t>(div_up(n, 256), (int64_t)kNumSMs * 8);
        k_red_pool<<<grid, 256, 0, c->stream>>>(null_dev, T, P, results_dev, expct, k, count, c->hist.as<uint32_t>());
        count_launch(c);
    }
    uint32_t* hh = c->pin_small.as<uint32_t>();
    PCT_CUDA(cudaMemcpyAsync(hh, c->hist.p, 8 * 256 * 4, cudaMemcpyDeviceToHost, c->stream));
    PCT_CUDA(cudaMemcpyAsync(hh + 8 * 256, count, 8, cudaMemcpyDeviceToHost, c->stream));
    PCT_CUDA(cudaStreamSynchronize(c->stream));
    unsigned long long nn;
    memcpy(&nn, hh + 8 * 256, 8);
    if (nn > 0x7fffffffull) throw Error(PCTSEA_EINVAL, "pooled null larger than a Java int");
    if (nn > 1) radix_passes_public(c, (int64_t)nn, hh, 4, &k, &k2, &v, &v2);
    k_red_fdr<<<1, 1024, 0, c->stream>>>(T, results_dev, k, count, Rsorted);
    count_launch(c);
    PCT_CUDA(cudaGetLastError());
}

}  // namespace pctsea
