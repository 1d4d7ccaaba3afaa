// common.cuh — shared internals of libpctsea_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <utility>
#include <mutex>
#include <stdexcept>
#include <cstdio>
#include <cmath>

#include "../../include/pctsea_b200.h"

namespace pctsea {


struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define PCT_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            char _b[512];                                                                           \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),         \
                     __FILE__, __LINE__);                                                           \
            throw ::pctsea::Error(PCTSEA_ECUDA, _b);                                                \
        }                                                                                           \
    } while (0)

#define PCT_REQUIRE(cond, msg)                                                                      \
    do {                                                                                            \
        if (!(cond)) throw ::pctsea::Error(PCTSEA_EINVAL, std::string(msg));                        \
    } while (0)

// Grow-only device buffer.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            throw Error(PCTSEA_ENOMEM, std::string("cudaMalloc of ") + std::to_string(want) + " bytes failed: " + cudaGetErrorString(e));
        }
        cap = want;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + (bytes >> 3) + 256;
        PCT_CUDA(cudaMallocHost(&p, want));
        cap = want;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

enum Ev { EV_BEGIN = 0, EV_SETUP, EV_SCORE, EV_RANK, EV_OBS, EV_LABELS, EV_NULL, EV_GATHER, EV_REDUCE, EV_END, EV_COUNT };

}  // namespace pctsea

struct pctsea_matrix {
    int64_t n_cells = 0;
    int32_t n_genes = 0;
    int64_t nnz = 0;
    const int64_t* row_ptr = nullptr;  // device
    const int32_t* col = nullptr;      // device
    const float* val = nullptr;        // device
    bool owned = false;
    bool all_positive = false;  // every stored value > 0 (found by the load-time validation)
    int32_t* label = nullptr;  // device, [n_cells], owned
    // gene-major copy for the lane-per-cell scoring kernel (score_gm.cu), built on request, owned
    uint2* gm_mo = nullptr;    // [n_groups][n_genes] {cell mask of the gene in the group, offset of its values}
    float* gm_val = nullptr;   // [nnz] values in (group, gene, cell) order
    int32_t* gm_slot_cell = nullptr;   // [n_groups * 32] cell of every (group, lane) slot, -1 = empty
    int64_t* gm_base = nullptr;        // [n_groups + 1] first value of each group in gm_val
    int64_t n_groups = 0;      // groups of 32 cells
    int32_t n_types = 0;
};

struct pctsea_ctx {
    int device = 0;
    // device properties, queried once in pctsea_create
    int num_sms = 0;
    int smem_per_sm = 0;             // bytes of shared memory per SM
    int smem_per_block_optin = 0;    // largest dynamic shared memory a CTA may opt in to
    // pctsea_ctx_set_option (testing and tuning; none is needed in production)
    int opt_obs_overlap = 2;         // observed-statistics kernels on the auxiliary stream beside the null (2: queued before it)
    int opt_score_path = 0;          // 1 = CSR scan even when the matrix has its gene-major index
    int opt_gm_unroll = 0;           // value loads in flight per lane of the gene-major kernel (0 = default)
    int opt_score_warps = 0;         // warps per CTA of the CSR-scan kernel (0 = as many as fit)
    int opt_perm_chunk = 0;          // permutations per label chunk of the EXACT null (0 = memory budget)
    int opt_null_threads = 0;        // threads (segments) per permutation of the fused null kernel (0 = as many as fit)
    int opt_null_ctas_per_sm = 0;    // resident CTAs per SM of the fused null kernel (0 = as many as fit)
    int opt_null_min_threads = 0;    // below this many threads per permutation the cell types are sliced (0 = 192)
    int opt_shard_observed = 1;      // multi-GPU analysis: observed statistics sharded by cell type over the ranks
    int opt_null_fine_entries = -1;  // cap on the fine (region A) entries of the null kernel's class table (-1 = what fits)
    int opt_obs_chain_mb = 0;        // budget of the observed accumulator history in MiB (0 = 6144); more types than fit go in slices
    int opt_shard_score = 1;         // multi-GPU analysis: stage 1 sharded over the ranks (0: every rank scores all cells)
    int opt_debug_gaps = 0;          // print a line to stderr when the stream sat idle between stages
    cudaStream_t stream = nullptr;
    std::string err;
    cudaStream_t stream_aux = nullptr;  // the observed-statistics kernels run here, beside the null on `stream`
    cudaEvent_t ev[pctsea::EV_COUNT];
    cudaEvent_t ev_obs[3] = { nullptr, nullptr, nullptr };  // ranked arrays ready; begin / end of the observed kernels
    bool obs_span_valid = false;
    std::vector<cudaEvent_t> ev_pool;   // reusable events of the per-chunk label / null spans
    size_t ev_used = 0;
    bool ev_rec[pctsea::EV_COUNT];
    pctsea_timings tm;
    int64_t launches = 0;
    bool nvtx_open = false;   // an NVTX stage range is open on the calling thread

    // workspaces (grow-only)
    pctsea::DevBuf list_gene, list_abund, ybg, bitmap, list_rank, score_sums;  // stage 1 lookup + per-cell centred sums
    pctsea::DevBuf score, nused, nnzc, kept;                    // stage 1 outputs [N]
    pctsea::DevBuf gm_fix;                                      // gene-major scoring: zero-variance cells to jitter
    int32_t list_mapped = 0;                                    // list entries with a matrix column (set by the upload)
    pctsea::DevBuf keysA, keysB, valsA, valsB, hist, tile_hist, scan_tmp, counters;  // radix sort
    pctsea::DevBuf order;                                       // [N] int32
    pctsea::DevBuf type_counts;                                 // [2T+1] int32 (hypergeometric inputs)
    int32_t type_counts_T = -1;
    pctsea::DevBuf label_tmp;                                   // [N] int32 (when labels passed per call)
    pctsea::DevBuf s_rank, lab_rank, fx, Sfx, class_cnt;        // ranked arrays [Nu]
    pctsea::DevBuf obs_chain, obs_misc, obs_tiles;                       // observed KS: accumulator history [Nu][Tpad]
    pctsea::DevBuf labels;                                      // EXACT null: [Pc][Nu_pad] permuted labels; fused null: per-CTA label scratch
    pctsea::DevBuf lab_ticket;                                  // permutation ticket counter of the fused null kernel
    pctsea::DevBuf cls_tab;                                     // size-sorted classes: cumulative counts, order, header, bucket table (null_fused.cu)
    int32_t cls_tab_cap = 0;                                    // table capacity cls_tab was last built for
    pctsea::DevBuf replay;                                      // [Pc][Nu] int32
    pctsea::DevBuf fxT;                                         // FAST mode: lane layout of the fixed-point scores
    pctsea::DevBuf null_ews, null_d;                            // [T][Pc]
    pctsea::DevBuf results;                                     // [T] pctsea_type_result
    pctsea::DevBuf red_tmp, red_misc;  // reduce: staged null of pctsea_reduce_null; expected values, sorted scores, bins
    pctsea::PinnedBuf pin_in, pin_out, pin_small;
    int32_t last_T = 0, last_Pc = 0;
    // multi-GPU (comm.cu): NCCL communicator attached by pctsea_comm_init_rank; one ctx = one rank
    void* comm = nullptr;                                       // ncclComm_t
    int32_t comm_nranks = 1, comm_rank = 0;
    pctsea::DevBuf null_gather;                                 // [nranks][T][width] all-gathered null slices (x2 with dStat)
    pctsea::DevBuf null_pool, null_pool_d;                      // [T][P] complete null in global permutation order
    bool shard_score = false;                                   // set for the duration of a multi-GPU pctsea_analyze: stage 1
                                                                // (gene-major kernel) scores every nranks-th tile + all-gather
    pctsea::DevBuf score_pack;                                  // [nranks][tiles per rank][256] {score, nused, kept} in tile order
    int32_t last_pool_P = 0;                                    // P of the resident pooled null, 0 = none
    bool last_pool_dstat = false;
    int32_t obs_slice_t0 = 0, obs_slice_n = 0;   // cell types whose accumulator history is resident in obs_chain
    int last_lab_bytes = 1;                      // label width of the resident ranked arrays
    int64_t last_obs_Nu = -1;   // n_used of the resident observed accumulator history (obs_chain), -1 = none
};

namespace pctsea {

// ---- Philox4x32-10 (device + host) ------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Feistel geometry: digits (L, R) in Z_a x Z_b with a * b >= n and b a multiple of 4 (the null kernel generates four
// consecutive positions at a time from segments that start at multiples of 4: with such rows the four share their
// high digit and no per-position row-wrap arithmetic is needed). Starting from a0 = ceil(sqrt(n)), the a in
// [a0, a0 + max(a0/16, 8)] with the fewest out-of-range points a * b - n is taken (first one on ties): every such
// point costs a cycle-walk, and one walking lane stalls its warp.
inline void feistel_dims(int64_t n, uint32_t* a_out, uint32_t* b_out) {
    int64_t a0 = (int64_t)std::floor(std::sqrt((double)n));
    while (a0 * a0 < n) a0++;
    while (a0 > 1 && (a0 - 1) * (a0 - 1) >= n) a0--;
    if (a0 < 1) a0 = 1;
    auto b_of = [n](int64_t a) { return ((n + a - 1) / a + 3) / 4 * 4; };   // rows of a multiple of 4 positions
    int64_t best_a = a0, best_w = a0 * b_of(a0) - n;
    for (int64_t a = a0 + 1; a <= a0 + std::max<int64_t>(a0 / 16, 8) && best_w > 0; a++) {
        const int64_t w = a * b_of(a) - n;
        if (w < best_w) { best_w = w; best_a = a; }
    }
    *a_out = (uint32_t)best_a;
    *b_out = (uint32_t)b_of(best_a);
}

inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return div_up(a, b) * b; }

// ---- stage launchers (implemented in the per-stage .cu files) ----------------------------------
struct ScoreArgs {
    const pctsea_matrix* m;
    int32_t L;
    const int32_t* gene_idx_dev;
    const float* abund_dev;
    pctsea_score_params prm;
};
void launch_score(pctsea_ctx* c, const ScoreArgs& a);   // picks the gene-major kernel when the matrix has its index
void launch_list_lookup(pctsea_ctx* c, const ScoreArgs& a);   // by-gene abundance table + 2-bit membership bitmap
void launch_score_final(pctsea_ctx* c, const ScoreArgs& a);   // centred sums -> getR() + thresholds
void launch_score_gm(pctsea_ctx* c, const ScoreArgs& a);      // score_gm.cu
void build_gene_index(pctsea_ctx* c, pctsea_matrix* m);       // score_gm.cu; syncs the stream
// Checks a device CSR: bit 0 = row_ptr not 0-based / non-decreasing / not ending at nnz, bit 1 = a column outside
// [0, n_genes), bit 2 = some stored value is not > 0 (informational), bit 3 = a row repeats a column. Syncs the stream.
int validate_csr(pctsea_ctx* c, const pctsea_matrix* m);

// Threshold + stable radix sort; writes c->order ([Np]); returns Np (syncs the stream once).
int64_t launch_rank(pctsea_ctx* c, int64_t N, const double* score_dev, const uint8_t* kept_dev, double min_score);

void launch_type_counts(pctsea_ctx* c, int64_t N, const double* score_dev, const uint8_t* kept_dev,
                        const int32_t* label_dev, double min_score, int32_t T);

// Generic stable LSD radix sort of (u64 key, u32 value) pairs; result lands in keys_out/vals_out
// (which of the two supplied buffers that is, is returned through the pointers).
void radix_sort_pairs(pctsea_ctx* c, int64_t n, uint64_t** keys, uint64_t** keys_alt, uint32_t** vals,
                      uint32_t** vals_alt, int key_bits);

struct EnrichGeom {
    int64_t Nu = 0;
    int64_t Nu_pad = 0;  // label row stride in elements (multiple of 128)
    int32_t T = 0;       // cell types
    int32_t Tc = 0;      // classes = T + 1 (sentinel class T = unknown type)
    int lab_bytes = 1;   // 1 (u8) or 2 (u16)
    int frac_bits = 0;
};
// Gather ranked arrays (s_rank, lab_rank), class counts, max |score|.
void launch_prepare_ranked(pctsea_ctx* c, int64_t N, const double* score_dev, const int32_t* label_dev,
                           const int32_t* order_dev, EnrichGeom& g);
void reserve_ks_observed(pctsea_ctx* c, const EnrichGeom& g);
void launch_ks_observed(pctsea_ctx* c, const EnrichGeom& g, cudaStream_t st, int32_t type_begin = 0, int32_t type_end = -1);
// a / b series of one type from the resident accumulator history into two device arrays of Nu floats
void launch_obs_curve(pctsea_ctx* c, int32_t type_id, int64_t Nu, int32_t T, float* a_dev, float* b_dev);
void launch_labels_replay(pctsea_ctx* c, const EnrichGeom& g, const int32_t* replay_dev, int32_t Pc);
// Geometry of the fused null kernel (null_fused.cu) for one ranking.
struct NullPlanFast {
    int32_t Kp = 0;           // threads (segments) per permutation
    int32_t seg = 0;          // positions per thread (multiple of 16)
    int32_t Ts = 0;           // cell types per pass (== T unless sliced)
    int32_t ctas_per_sm = 1;
    int32_t nb = 0;           // coarse (region B) buckets of the class table
    int32_t nA_cap = 0;       // entries on top of nb that the CTA's shared memory has room for
    int32_t tab_cap = 0;      // capacity of the class table = nb + nA_cap (its fine / coarse split is chosen on the device)
    int shift = 0;            // log2 of the coarse bucket width
    int cum_smem = 0;         // cumulative class counts held in shared memory
    int word_bytes = 4;       // bytes of one label word (4 positions)
    size_t smem = 0;          // dynamic shared memory per CTA
};
NullPlanFast plan_null_fast(const pctsea_ctx* c, const EnrichGeom& g);
void launch_null_fast_prepare(pctsea_ctx* c, const EnrichGeom& g, const NullPlanFast& pl);
void launch_class_tables(pctsea_ctx* c, const EnrichGeom& g, int shift_min, int32_t cap);
void launch_labels_philox_rm(pctsea_ctx* c, const EnrichGeom& g, uint64_t seed, int32_t perm_begin, int32_t Pc, int rounds);
void launch_null_fused(pctsea_ctx* c, const EnrichGeom& g, const NullPlanFast& pl, uint64_t seed, int32_t perm_begin,
                       int32_t Pc, int rounds, int32_t P_stride, int32_t p_off);
void launch_ks_exact_null(pctsea_ctx* c, const EnrichGeom& g, int32_t Pc, int32_t P_stride, int32_t p_off,
                          size_t label_off);
void launch_reduce(pctsea_ctx* c, int32_t T, int32_t P, const float* null_dev, int mean_policy,
                   pctsea_type_result* results_dev);

// comm.cu — NCCL bound at run time
void comm_unique_id(void* id_out);
void comm_init_rank(pctsea_ctx* c, int32_t nranks, int32_t rank, const void* id_bytes);
void comm_destroy(pctsea_ctx* c);
void comm_share(int64_t n, int32_t nranks, int32_t rank, int64_t* begin, int64_t* count);
// all-gather of c->null_ews (and c->null_d) [T][width] over the ranks, reordered into c->null_pool[_d] [T][P]
void comm_pool_null(pctsea_ctx* c, int32_t T, int32_t P, int32_t width, bool with_dstat);
void comm_all_gather_in_place(pctsea_ctx* c, void* buf, size_t bytes_per_rank);

inline void count_launch(pctsea_ctx* c, int n = 1) { c->launches += n; }

// Opt a kernel in to `bytes` of dynamic shared memory, once per (device, kernel) and size: the attribute call
// reaches the driver, and a warm analysis call should make no driver round-trips beyond its launches (DESIGN.md
// "Host-side stalls"). The attribute belongs to the device function, not to a ctx, so the record is process-wide
// (several ctxs may share a device) and the limit is only ever raised.
struct SmemOptIn {
    std::mutex mu;
    std::vector<std::pair<std::pair<int, const void*>, size_t>> seen;
};
inline SmemOptIn& smem_opt_in() { static SmemOptIn s; return s; }

template <typename K>
inline void ensure_dyn_smem(pctsea_ctx* c, K kernel, size_t bytes) {
    const std::pair<int, const void*> key{ c->device, reinterpret_cast<const void*>(kernel) };
    SmemOptIn& r = smem_opt_in();
    std::lock_guard<std::mutex> lock(r.mu);
    for (auto& e : r.seen)
        if (e.first == key) {
            if (e.second >= bytes) return;
            PCT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            e.second = bytes;
            return;
        }
    PCT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    r.seen.push_back({ key, bytes });
}

}  // namespace pctsea
