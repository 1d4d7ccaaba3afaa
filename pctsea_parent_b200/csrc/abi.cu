// abi.cu — the extern "C" boundary of libpctsea_b200.so (see include/pctsea_b200.h).
// Host-side sequencing only: uploads, stage launches on one stream, per-stage CUDA events, downloads.
#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <limits>

#include <nvtx3/nvToolsExt.h>   // header-only: the calls are no-ops unless a profiler injects itself (nsys, ncu)

using namespace pctsea;

namespace {

struct EvTimer {
    // per-chunk events for the null loop, taken from a pool the ctx keeps (no event creation inside a call once
    // the pool is warm: every runtime call that reaches the driver is a chance for the launching thread to stall)
    pctsea_ctx* c;
    explicit EvTimer(pctsea_ctx* ctx) : c(ctx) { c->ev_used = 0; }
    cudaEvent_t mark(cudaStream_t s) {
        if (c->ev_used == c->ev_pool.size()) {
            cudaEvent_t e;
            PCT_CUDA(cudaEventCreate(&e));
            c->ev_pool.push_back(e);
        }
        cudaEvent_t e = c->ev_pool[c->ev_used++];
        PCT_CUDA(cudaEventRecord(e, s));
        return e;
    }
};

// Stage boundary: a CUDA event on the ctx stream (pctsea_timings) and an NVTX range on the calling thread, so that a
// timeline shows the path as setup | score | rank | enrich (observed + null) | reduce | results.
void nvtx_close(pctsea_ctx* c) {
    if (c->nvtx_open) { nvtxRangePop(); c->nvtx_open = false; }
}
void rec(pctsea_ctx* c, int which) {
    PCT_CUDA(cudaEventRecord(c->ev[which], c->stream));
    c->ev_rec[which] = true;
    const char* next = nullptr;
    switch (which) {
        case EV_BEGIN: next = "pctsea:setup"; break;
        case EV_SETUP: next = "pctsea:score"; break;
        case EV_SCORE: next = "pctsea:rank"; break;
        case EV_RANK: next = "pctsea:enrich (observed + null)"; break;
        case EV_NULL: next = "pctsea:reduce"; break;
        case EV_REDUCE: next = "pctsea:results"; break;
        case EV_END: nvtx_close(c); return;
        default: return;   // EV_OBS, EV_GATHER: inside a stage
    }
    nvtx_close(c);
    nvtxRangePushA(next);
    c->nvtx_open = true;
}

float between(pctsea_ctx* c, int a, int b) {
    if (!c->ev_rec[a] || !c->ev_rec[b]) return 0.0f;
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) != cudaSuccess) { cudaGetLastError(); return 0.0f; }
    return ms;
}

void begin_call(pctsea_ctx* c) {
    PCT_CUDA(cudaSetDevice(c->device));
    nvtx_close(c);                         // open only if a previous call failed half-way
    cudaStreamSynchronize(c->stream_aux);  // idle unless a previous call failed half-way
    c->err.clear();
    memset(&c->tm, 0, sizeof c->tm);
    for (int i = 0; i < EV_COUNT; i++) c->ev_rec[i] = false;
    c->launches = 0;
}

template <typename F>
int guarded(pctsea_ctx* c, F&& f) {
    if (!c) return PCTSEA_EINVAL;
    try {
        begin_call(c);
        f();
        return PCTSEA_OK;
    } catch (const Error& e) {
        c->err = e.what();
        cudaGetLastError();
        return e.code;
    } catch (const std::bad_alloc&) {
        c->err = "host allocation failed";
        return PCTSEA_ENOMEM;
    } catch (const std::exception& e) {
        c->err = e.what();
        return PCTSEA_EINVAL;
    } catch (...) {
        c->err = "unknown error";
        return PCTSEA_EINVAL;
    }
}

void h2d(pctsea_ctx* c, void* dst, const void* src, size_t bytes) {
    if (!bytes) return;
    PCT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    c->tm.h2d_bytes += (int64_t)bytes;
}
void d2h(pctsea_ctx* c, void* dst, const void* src, size_t bytes) {
    if (!bytes) return;
    PCT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    c->tm.d2h_bytes += (int64_t)bytes;
}

void validate_score_params(const pctsea_score_params* p, int32_t L) {
    PCT_REQUIRE(p != nullptr, "score params are NULL");
    PCT_REQUIRE(p->method == PCTSEA_METHOD_PEARSONS_CORRELATION || p->method == PCTSEA_METHOD_SIMPLE_SCORE,
                "scoring method not on the accelerated path (only PEARSONS_CORRELATION and SIMPLE_SCORE)");
    PCT_REQUIRE(p->jitter_mode == PCTSEA_JITTER_NAN || p->jitter_mode == PCTSEA_JITTER_PHILOX, "bad jitter_mode");
    if (p->min_genes_cells > L) {
        // model/SingleCell.java:347-352
        throw Error(PCTSEA_EMINGENES, "Minimum number of genes to be expressed in the cell lines (" +
                                          std::to_string(p->min_genes_cells) +
                                          ") is higher than the actual number of genes in the experimental input data (" +
                                          std::to_string(L) + ")");
    }
}

void upload_list(pctsea_ctx* c, const pctsea_matrix* m, int32_t L, const int32_t* gene_idx, const float* abundance) {
    PCT_REQUIRE(L >= 0, "negative list length");
    PCT_REQUIRE(L == 0 || (gene_idx && abundance), "gene_idx / abundance are NULL");
    PCT_REQUIRE(L <= 32767, "list longer than the reference's short gene ids allow (32767)");
    // list entries must map to distinct columns (one matrix column per list entry)
    std::vector<int32_t> tmp;
    tmp.reserve(L);
    for (int32_t l = 0; l < L; l++) {
        PCT_REQUIRE(gene_idx[l] >= -1 && gene_idx[l] < m->n_genes, "gene_idx out of range");
        if (gene_idx[l] >= 0) tmp.push_back(gene_idx[l]);
    }
    c->list_mapped = (int32_t)tmp.size();
    std::sort(tmp.begin(), tmp.end());
    PCT_REQUIRE(std::adjacent_find(tmp.begin(), tmp.end()) == tmp.end(), "duplicate matrix column in gene_idx");
    c->list_gene.reserve((size_t)std::max(L, 1) * 4);
    c->list_abund.reserve((size_t)std::max(L, 1) * 4);
    c->pin_in.reserve((size_t)std::max(L, 1) * 8);
    memcpy(c->pin_in.p, gene_idx, (size_t)L * 4);
    memcpy(c->pin_in.as<unsigned char>() + (size_t)L * 4, abundance, (size_t)L * 4);
    h2d(c, c->list_gene.p, c->pin_in.p, (size_t)L * 4);
    h2d(c, c->list_abund.p, c->pin_in.as<unsigned char>() + (size_t)L * 4, (size_t)L * 4);
}

struct NullPlan {
    int32_t P = 0, begin = 0, count = 0;
    bool full = false;
    int rounds = PCTSEA_DEFAULT_FEISTEL_ROUNDS;
    // multi-GPU call (communicator attached): this rank computes [begin, begin + count), the slices are all-gathered
    // into the pooled null [T][P]; stride = permutations per row of the local slice buffers (the widest rank's count)
    bool pooled = false, pool_dstat = false;
    int32_t stride = 0;
};

NullPlan validate_null_params(const pctsea_null_params* p) {
    PCT_REQUIRE(p != nullptr, "null params are NULL");
    PCT_REQUIRE(p->n_perm >= 0, "negative n_perm");
    PCT_REQUIRE(p->perm_mode == PCTSEA_PERM_REPLAY || p->perm_mode == PCTSEA_PERM_PHILOX, "bad perm_mode");
    PCT_REQUIRE(p->ks_mode == PCTSEA_KS_EXACT || p->ks_mode == PCTSEA_KS_FAST, "bad ks_mode");
    NullPlan np;
    np.P = p->n_perm;
    np.begin = p->perm_begin;
    np.count = p->perm_count == 0 ? (p->n_perm - p->perm_begin) : p->perm_count;
    PCT_REQUIRE(np.begin >= 0 && np.count >= 0 && np.begin + np.count <= np.P, "permutation slice out of range");
    np.full = (np.begin == 0 && np.count == np.P);
    np.rounds = p->feistel_rounds == 0 ? PCTSEA_DEFAULT_FEISTEL_ROUNDS : p->feistel_rounds;
    PCT_REQUIRE(np.rounds >= 2 && np.rounds <= 32, "feistel_rounds out of range [2,32]");
    np.stride = np.count;
    return np;
}

// With a communicator: the call covers all n_perm permutations, this rank takes its contiguous share.
void shard_over_ranks(const pctsea_ctx* c, const pctsea_null_params* p, NullPlan* np, bool want_dstat) {
    if (!c->comm) return;
    PCT_REQUIRE(np->full, "with a communicator attached pctsea_analyze covers all permutations: perm_begin = perm_count = 0");
    PCT_REQUIRE(p->perm_mode == PCTSEA_PERM_PHILOX, "multi-GPU analysis generates its permutations (perm_mode PHILOX)");
    int64_t b = 0, n = 0;
    comm_share(np->P, c->comm_nranks, c->comm_rank, &b, &n);
    np->begin = (int32_t)b;
    np->count = (int32_t)n;
    np->stride = (int32_t)div_up(np->P, c->comm_nranks);
    np->pooled = true;
    np->pool_dstat = want_dstat;
    np->full = false;
}

// Stage 3 on device-resident score / label / order. Leaves results in c->results, the null slice in
// c->null_ews / c->null_d ([T][count]).
void enrich_device(pctsea_ctx* c, int64_t N, const double* score_dev, const int32_t* label_dev,
                   const int32_t* order_dev, int64_t Nu, int32_t T, const pctsea_null_params* prm, const NullPlan& np,
                   const int32_t* replay_host, EvTimer& et, float* labels_ms, float* null_ms,
                   std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& lab_spans,
                   std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& null_spans) {
    (void)labels_ms; (void)null_ms;
    PCT_REQUIRE(T >= 0 && T <= 65534, "n_types out of range");
    PCT_REQUIRE(Nu >= 1, "no ranked cells");
    PCT_REQUIRE(Nu < (int64_t)1 << 31, "too many ranked cells");
    EnrichGeom g;
    g.Nu = Nu;
    g.T = T;
    const bool fast = prm->ks_mode == PCTSEA_KS_FAST;
    const bool philox = prm->perm_mode == PCTSEA_PERM_PHILOX;
    if (fast && !philox) throw Error(PCTSEA_EINVAL, "REPLAY permutations are a validation mode: use ks_mode EXACT");
    launch_prepare_ranked(c, N, score_dev, label_dev, order_dev, g);
    const int32_t Pc_all = np.count;
    const int32_t Pstride = np.stride;   // == Pc_all unless the slices of a multi-GPU call differ in size
    const bool want_null = Pc_all > 0 && T > 0;
    // Everything the call will allocate is reserved here, before any memory budget is taken (the grow-only buffers
    // over-allocate by 1/8): the observed accumulator history first, then the null.
    // A multi-GPU analysis shards the observed statistics by cell type (every type is an independent walk of the
    // ranking): rank r walks the r-th block of ceil(T / R) types, and the result records are all-gathered in place
    // next to the null (PCTSEA_OPT_SHARD_OBSERVED 0: every rank walks all types).
    const bool shard_obs = np.pooled && c->comm_nranks > 1 && T > 0 && c->opt_shard_observed;
    const int32_t obs_w = shard_obs ? (int32_t)div_up(T, c->comm_nranks) : T;
    const int32_t obs_tb = shard_obs ? std::min(T, c->comm_rank * obs_w) : 0;
    const int32_t obs_te = shard_obs ? std::min(T, obs_tb + obs_w) : T;
    c->results.reserve((size_t)std::max(shard_obs ? obs_w * c->comm_nranks : T, 1) * sizeof(pctsea_type_result) + 64);
    reserve_ks_observed(c, g);
    c->null_ews.reserve((size_t)std::max(T, 1) * std::max(Pstride, 1) * 4 + 64);
    c->null_d.reserve((size_t)std::max(T, 1) * std::max(Pstride, 1) * 4 + 64);
    // The observed statistics and the permutation null both depend only on the ranked arrays. The observed stage is
    // a dependent chain of small, latency-bound launches (k_obs_chain: one 512-thread CTA per cell type walking the
    // ranking), so by default it goes to the auxiliary high-priority stream and runs beside the null stage; the
    // reduction waits for both. PCTSEA_OPT_OBS_OVERLAP: 0 keeps the observed kernels on the main stream (before the
    // null), 1 queues them behind the launch of the null kernel, 2 (default) queues them BEFORE it: the null kernel is
    // persistent and fills every SM it is given for the whole stage, so the chain has to be resident first — its
    // CTAs then share their SMs with one null CTA each, the null CTAs that found no room start as the chain's CTAs
    // retire, and the ticket counter of the null kernel balances the work whatever the start order was. With the
    // overlap observed_ms is the span of the observed kernels on their own stream, not a share of total_ms.
    const int obs_mode = (want_null && philox) ? c->opt_obs_overlap : 0;
    const bool obs_overlap = obs_mode != 0;
    c->obs_span_valid = false;
    bool obs_queued = false;
    auto queue_observed_aux = [&]() {
        PCT_CUDA(cudaStreamWaitEvent(c->stream_aux, c->ev_obs[0], 0));
        PCT_CUDA(cudaEventRecord(c->ev_obs[1], c->stream_aux));
        launch_ks_observed(c, g, c->stream_aux, obs_tb, obs_te);
        PCT_CUDA(cudaEventRecord(c->ev_obs[2], c->stream_aux));
        c->obs_span_valid = true;
        obs_queued = true;
    };
    if (obs_overlap) {
        PCT_CUDA(cudaEventRecord(c->ev_obs[0], c->stream));   // ranked arrays and class counts are ready
        if (obs_mode == 2) queue_observed_aux();
    }
    NullPlanFast plan;
    if (want_null && philox) {
        plan = plan_null_fast(c, g);
        launch_class_tables(c, g, plan.shift, plan.tab_cap);
        if (fast) launch_null_fast_prepare(c, g, plan);
    }
    rec(c, EV_RANK);
    if (!obs_overlap) launch_ks_observed(c, g, c->stream, obs_tb, obs_te);
    c->last_obs_Nu = T > 0 ? Nu : -1;
    rec(c, EV_OBS);

    c->last_T = T;
    c->last_Pc = Pc_all;
    if (want_null && fast) {
        // FAST arithmetic: one fused kernel generates the labels and evaluates them; nothing of size P x Nu exists
        cudaEvent_t e0 = et.mark(c->stream);
        launch_null_fused(c, g, plan, prm->seed, np.begin, Pc_all, np.rounds, Pstride, 0);
        cudaEvent_t e1 = et.mark(c->stream);
        if (obs_overlap && !obs_queued) queue_observed_aux();   // mode 1: behind the null launch
        null_spans.push_back({ e0, e1 });
    } else if (want_null) {
        if (!philox) PCT_REQUIRE(replay_host != nullptr, "REPLAY mode needs replay_perm");
        // EXACT (validation) null: chunk the permutations so the label matrix stays within half of the free memory
        const size_t row_bytes = (size_t)g.Nu_pad * g.lab_bytes;
        const size_t per_perm = row_bytes + (philox ? 0 : (size_t)g.Nu * 4);
        int32_t chunk = Pc_all;
        // The free-memory query goes to the kernel-mode driver and was measured to stall the launching thread for
        // 1-70 ms now and then, with the stream idle behind it: ask only when the label buffers have to grow.
        const bool fits = (size_t)Pc_all * row_bytes + 64 <= c->labels.cap &&
                          (philox || (size_t)Pc_all * g.Nu * 4 + 64 <= c->replay.cap);
        if (!fits) {
            size_t free_b = 0, total_b = 0;
            PCT_CUDA(cudaMemGetInfo(&free_b, &total_b));
            // the grow-only buffers over-allocate by 1/8 (DevBuf::reserve): budget for what will really be requested
            const size_t budget = std::max<size_t>((free_b / 2 + c->labels.cap + c->replay.cap) / 9 * 8, (size_t)64 << 20);
            chunk = (int32_t)std::max<size_t>(1, std::min<size_t>((size_t)Pc_all, budget / std::max<size_t>(per_perm, 1)));
        }
        if (c->opt_perm_chunk > 0) chunk = std::max(1, std::min(chunk, c->opt_perm_chunk));   // PCTSEA_OPT_PERM_CHUNK
        c->labels.reserve((size_t)chunk * row_bytes + 64);
        for (int32_t p0 = 0; p0 < Pc_all; p0 += chunk) {
            const int32_t pc = std::min(chunk, Pc_all - p0);
            cudaEvent_t e0 = et.mark(c->stream);
            if (!philox) {
                c->replay.reserve((size_t)pc * g.Nu * 4 + 64);
                h2d(c, c->replay.p, replay_host + (size_t)p0 * g.Nu, (size_t)pc * g.Nu * 4);
                launch_labels_replay(c, g, c->replay.as<int32_t>(), pc);
            } else {
                launch_labels_philox_rm(c, g, prm->seed, np.begin + p0, pc, np.rounds);
            }
            cudaEvent_t e1 = et.mark(c->stream);
            if (obs_overlap && !obs_queued) queue_observed_aux();
            launch_ks_exact_null(c, g, pc, Pstride, p0, 0);
            cudaEvent_t e2 = et.mark(c->stream);
            lab_spans.push_back({ e0, e1 });
            null_spans.push_back({ e1, e2 });
        }
    }
    if (obs_overlap && !obs_queued) queue_observed_aux();
    if (obs_queued) PCT_CUDA(cudaStreamWaitEvent(c->stream, c->ev_obs[2], 0));
    rec(c, EV_NULL);
    c->last_pool_P = 0;
    if (np.pooled && T > 0) {
        // the one exchange step of the path: all ranks' slices -> the complete null, in global permutation order
        // (and the observed statistics of the other ranks' cell types)
        if (shard_obs) comm_all_gather_in_place(c, c->results.p, (size_t)obs_w * sizeof(pctsea_type_result));
        comm_pool_null(c, T, np.P, Pstride, np.pool_dstat);
        c->last_pool_P = np.P;
        c->last_pool_dstat = np.pool_dstat;
        rec(c, EV_GATHER);
        launch_reduce(c, T, np.P, c->null_pool.as<float>(), prm->mean_policy, c->results.as<pctsea_type_result>());
    } else if (np.full && T > 0) {
        launch_reduce(c, T, np.P, c->null_ews.as<float>(), prm->mean_policy, c->results.as<pctsea_type_result>());
    }
    rec(c, EV_REDUCE);
}

void finish_timings(pctsea_ctx* c, const std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& lab_spans,
                    const std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& null_spans) {
    PCT_CUDA(cudaStreamSynchronize(c->stream));
    c->tm.setup_ms = between(c, EV_BEGIN, EV_SETUP);
    c->tm.score_ms = between(c, EV_SETUP, EV_SCORE);
    c->tm.rank_ms = between(c, EV_SCORE, EV_RANK);
    c->tm.observed_ms = between(c, EV_RANK, EV_OBS);
    if (c->obs_span_valid) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->ev_obs[1], c->ev_obs[2]) == cudaSuccess) c->tm.observed_ms = ms;
    }
    c->tm.gather_ms = between(c, EV_NULL, EV_GATHER);
    c->tm.reduce_ms = between(c, c->ev_rec[EV_GATHER] ? EV_GATHER : EV_NULL, EV_REDUCE);
    c->tm.total_ms = between(c, EV_BEGIN, EV_END);
    float lm = 0, nm = 0;
    for (auto& s : lab_spans) { float ms = 0; if (cudaEventElapsedTime(&ms, s.first, s.second) == cudaSuccess) lm += ms; }
    for (auto& s : null_spans) { float ms = 0; if (cudaEventElapsedTime(&ms, s.first, s.second) == cudaSuccess) nm += ms; }
    c->tm.labels_ms = lm;
    c->tm.null_ms = nm;
    c->tm.launches = c->launches;
    if (c->opt_debug_gaps && !null_spans.empty()) {   // PCTSEA_OPT_DEBUG_GAPS: where the stream sat idle
        auto el = [](cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; };
        const float pre = el(c->ev[EV_RANK], lab_spans.empty() ? null_spans.front().first : lab_spans.front().first);
        const float post = el(null_spans.back().second, c->ev[EV_NULL]);
        const float tail = el(c->ev[EV_REDUCE], c->ev[EV_END]);
        if (c->tm.total_ms > 1.3f * (c->tm.setup_ms + c->tm.score_ms + c->tm.rank_ms + lm + nm + c->tm.reduce_ms))
            fprintf(stderr, "[pctsea gaps] total %.3f setup %.3f score %.3f rank %.3f | rank->labels %.3f | labels %.3f null %.3f | null->join %.3f (observed span %.3f) | reduce %.3f | results d2h %.3f\n",
                    c->tm.total_ms, c->tm.setup_ms, c->tm.score_ms, c->tm.rank_ms, pre, lm, nm, post, c->tm.observed_ms, c->tm.reduce_ms, tail);
    }
}

NullPlan validate_analysis(const pctsea_matrix* m, double min_score, const int32_t* label, int32_t n_types,
                           const pctsea_null_params* nprm) {
    PCT_REQUIRE(min_score == min_score, "min_score is NaN");
    PCT_REQUIRE(m->n_cells >= 1, "empty matrix");
    if (!label) {
        PCT_REQUIRE(m->label != nullptr, "no labels: pass label[] or call pctsea_matrix_set_labels");
        PCT_REQUIRE(n_types == m->n_types, "n_types differs from the matrix's resident labels");
    }
    return validate_null_params(nprm);
}

// The matrix's resident labels, or the caller's uploaded to the ctx.
const int32_t* resident_labels(pctsea_ctx* c, const pctsea_matrix* m, const int32_t* label) {
    if (!label) return m->label;
    c->label_tmp.reserve((size_t)m->n_cells * 4);
    h2d(c, c->label_tmp.p, label, (size_t)m->n_cells * 4);
    return c->label_tmp.as<int32_t>();
}

void fill_nan_results(pctsea_type_result* out, int32_t T) {
    const float fn = std::numeric_limits<float>::quiet_NaN();
    const double dn = std::numeric_limits<double>::quiet_NaN();
    for (int32_t t = 0; t < T; t++) {
        pctsea_type_result r;
        memset(&r, 0, sizeof r);
        r.ews = fn; r.dstat = fn; r.supX = -1; r.norm_supX = -1.0; r.raw_sup = fn; r.norm_ews = fn;
        r.p_emp = dn; r.fdr = dn; r.ks_d = dn;
        out[t] = r;
    }
}

// One list through the whole path (the per-schema loop body, PCTSEA.java:348-434) on resident labels; EV_BEGIN is
// already recorded. Ends with the stream synchronised and the ctx timings filled.
void analyze_core(pctsea_ctx* c, const pctsea_matrix* m, int32_t L, const int32_t* gene_idx, const float* abundance,
                  const pctsea_score_params* sprm, double min_score, const int32_t* label_dev, int32_t n_types,
                  const pctsea_null_params* nprm, const NullPlan& np, const int32_t* replay_perm,
                  pctsea_type_result* out, float* null_out, float* nullD_out, double* score_out,
                  int32_t* n_genes_used_out, uint8_t* kept_out, int32_t* order_out, int64_t* n_pass_out) {
    const int64_t N = m->n_cells;
    upload_list(c, m, L, gene_idx, abundance);
    validate_score_params(sprm, L);
    rec(c, EV_SETUP);
    ScoreArgs sa{ m, L, c->list_gene.as<int32_t>(), c->list_abund.as<float>(), *sprm };
    launch_score(c, sa);
    rec(c, EV_SCORE);
    launch_type_counts(c, N, c->score.as<double>(), c->kept.as<uint8_t>(), label_dev, min_score, n_types);
    const int64_t Np = launch_rank(c, N, c->score.as<double>(), c->kept.as<uint8_t>(), min_score);
    if (n_pass_out) *n_pass_out = Np;
    c->tm.n_pass = Np;
    c->tm.n_used = Np > 0 ? Np - 1 : 0;
    if (score_out) d2h(c, score_out, c->score.p, (size_t)N * 8);
    if (n_genes_used_out) d2h(c, n_genes_used_out, c->nused.p, (size_t)N * 4);
    if (kept_out) d2h(c, kept_out, c->kept.p, (size_t)N);
    if (order_out) d2h(c, order_out, c->order.p, (size_t)Np * 4);
    if (Np < (int64_t)nprm->min_pass || Np < 2) {
        PCT_CUDA(cudaStreamSynchronize(c->stream));
        // PCTSEA.java:389-395
        throw Error(PCTSEA_ETOOFEW, "There is not enough cells passing the threshold (only " + std::to_string(Np) +
                                        " pass the threshold and the minimum is " +
                                        std::to_string(std::max(nprm->min_pass, 2)) + ")");
    }
    const int64_t Nu = Np - 1;
    EvTimer et(c);
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ls, ns;
    float lm = 0, nm = 0;
    enrich_device(c, N, c->score.as<double>(), label_dev, c->order.as<int32_t>(), Nu, n_types, nprm, np, replay_perm,
                  et, &lm, &nm, ls, ns);
    d2h(c, out, c->results.p, (size_t)n_types * sizeof(pctsea_type_result));
    if (np.pooled) {
        if (null_out) d2h(c, null_out, c->null_pool.p, (size_t)n_types * np.P * 4);
        if (nullD_out) d2h(c, nullD_out, c->null_pool_d.p, (size_t)n_types * np.P * 4);
    } else {
        if (null_out) d2h(c, null_out, c->null_ews.p, (size_t)n_types * np.count * 4);
        if (nullD_out) d2h(c, nullD_out, c->null_d.p, (size_t)n_types * np.count * 4);
    }
    rec(c, EV_END);
    finish_timings(c, ls, ns);
    c->tm.n_pass = Np;
    c->tm.n_used = Nu;
}

}  // namespace

extern "C" {

int pctsea_abi_version(void) { return PCTSEA_ABI_VERSION; }

int pctsea_create(int device_id, pctsea_ctx** out) {
    if (!out) return PCTSEA_EINVAL;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device_id < 0 || device_id >= n) {
        cudaGetLastError();
        return PCTSEA_ECUDA;  // no CPU fallback
    }
    pctsea_ctx* c = new (std::nothrow) pctsea_ctx();
    if (!c) return PCTSEA_ENOMEM;
    c->device = device_id;
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) { cudaGetLastError(); delete c; return PCTSEA_ECUDA; }
        c->num_sms = prop.multiProcessorCount;
        c->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
        c->smem_per_block_optin = (int)prop.sharedMemPerBlockOptin;
    }
    // the auxiliary stream carries the observed-statistics kernels: a chain of small launches that shares the SMs with
    // the big null kernels; its CTAs go first whenever resources free up, or the chain becomes the critical path
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaSetDevice(device_id) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->stream_aux, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
        cudaGetLastError();
        delete c;
        return PCTSEA_ECUDA;
    }
    for (int i = 0; i < EV_COUNT; i++) {
        if (cudaEventCreate(&c->ev[i]) != cudaSuccess) { cudaGetLastError(); delete c; return PCTSEA_ECUDA; }
        c->ev_rec[i] = false;
    }
    for (int i = 0; i < 3; i++)
        if (cudaEventCreate(&c->ev_obs[i]) != cudaSuccess) { cudaGetLastError(); delete c; return PCTSEA_ECUDA; }
    memset(&c->tm, 0, sizeof c->tm);
    *out = c;
    return PCTSEA_OK;
}

int pctsea_destroy(pctsea_ctx* c) {
    if (!c) return PCTSEA_EINVAL;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    comm_destroy(c);
    DevBuf* bufs[] = { &c->list_gene, &c->list_abund, &c->ybg, &c->bitmap, &c->list_rank, &c->score_sums, &c->score, &c->nused, &c->nnzc, &c->kept, &c->gm_fix,
                       &c->keysA, &c->keysB, &c->valsA, &c->valsB, &c->hist, &c->tile_hist, &c->scan_tmp, &c->counters,
                       &c->order, &c->type_counts, &c->label_tmp, &c->obs_chain, &c->obs_misc, &c->obs_tiles, &c->s_rank, &c->lab_rank, &c->fx, &c->Sfx, &c->class_cnt, &c->labels, &c->lab_ticket, &c->cls_tab,
                       &c->replay, &c->fxT, &c->null_ews, &c->null_d, &c->null_gather, &c->null_pool, &c->null_pool_d, &c->score_pack, &c->results, &c->red_tmp,
                       &c->red_misc };
    for (DevBuf* b : bufs) b->release();
    c->pin_in.release(); c->pin_out.release(); c->pin_small.release();
    for (int i = 0; i < EV_COUNT; i++) cudaEventDestroy(c->ev[i]);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < 3; i++) if (c->ev_obs[i]) cudaEventDestroy(c->ev_obs[i]);
    cudaStreamSynchronize(c->stream_aux);
    cudaStreamDestroy(c->stream_aux);
    cudaStreamDestroy(c->stream);
    delete c;
    return PCTSEA_OK;
}

const char* pctsea_last_error(pctsea_ctx* c) { return c ? c->err.c_str() : "NULL ctx"; }

int pctsea_ctx_set_option(pctsea_ctx* c, int32_t option, int64_t value) {
    if (!c) return PCTSEA_EINVAL;
    auto bad = [&](const char* what) { c->err = std::string("pctsea_ctx_set_option: ") + what; return PCTSEA_EINVAL; };
    if (value < 0 || value > (1 << 20)) return bad("value out of range");
    const int v = (int)value;
    switch (option) {
        case PCTSEA_OPT_OBS_OVERLAP: if (v > 2) return bad("0, 1 or 2"); c->opt_obs_overlap = v; break;
        case PCTSEA_OPT_SCORE_PATH: if (v > 1) return bad("0 or 1"); c->opt_score_path = v; break;
        case PCTSEA_OPT_GM_UNROLL: if (v > 8) return bad("at most 8"); c->opt_gm_unroll = v; break;
        case PCTSEA_OPT_SCORE_WARPS: if (v > 32) return bad("at most 32"); c->opt_score_warps = v; break;
        case PCTSEA_OPT_PERM_CHUNK: c->opt_perm_chunk = v; break;
        case PCTSEA_OPT_NULL_THREADS: if (v > 512) return bad("at most 512"); c->opt_null_threads = v; break;
        case PCTSEA_OPT_NULL_CTAS_PER_SM: if (v > 4) return bad("at most 4"); c->opt_null_ctas_per_sm = v; break;
        case PCTSEA_OPT_NULL_MIN_THREADS: if (v > 512) return bad("at most 512"); c->opt_null_min_threads = v; break;
        case PCTSEA_OPT_DEBUG_GAPS: if (v > 1) return bad("0 or 1"); c->opt_debug_gaps = v; break;
        case PCTSEA_OPT_SHARD_SCORE: if (v > 1) return bad("0 or 1"); c->opt_shard_score = v; break;
        case PCTSEA_OPT_OBS_CHAIN_MB: c->opt_obs_chain_mb = v; break;
        case PCTSEA_OPT_NULL_FINE_ENTRIES: c->opt_null_fine_entries = v - 1; break;
        case PCTSEA_OPT_SHARD_OBSERVED: if (v > 1) return bad("0 or 1"); c->opt_shard_observed = v; break;
        default: return bad("unknown option");
    }
    return PCTSEA_OK;
}

int pctsea_get_timings(pctsea_ctx* c, pctsea_timings* out) {
    if (!c || !out) return PCTSEA_EINVAL;
    *out = c->tm;
    return PCTSEA_OK;
}

void* pctsea_stream(pctsea_ctx* c) { return c ? (void*)c->stream : nullptr; }

int pctsea_matrix_load_csr(pctsea_ctx* c, int64_t n_cells, int32_t n_genes, const int64_t* row_ptr,
                           const int32_t* col_idx, const float* val, pctsea_matrix** out) {
    return guarded(c, [&] {
        PCT_REQUIRE(out != nullptr, "out is NULL");
        *out = nullptr;
        PCT_REQUIRE(n_cells >= 0 && n_cells < ((int64_t)1 << 31), "n_cells out of range");
        PCT_REQUIRE(n_genes >= 0, "n_genes out of range");
        PCT_REQUIRE(row_ptr != nullptr, "row_ptr is NULL");
        PCT_REQUIRE(row_ptr[0] == 0, "row_ptr[0] must be 0");
        for (int64_t i = 0; i < n_cells; i++) PCT_REQUIRE(row_ptr[i + 1] >= row_ptr[i], "row_ptr must be non-decreasing");
        const int64_t nnz = row_ptr[n_cells];
        PCT_REQUIRE(nnz == 0 || (col_idx && val), "col_idx / val are NULL");
        pctsea_matrix* m = new pctsea_matrix();
        m->n_cells = n_cells; m->n_genes = n_genes; m->nnz = nnz; m->owned = true;
        void *rp = nullptr, *ci = nullptr, *vv = nullptr;
        try {
            PCT_CUDA(cudaMalloc(&rp, (size_t)(n_cells + 1) * 8));
            PCT_CUDA(cudaMalloc(&ci, (size_t)std::max<int64_t>(nnz, 1) * 4));
            PCT_CUDA(cudaMalloc(&vv, (size_t)std::max<int64_t>(nnz, 1) * 4));
            PCT_CUDA(cudaMemcpyAsync(rp, row_ptr, (size_t)(n_cells + 1) * 8, cudaMemcpyHostToDevice, c->stream));
            if (nnz) {
                PCT_CUDA(cudaMemcpyAsync(ci, col_idx, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream));
                PCT_CUDA(cudaMemcpyAsync(vv, val, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream));
            }
            PCT_CUDA(cudaStreamSynchronize(c->stream));
        } catch (...) {
            if (rp) cudaFree(rp);
            if (ci) cudaFree(ci);
            if (vv) cudaFree(vv);
            delete m;
            throw;
        }
        m->row_ptr = (const int64_t*)rp; m->col = (const int32_t*)ci; m->val = (const float*)vv;
        int bad = 0;
        try { bad = validate_csr(c, m); } catch (...) { pctsea_matrix_free(c, m); throw; }
        m->all_positive = !(bad & 4);
        if (bad & 3) {
            pctsea_matrix_free(c, m);
            throw Error(PCTSEA_EINVAL, "col_idx entries must lie in [0, n_genes)");
        }
        if (bad & 8) {
            pctsea_matrix_free(c, m);
            throw Error(PCTSEA_EINVAL, "a row stores the same column twice: one value per (cell, gene) (model/SingleCell.java:80)");
        }
        *out = m;
    });
}

int pctsea_matrix_wrap_device(pctsea_ctx* c, int64_t n_cells, int32_t n_genes, const int64_t* row_ptr_dev,
                              const int32_t* col_idx_dev, const float* val_dev, pctsea_matrix** out) {
    return guarded(c, [&] {
        PCT_REQUIRE(out != nullptr, "out is NULL");
        *out = nullptr;
        PCT_REQUIRE(n_cells >= 0 && n_cells < ((int64_t)1 << 31) && n_genes >= 0, "shape out of range");
        PCT_REQUIRE(row_ptr_dev != nullptr, "row_ptr_dev is NULL");
        int64_t nnz = 0;
        PCT_CUDA(cudaMemcpy(&nnz, row_ptr_dev + n_cells, 8, cudaMemcpyDeviceToHost));
        pctsea_matrix* m = new pctsea_matrix();
        m->n_cells = n_cells; m->n_genes = n_genes; m->nnz = nnz; m->owned = false;
        m->row_ptr = row_ptr_dev; m->col = col_idx_dev; m->val = val_dev;
        int bad = 0;
        try {
            PCT_REQUIRE(nnz >= 0, "row_ptr_dev[n_cells] is negative");
            PCT_REQUIRE(nnz == 0 || (col_idx_dev && val_dev), "col_idx_dev / val_dev are NULL");
            PCT_REQUIRE(((uintptr_t)col_idx_dev & 15) == 0 && ((uintptr_t)val_dev & 15) == 0,
                        "col_idx_dev / val_dev must be 16-byte aligned");
            bad = validate_csr(c, m);
            m->all_positive = !(bad & 4);
            if (bad & 1) throw Error(PCTSEA_EINVAL, "row_ptr_dev must start at 0 and be non-decreasing");
            if (bad & 2) throw Error(PCTSEA_EINVAL, "col_idx_dev entries must lie in [0, n_genes)");
            if (bad & 8) throw Error(PCTSEA_EINVAL, "a row stores the same column twice: one value per (cell, gene) (model/SingleCell.java:80)");
        } catch (...) { delete m; throw; }
        *out = m;
    });
}

int pctsea_matrix_build_gene_index(pctsea_ctx* c, pctsea_matrix* m) {
    return guarded(c, [&] {
        PCT_REQUIRE(m != nullptr, "matrix is NULL");
        build_gene_index(c, m);
    });
}

int pctsea_matrix_set_labels(pctsea_ctx* c, pctsea_matrix* m, const int32_t* label, int32_t n_types) {
    return guarded(c, [&] {
        PCT_REQUIRE(m != nullptr && label != nullptr, "matrix / label are NULL");
        PCT_REQUIRE(n_types >= 0 && n_types <= 65534, "n_types out of range");
        if (!m->label) { void* p = nullptr; PCT_CUDA(cudaMalloc(&p, (size_t)std::max<int64_t>(m->n_cells, 1) * 4)); m->label = (int32_t*)p; }
        PCT_CUDA(cudaMemcpy(m->label, label, (size_t)m->n_cells * 4, cudaMemcpyHostToDevice));
        m->n_types = n_types;
    });
}

int pctsea_matrix_free(pctsea_ctx* c, pctsea_matrix* m) {
    if (!m) return PCTSEA_EINVAL;
    if (c) cudaSetDevice(c->device);
    if (m->owned) { cudaFree((void*)m->row_ptr); cudaFree((void*)m->col); cudaFree((void*)m->val); }
    if (m->label) cudaFree(m->label);
    if (m->gm_mo) cudaFree(m->gm_mo);
    if (m->gm_val) cudaFree(m->gm_val);
    if (m->gm_slot_cell) cudaFree(m->gm_slot_cell);
    if (m->gm_base) cudaFree(m->gm_base);
    delete m;
    return PCTSEA_OK;
}

int pctsea_score(pctsea_ctx* c, const pctsea_matrix* m, int32_t L, const int32_t* gene_idx, const float* abundance,
                 const pctsea_score_params* prm, double* score_out, int32_t* n_genes_used_out, int32_t* n_nonzero_out,
                 uint8_t* kept_out) {
    return guarded(c, [&] {
        PCT_REQUIRE(m != nullptr, "matrix is NULL");
        PCT_REQUIRE(score_out != nullptr, "score_out is NULL");
        rec(c, EV_BEGIN);
        upload_list(c, m, L, gene_idx, abundance);
        validate_score_params(prm, L);
        rec(c, EV_SETUP);
        ScoreArgs sa{ m, L, c->list_gene.as<int32_t>(), c->list_abund.as<float>(), *prm };
        launch_score(c, sa);
        rec(c, EV_SCORE);
        const int64_t N = m->n_cells;
        d2h(c, score_out, c->score.p, (size_t)N * 8);
        if (n_genes_used_out) d2h(c, n_genes_used_out, c->nused.p, (size_t)N * 4);
        if (n_nonzero_out) d2h(c, n_nonzero_out, c->nnzc.p, (size_t)N * 4);
        if (kept_out) d2h(c, kept_out, c->kept.p, (size_t)N);
        rec(c, EV_END);
        finish_timings(c, {}, {});
    });
}

int pctsea_rank(pctsea_ctx* c, int64_t n_cells, const double* score, const uint8_t* kept, double min_score,
                int32_t* order_out, int64_t* n_pass_out) {
    return guarded(c, [&] {
        PCT_REQUIRE(n_cells >= 0 && n_cells < ((int64_t)1 << 31), "n_cells out of range");
        PCT_REQUIRE(n_cells == 0 || score != nullptr, "score is NULL");
        PCT_REQUIRE(order_out != nullptr && n_pass_out != nullptr, "outputs are NULL");
        PCT_REQUIRE(min_score == min_score, "min_score is NaN");
        rec(c, EV_BEGIN);
        c->score.reserve((size_t)std::max<int64_t>(n_cells, 1) * 8);
        c->kept.reserve((size_t)std::max<int64_t>(n_cells, 1));
        h2d(c, c->score.p, score, (size_t)n_cells * 8);
        if (kept) h2d(c, c->kept.p, kept, (size_t)n_cells);
        rec(c, EV_SETUP);
        rec(c, EV_SCORE);
        int64_t np = launch_rank(c, n_cells, c->score.as<double>(), kept ? c->kept.as<uint8_t>() : nullptr, min_score);
        rec(c, EV_RANK);
        d2h(c, order_out, c->order.p, (size_t)np * 4);
        rec(c, EV_END);
        finish_timings(c, {}, {});
        *n_pass_out = np;
        c->tm.n_pass = np;
        c->tm.n_used = np > 0 ? np - 1 : 0;
    });
}

int pctsea_enrich(pctsea_ctx* c, int64_t n_cells, const double* score, int64_t n_used, const int32_t* order,
                  const int32_t* label, int32_t n_types, const pctsea_null_params* prm, const int32_t* replay_perm,
                  pctsea_type_result* out, float* null_out, float* nullD_out) {
    return guarded(c, [&] {
        PCT_REQUIRE(score && order && label && out, "score / order / label / out are NULL");
        PCT_REQUIRE(n_cells >= 1 && n_cells < ((int64_t)1 << 31), "n_cells out of range");
        PCT_REQUIRE(n_used >= 1 && n_used <= n_cells, "n_used out of range");
        for (int64_t i = 0; i < n_used; i++) PCT_REQUIRE(order[i] >= 0 && order[i] < n_cells, "order entry out of range");
        NullPlan np = validate_null_params(prm);
        rec(c, EV_BEGIN);
        c->score.reserve((size_t)n_cells * 8);
        c->label_tmp.reserve((size_t)n_cells * 4);
        c->order.reserve((size_t)n_cells * 4);
        h2d(c, c->score.p, score, (size_t)n_cells * 8);
        h2d(c, c->label_tmp.p, label, (size_t)n_cells * 4);
        h2d(c, c->order.p, order, (size_t)n_used * 4);
        rec(c, EV_SETUP);
        rec(c, EV_SCORE);
        EvTimer et(c);
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ls, ns;
        float lm = 0, nm = 0;
        enrich_device(c, n_cells, c->score.as<double>(), c->label_tmp.as<int32_t>(), c->order.as<int32_t>(), n_used,
                      n_types, prm, np, replay_perm, et, &lm, &nm, ls, ns);
        d2h(c, out, c->results.p, (size_t)n_types * sizeof(pctsea_type_result));
        if (null_out) d2h(c, null_out, c->null_ews.p, (size_t)n_types * np.count * 4);
        if (nullD_out) d2h(c, nullD_out, c->null_d.p, (size_t)n_types * np.count * 4);
        rec(c, EV_END);
        finish_timings(c, ls, ns);
        c->tm.n_used = n_used;
        c->tm.n_pass = n_used + 1;
    });
}

static void reduce_common(pctsea_ctx* c, int32_t T, int32_t P, const float* null_dev, int32_t mean_policy,
                          pctsea_type_result* inout) {
    c->results.reserve((size_t)std::max(T, 1) * sizeof(pctsea_type_result));
    h2d(c, c->results.p, inout, (size_t)T * sizeof(pctsea_type_result));
    rec(c, EV_NULL);
    launch_reduce(c, T, P, null_dev, mean_policy, c->results.as<pctsea_type_result>());
    rec(c, EV_REDUCE);
    d2h(c, inout, c->results.p, (size_t)T * sizeof(pctsea_type_result));
    rec(c, EV_END);
    finish_timings(c, {}, {});
}

int pctsea_reduce_null(pctsea_ctx* c, int32_t n_types, int32_t n_perm, const float* null_ews, int32_t mean_policy,
                       pctsea_type_result* inout) {
    return guarded(c, [&] {
        PCT_REQUIRE(n_types >= 0 && n_perm >= 0 && inout != nullptr, "bad arguments");
        PCT_REQUIRE((int64_t)n_types * n_perm == 0 || null_ews != nullptr, "null_ews is NULL");
        rec(c, EV_BEGIN);
        // keep it apart from the ctx's own null slice
        c->red_tmp.reserve((size_t)std::max<int64_t>((int64_t)n_types * n_perm, 1) * 4);
        h2d(c, c->red_tmp.p, null_ews, (size_t)n_types * n_perm * 4);
        reduce_common(c, n_types, n_perm, c->red_tmp.as<float>(), mean_policy, inout);
    });
}

int pctsea_reduce_null_dev(pctsea_ctx* c, int32_t n_types, int32_t n_perm, const float* null_ews_dev,
                           int32_t mean_policy, pctsea_type_result* inout) {
    return guarded(c, [&] {
        PCT_REQUIRE(n_types >= 0 && n_perm >= 0 && inout != nullptr, "bad arguments");
        PCT_REQUIRE((int64_t)n_types * n_perm == 0 || null_ews_dev != nullptr, "null_ews_dev is NULL");
        rec(c, EV_BEGIN);
        reduce_common(c, n_types, n_perm, null_ews_dev, mean_policy, inout);
    });
}

int pctsea_analyze(pctsea_ctx* c, const pctsea_matrix* m, int32_t L, const int32_t* gene_idx, const float* abundance,
                   const pctsea_score_params* sprm, double min_score, const int32_t* label, int32_t n_types,
                   const pctsea_null_params* nprm, const int32_t* replay_perm, pctsea_type_result* out,
                   float* null_out, float* nullD_out, double* score_out, int32_t* n_genes_used_out, uint8_t* kept_out,
                   int32_t* order_out, int64_t* n_pass_out) {
    return guarded(c, [&] {
        PCT_REQUIRE(m != nullptr && out != nullptr, "matrix / out are NULL");
        NullPlan np = validate_analysis(m, min_score, label, n_types, nprm);
        shard_over_ranks(c, nprm, &np, nullD_out != nullptr);
        rec(c, EV_BEGIN);
        const int32_t* label_dev = resident_labels(c, m, label);
        // stage 1 shards over the ranks as well (gene-major kernel): every rank scores 1 / nranks of the cells and the
        // scores are all-gathered, instead of every rank scoring all of them
        struct Guard { pctsea_ctx* c; ~Guard() { c->shard_score = false; } } guard{ c };
        c->shard_score = np.pooled && c->comm_nranks > 1 && c->opt_shard_score;
        analyze_core(c, m, L, gene_idx, abundance, sprm, min_score, label_dev, n_types, nprm, np, replay_perm, out,
                     null_out, nullD_out, score_out, n_genes_used_out, kept_out, order_out, n_pass_out);
    });
}

int pctsea_analyze_batch(pctsea_ctx* c, const pctsea_matrix* m, int32_t n_lists, const int64_t* list_ptr,
                         const int32_t* gene_idx, const float* abundance, const pctsea_score_params* sprm,
                         double min_score, const int32_t* label, int32_t n_types, const pctsea_null_params* nprm,
                         pctsea_type_result* out, int64_t* n_pass_out, int32_t* status_out, float* null_out,
                         float* nullD_out) {
    return guarded(c, [&] {
        PCT_REQUIRE(m != nullptr && out != nullptr && status_out != nullptr, "matrix / out / status_out are NULL");
        PCT_REQUIRE(n_lists >= 0, "negative n_lists");
        PCT_REQUIRE(n_lists == 0 || list_ptr != nullptr, "list_ptr is NULL");
        PCT_REQUIRE(n_lists == 0 || list_ptr[0] == 0, "list_ptr must start at 0");
        for (int32_t i = 0; i < n_lists; i++) {
            PCT_REQUIRE(list_ptr[i + 1] >= list_ptr[i], "list_ptr is decreasing");
            PCT_REQUIRE(list_ptr[i + 1] - list_ptr[i] <= 32767, "list longer than the reference's short gene ids allow (32767)");
        }
        const NullPlan np = validate_analysis(m, min_score, label, n_types, nprm);
        PCT_REQUIRE(nprm->perm_mode == PCTSEA_PERM_PHILOX, "the batch entry point generates its permutations (perm_mode PHILOX)");
        rec(c, EV_BEGIN);
        const int32_t* label_dev = resident_labels(c, m, label);   // uploaded once for the whole batch
        pctsea_timings acc;
        memset(&acc, 0, sizeof acc);
        const size_t null_stride = (size_t)n_types * np.count;
        for (int32_t i = 0; i < n_lists; i++) {
            const int64_t b = list_ptr[i];
            const int32_t L = (int32_t)(list_ptr[i + 1] - b);
            pctsea_type_result* out_i = out + (size_t)i * n_types;
            int64_t npass = 0;
            status_out[i] = PCTSEA_OK;
            if (i > 0) rec(c, EV_BEGIN);
            try {
                analyze_core(c, m, L, gene_idx + b, abundance + b, sprm, min_score, label_dev, n_types, nprm, np, nullptr,
                             out_i, null_out ? null_out + i * null_stride : nullptr,
                             nullD_out ? nullD_out + i * null_stride : nullptr, nullptr, nullptr, nullptr, nullptr, &npass);
            } catch (const Error& e) {
                // A list that fails one of the reference's own input checks (PCTSEA.java:389-395,
                // model/SingleCell.java:347-352) fails alone, as one analysis of the web app's pool does.
                if (e.code != PCTSEA_ETOOFEW && e.code != PCTSEA_EMINGENES) throw;
                PCT_CUDA(cudaStreamSynchronize(c->stream));
                status_out[i] = e.code;
                c->err = "list " + std::to_string(i) + ": " + e.what();
                fill_nan_results(out_i, n_types);
                if (null_out) std::fill_n(null_out + i * null_stride, null_stride, std::numeric_limits<float>::quiet_NaN());
                if (nullD_out) std::fill_n(nullD_out + i * null_stride, null_stride, std::numeric_limits<float>::quiet_NaN());
            }
            if (n_pass_out) n_pass_out[i] = npass;
            // per-stage device times, copies and launches add up over the lists; n_pass / n_used are the last list's
            acc.setup_ms += c->tm.setup_ms; acc.score_ms += c->tm.score_ms; acc.rank_ms += c->tm.rank_ms;
            acc.observed_ms += c->tm.observed_ms; acc.labels_ms += c->tm.labels_ms; acc.null_ms += c->tm.null_ms;
            acc.reduce_ms += c->tm.reduce_ms; acc.total_ms += c->tm.total_ms;
            acc.launches += c->launches; acc.h2d_bytes += c->tm.h2d_bytes; acc.d2h_bytes += c->tm.d2h_bytes;
            acc.n_pass = c->tm.n_pass; acc.n_used = c->tm.n_used;
            memset(&c->tm, 0, sizeof c->tm);
            for (int k = 0; k < EV_COUNT; k++) c->ev_rec[k] = false;
            c->launches = 0;
        }
        c->tm = acc;
        c->launches = acc.launches;
    });
}

int pctsea_last_type_counts(pctsea_ctx* c, int32_t n_types, int32_t* n_cells_of_type, int32_t* n_cells_of_type_passing,
                            int64_t* n_cells_kept) {
    return guarded(c, [&] {
        PCT_REQUIRE(c->type_counts_T == n_types && n_types >= 0, "no type counts for this n_types: call pctsea_analyze first");
        std::vector<int32_t> tmp((size_t)2 * n_types + 1);
        PCT_CUDA(cudaMemcpyAsync(tmp.data(), c->type_counts.p, tmp.size() * 4, cudaMemcpyDeviceToHost, c->stream));
        PCT_CUDA(cudaStreamSynchronize(c->stream));
        if (n_cells_of_type) memcpy(n_cells_of_type, tmp.data(), (size_t)n_types * 4);
        if (n_cells_of_type_passing) memcpy(n_cells_of_type_passing, tmp.data() + n_types, (size_t)n_types * 4);
        if (n_cells_kept) *n_cells_kept = tmp[2 * (size_t)n_types];
    });
}

int pctsea_last_observed_curve(pctsea_ctx* c, int32_t type_id, int64_t n_used, float* a_out, float* b_out) {
    return guarded(c, [&] {
        PCT_REQUIRE(c->last_obs_Nu >= 0, "no observed enrichment on this ctx yet: call pctsea_analyze / pctsea_enrich first");
        PCT_REQUIRE(n_used == c->last_obs_Nu, "n_used differs from the last enrichment call");
        PCT_REQUIRE(type_id >= 0 && type_id < c->last_T, "type_id out of range");
        PCT_REQUIRE(a_out != nullptr && b_out != nullptr, "outputs are NULL");
        if (n_used == 0) return;
        // the fixed-point score buffers are free between calls: reuse them as the two staging arrays
        c->fx.reserve((size_t)n_used * 4 + 64);
        c->Sfx.reserve((size_t)n_used * 8 + 64);
        float* a_dev = c->fx.as<float>();
        float* b_dev = c->Sfx.as<float>();
        launch_obs_curve(c, type_id, n_used, c->last_T, a_dev, b_dev);
        d2h(c, a_out, a_dev, (size_t)n_used * 4);
        d2h(c, b_out, b_dev, (size_t)n_used * 4);
        PCT_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int pctsea_comm_unique_id(void* id_out) {
    if (!id_out) return PCTSEA_EINVAL;
    try { comm_unique_id(id_out); return PCTSEA_OK; }
    catch (const Error& e) { return e.code; }
    catch (...) { return PCTSEA_ENCCL; }
}

int pctsea_comm_init_rank(pctsea_ctx* c, int32_t nranks, int32_t rank, const void* id) {
    return guarded(c, [&] { comm_init_rank(c, nranks, rank, id); });
}

int pctsea_comm_destroy(pctsea_ctx* c) {
    return guarded(c, [&] { comm_destroy(c); });
}

int pctsea_last_null_dev(pctsea_ctx* c, float** null_dev, float** nullD_dev, int32_t* n_types, int32_t* perm_count) {
    if (!c) return PCTSEA_EINVAL;
    if (c->last_pool_P > 0) {   // multi-GPU analysis: the complete null, identical on every rank
        if (null_dev) *null_dev = c->null_pool.as<float>();
        if (nullD_dev) *nullD_dev = c->last_pool_dstat ? c->null_pool_d.as<float>() : nullptr;
        if (n_types) *n_types = c->last_T;
        if (perm_count) *perm_count = c->last_pool_P;
        return PCTSEA_OK;
    }
    if (null_dev) *null_dev = c->null_ews.as<float>();
    if (nullD_dev) *nullD_dev = c->null_d.as<float>();
    if (n_types) *n_types = c->last_T;
    if (perm_count) *perm_count = c->last_Pc;
    return PCTSEA_OK;
}

}  // extern "C"
