/*
 * pctsea_b200.h — C ABI of libpctsea_b200.so: the B200 (sm_100a) implementation of PCTSEA's hot path
 * (per-cell scoring -> ranking -> weighted-KS enrichment + label-permutation null -> norm/p/FDR).
 *
 * The reference (proteomicsyates/pctsea-parent) has no FFI/plugin interface for this path: the path is a
 * set of private methods of pctsea-core's PCTSEA class working on Java object graphs. Each entry point
 * below names the reference code whose *body* it replaces; a Java host binds them through JNI or Panama
 * (INTEGRATION.md shows the stubs). Paths are relative to
 * pctsea-core/src/main/java/edu/scripps/yates/pctsea/ of the reference.
 *
 * Conventions: every function returns 0 (PCTSEA_OK) or a negative PCTSEA_E* code; pctsea_last_error(ctx)
 * returns a message valid until the next call on that ctx. All pointers are HOST pointers unless the
 * parameter name ends in _dev. Caller allocates every output buffer. No C++ exceptions cross the ABI.
 * There is NO CPU fallback: without a CUDA device pctsea_create fails with PCTSEA_ECUDA.
 * A ctx is bound to one device and is not re-entrant; multi-GPU = one ctx (one process or thread) per GPU joined
 * by pctsea_comm_init_rank: pctsea_analyze then shards the permutations over the ranks and all-gathers the null
 * slices with NCCL inside the call. (A host that prefers its own collective leaves the communicator out and
 * passes explicit slices in pctsea_null_params.perm_begin/perm_count.)
 */
#ifndef PCTSEA_B200_H
#define PCTSEA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCTSEA_ABI_VERSION 2

#define PCTSEA_OK         0
#define PCTSEA_EINVAL    (-1)
#define PCTSEA_ECUDA     (-2)
#define PCTSEA_ENCCL     (-3)  /* libnccl could not be loaded, or an NCCL call failed (pctsea_comm_*, multi-GPU pctsea_analyze) */
#define PCTSEA_ETOOFEW   (-4)  /* fewer than min_pass cells pass the score threshold, PCTSEA.java:389-395 */
#define PCTSEA_EMINGENES (-5)  /* min_genes_cells > list size, model/SingleCell.java:347-352 */
#define PCTSEA_ENOMEM    (-6)

/* model/ScoringMethod.java:8-23 (only the two methods on the hot path) */
#define PCTSEA_METHOD_PEARSONS_CORRELATION 0
#define PCTSEA_METHOD_SIMPLE_SCORE         1

/* zero-variance policy for model/SingleCell.java:433-448 (the reference adds unseeded Math.random()/1e6) */
#define PCTSEA_JITTER_NAN    0  /* such cells score NaN */
#define PCTSEA_JITTER_PHILOX 1  /* U[0,1)/1e6 from Philox4x32-10 keyed by (seed, cell, vector, element) */

#define PCTSEA_PERM_REPLAY 0    /* consume uploaded permutation index arrays (validation) */
#define PCTSEA_PERM_PHILOX 1    /* generated: position x of permutation p carries the class at index pi_p(x) of the
                                   SIZE-SORTED arrangement of the ranked labels (the n_types + 1 classes — the last one
                                   the unknown-type cells — in ascending order of their cell counts, ties by id, each a
                                   contiguous run); pi_p = keyed Feistel bijection of [0, n_used) with arithmetic round
                                   functions, one Philox4x32-10 key per round and permutation. A pure function of
                                   (seed, global permutation index, x): identical for any GPU count or slicing. */

/* Rounds of the keyed Feistel bijection behind PCTSEA_PERM_PHILOX. The class of a position is a range test on the
 * high digit of pi_p(x), which exposes the unevenness of a random round function: measured on the null of the
 * enrichment score against the JDK shuffle (standard deviation, 99th percentile and two-sample KS per cell type,
 * DESIGN.md), 3 and 4 rounds inflate the null's spread by 2 %, 5 and more are indistinguishable. */
#define PCTSEA_DEFAULT_FEISTEL_ROUNDS 5

#define PCTSEA_KS_EXACT 0       /* sequential fp32/fp64 emulation of EnrichmentWeigthedScoreParallel.run */
#define PCTSEA_KS_FAST  1       /* needs PERM_PHILOX. Scores quantised to int32 fixed point (scale 2^F, F the largest
                                   with the sum of |fx| over the ranking below 2^31), exact integer running sums,
                                   supremum evaluated at hit / pre-hit positions only, each candidate one fp32 FMA.
                                   Not the reference's arithmetic: within 2e-5 of the real-number value for rankings
                                   of up to 2 M cells (the reference's own fp32 accumulators drift by 1e-5 at 220 k
                                   cells and 2e-3 at 2 M). Semantic differences beyond rounding: a zero denominator
                                   gives ES 0 (reference: NaN / Inf); equal |diff| resolves to the positive one
                                   (reference: first position); assumes all ranked scores share a sign (true on the
                                   reference's path: ScoreThreshold passes one side of the threshold). The observed
                                   ES always uses the reference's float chain. */

typedef struct pctsea_ctx    pctsea_ctx;
typedef struct pctsea_matrix pctsea_matrix;

typedef struct pctsea_score_params {
    int32_t  method;           /* PCTSEA_METHOD_* */
    int32_t  min_genes_cells;  /* InputParameters.MIN_GENES_CELLS */
    double   min_corr;         /* InputParameters.MIN_CORRELATION */
    int32_t  has_min_corr;     /* 0 = null Double (never on the reference's path) */
    int32_t  jitter_mode;      /* PCTSEA_JITTER_* */
    uint64_t seed;
} pctsea_score_params;

typedef struct pctsea_null_params {
    int32_t  n_perm;           /* P: InputParameters.PERM, total over all ranks */
    int32_t  perm_begin;       /* first GLOBAL permutation index computed by this call */
    int32_t  perm_count;       /* permutations computed by this call; 0 = all n_perm */
    int32_t  perm_mode;        /* PCTSEA_PERM_* */
    int32_t  ks_mode;          /* PCTSEA_KS_* */
    int32_t  feistel_rounds;   /* 0 = default (PCTSEA_DEFAULT_FEISTEL_ROUNDS); 2..32 */
    uint64_t seed;
    int32_t  mean_policy;      /* Maths.mean(TFloatList): 0 = fp32 running sum (default), 1 = fp64 sum */
    int32_t  min_pass;         /* PCTSEA_ETOOFEW guard; the reference uses 1000 (PCTSEA.java:138) */
} pctsea_null_params;

/* One per cell type: the fields model/CellTypeClassification.java receives through setEnrichment (:91),
 * setSizeA/B, setNormalizedEnrichmentScore (:443), setEnrichmentSignificance (:246), setEnrichmentFDR (:447). */
typedef struct pctsea_type_result {
    float   ews;        /* observed enrichment score (compensated supremum) */
    float   dstat;      /* raw supremum * sqrt(sizeA*sizeB/(sizeA+sizeB)) */
    int32_t supX;       /* 1-based position of the supremum; -1 when NaN */
    double  norm_supX;  /* supX / Nu; -1 when NaN */
    int32_t sizeA;      /* cells of the type among the Nu ranked cells */
    int32_t sizeB;      /* Nu - sizeA */
    float   raw_sup;    /* supremum before compensation */
    float   norm_ews;   /* a11 */
    double  p_emp;      /* a10 */
    double  fdr;        /* a12 */
    /* observed-only extras of EnrichmentWeigthedScoreParallel.run (set for valid types, else 0 / 0 / NaN) */
    float   sec_ews;    /* setSecondaryEnrichment (:244-312): secondary supremum, 0 when none */
    int32_t sec_supX;   /* its 1-based position, 0 when none */
    double  ks_d;       /* setKSTestPvalue (:211-212,228) = commons-math kolmogorovSmirnovStatistic of the type's
                           scores vs the others; (float) of it is setEnrichmentUnweightedScore (PCTSEA.java:2606-2616) */
} pctsea_type_result;

/* Per-stage device times (CUDA events on the ctx stream) of the most recent call. */
typedef struct pctsea_timings {
    float   setup_ms;     /* list upload + lookup tables */
    float   score_ms;     /* stage 1 kernel */
    float   rank_ms;      /* stage 2: compaction + radix sort + gather */
    float   observed_ms;  /* stage 3 observed (exact): span of its kernels, which run on a second stream beside
                             labels_ms + null_ms (so the stage times do not add up to total_ms) */
    float   labels_ms;    /* EXACT null only: permuted label generation (the FAST null generates them in-kernel) */
    float   null_ms;      /* the null: labels + KS of every (type, permutation) */
    float   reduce_ms;    /* a10-a12 */
    float   total_ms;     /* first to last event, includes the copies issued by the call */
    int64_t n_pass;       /* Np */
    int64_t n_used;       /* Nu = Np - 1 */
    int64_t launches;     /* kernels launched by the call */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    float   gather_ms;    /* multi-GPU: NCCL all-gather of the null slices + reorder (0 without a communicator) */
    float   reserved0;
} pctsea_timings;

int         pctsea_abi_version(void);
int         pctsea_create(int device_id, pctsea_ctx** out);
int         pctsea_destroy(pctsea_ctx* ctx);
const char* pctsea_last_error(pctsea_ctx* ctx);
int         pctsea_get_timings(pctsea_ctx* ctx, pctsea_timings* out);

/* Testing and tuning switches of a ctx (none is needed in production; results never depend on them except where
 * noted). Returns PCTSEA_EINVAL for an unknown option or a value out of range. */
#define PCTSEA_OPT_OBS_OVERLAP       1  /* observed-statistics kernels: 0 = on the main stream before the null; on a second
                                           stream beside the null, queued 1 = behind / 2 (default) = ahead of the null kernel */
#define PCTSEA_OPT_SCORE_PATH        2  /* 0 (default): gene-major kernel when the matrix has its index; 1: always the CSR
                                           scan (the two differ in the last bits of a score, see build_gene_index) */
#define PCTSEA_OPT_GM_UNROLL         3  /* value loads in flight per lane of the gene-major kernel: 1, 2, 4, 8 (0 = default) */
#define PCTSEA_OPT_SCORE_WARPS       4  /* warps per CTA of the CSR-scan kernel (0 = as many as fit) */
#define PCTSEA_OPT_PERM_CHUNK        5  /* permutations per label chunk of the EXACT null (0 = from the free memory) */
#define PCTSEA_OPT_NULL_THREADS      6  /* threads (ranking segments) per permutation of the null kernel (0 = as many as fit) */
#define PCTSEA_OPT_NULL_CTAS_PER_SM  7  /* resident CTAs per SM of the null kernel (0 = as many as fit) */
#define PCTSEA_OPT_NULL_MIN_THREADS  8  /* cell types are processed in slices when fewer threads would fit (0 = 192) */
#define PCTSEA_OPT_DEBUG_GAPS        9  /* 1: report idle gaps between the stages on stderr */
#define PCTSEA_OPT_SHARD_SCORE      10  /* multi-GPU pctsea_analyze with the gene-major index: 1 (default) every rank scores
                                           1 / nranks of the cells and the scores are all-gathered; 0: every rank scores all */
#define PCTSEA_OPT_OBS_CHAIN_MB     11  /* device memory for the observed accumulator history in MiB (0 = 6144): 8 bytes per
                                           (ranked cell, type); the types go through it in slices when they do not fit */
#define PCTSEA_OPT_NULL_FINE_ENTRIES 12 /* null kernel's class table: at most value - 1 fine entries for the small classes
                                           (0 = as many as fit the CTA's shared memory; 1 = none: every bucket coarse) */
#define PCTSEA_OPT_SHARD_OBSERVED   13  /* multi-GPU pctsea_analyze: 1 (default) every rank computes the observed statistics of
                                           1 / nranks of the cell types and the records are all-gathered; 0: every rank all */
int         pctsea_ctx_set_option(pctsea_ctx* ctx, int32_t option, int64_t value);
/* cudaStream_t the ctx launches on (so a caller can time or order against it). */
void*       pctsea_stream(pctsea_ctx* ctx);

/* ---- matrix: loaded once per dataset, device-resident CSR ---------------------------------------
 * Replaces GeneExpressionsRetriever.getSingleCellExpressionsFromDB (GeneExpressionsRetriever.java:120-274)
 * + the per-cell sparse vectors of model/SingleCell.java:49-104. Rows are cells in DB order (= tie-break
 * order), columns are dataset genes. Columns within a row must be unique; storage order of a row is the
 * canonical order of the Pearson update. Both entry points check the CSR once on the device (row_ptr starts at
 * 0, is non-decreasing and ends at nnz; every column lies in [0, n_genes)) and return PCTSEA_EINVAL otherwise:
 * the scoring kernel indexes its on-chip gene tables with the stored columns unchecked. */
int pctsea_matrix_load_csr(pctsea_ctx* ctx, int64_t n_cells, int32_t n_genes, const int64_t* row_ptr,
                           const int32_t* col_idx, const float* val, pctsea_matrix** out);
/* Adopt device arrays without copying (they must outlive the matrix and stay unmodified; col_idx_dev and
 * val_dev must be 16-byte aligned, as any cudaMalloc / framework allocation is). */
int pctsea_matrix_wrap_device(pctsea_ctx* ctx, int64_t n_cells, int32_t n_genes, const int64_t* row_ptr_dev,
                              const int32_t* col_idx_dev, const float* val_dev, pctsea_matrix** out);
/* Optional, once per matrix: build the gene-major copy that lets the scoring stage read only the list's genes
 * instead of scanning the whole CSR (the role of the per-list SingleCell.expressionIndexes the reference fills from
 * MongoDB, model/SingleCell.java:49-104, without a per-list rebuild). Costs 8 B * n_genes * n_cells/32 + 4 B * nnz
 * of device memory (config B: 7 GB next to the 6.5 GB CSR). With it pctsea_score / pctsea_analyze[_batch] run the
 * lane-per-cell kernel, whose pair order is ascending gene (matrix column) order and whose Pearson sums are the
 * shifted-data form (oracle: orc_pearson_shifted); without it they run the CSR scan (pair order = row storage
 * order, two-pass centred sums; oracle: orc_pearson_strided32). Both are within 1e-9 relative of the reference's
 * sequential SimpleRegression update; they differ from each other in the last bits. */
int pctsea_matrix_build_gene_index(pctsea_ctx* ctx, pctsea_matrix* m);
/* Cell-type ids per cell in [-1, n_types); -1 = unknown type (model/SingleCell.java:224-229). */
int pctsea_matrix_set_labels(pctsea_ctx* ctx, pctsea_matrix* m, const int32_t* label, int32_t n_types);
int pctsea_matrix_free(pctsea_ctx* ctx, pctsea_matrix* m);

/* ---- stage 1: PCTSEA.filterByMinimumNumberOfGenesExpressedInEachCell (PCTSEA.java:544-572) +
 * calculateScoresToRankSingleCells loop (PCTSEA.java:2305-2343) -> SingleCell.hasMinNumberOfExpressedGenes
 * (model/SingleCell.java:344-369), calculateCorrelation / calculateSimpleScore (:380-476).
 * gene_idx[L]: matrix column per list entry (-1 = gene absent from the dataset); entries must be distinct. */
int pctsea_score(pctsea_ctx* ctx, const pctsea_matrix* m, int32_t L, const int32_t* gene_idx,
                 const float* abundance, const pctsea_score_params* prm, double* score_out /*[N]*/,
                 int32_t* n_genes_used_out /*[N] or NULL*/, int32_t* n_nonzero_out /*[N] or NULL*/,
                 uint8_t* kept_out /*[N] or NULL*/);

/* ---- stage 2: ScoreThreshold.passThreshold / getSingleCellsPassingThresholdSortedByScore
 * (scoring/ScoreThreshold.java:57-75,98-106), PCTSEAUtils.sortByScoreDescending (utils/PCTSEAUtils.java:44-53).
 * order_out[0 .. n_pass-2] = KS list (score desc by Double.compare, ties by cell index asc);
 * order_out[n_pass-1] = the cell dropped by subList(0,size-1) (PCTSEA.java:421-424). */
int pctsea_rank(pctsea_ctx* ctx, int64_t n_cells, const double* score, const uint8_t* kept /*or NULL*/,
                double min_score, int32_t* order_out /*[N]*/, int64_t* n_pass_out);

/* ---- stage 3: PCTSEA.calculateEnrichmentScore (PCTSEA.java:2454-2488) ->
 * EnrichmentWeigthedScoreParallel.run (utils/parallel/EnrichmentWeigthedScoreParallel.java:72-353) for the
 * observed labels and, per permutation, PCTSEA.calculateSignificanceByCellTypesPermutations
 * (PCTSEA.java:1380-1549). order[n_used] indexes score/label. replay_perm: [perm_count][n_used] int32 index
 * arrays (label at rank i = label of rank replay[p][i]) when perm_mode == REPLAY, else NULL.
 * null_out / nullD_out: [n_types][perm_count] row-major or NULL. When perm_count covers all n_perm the
 * a10-a12 fields of out[] are filled; otherwise they are NaN and the caller gathers the null slices and
 * calls pctsea_reduce_null. */
int pctsea_enrich(pctsea_ctx* ctx, int64_t n_cells, const double* score, int64_t n_used, const int32_t* order,
                  const int32_t* label, int32_t n_types, const pctsea_null_params* prm,
                  const int32_t* replay_perm, pctsea_type_result* out /*[n_types]*/, float* null_out,
                  float* nullD_out);

/* a10-a12 from a supplied null (also the -load_random path, PCTSEA.java:1385-1387): reads out[t].ews,
 * fills norm_ews, p_emp, fdr. null is [n_types][n_perm]. */
int pctsea_reduce_null(pctsea_ctx* ctx, int32_t n_types, int32_t n_perm, const float* null_ews,
                       int32_t mean_policy, pctsea_type_result* inout);
int pctsea_reduce_null_dev(pctsea_ctx* ctx, int32_t n_types, int32_t n_perm, const float* null_ews_dev,
                           int32_t mean_policy, pctsea_type_result* inout);

/* ---- the whole path in one call: the body of the per-schema loop PCTSEA.java:348-434 between the list
 * being known and the CellTypeClassification fields being set. Matrix (and, if label == NULL, its labels)
 * are resident; list in, per-type results out. Optional outputs may be NULL. */
int pctsea_analyze(pctsea_ctx* ctx, const pctsea_matrix* m, int32_t L, const int32_t* gene_idx,
                   const float* abundance, const pctsea_score_params* sprm, double min_score,
                   const int32_t* label /*[N] or NULL*/, int32_t n_types, const pctsea_null_params* nprm,
                   const int32_t* replay_perm, pctsea_type_result* out /*[n_types]*/,
                   float* null_out, float* nullD_out, double* score_out, int32_t* n_genes_used_out,
                   uint8_t* kept_out, int32_t* order_out, int64_t* n_pass_out);
/* ---- many input lists against one resident matrix (BASELINE config E; SURVEY.md §8 b "batched"). The reference
 * has no batch call: the web front-end submits one PCTSEA.run() per list to a 40-thread pool
 * (pctseaweb/src/main/java/edu/scripps/yates/pctsea/views/analyze/AnalyzeView.java:163,747-766) and each run repeats the per-schema loop PCTSEA.java:348-434. Here
 * the lists of one GPU's share go through one call: list i is gene_idx / abundance[list_ptr[i] .. list_ptr[i+1]),
 * labels are uploaded (or taken from the matrix) once, and every list gets the same parameters and the same
 * permutation stream, so out[i] is bit-identical to what pctsea_analyze returns for list i alone.
 * out: [n_lists][n_types]; n_pass_out: [n_lists] or NULL; null_out / nullD_out: [n_lists][n_types][perm_count] or
 * NULL. status_out[i] = PCTSEA_OK, or PCTSEA_ETOOFEW / PCTSEA_EMINGENES when list i fails the reference's own input
 * checks (its results are then NaN; the other lists are unaffected). Any other error aborts the call. Lists are
 * independent, so multi-GPU = each rank passes its share of the lists (no collective). pctsea_get_timings reports
 * the sums over the lists. */
int pctsea_analyze_batch(pctsea_ctx* ctx, const pctsea_matrix* m, int32_t n_lists, const int64_t* list_ptr /*[n_lists+1]*/,
                         const int32_t* gene_idx, const float* abundance, const pctsea_score_params* sprm,
                         double min_score, const int32_t* label /*[N] or NULL*/, int32_t n_types,
                         const pctsea_null_params* nprm, pctsea_type_result* out, int64_t* n_pass_out,
                         int32_t* status_out /*[n_lists]*/, float* null_out, float* nullD_out);

/* Inputs of PCTSEA.calculateHyperGeometricStatistics (PCTSEA.java:2169-2264) from the last pctsea_analyze on
 * this ctx: per type, the cells surviving the min_genes_cells filter (numCellsOfType) and those of them passing
 * the score threshold (numCellsOfTypeWithPositiveCorrelation); n_cells_kept = the population size N. A type
 * with n_cells_of_type == 0 has no CellTypeClassification in the reference (PCTSEA.java:2172-2174). */
int pctsea_last_type_counts(pctsea_ctx* ctx, int32_t n_types, int32_t* n_cells_of_type /*[n_types] or NULL*/,
                            int32_t* n_cells_of_type_passing /*[n_types] or NULL*/, int64_t* n_cells_kept /*or NULL*/);

/* The observed running-sum curves of one cell type from the last pctsea_analyze / pctsea_enrich on this ctx
 * (the accumulator history is still resident): a_out[i] = weighted fraction of the type's cells at ranked
 * positions <= i, b_out[i] = the same for all other cells, i in [0, n_used) — the two series
 * EnrichmentWeigthedScoreParallel.java:138-186 builds and :355-410 writes to "<type>_ews" (position i is the
 * reference's x = i + 1). Bit-identical to the reference's a / b, including the 0.0f a series keeps until the
 * type's first cell. n_used must be the n_used of that call. */
int pctsea_last_observed_curve(pctsea_ctx* ctx, int32_t type_id, int64_t n_used, float* a_out, float* b_out);

/* ---- multi-GPU: permutations shard over the GPUs of one box; the null slices are the only data exchanged
 * (SURVEY.md §8 e; the reference is single-process, PCTSEA.java:1398-1429 runs all permutations in one JVM).
 * One ctx per GPU (one process per GPU, or one thread per GPU inside one JVM). Rank 0 obtains an id with
 * pctsea_comm_unique_id and hands the 128 bytes to the other ranks by any means the host has (a socket, a file,
 * MPI, torch.distributed); then EVERY rank calls pctsea_comm_init_rank (collective, blocks until all ranks joined).
 * NCCL is loaded at run time (libnccl.so.2); without it these calls return PCTSEA_ENCCL and everything else works.
 *
 * With a communicator attached, pctsea_analyze must be called by all ranks with the same list and parameters,
 * perm_begin = 0 and perm_count = 0 ("all n_perm permutations"): rank r computes the r-th contiguous share of the
 * global permutation indices (sizes differ by at most one), the slices are all-gathered over NVLink on the ctx
 * stream, and a10-a12 run on the complete null on every rank (with the gene-major index stage 1 shards over the
 * ranks too: every rank scores 1 / nranks of the cells, 16 bytes per cell are all-gathered) — out[], null_out and nullD_out ([n_types][n_perm])
 * are bit-identical on every rank to what a single-GPU call returns. pctsea_analyze_batch, pctsea_enrich and the
 * stage calls stay per-rank (independent lists shard without a collective). */
#define PCTSEA_COMM_ID_BYTES 128
int pctsea_comm_unique_id(void* id_out /*[PCTSEA_COMM_ID_BYTES]*/);
int pctsea_comm_init_rank(pctsea_ctx* ctx, int32_t nranks, int32_t rank, const void* id /*[PCTSEA_COMM_ID_BYTES]*/);
int pctsea_comm_destroy(pctsea_ctx* ctx);   /* also done by pctsea_destroy */

/* Device pointers of the null slice left by the last pctsea_analyze / pctsea_enrich on this ctx:
 * [n_types][perm_count] fp32 each. Valid until the next call on the ctx (for a host that runs its own collective).
 * After a multi-GPU pctsea_analyze: the complete null [n_types][n_perm] (nullD_dev is NULL unless that call was
 * given a nullD_out). */
int pctsea_last_null_dev(pctsea_ctx* ctx, float** null_dev, float** nullD_dev, int32_t* n_types,
                         int32_t* perm_count);

#ifdef __cplusplus
}
#endif
#endif /* PCTSEA_B200_H */
