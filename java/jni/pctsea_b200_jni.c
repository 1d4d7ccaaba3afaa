/*
 * pctsea_b200_jni.c — the C side of edu.scripps.yates.pctsea.gpu.PctseaB200: one JNI function per native method, each
 * a direct call of the entry point of include/pctsea_b200.h named in its comment. Built into libpctsea_b200_jni.so
 * (java/jni/Makefile), which links against libpctsea_b200.so. No logic lives here beyond argument marshalling:
 * Java arrays are pinned with GetPrimitiveArrayCritical for the duration of the call (the library copies what it
 * needs to pinned staging memory / the device before returning), result structs are unpacked into double[13] rows.
 *
 * Errors: PCTSEA_EINVAL / PCTSEA_ETOOFEW / PCTSEA_EMINGENES -> IllegalArgumentException carrying pctsea_last_error
 * (the reference throws IllegalArgumentException with the same messages, PCTSEA.java:389-395,
 * model/SingleCell.java:347-352); PCTSEA_ENOMEM -> OutOfMemoryError; anything else -> IllegalStateException.
 */
#include <jni.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pctsea_b200.h"

#define JFN(ret, name) JNIEXPORT ret JNICALL Java_edu_scripps_yates_pctsea_gpu_PctseaB200_##name
#define CTX(h) ((pctsea_ctx*)(intptr_t)(h))
#define MAT(h) ((pctsea_matrix*)(intptr_t)(h))
#define RESULT_FIELDS 13

static void throw_rc(JNIEnv* env, pctsea_ctx* ctx, int rc, const char* what) {
    const char* cls = "java/lang/IllegalStateException";
    if (rc == PCTSEA_EINVAL || rc == PCTSEA_ETOOFEW || rc == PCTSEA_EMINGENES) cls = "java/lang/IllegalArgumentException";
    else if (rc == PCTSEA_ENOMEM) cls = "java/lang/OutOfMemoryError";
    const char* msg = ctx ? pctsea_last_error(ctx) : NULL;
    char buf[1024];
    snprintf(buf, sizeof buf, "%s: [%d] %s", what, rc, (msg && msg[0]) ? msg : "no message");
    jclass c = (*env)->FindClass(env, cls);
    if (c) (*env)->ThrowNew(env, c, buf);
}

/* pin / unpin a (possibly null) primitive array */
static void* pin(JNIEnv* env, jarray a) { return a ? (*env)->GetPrimitiveArrayCritical(env, a, NULL) : NULL; }
static void unpin(JNIEnv* env, jarray a, void* p, jint mode) { if (a && p) (*env)->ReleasePrimitiveArrayCritical(env, a, p, mode); }

static void result_to_rows(JNIEnv* env, const pctsea_type_result* r, jint n_types, jobjectArray rows) {
    for (jint t = 0; t < n_types; t++) {
        jdoubleArray row = (jdoubleArray)(*env)->GetObjectArrayElement(env, rows, t);
        if (!row) continue;
        const jdouble v[RESULT_FIELDS] = { r[t].ews, r[t].dstat, (double)r[t].supX, r[t].norm_supX, (double)r[t].sizeA,
                                           (double)r[t].sizeB, r[t].raw_sup, r[t].norm_ews, r[t].p_emp, r[t].fdr,
                                           r[t].sec_ews, (double)r[t].sec_supX, r[t].ks_d };
        (*env)->SetDoubleArrayRegion(env, row, 0, RESULT_FIELDS, v);
        (*env)->DeleteLocalRef(env, row);
    }
}

static void rows_to_result(JNIEnv* env, jobjectArray rows, jint n_types, pctsea_type_result* r) {
    memset(r, 0, (size_t)n_types * sizeof *r);
    for (jint t = 0; t < n_types; t++) {
        jdoubleArray row = (jdoubleArray)(*env)->GetObjectArrayElement(env, rows, t);
        if (!row) continue;
        jdouble v[RESULT_FIELDS];
        (*env)->GetDoubleArrayRegion(env, row, 0, RESULT_FIELDS, v);
        r[t].ews = (float)v[0]; r[t].dstat = (float)v[1]; r[t].supX = (int32_t)v[2]; r[t].norm_supX = v[3];
        r[t].sizeA = (int32_t)v[4]; r[t].sizeB = (int32_t)v[5]; r[t].raw_sup = (float)v[6]; r[t].norm_ews = (float)v[7];
        r[t].p_emp = v[8]; r[t].fdr = v[9]; r[t].sec_ews = (float)v[10]; r[t].sec_supX = (int32_t)v[11]; r[t].ks_d = v[12];
        (*env)->DeleteLocalRef(env, row);
    }
}

/* ---- ctx ------------------------------------------------------------------------------------------ */
JFN(jint, abiVersion)(JNIEnv* env, jclass cls) { (void)env; (void)cls; return pctsea_abi_version(); }

JFN(jlong, create)(JNIEnv* env, jclass cls, jint device) {   /* pctsea_create */
    (void)cls;
    pctsea_ctx* c = NULL;
    const int rc = pctsea_create(device, &c);
    if (rc != PCTSEA_OK) { throw_rc(env, NULL, rc, "pctsea_create (no usable CUDA device; there is no CPU fallback)"); return 0; }
    return (jlong)(intptr_t)c;
}

JFN(void, destroy)(JNIEnv* env, jclass cls, jlong ctx) { (void)env; (void)cls; pctsea_destroy(CTX(ctx)); }   /* pctsea_destroy */

JFN(jstring, lastError)(JNIEnv* env, jclass cls, jlong ctx) {   /* pctsea_last_error */
    (void)cls;
    return (*env)->NewStringUTF(env, pctsea_last_error(CTX(ctx)));
}

JFN(void, getTimings)(JNIEnv* env, jclass cls, jlong ctx, jfloatArray ms9, jlongArray counts5) {   /* pctsea_get_timings */
    (void)cls;
    pctsea_timings t;
    const int rc = pctsea_get_timings(CTX(ctx), &t);
    if (rc != PCTSEA_OK) { throw_rc(env, CTX(ctx), rc, "pctsea_get_timings"); return; }
    const jfloat ms[9] = { t.setup_ms, t.score_ms, t.rank_ms, t.observed_ms, t.labels_ms, t.null_ms, t.reduce_ms, t.total_ms, t.gather_ms };
    const jlong n[5] = { t.n_pass, t.n_used, t.launches, t.h2d_bytes, t.d2h_bytes };
    (*env)->SetFloatArrayRegion(env, ms9, 0, 9, ms);
    (*env)->SetLongArrayRegion(env, counts5, 0, 5, n);
}

JFN(void, setOption)(JNIEnv* env, jclass cls, jlong ctx, jint option, jlong value) {   /* pctsea_ctx_set_option */
    (void)cls;
    const int rc = pctsea_ctx_set_option(CTX(ctx), option, value);
    if (rc != PCTSEA_OK) throw_rc(env, CTX(ctx), rc, "pctsea_ctx_set_option");
}

/* ---- matrix --------------------------------------------------------------------------------------- */
JFN(jlong, loadCsr)(JNIEnv* env, jclass cls, jlong ctx, jlong nCells, jint nGenes, jlongArray rowPtr, jintArray colIdx,
                    jfloatArray val) {   /* pctsea_matrix_load_csr */
    (void)cls;
    pctsea_matrix* m = NULL;
    jlong* rp = pin(env, rowPtr);
    jint* ci = pin(env, colIdx);
    jfloat* vv = pin(env, val);
    const int rc = pctsea_matrix_load_csr(CTX(ctx), nCells, nGenes, (const int64_t*)rp, (const int32_t*)ci, vv, &m);
    unpin(env, val, vv, JNI_ABORT);
    unpin(env, colIdx, ci, JNI_ABORT);
    unpin(env, rowPtr, rp, JNI_ABORT);
    if (rc != PCTSEA_OK) { throw_rc(env, CTX(ctx), rc, "pctsea_matrix_load_csr"); return 0; }
    return (jlong)(intptr_t)m;
}

JFN(jlong, loadCsrDirect)(JNIEnv* env, jclass cls, jlong ctx, jlong nCells, jint nGenes, jobject rowPtr, jobject colIdx,
                          jobject val) {   /* pctsea_matrix_load_csr from direct buffers (more than 2^31 - 1 non-zeros) */
    (void)cls;
    pctsea_matrix* m = NULL;
    const int64_t* rp = (const int64_t*)(*env)->GetDirectBufferAddress(env, rowPtr);
    const int32_t* ci = (const int32_t*)(*env)->GetDirectBufferAddress(env, colIdx);
    const float* vv = (const float*)(*env)->GetDirectBufferAddress(env, val);
    if (!rp || (*env)->GetDirectBufferCapacity(env, rowPtr) < (jlong)(nCells + 1) * 8) {
        throw_rc(env, NULL, PCTSEA_EINVAL, "loadCsrDirect: rowPtr must be a direct buffer of (nCells + 1) longs in native byte order");
        return 0;
    }
    const int rc = pctsea_matrix_load_csr(CTX(ctx), nCells, nGenes, rp, ci, vv, &m);
    if (rc != PCTSEA_OK) { throw_rc(env, CTX(ctx), rc, "pctsea_matrix_load_csr"); return 0; }
    return (jlong)(intptr_t)m;
}

JFN(void, setLabels)(JNIEnv* env, jclass cls, jlong ctx, jlong matrix, jintArray label, jint nTypes) {   /* pctsea_matrix_set_labels */
    (void)cls;
    jint* lb = pin(env, label);
    const int rc = pctsea_matrix_set_labels(CTX(ctx), MAT(matrix), (const int32_t*)lb, nTypes);
    unpin(env, label, lb, JNI_ABORT);
    if (rc != PCTSEA_OK) throw_rc(env, CTX(ctx), rc, "pctsea_matrix_set_labels");
}

JFN(void, buildGeneIndex)(JNIEnv* env, jclass cls, jlong ctx, jlong matrix) {   /* pctsea_matrix_build_gene_index */
    (void)cls;
    const int rc = pctsea_matrix_build_gene_index(CTX(ctx), MAT(matrix));
    if (rc != PCTSEA_OK) throw_rc(env, CTX(ctx), rc, "pctsea_matrix_build_gene_index");
}

JFN(void, freeMatrix)(JNIEnv* env, jclass cls, jlong ctx, jlong matrix) {   /* pctsea_matrix_free */
    (void)env; (void)cls;
    pctsea_matrix_free(CTX(ctx), MAT(matrix));
}

/* ---- stages --------------------------------------------------------------------------------------- */
JFN(void, score)(JNIEnv* env, jclass cls, jlong ctx, jlong matrix, jintArray geneIdx, jfloatArray abundance, jint method,
                 jint minGenesCells, jdouble minCorr, jint jitterMode, jlong seed, jdoubleArray scoreOut,
                 jintArray nGenesUsedOut, jintArray nNonZeroOut, jbyteArray keptOut) {   /* pctsea_score */
    (void)cls;
    const jsize L = (*env)->GetArrayLength(env, geneIdx);
    pctsea_score_params prm = { method, minGenesCells, minCorr, 1, jitterMode, (uint64_t)seed };
    jint* gi = pin(env, geneIdx);
    jfloat* ab = pin(env, abundance);
    jdouble* so = pin(env, scoreOut);
    jint* nu = pin(env, nGenesUsedOut);
    jint* nz = pin(env, nNonZeroOut);
    jbyte* kp = pin(env, keptOut);
    const int rc = pctsea_score(CTX(ctx), MAT(matrix), L, (const int32_t*)gi, ab, &prm, so, (int32_t*)nu, (int32_t*)nz, (uint8_t*)kp);
    unpin(env, keptOut, kp, 0);
    unpin(env, nNonZeroOut, nz, 0);
    unpin(env, nGenesUsedOut, nu, 0);
    unpin(env, scoreOut, so, 0);
    unpin(env, abundance, ab, JNI_ABORT);
    unpin(env, geneIdx, gi, JNI_ABORT);
    if (rc != PCTSEA_OK) throw_rc(env, CTX(ctx), rc, "pctsea_score");
}

JFN(jlong, rank)(JNIEnv* env, jclass cls, jlong ctx, jdoubleArray score, jbyteArray kept, jdouble minScore,
                 jintArray orderOut) {   /* pctsea_rank */
    (void)cls;
    const jsize n = (*env)->GetArrayLength(env, score);
    int64_t npass = 0;
    jdouble* sc = pin(env, score);
    jbyte* kp = pin(env, kept);
    jint* od = pin(env, orderOut);
    const int rc = pctsea_rank(CTX(ctx), n, sc, (const uint8_t*)kp, minScore, (int32_t*)od, &npass);
    unpin(env, orderOut, od, 0);
    unpin(env, kept, kp, JNI_ABORT);
    unpin(env, score, sc, JNI_ABORT);
    if (rc != PCTSEA_OK) { throw_rc(env, CTX(ctx), rc, "pctsea_rank"); return 0; }
    return (jlong)npass;
}

JFN(void, enrich)(JNIEnv* env, jclass cls, jlong ctx, jdoubleArray score, jlong nUsed, jintArray order, jintArray label,
                  jint nTypes, jint nPerm, jint permBegin, jint permCount, jint permMode, jint ksMode, jint feistelRounds,
                  jlong seed, jint meanPolicy, jintArray replayPerm, jobjectArray result, jfloatArray nullOut,
                  jfloatArray nullDOut) {   /* pctsea_enrich */
    (void)cls;
    const jsize n = (*env)->GetArrayLength(env, score);
    pctsea_null_params prm = { nPerm, permBegin, permCount, permMode, ksMode, feistelRounds, (uint64_t)seed, meanPolicy, 0 };
    pctsea_type_result* res = (pctsea_type_result*)calloc((size_t)(nTypes > 0 ? nTypes : 1), sizeof *res);
    if (!res) { throw_rc(env, NULL, PCTSEA_ENOMEM, "enrich"); return; }
    jdouble* sc = pin(env, score);
    jint* od = pin(env, order);
    jint* lb = pin(env, label);
    jint* rp = pin(env, replayPerm);
    jfloat* ne = pin(env, nullOut);
    jfloat* nd = pin(env, nullDOut);
    const int rc = pctsea_enrich(CTX(ctx), n, sc, nUsed, (const int32_t*)od, (const int32_t*)lb, nTypes, &prm,
                                 (const int32_t*)rp, res, ne, nd);
    unpin(env, nullDOut, nd, 0);
    unpin(env, nullOut, ne, 0);
    unpin(env, replayPerm, rp, JNI_ABORT);
    unpin(env, label, lb, JNI_ABORT);
    unpin(env, order, od, JNI_ABORT);
    unpin(env, score, sc, JNI_ABORT);
    if (rc != PCTSEA_OK) throw_rc(env, CTX(ctx), rc, "pctsea_enrich");
    else result_to_rows(env, res, nTypes, result);
    free(res);
}

JFN(void, reduceNull)(JNIEnv* env, jclass cls, jlong ctx, jint nTypes, jint nPerm, jfloatArray nullEws, jint meanPolicy,
                      jobjectArray resultInOut) {   /* pctsea_reduce_null */
    (void)cls;
    pctsea_type_result* res = (pctsea_type_result*)calloc((size_t)(nTypes > 0 ? nTypes : 1), sizeof *res);
    if (!res) { throw_rc(env, NULL, PCTSEA_ENOMEM, "reduceNull"); return; }
    rows_to_result(env, resultInOut, nTypes, res);
    jfloat* ne = pin(env, nullEws);
    const int rc = pctsea_reduce_null(CTX(ctx), nTypes, nPerm, ne, meanPolicy, res);
    unpin(env, nullEws, ne, JNI_ABORT);
    if (rc != PCTSEA_OK) throw_rc(env, CTX(ctx), rc, "pctsea_reduce_null");
    else result_to_rows(env, res, nTypes, resultInOut);
    free(res);
}

/* ---- whole path ----------------------------------------------------------------------------------- */
JFN(jlong, analyze)(JNIEnv* env, jclass cls, jlong ctx, jlong matrix, jintArray geneIdx, jfloatArray abundance, jint method,
                    jint minGenesCells, jdouble minCorr, jint jitterMode, jlong scoreSeed, jdouble minScore, jintArray label,
                    jint nTypes, jint nPerm, jint permBegin, jint permCount, jint permMode, jint ksMode, jint feistelRounds,
                    jlong permSeed, jint meanPolicy, jint minPass, jintArray replayPerm, jobjectArray result,
                    jfloatArray nullOut, jfloatArray nullDOut, jdoubleArray scoreOut, jintArray nGenesUsedOut,
                    jbyteArray keptOut, jintArray orderOut) {   /* pctsea_analyze */
    (void)cls;
    const jsize L = (*env)->GetArrayLength(env, geneIdx);
    pctsea_score_params sprm = { method, minGenesCells, minCorr, 1, jitterMode, (uint64_t)scoreSeed };
    pctsea_null_params nprm = { nPerm, permBegin, permCount, permMode, ksMode, feistelRounds, (uint64_t)permSeed, meanPolicy, minPass };
    pctsea_type_result* res = (pctsea_type_result*)calloc((size_t)(nTypes > 0 ? nTypes : 1), sizeof *res);
    if (!res) { throw_rc(env, NULL, PCTSEA_ENOMEM, "analyze"); return 0; }
    int64_t npass = 0;
    jint* gi = pin(env, geneIdx);
    jfloat* ab = pin(env, abundance);
    jint* lb = pin(env, label);
    jint* rp = pin(env, replayPerm);
    jfloat* ne = pin(env, nullOut);
    jfloat* nd = pin(env, nullDOut);
    jdouble* so = pin(env, scoreOut);
    jint* nu = pin(env, nGenesUsedOut);
    jbyte* kp = pin(env, keptOut);
    jint* od = pin(env, orderOut);
    const int rc = pctsea_analyze(CTX(ctx), MAT(matrix), L, (const int32_t*)gi, ab, &sprm, minScore, (const int32_t*)lb, nTypes,
                                  &nprm, (const int32_t*)rp, res, ne, nd, so, (int32_t*)nu, (uint8_t*)kp, (int32_t*)od, &npass);
    unpin(env, orderOut, od, 0);
    unpin(env, keptOut, kp, 0);
    unpin(env, nGenesUsedOut, nu, 0);
    unpin(env, scoreOut, so, 0);
    unpin(env, nullDOut, nd, 0);
    unpin(env, nullOut, ne, 0);
    unpin(env, replayPerm, rp, JNI_ABORT);
    unpin(env, label, lb, JNI_ABORT);
    unpin(env, abundance, ab, JNI_ABORT);
    unpin(env, geneIdx, gi, JNI_ABORT);
    if (rc != PCTSEA_OK) throw_rc(env, CTX(ctx), rc, "pctsea_analyze");
    else result_to_rows(env, res, nTypes, result);
    free(res);
    return (jlong)npass;
}

JFN(void, analyzeBatch)(JNIEnv* env, jclass cls, jlong ctx, jlong matrix, jlongArray listPtr, jintArray geneIdx,
                        jfloatArray abundance, jint method, jint minGenesCells, jdouble minCorr, jint jitterMode,
                        jlong scoreSeed, jdouble minScore, jintArray label, jint nTypes, jint nPerm, jint ksMode,
                        jint feistelRounds, jlong permSeed, jint meanPolicy, jint minPass, jobjectArray result,
                        jlongArray nPassOut, jintArray statusOut, jfloatArray nullOut, jfloatArray nullDOut) {   /* pctsea_analyze_batch */
    (void)cls;
    const jsize nLists = (*env)->GetArrayLength(env, listPtr) - 1;
    pctsea_score_params sprm = { method, minGenesCells, minCorr, 1, jitterMode, (uint64_t)scoreSeed };
    pctsea_null_params nprm = { nPerm, 0, 0, PCTSEA_PERM_PHILOX, ksMode, feistelRounds, (uint64_t)permSeed, meanPolicy, minPass };
    const size_t nres = (size_t)(nLists > 0 ? nLists : 1) * (size_t)(nTypes > 0 ? nTypes : 1);
    pctsea_type_result* res = (pctsea_type_result*)calloc(nres, sizeof *res);
    if (!res) { throw_rc(env, NULL, PCTSEA_ENOMEM, "analyzeBatch"); return; }
    jlong* lp = pin(env, listPtr);
    jint* gi = pin(env, geneIdx);
    jfloat* ab = pin(env, abundance);
    jint* lb = pin(env, label);
    jlong* np = pin(env, nPassOut);
    jint* st = pin(env, statusOut);
    jfloat* ne = pin(env, nullOut);
    jfloat* nd = pin(env, nullDOut);
    const int rc = pctsea_analyze_batch(CTX(ctx), MAT(matrix), nLists, (const int64_t*)lp, (const int32_t*)gi, ab, &sprm, minScore,
                                        (const int32_t*)lb, nTypes, &nprm, res, (int64_t*)np, (int32_t*)st, ne, nd);
    unpin(env, nullDOut, nd, 0);
    unpin(env, nullOut, ne, 0);
    unpin(env, statusOut, st, 0);
    unpin(env, nPassOut, np, 0);
    unpin(env, label, lb, JNI_ABORT);
    unpin(env, abundance, ab, JNI_ABORT);
    unpin(env, geneIdx, gi, JNI_ABORT);
    unpin(env, listPtr, lp, JNI_ABORT);
    if (rc != PCTSEA_OK) throw_rc(env, CTX(ctx), rc, "pctsea_analyze_batch");
    else
        for (jsize i = 0; i < nLists; i++) {
            jobjectArray rows = (jobjectArray)(*env)->GetObjectArrayElement(env, result, i);
            if (!rows) continue;
            result_to_rows(env, res + (size_t)i * nTypes, nTypes, rows);
            (*env)->DeleteLocalRef(env, rows);
        }
    free(res);
}

/* ---- by-products of the last analyze ---------------------------------------------------------------- */
JFN(jlong, lastTypeCounts)(JNIEnv* env, jclass cls, jlong ctx, jint nTypes, jintArray nCellsOfType,
                           jintArray nCellsOfTypePassing) {   /* pctsea_last_type_counts */
    (void)cls;
    int64_t kept = 0;
    jint* a = pin(env, nCellsOfType);
    jint* b = pin(env, nCellsOfTypePassing);
    const int rc = pctsea_last_type_counts(CTX(ctx), nTypes, (int32_t*)a, (int32_t*)b, &kept);
    unpin(env, nCellsOfTypePassing, b, 0);
    unpin(env, nCellsOfType, a, 0);
    if (rc != PCTSEA_OK) { throw_rc(env, CTX(ctx), rc, "pctsea_last_type_counts"); return 0; }
    return (jlong)kept;
}

JFN(void, lastObservedCurve)(JNIEnv* env, jclass cls, jlong ctx, jint typeId, jlong nUsed, jfloatArray aOut,
                             jfloatArray bOut) {   /* pctsea_last_observed_curve */
    (void)cls;
    jfloat* a = pin(env, aOut);
    jfloat* b = pin(env, bOut);
    const int rc = pctsea_last_observed_curve(CTX(ctx), typeId, nUsed, a, b);
    unpin(env, bOut, b, 0);
    unpin(env, aOut, a, 0);
    if (rc != PCTSEA_OK) throw_rc(env, CTX(ctx), rc, "pctsea_last_observed_curve");
}

/* ---- multi-GPU -------------------------------------------------------------------------------------- */
JFN(void, commUniqueId)(JNIEnv* env, jclass cls, jbyteArray idOut) {   /* pctsea_comm_unique_id */
    (void)cls;
    unsigned char id[PCTSEA_COMM_ID_BYTES];
    const int rc = pctsea_comm_unique_id(id);
    if (rc != PCTSEA_OK) { throw_rc(env, NULL, rc, "pctsea_comm_unique_id (libnccl.so.2 not loadable?)"); return; }
    (*env)->SetByteArrayRegion(env, idOut, 0, PCTSEA_COMM_ID_BYTES, (const jbyte*)id);
}

JFN(void, commInitRank)(JNIEnv* env, jclass cls, jlong ctx, jint nRanks, jint rank, jbyteArray id) {   /* pctsea_comm_init_rank */
    (void)cls;
    unsigned char buf[PCTSEA_COMM_ID_BYTES];
    (*env)->GetByteArrayRegion(env, id, 0, PCTSEA_COMM_ID_BYTES, (jbyte*)buf);   /* copied: the call blocks on the other ranks */
    const int rc = pctsea_comm_init_rank(CTX(ctx), nRanks, rank, buf);
    if (rc != PCTSEA_OK) throw_rc(env, CTX(ctx), rc, "pctsea_comm_init_rank");
}

JFN(void, commDestroy)(JNIEnv* env, jclass cls, jlong ctx) {   /* pctsea_comm_destroy */
    (void)cls;
    const int rc = pctsea_comm_destroy(CTX(ctx));
    if (rc != PCTSEA_OK) throw_rc(env, CTX(ctx), rc, "pctsea_comm_destroy");
}
