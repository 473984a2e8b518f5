/*
 * SealCuda.c -- JNI shim between a GPU-backed jgi.Seal (Java) and the Seal entry points of libbbduk_b200.so
 * (include/seal_b200.h). Pure marshalling, same conventions as jni/BBDukCuda.c: arrays are copied with
 * Get / Set*ArrayRegion into C buffers, so no critical section is held across a CUDA call.
 * NOT compiled against a real JDK in this repository's image (tests compile it against tests/stubs/jni.h).
 *
 * Java side: Seal.spawnLoadThreads (jgi/Seal.java:1261-1553) hands every reference sequence to addRefNative and
 * calls finalizeNative instead of running LoadThreads; ProcessThread.run (jgi/Seal.java:2011-2310) flattens a
 * ListNum<Read> into bases[] + offsets[] after its quality / length preamble and replaces the block from
 * "Do kmer matching" (:2186-2276) by one processNative call; scaffoldReadCounts / BaseCounts / FragCounts /
 * AmbigReadCounts (:2431-2441) are read once at the end with scaffoldCountsNative.
 */
#include <jni.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "seal_b200.h"

#define H(x) ((seal_handle *)(intptr_t)(x))

/* cfg = {k, rcomp, maskMiddle, midMaskLen, forbidNs, hdist, speed, qskip, rskip, restrictLeft, restrictRight,
 *        ambigMode, matchMode, keepPairsTogether, clearzone, minKmerHits, idsStride}; fractions separately */
JNIEXPORT jlong JNICALL Java_jgi_SealGPU_createNative(JNIEnv *env, jclass cls, jintArray jcfg, jfloat clearzoneFraction,
                                                      jfloat minKmerFraction, jint device) {
    jint v[17];
    seal_cfg c;
    seal_handle *h = NULL;
    if ((*env)->GetArrayLength(env, jcfg) < 17) return 0;
    (*env)->GetIntArrayRegion(env, jcfg, 0, 17, v);
    seal_b200_cfg_default(&c);
    c.k = v[0];
    c.rcomp = v[1];
    c.mask_middle = v[2];
    c.mid_mask_len = v[3];
    c.forbid_ns = v[4];
    c.hdist = v[5];
    c.speed = v[6];
    c.qskip = v[7];
    c.rskip = v[8];
    c.restrict_left = v[9];
    c.restrict_right = v[10];
    c.ambig_mode = v[11];
    c.match_mode = v[12];
    c.keep_pairs_together = v[13];
    c.clearzone = v[14];
    c.min_kmer_hits = v[15];
    c.ids_stride = v[16];
    c.clearzone_fraction = clearzoneFraction;
    c.min_kmer_fraction = minKmerFraction;
    c.device = device;
    if (seal_b200_create(&c, &h)) return 0;
    return (jlong)(intptr_t)h;
}

JNIEXPORT jint JNICALL Java_jgi_SealGPU_addRefNative(JNIEnv *env, jclass cls, jlong handle, jbyteArray jbases) {
    const jsize n = (*env)->GetArrayLength(env, jbases);
    int64_t off[2] = {0, n};
    jbyte *b = (jbyte *)malloc(n > 0 ? (size_t)n : 1);
    jint rc;
    if (!b) return 1;
    (*env)->GetByteArrayRegion(env, jbases, 0, n, b);
    rc = seal_b200_add_ref(H(handle), (const uint8_t *)b, off, 1);
    free(b);
    return rc;
}

/* out3 = {storedKmers, (k-mer, id) entries, refKmers} */
JNIEXPORT jint JNICALL Java_jgi_SealGPU_finalizeNative(JNIEnv *env, jclass cls, jlong handle, jlongArray jout3) {
    int64_t v[3];
    const jint rc = seal_b200_finalize(H(handle), v);
    if (!rc) (*env)->SetLongArrayRegion(env, jout3, 0, 3, (const jlong *)v);
    return rc;
}

/* Per-unit results into nAssigned / firstId / nSites / maxHits (length = units) and ids (units * idsStride, may be
 * null); stats8 = seal_stats. Units = pairs when paired && kpt, else reads. */
JNIEXPORT jint JNICALL Java_jgi_SealGPU_processNative(JNIEnv *env, jclass cls, jlong handle, jbyteArray jbases, jlongArray joffsets,
                                                      jlong nReads, jboolean paired, jlong firstNumericId, jint idsStride,
                                                      jintArray jnAssigned, jintArray jfirstId, jintArray jnSites, jintArray jmaxHits,
                                                      jintArray jids, jlongArray jstats8) {
    const jsize nb = (*env)->GetArrayLength(env, jbases);
    const int64_t nu = seal_b200_n_units(H(handle), nReads, paired ? 1 : 0);
    jint rc = 1;
    seal_out o;
    seal_stats st;
    jbyte *b = (jbyte *)malloc(nb > 0 ? (size_t)nb : 1);
    jlong *off = (jlong *)malloc((size_t)(nReads + 1) * sizeof(jlong));
    int32_t *res = (int32_t *)malloc((size_t)(nu > 0 ? nu : 1) * (4 + (size_t)(jids ? idsStride : 0)) * sizeof(int32_t));
    if (b && off && res && nu >= 0) {
        (*env)->GetByteArrayRegion(env, jbases, 0, nb, b);
        (*env)->GetLongArrayRegion(env, joffsets, 0, (jsize)(nReads + 1), off);
        o.n_assigned = res;
        o.first_id = res + nu;
        o.n_sites = res + 2 * nu;
        o.max_hits = res + 3 * nu;
        o.ids = jids ? res + 4 * nu : NULL;
        rc = seal_b200_process(H(handle), (const uint8_t *)b, (const int64_t *)off, nReads, paired ? 1 : 0, firstNumericId, &o, &st);
        if (!rc) {
            (*env)->SetIntArrayRegion(env, jnAssigned, 0, (jsize)nu, (const jint *)o.n_assigned);
            (*env)->SetIntArrayRegion(env, jfirstId, 0, (jsize)nu, (const jint *)o.first_id);
            (*env)->SetIntArrayRegion(env, jnSites, 0, (jsize)nu, (const jint *)o.n_sites);
            (*env)->SetIntArrayRegion(env, jmaxHits, 0, (jsize)nu, (const jint *)o.max_hits);
            if (jids) (*env)->SetIntArrayRegion(env, jids, 0, (jsize)(nu * idsStride), (const jint *)o.ids);
            (*env)->SetLongArrayRegion(env, jstats8, 0, 8, (const jlong *)&st);
        }
    }
    free(b);
    free(off);
    free(res);
    return rc;
}

/* four arrays of n = scaffolds + 1 longs (index = scaffold id) */
JNIEXPORT jint JNICALL Java_jgi_SealGPU_scaffoldCountsNative(JNIEnv *env, jclass cls, jlong handle, jlongArray jreads, jlongArray jbases,
                                                             jlongArray jfrags, jlongArray jambig) {
    const jsize n = (*env)->GetArrayLength(env, jreads);
    int64_t *v = (int64_t *)malloc((size_t)(n > 0 ? n : 1) * 4 * sizeof(int64_t));
    jint rc;
    if (!v) return 1;
    rc = seal_b200_scaffold_counts(H(handle), v, v + n, v + 2 * (size_t)n, v + 3 * (size_t)n, n);
    if (!rc) {
        (*env)->SetLongArrayRegion(env, jreads, 0, n, (const jlong *)v);
        (*env)->SetLongArrayRegion(env, jbases, 0, n, (const jlong *)(v + n));
        (*env)->SetLongArrayRegion(env, jfrags, 0, n, (const jlong *)(v + 2 * (size_t)n));
        (*env)->SetLongArrayRegion(env, jambig, 0, n, (const jlong *)(v + 3 * (size_t)n));
    }
    free(v);
    return rc;
}

JNIEXPORT jstring JNICALL Java_jgi_SealGPU_lastErrorNative(JNIEnv *env, jclass cls, jlong handle) {
    return (*env)->NewStringUTF(env, seal_b200_last_error(H(handle)));
}

JNIEXPORT void JNICALL Java_jgi_SealGPU_destroyNative(JNIEnv *env, jclass cls, jlong handle) { seal_b200_destroy(H(handle)); }
