/*
 * BBDukCuda.c -- JNI shim between bbduk.BBDukIndexGPU (Java) and libbbduk_b200.so (C ABI).
 *
 * Pure marshalling, modelled on jni/BBMergeOverlapper.c:389-437 of the reference: primitive arrays are
 * pinned with GetPrimitiveArrayCritical, inputs released with JNI_ABORT, outputs with 0; scalar status
 * return; no Java exception is thrown from native code; no global native state (the handle is a jlong
 * owned by the Java object). NOT compiled in this repository's image (no JDK / jni.h); build next to
 * the reference's jni/ directory:
 *   gcc -O3 -std=c99 -fPIC -shared -I$JAVA_HOME/include -I$JAVA_HOME/include/linux \
 *       -I<repo>/include BBDukCuda.c -L<repo>/bbtools_b200 -lbbduk_b200 -o libbbdukcuda.so
 */
#include <jni.h>
#include <string.h>

#include "bbduk_b200.h"

/* int[] cfg carries the bbduk_cfg fields in declaration order; float fields are passed as raw bits */
JNIEXPORT jlong JNICALL Java_bbduk_BBDukIndexGPU_createNative(JNIEnv *env, jclass cls, jintArray jcfg) {
    bbduk_cfg cfg;
    bbduk_b200_cfg_default(&cfg);
    const jint n = (*env)->GetArrayLength(env, jcfg);
    jint *c = (jint *)(*env)->GetPrimitiveArrayCritical(env, jcfg, NULL);
    const size_t bytes = (size_t)n * sizeof(jint);
    memcpy(&cfg, c, bytes < sizeof cfg ? bytes : sizeof cfg);
    (*env)->ReleasePrimitiveArrayCritical(env, jcfg, c, JNI_ABORT);
    cfg.struct_size = (int32_t)sizeof cfg;
    bbduk_handle *h = NULL;
    if (bbduk_b200_create(&cfg, &h)) return 0;
    return (jlong)(intptr_t)h;
}

JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_addRefNative(JNIEnv *env, jclass cls, jlong handle, jbyteArray jbases,
                                                             jlongArray joffsets, jint nSeqs) {
    jbyte *b = (jbyte *)(*env)->GetPrimitiveArrayCritical(env, jbases, NULL);
    jlong *o = (jlong *)(*env)->GetPrimitiveArrayCritical(env, joffsets, NULL);
    const jint rc = bbduk_b200_add_ref((bbduk_handle *)(intptr_t)handle, (const uint8_t *)b, (const int64_t *)o, nSeqs);
    (*env)->ReleasePrimitiveArrayCritical(env, joffsets, o, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jbases, b, JNI_ABORT);
    return rc;
}

JNIEXPORT jlong JNICALL Java_bbduk_BBDukIndexGPU_finalizeNative(JNIEnv *env, jclass cls, jlong handle) {
    int64_t stored = -1;
    if (bbduk_b200_finalize((bbduk_handle *)(intptr_t)handle, &stored)) return -1;
    return (jlong)stored;
}

/* One aggregated batch (>= ~1 M reads) in, struct-of-arrays results out.
 * The library copies out of the pinned Java arrays into its own staging before it returns from the
 * critical section's memcpy; it never holds a critical array across a CUDA synchronisation point of
 * another thread because each call owns a private stream + staging slot. */
JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_processNative(JNIEnv *env, jclass cls, jlong handle, jbyteArray jbases,
                                                              jlongArray joffsets, jlong nReads, jboolean paired,
                                                              jintArray jid0, jintArray jlo, jintArray jhi,
                                                              jbyteArray jflags, jintArray jcount, jlongArray jstats) {
    bbduk_out out;
    bbduk_stats st;
    memset(&out, 0, sizeof out);
    jbyte *b = (jbyte *)(*env)->GetPrimitiveArrayCritical(env, jbases, NULL);
    jlong *o = (jlong *)(*env)->GetPrimitiveArrayCritical(env, joffsets, NULL);
    if (jid0) out.id0 = (int32_t *)(*env)->GetPrimitiveArrayCritical(env, jid0, NULL);
    if (jlo) out.lo = (int32_t *)(*env)->GetPrimitiveArrayCritical(env, jlo, NULL);
    if (jhi) out.hi = (int32_t *)(*env)->GetPrimitiveArrayCritical(env, jhi, NULL);
    if (jflags) out.flags = (uint8_t *)(*env)->GetPrimitiveArrayCritical(env, jflags, NULL);
    if (jcount) out.count = (int32_t *)(*env)->GetPrimitiveArrayCritical(env, jcount, NULL);
    const jint rc = bbduk_b200_process((bbduk_handle *)(intptr_t)handle, (const uint8_t *)b, (const int64_t *)o,
                                       (int64_t)nReads, paired ? 1 : 0, &out, &st);
    if (jcount) (*env)->ReleasePrimitiveArrayCritical(env, jcount, out.count, 0);
    if (jflags) (*env)->ReleasePrimitiveArrayCritical(env, jflags, out.flags, 0);
    if (jhi) (*env)->ReleasePrimitiveArrayCritical(env, jhi, out.hi, 0);
    if (jlo) (*env)->ReleasePrimitiveArrayCritical(env, jlo, out.lo, 0);
    if (jid0) (*env)->ReleasePrimitiveArrayCritical(env, jid0, out.id0, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, joffsets, o, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jbases, b, JNI_ABORT);
    if (jstats && !rc) (*env)->SetLongArrayRegion(env, jstats, 0, 8, (const jlong *)&st);
    return rc;
}

/* Replaces the tbo block of BBDukProcessorS.processList (bbduk/BBDukProcessorS.java:1096-1143 = jgi/BBDuk.java:2878-2926)
 * for the batch that processNative just answered: hi[] and flags[] are updated in place, insert[] (may be null) gets
 * the insert size used per pair (-1 none, -2 ambiguous), stats2 += {readsTrimmedByOverlap, basesTrimmedByOverlap}.
 * tboCfg = {strictOverlap, minOverlap0, minOverlap, minInsert0, minInsert, qualOffset} (-1 / 0 = the reference's
 * defaults), meeFilter 0 = default. quals may be null (reads without qualities). */
JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_tboNative(JNIEnv *env, jclass cls, jlong handle, jintArray jcfg, jfloat meeFilter,
                                                          jbyteArray jbases, jbyteArray jquals, jlongArray joffsets, jlong nReads,
                                                          jintArray jlo, jintArray jhi, jbyteArray jflags, jintArray jinsert,
                                                          jlongArray jstats2) {
    bbduk_tbo_cfg cfg;
    jint c[6];
    int64_t st[2] = {0, 0};
    bbduk_b200_tbo_cfg_default(&cfg);
    (*env)->GetIntArrayRegion(env, jcfg, 0, 6, c);
    cfg.strict_overlap = c[0];
    cfg.min_overlap0 = c[1];
    cfg.min_overlap = c[2];
    cfg.min_insert0 = c[3];
    cfg.min_insert = c[4];
    cfg.qual_offset = c[5];
    cfg.mee_filter = meeFilter;
    jbyte *b = (jbyte *)(*env)->GetPrimitiveArrayCritical(env, jbases, NULL);
    jbyte *q = jquals ? (jbyte *)(*env)->GetPrimitiveArrayCritical(env, jquals, NULL) : NULL;
    jlong *o = (jlong *)(*env)->GetPrimitiveArrayCritical(env, joffsets, NULL);
    jint *lo = (jint *)(*env)->GetPrimitiveArrayCritical(env, jlo, NULL);
    jint *hi = (jint *)(*env)->GetPrimitiveArrayCritical(env, jhi, NULL);
    jbyte *fl = (jbyte *)(*env)->GetPrimitiveArrayCritical(env, jflags, NULL);
    jint *ins = jinsert ? (jint *)(*env)->GetPrimitiveArrayCritical(env, jinsert, NULL) : NULL;
    const jint rc = bbduk_b200_tbo((bbduk_handle *)(intptr_t)handle, &cfg, (const uint8_t *)b, (const uint8_t *)q,
                                   (const int64_t *)o, (int64_t)nReads, (const int32_t *)lo, (int32_t *)hi, (uint8_t *)fl,
                                   (int32_t *)ins, st);
    if (jinsert) (*env)->ReleasePrimitiveArrayCritical(env, jinsert, ins, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, jflags, fl, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, jhi, hi, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, jlo, lo, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, joffsets, o, JNI_ABORT);
    if (jquals) (*env)->ReleasePrimitiveArrayCritical(env, jquals, q, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jbases, b, JNI_ABORT);
    if (jstats2 && !rc) {
        jlong v[2];
        (*env)->GetLongArrayRegion(env, jstats2, 0, 2, v);
        v[0] += st[0];
        v[1] += st[1];
        (*env)->SetLongArrayRegion(env, jstats2, 0, 2, v);
    }
    return rc;
}

/* Replaces the poly-X, quality-trimming and quality-filtering blocks of the per-pair loop (jgi/BBDuk.java:2954-3052,
 * :3074-3170) for the batch that processNative (and tboNative) answered: lo[], hi[] and flags[] are updated in place,
 * stats8 += {readsQTrimmed, basesQTrimmed, readsQFiltered, basesQFiltered, readsNFiltered, basesNFiltered, readsPolyTrimmed,
 * basesPolyTrimmed}. qCfg = {qtrimLeft, qtrimRight, minBaseQuality, maxNs (-1 off), maxReadLength (0 unlimited), qualOffset
 * (0 for Read.quality), trimPolyA, trimPolyGLeft, trimPolyGRight, filterPolyG, trimPolyCLeft, trimPolyCRight, filterPolyC,
 * maxNonPoly}. */
JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_qtrimNative(JNIEnv *env, jclass cls, jlong handle, jintArray jcfg, jfloat trimq,
                                                            jbyteArray jbases, jbyteArray jquals, jlongArray joffsets, jlong nReads,
                                                            jboolean paired, jintArray jlo, jintArray jhi, jbyteArray jflags,
                                                            jlongArray jstats8) {
    bbduk_qtrim_cfg cfg;
    jint c[14];
    int64_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bbduk_b200_qtrim_cfg_default(&cfg);
    (*env)->GetIntArrayRegion(env, jcfg, 0, 14, c);
    cfg.qtrim_left = c[0];
    cfg.qtrim_right = c[1];
    cfg.min_base_quality = c[2];
    cfg.max_ns = c[3];
    cfg.max_read_length = c[4];
    cfg.qual_offset = c[5];
    cfg.trim_poly_a = c[6];
    cfg.trim_poly_g_left = c[7];
    cfg.trim_poly_g_right = c[8];
    cfg.filter_poly_g = c[9];
    cfg.trim_poly_c_left = c[10];
    cfg.trim_poly_c_right = c[11];
    cfg.filter_poly_c = c[12];
    cfg.max_non_poly = c[13];
    cfg.trimq = trimq;
    jbyte *b = (jbyte *)(*env)->GetPrimitiveArrayCritical(env, jbases, NULL);
    jbyte *q = jquals ? (jbyte *)(*env)->GetPrimitiveArrayCritical(env, jquals, NULL) : NULL;
    jlong *o = (jlong *)(*env)->GetPrimitiveArrayCritical(env, joffsets, NULL);
    jint *lo = (jint *)(*env)->GetPrimitiveArrayCritical(env, jlo, NULL);
    jint *hi = (jint *)(*env)->GetPrimitiveArrayCritical(env, jhi, NULL);
    jbyte *fl = (jbyte *)(*env)->GetPrimitiveArrayCritical(env, jflags, NULL);
    const jint rc = bbduk_b200_qtrim((bbduk_handle *)(intptr_t)handle, &cfg, (const uint8_t *)b, (const uint8_t *)q,
                                     (const int64_t *)o, (int64_t)nReads, paired ? 1 : 0, (int32_t *)lo, (int32_t *)hi,
                                     (uint8_t *)fl, st);
    (*env)->ReleasePrimitiveArrayCritical(env, jflags, fl, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, jhi, hi, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, jlo, lo, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, joffsets, o, JNI_ABORT);
    if (jquals) (*env)->ReleasePrimitiveArrayCritical(env, jquals, q, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jbases, b, JNI_ABORT);
    if (jstats8 && !rc) {
        jlong v[8];
        (*env)->GetLongArrayRegion(env, jstats8, 0, 8, v);
        for (int i = 0; i < 8; i++) v[i] += st[i];
        (*env)->SetLongArrayRegion(env, jstats8, 0, 8, v);
    }
    return rc;
}

/* Replaces the "Test entropy" block of the per-pair loop (jgi/BBDuk.java:3175-3186; eTrackerT.passes(r.bases, true)) for the
 * batch that processNative / tboNative / qtrimNative answered: flags[] (and hi[] with trimfailuresto1bp) are updated in
 * place, stats2 += {readsEFiltered, basesEFiltered}. eCfg = {k, window, highPass}. */
JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_entropyNative(JNIEnv *env, jclass cls, jlong handle, jintArray jcfg, jfloat cutoff,
                                                              jbyteArray jbases, jlongArray joffsets, jlong nReads, jboolean paired,
                                                              jintArray jlo, jintArray jhi, jbyteArray jflags, jlongArray jstats2) {
    bbduk_entropy_cfg cfg;
    jint c[3];
    int64_t st[2] = {0, 0};
    bbduk_b200_entropy_cfg_default(&cfg);
    (*env)->GetIntArrayRegion(env, jcfg, 0, 3, c);
    cfg.k = c[0];
    cfg.window = c[1];
    cfg.high_pass = c[2];
    cfg.cutoff = cutoff;
    jbyte *b = (jbyte *)(*env)->GetPrimitiveArrayCritical(env, jbases, NULL);
    jlong *o = (jlong *)(*env)->GetPrimitiveArrayCritical(env, joffsets, NULL);
    jint *lo = (jint *)(*env)->GetPrimitiveArrayCritical(env, jlo, NULL);
    jint *hi = (jint *)(*env)->GetPrimitiveArrayCritical(env, jhi, NULL);
    jbyte *fl = (jbyte *)(*env)->GetPrimitiveArrayCritical(env, jflags, NULL);
    const jint rc = bbduk_b200_entropy((bbduk_handle *)(intptr_t)handle, &cfg, (const uint8_t *)b, (const int64_t *)o, (int64_t)nReads,
                                       paired ? 1 : 0, (const int32_t *)lo, (int32_t *)hi, (uint8_t *)fl, st);
    (*env)->ReleasePrimitiveArrayCritical(env, jflags, fl, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, jhi, hi, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, jlo, lo, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, joffsets, o, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jbases, b, JNI_ABORT);
    if (jstats2 && !rc) {
        jlong v[2];
        (*env)->GetLongArrayRegion(env, jstats2, 0, 2, v);
        v[0] += st[0];
        v[1] += st[1];
        (*env)->SetLongArrayRegion(env, jstats2, 0, 2, v);
    }
    return rc;
}

/* The whole device part of the per-pair loop in one call (bbduk_b200_process_chain): k-mer block, then tbo / poly-X + quality
 * trimming + filters / entropy filter as switched on in `steps` = {doTbo, doQtrim, doEntropy}. tboCfg, qCfg, eCfg and the
 * three floats as in tboNative / qtrimNative / entropyNative. stats28 = 8 (k-mer block, overwritten) + 2 (tbo) + 8 (qtrim) +
 * 2 (entropy) + 8 spare longs. */
JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_processChainNative(JNIEnv *env, jclass cls, jlong handle, jintArray jsteps,
                                                                   jintArray jtbo, jintArray jq, jintArray je, jfloatArray jfl,
                                                                   jbyteArray jbases, jbyteArray jquals, jlongArray joffsets,
                                                                   jlong nReads, jboolean paired, jintArray jid0, jintArray jlo,
                                                                   jintArray jhi, jbyteArray jflags, jlongArray jstats28) {
    bbduk_chain_cfg cfg;
    jint st3[3], t[6], q[14], e[3];
    jfloat f[3];
    bbduk_out out;
    bbduk_stats st;
    int64_t extra[12];
    memset(&out, 0, sizeof out);
    memset(extra, 0, sizeof extra);
    bbduk_b200_chain_cfg_default(&cfg);
    (*env)->GetIntArrayRegion(env, jsteps, 0, 3, st3);
    (*env)->GetIntArrayRegion(env, jtbo, 0, 6, t);
    (*env)->GetIntArrayRegion(env, jq, 0, 14, q);
    (*env)->GetIntArrayRegion(env, je, 0, 3, e);
    (*env)->GetFloatArrayRegion(env, jfl, 0, 3, f); /* {meeFilter, trimq, entropyCutoff} */
    cfg.do_tbo = st3[0];
    cfg.do_qtrim = st3[1];
    cfg.do_entropy = st3[2];
    cfg.tbo.strict_overlap = t[0];
    cfg.tbo.min_overlap0 = t[1];
    cfg.tbo.min_overlap = t[2];
    cfg.tbo.min_insert0 = t[3];
    cfg.tbo.min_insert = t[4];
    cfg.tbo.qual_offset = t[5];
    cfg.tbo.mee_filter = f[0];
    cfg.qtrim.qtrim_left = q[0];
    cfg.qtrim.qtrim_right = q[1];
    cfg.qtrim.min_base_quality = q[2];
    cfg.qtrim.max_ns = q[3];
    cfg.qtrim.max_read_length = q[4];
    cfg.qtrim.qual_offset = q[5];
    cfg.qtrim.trim_poly_a = q[6];
    cfg.qtrim.trim_poly_g_left = q[7];
    cfg.qtrim.trim_poly_g_right = q[8];
    cfg.qtrim.filter_poly_g = q[9];
    cfg.qtrim.trim_poly_c_left = q[10];
    cfg.qtrim.trim_poly_c_right = q[11];
    cfg.qtrim.filter_poly_c = q[12];
    cfg.qtrim.max_non_poly = q[13];
    cfg.qtrim.trimq = f[1];
    cfg.entropy.k = e[0];
    cfg.entropy.window = e[1];
    cfg.entropy.high_pass = e[2];
    cfg.entropy.cutoff = f[2];
    jbyte *b = (jbyte *)(*env)->GetPrimitiveArrayCritical(env, jbases, NULL);
    jbyte *qq = jquals ? (jbyte *)(*env)->GetPrimitiveArrayCritical(env, jquals, NULL) : NULL;
    jlong *o = (jlong *)(*env)->GetPrimitiveArrayCritical(env, joffsets, NULL);
    if (jid0) out.id0 = (int32_t *)(*env)->GetPrimitiveArrayCritical(env, jid0, NULL);
    out.lo = (int32_t *)(*env)->GetPrimitiveArrayCritical(env, jlo, NULL);
    out.hi = (int32_t *)(*env)->GetPrimitiveArrayCritical(env, jhi, NULL);
    out.flags = (uint8_t *)(*env)->GetPrimitiveArrayCritical(env, jflags, NULL);
    const jint rc = bbduk_b200_process_chain((bbduk_handle *)(intptr_t)handle, &cfg, (const uint8_t *)b, (const uint8_t *)qq,
                                             (const int64_t *)o, (int64_t)nReads, paired ? 1 : 0, &out, &st, extra, extra + 2,
                                             extra + 10);
    (*env)->ReleasePrimitiveArrayCritical(env, jflags, out.flags, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, jhi, out.hi, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, jlo, out.lo, 0);
    if (jid0) (*env)->ReleasePrimitiveArrayCritical(env, jid0, out.id0, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, joffsets, o, JNI_ABORT);
    if (jquals) (*env)->ReleasePrimitiveArrayCritical(env, jquals, qq, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jbases, b, JNI_ABORT);
    if (jstats28 && !rc) {
        (*env)->SetLongArrayRegion(env, jstats28, 0, 8, (const jlong *)&st);
        (*env)->SetLongArrayRegion(env, jstats28, 8, 12, (const jlong *)extra);
    }
    return rc;
}

JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_scaffoldCountsNative(JNIEnv *env, jclass cls, jlong handle,
                                                                     jlongArray jreads, jlongArray jbases) {
    const jint n = (*env)->GetArrayLength(env, jreads);
    jlong *r = (jlong *)(*env)->GetPrimitiveArrayCritical(env, jreads, NULL);
    jlong *b = (jlong *)(*env)->GetPrimitiveArrayCritical(env, jbases, NULL);
    const jint rc = bbduk_b200_scaffold_counts((bbduk_handle *)(intptr_t)handle, (int64_t *)r, (int64_t *)b, n);
    (*env)->ReleasePrimitiveArrayCritical(env, jbases, b, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, jreads, r, 0);
    return rc;
}

JNIEXPORT jstring JNICALL Java_bbduk_BBDukIndexGPU_lastErrorNative(JNIEnv *env, jclass cls, jlong handle) {
    return (*env)->NewStringUTF(env, bbduk_b200_last_error((bbduk_handle *)(intptr_t)handle));
}

JNIEXPORT void JNICALL Java_bbduk_BBDukIndexGPU_destroyNative(JNIEnv *env, jclass cls, jlong handle) {
    bbduk_b200_destroy((bbduk_handle *)(intptr_t)handle);
}
