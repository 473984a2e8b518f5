/*
 * BBDukCuda.c -- JNI shim between bbduk.BBDukIndexGPU (Java) and libbbduk_b200.so (C ABI).
 *
 * Pure marshalling. Conventions of the reference's own shim (jni/BBMergeOverlapper.c:389-437): static native methods, scalar
 * status return, caller-allocated result arrays, no Java exception thrown from native code, no global native state (the
 * handle is a jlong owned by the Java object). One deliberate difference: BBMergeOverlapper pins its arrays with
 * GetPrimitiveArrayCritical for a call that lasts microseconds; a batch call here runs CUDA work for tens of milliseconds
 * (allocation, stream synchronisation, host worker threads), and the JNI specification forbids blocking inside a critical
 * region (the collector is locked out JVM-wide). So every array is COPIED with Get / Set<Type>ArrayRegion into thread-local
 * staging owned by this shim, the library runs on the copies, and results are copied back: no critical region anywhere.
 * NOT compiled in this repository's image (no JDK / jni.h; tests compile it against tests/stubs/jni.h); build next to the
 * reference's jni/ directory:
 *   gcc -O3 -std=c99 -fPIC -shared -I$JAVA_HOME/include -I$JAVA_HOME/include/linux \
 *       -I<repo>/include BBDukCuda.c -L<repo>/bbtools_b200 -lbbduk_b200 -o libbbdukcuda.so
 */
#include <jni.h>
#include <stdlib.h>
#include <string.h>

#include "bbduk_b200.h"

#define H(x) ((bbduk_handle *)(intptr_t)(x))

/* thread-local staging: one growable buffer per role, reused from batch to batch by the calling ProcessThread */
enum { B_BASES, B_QUALS, B_OFF, B_ID0, B_ID0B, B_LO, B_HI, B_FLAGS, B_COUNT, B_MASK, B_MASKOFF, B_INSERT, B_N };
static __thread struct { void *p; size_t cap; } tl[B_N];

static void *need(int slot, size_t bytes) {
    if (bytes > tl[slot].cap) {
        free(tl[slot].p);
        tl[slot].cap = bytes + bytes / 4 + 4096;
        tl[slot].p = malloc(tl[slot].cap);
        if (!tl[slot].p) tl[slot].cap = 0;
    }
    return tl[slot].p;
}

/* offsets[0..nReads] and the bases (and qualities) they cover, copied out of the Java arrays; 0 on success */
static int stage_in(JNIEnv *env, jbyteArray jbases, jbyteArray jquals, jlongArray joffsets, jlong nReads, uint8_t **b, uint8_t **q,
                    int64_t **o) {
    if (nReads < 0 || !jbases || !joffsets) return 1;
    *o = (int64_t *)need(B_OFF, sizeof(int64_t) * (size_t)(nReads + 1));
    if (!*o) return 1;
    (*env)->GetLongArrayRegion(env, joffsets, 0, (jsize)(nReads + 1), (jlong *)*o);
    const int64_t nb = (*o)[nReads];
    if (nb < 0 || nb > (*env)->GetArrayLength(env, jbases)) return 1;
    *b = (uint8_t *)need(B_BASES, (size_t)nb + 64);
    if (!*b) return 1;
    (*env)->GetByteArrayRegion(env, jbases, 0, (jsize)nb, (jbyte *)*b);
    if (q) {
        *q = NULL;
        if (jquals) {
            *q = (uint8_t *)need(B_QUALS, (size_t)nb + 64);
            if (!*q) return 1;
            (*env)->GetByteArrayRegion(env, jquals, 0, (jsize)nb, (jbyte *)*q);
        }
    }
    return 0;
}

static void add_longs(JNIEnv *env, jlongArray j, int n, const int64_t *st) {
    jlong v[16];
    if (!j) return;
    (*env)->GetLongArrayRegion(env, j, 0, n, v);
    for (int i = 0; i < n; i++) v[i] += st[i];
    (*env)->SetLongArrayRegion(env, j, 0, n, v);
}

/* int[] cfg carries the bbduk_cfg fields in declaration order; float fields are passed as raw bits */
static void cfg_from(JNIEnv *env, jintArray jcfg, bbduk_cfg *cfg) {
    jint c[64];
    bbduk_b200_cfg_default(cfg);
    jsize n = (*env)->GetArrayLength(env, jcfg);
    if (n > 64) n = 64;
    (*env)->GetIntArrayRegion(env, jcfg, 0, n, c);
    const size_t bytes = (size_t)n * sizeof(jint);
    memcpy(cfg, c, bytes < sizeof *cfg ? bytes : sizeof *cfg);
    cfg->struct_size = (int32_t)sizeof *cfg;
}

JNIEXPORT jlong JNICALL Java_bbduk_BBDukIndexGPU_createNative(JNIEnv *env, jclass cls, jintArray jcfg) {
    bbduk_cfg cfg;
    bbduk_handle *h = NULL;
    cfg_from(env, jcfg, &cfg);
    if (bbduk_b200_create(&cfg, &h)) return 0;
    return (jlong)(intptr_t)h;
}

/* v16 = the constants the library derives from cfg (bbduk_b200_describe_cfg), for the Java side's cross-check against BBDukParser */
JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_describeNative(JNIEnv *env, jclass cls, jintArray jcfg, jlongArray jv) {
    bbduk_cfg cfg;
    int64_t v[16];
    cfg_from(env, jcfg, &cfg);
    const jint rc = bbduk_b200_describe_cfg(&cfg, v);
    if (!rc) (*env)->SetLongArrayRegion(env, jv, 0, 16, (const jlong *)v);
    return rc;
}

JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_addRefNative(JNIEnv *env, jclass cls, jlong handle, jbyteArray jbases,
                                                             jlongArray joffsets, jint nSeqs) {
    uint8_t *b;
    int64_t *o;
    if (stage_in(env, jbases, NULL, joffsets, nSeqs, &b, NULL, &o)) return 1;
    return bbduk_b200_add_ref(H(handle), b, o, nSeqs);
}

/* out = {storedKmers ("Added N kmers"), refKmers} */
JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_finalizeNative(JNIEnv *env, jclass cls, jlong handle, jlongArray jout) {
    int64_t v[2] = {-1, -1};
    const jint rc = bbduk_b200_finalize(H(handle), &v[0]);
    if (!rc) {
        v[1] = bbduk_b200_ref_kmers(H(handle));
        (*env)->SetLongArrayRegion(env, jout, 0, 2, (const jlong *)v);
    }
    return rc;
}

/* bbduk_b200_replicate: one new handle per device id, the table copied by NCCL broadcast / peer copies inside the library */
JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_replicateNative(JNIEnv *env, jclass cls, jlong handle, jintArray jdevs, jlongArray jout) {
    jint d[64];
    bbduk_handle *hs[64];
    jlong o[64];
    const jsize n = (*env)->GetArrayLength(env, jdevs);
    if (n < 1 || n > 64 || (*env)->GetArrayLength(env, jout) < n) return 1;
    (*env)->GetIntArrayRegion(env, jdevs, 0, n, d);
    const jint rc = bbduk_b200_replicate(H(handle), (const int32_t *)d, n, hs);
    if (!rc) {
        for (jsize i = 0; i < n; i++) o[i] = (jlong)(intptr_t)hs[i];
        (*env)->SetLongArrayRegion(env, jout, 0, n, o);
    }
    return rc;
}

/* BBDukIndex.dump: the stored (key, id) pairs; returns their number or -1 */
JNIEXPORT jlong JNICALL Java_bbduk_BBDukIndexGPU_dumpNative(JNIEnv *env, jclass cls, jlong handle, jlongArray jkeys, jintArray jids) {
    const jsize cap = (*env)->GetArrayLength(env, jkeys);
    int64_t n = 0;
    uint64_t *k = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(cap > 0 ? cap : 1));
    int32_t *v = (int32_t *)malloc(sizeof(int32_t) * (size_t)(cap > 0 ? cap : 1));
    if (!k || !v || (*env)->GetArrayLength(env, jids) < cap || bbduk_b200_table_export(H(handle), k, v, cap, &n)) {
        free(k);
        free(v);
        return -1;
    }
    const jsize m = (jsize)(n < cap ? n : cap);
    (*env)->SetLongArrayRegion(env, jkeys, 0, m, (const jlong *)k);
    (*env)->SetIntArrayRegion(env, jids, 0, m, (const jint *)v);
    free(k);
    free(v);
    return (jlong)n;
}

/* One aggregated batch (a whole read list) in, struct-of-arrays results out. Any result array may be
 * null (not wanted); maskBits / maskOff only in kmask mode (maskOff[nReads+1] is an INPUT: word offsets per read). */
JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_processNative(JNIEnv *env, jclass cls, jlong handle, jbyteArray jbases,
                                                              jlongArray joffsets, jlong nReads, jboolean paired,
                                                              jintArray jid0, jintArray jid0b, jintArray jlo, jintArray jhi,
                                                              jbyteArray jflags, jintArray jcount, jintArray jmask,
                                                              jlongArray jmaskoff, jlongArray jstats) {
    bbduk_out out;
    bbduk_stats st;
    uint8_t *b;
    int64_t *o;
    const size_t n = (size_t)(nReads > 0 ? nReads : 0);
    memset(&out, 0, sizeof out);
    if (stage_in(env, jbases, NULL, joffsets, nReads, &b, NULL, &o)) return 1;
    if (jid0) out.id0 = (int32_t *)need(B_ID0, 4 * n + 4);
    if (jid0b) out.id0b = (int32_t *)need(B_ID0B, 4 * n + 4);
    if (jlo) out.lo = (int32_t *)need(B_LO, 4 * n + 4);
    if (jhi) out.hi = (int32_t *)need(B_HI, 4 * n + 4);
    if (jflags) out.flags = (uint8_t *)need(B_FLAGS, n + 4);
    if (jcount) out.count = (int32_t *)need(B_COUNT, 4 * n + 4);
    int64_t words = 0;
    if (jmask && jmaskoff) {
        int64_t *mo = (int64_t *)need(B_MASKOFF, sizeof(int64_t) * (n + 1));
        if (!mo) return 1;
        (*env)->GetLongArrayRegion(env, jmaskoff, 0, (jsize)(n + 1), (jlong *)mo);
        words = mo[n];
        if (words < 0 || words > (*env)->GetArrayLength(env, jmask)) return 1;
        out.mask_off = mo;
        out.maskbits = (uint32_t *)need(B_MASK, 4 * (size_t)words + 4);
        if (!out.maskbits) return 1;
    }
    const jint rc = bbduk_b200_process(H(handle), b, o, (int64_t)nReads, paired ? 1 : 0, &out, &st);
    if (rc) return rc;
    if (jid0) (*env)->SetIntArrayRegion(env, jid0, 0, (jsize)n, (const jint *)out.id0);
    if (jid0b) (*env)->SetIntArrayRegion(env, jid0b, 0, (jsize)n, (const jint *)out.id0b);
    if (jlo) (*env)->SetIntArrayRegion(env, jlo, 0, (jsize)n, (const jint *)out.lo);
    if (jhi) (*env)->SetIntArrayRegion(env, jhi, 0, (jsize)n, (const jint *)out.hi);
    if (jflags) (*env)->SetByteArrayRegion(env, jflags, 0, (jsize)n, (const jbyte *)out.flags);
    if (jcount) (*env)->SetIntArrayRegion(env, jcount, 0, (jsize)n, (const jint *)out.count);
    if (out.maskbits) (*env)->SetIntArrayRegion(env, jmask, 0, (jsize)words, (const jint *)out.maskbits);
    if (jstats) (*env)->SetLongArrayRegion(env, jstats, 0, 8, (const jlong *)&st);
    return 0;
}

/* lo / hi / flags of a batch: in (copied from Java), and back out after the step */
static int stage_state(JNIEnv *env, jintArray jlo, jintArray jhi, jbyteArray jflags, size_t n, int32_t **lo, int32_t **hi, uint8_t **fl) {
    *lo = (int32_t *)need(B_LO, 4 * n + 4);
    *hi = (int32_t *)need(B_HI, 4 * n + 4);
    *fl = (uint8_t *)need(B_FLAGS, n + 4);
    if (!*lo || !*hi || !*fl) return 1;
    (*env)->GetIntArrayRegion(env, jlo, 0, (jsize)n, (jint *)*lo);
    (*env)->GetIntArrayRegion(env, jhi, 0, (jsize)n, (jint *)*hi);
    (*env)->GetByteArrayRegion(env, jflags, 0, (jsize)n, (jbyte *)*fl);
    return 0;
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
    uint8_t *b, *q, *fl;
    int64_t *o;
    int32_t *lo, *hi, *ins = NULL;
    const size_t n = (size_t)(nReads > 0 ? nReads : 0);
    bbduk_b200_tbo_cfg_default(&cfg);
    (*env)->GetIntArrayRegion(env, jcfg, 0, 6, c);
    cfg.strict_overlap = c[0];
    cfg.min_overlap0 = c[1];
    cfg.min_overlap = c[2];
    cfg.min_insert0 = c[3];
    cfg.min_insert = c[4];
    cfg.qual_offset = c[5];
    cfg.mee_filter = meeFilter;
    if (stage_in(env, jbases, jquals, joffsets, nReads, &b, &q, &o) || stage_state(env, jlo, jhi, jflags, n, &lo, &hi, &fl)) return 1;
    if (jinsert && !(ins = (int32_t *)need(B_INSERT, 2 * n + 8))) return 1;
    const jint rc = bbduk_b200_tbo(H(handle), &cfg, b, q, o, (int64_t)nReads, lo, hi, fl, ins, st);
    if (rc) return rc;
    (*env)->SetIntArrayRegion(env, jhi, 0, (jsize)n, (const jint *)hi);
    (*env)->SetByteArrayRegion(env, jflags, 0, (jsize)n, (const jbyte *)fl);
    if (jinsert) (*env)->SetIntArrayRegion(env, jinsert, 0, (jsize)(n / 2), (const jint *)ins);
    add_longs(env, jstats2, 2, st);
    return 0;
}

/* Replaces the poly-X, quality-trimming and quality-filtering blocks of the per-pair loop (jgi/BBDuk.java:2954-3052,
 * :3074-3170) for the batch that processNative (and tboNative) answered: lo[], hi[] and flags[] are updated in place,
 * stats8 += {readsQTrimmed, basesQTrimmed, readsQFiltered, basesQFiltered, readsNFiltered, basesNFiltered, readsPolyTrimmed,
 * basesPolyTrimmed}. qCfg = {qtrimLeft, qtrimRight, minBaseQuality, maxNs (-1 off), maxReadLength (0 unlimited), qualOffset
 * (0 for Read.quality), trimPolyA, trimPolyGLeft, trimPolyGRight, filterPolyG, trimPolyCLeft, trimPolyCRight, filterPolyC,
 * maxNonPoly, minConsecutiveBases, floatToRawIntBits(maxNRate), floatToRawIntBits(minBaseFrequency)} (17 ints; the trimming rule stays TrimRead's default, optimal). */
static void qcfg_from(const jint *c, jfloat trimq, bbduk_qtrim_cfg *cfg) {
    bbduk_b200_qtrim_cfg_default(cfg);
    cfg->qtrim_left = c[0];
    cfg->qtrim_right = c[1];
    cfg->min_base_quality = c[2];
    cfg->max_ns = c[3];
    cfg->max_read_length = c[4];
    cfg->qual_offset = c[5];
    cfg->trim_poly_a = c[6];
    cfg->trim_poly_g_left = c[7];
    cfg->trim_poly_g_right = c[8];
    cfg->filter_poly_g = c[9];
    cfg->trim_poly_c_left = c[10];
    cfg->trim_poly_c_right = c[11];
    cfg->filter_poly_c = c[12];
    cfg->max_non_poly = c[13];
    cfg->min_consecutive_bases = c[14];
    { /* the two rates travel as Float.floatToRawIntBits in the int vector */
        union { jint i; float f; } u;
        u.i = c[15];
        cfg->max_n_rate = u.f;
        u.i = c[16];
        cfg->min_base_frequency = u.f;
    }
    cfg->trimq = trimq;
}

JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_qtrimNative(JNIEnv *env, jclass cls, jlong handle, jintArray jcfg, jfloat trimq,
                                                            jbyteArray jbases, jbyteArray jquals, jlongArray joffsets, jlong nReads,
                                                            jboolean paired, jintArray jlo, jintArray jhi, jbyteArray jflags,
                                                            jlongArray jstats8) {
    bbduk_qtrim_cfg cfg;
    jint c[17];
    int64_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint8_t *b, *q, *fl;
    int64_t *o;
    int32_t *lo, *hi;
    const size_t n = (size_t)(nReads > 0 ? nReads : 0);
    (*env)->GetIntArrayRegion(env, jcfg, 0, 17, c);
    qcfg_from(c, trimq, &cfg);
    if (stage_in(env, jbases, jquals, joffsets, nReads, &b, &q, &o) || stage_state(env, jlo, jhi, jflags, n, &lo, &hi, &fl)) return 1;
    const jint rc = bbduk_b200_qtrim(H(handle), &cfg, b, q, o, (int64_t)nReads, paired ? 1 : 0, lo, hi, fl, st);
    if (rc) return rc;
    (*env)->SetIntArrayRegion(env, jlo, 0, (jsize)n, (const jint *)lo);
    (*env)->SetIntArrayRegion(env, jhi, 0, (jsize)n, (const jint *)hi);
    (*env)->SetByteArrayRegion(env, jflags, 0, (jsize)n, (const jbyte *)fl);
    add_longs(env, jstats8, 8, st);
    return 0;
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
    uint8_t *b, *fl;
    int64_t *o;
    int32_t *lo, *hi;
    const size_t n = (size_t)(nReads > 0 ? nReads : 0);
    bbduk_b200_entropy_cfg_default(&cfg);
    (*env)->GetIntArrayRegion(env, jcfg, 0, 3, c);
    cfg.k = c[0];
    cfg.window = c[1];
    cfg.high_pass = c[2];
    cfg.cutoff = cutoff;
    if (stage_in(env, jbases, NULL, joffsets, nReads, &b, NULL, &o) || stage_state(env, jlo, jhi, jflags, n, &lo, &hi, &fl)) return 1;
    const jint rc = bbduk_b200_entropy(H(handle), &cfg, b, o, (int64_t)nReads, paired ? 1 : 0, lo, hi, fl, st);
    if (rc) return rc;
    (*env)->SetIntArrayRegion(env, jhi, 0, (jsize)n, (const jint *)hi);
    (*env)->SetByteArrayRegion(env, jflags, 0, (jsize)n, (const jbyte *)fl);
    add_longs(env, jstats2, 2, st);
    return 0;
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
    jint st3[3], t[6], q[17], e[3];
    jfloat f[3];
    bbduk_out out;
    bbduk_stats st;
    int64_t extra[12];
    uint8_t *b, *qq;
    int64_t *o;
    const size_t n = (size_t)(nReads > 0 ? nReads : 0);
    memset(&out, 0, sizeof out);
    memset(extra, 0, sizeof extra);
    bbduk_b200_chain_cfg_default(&cfg);
    (*env)->GetIntArrayRegion(env, jsteps, 0, 3, st3);
    (*env)->GetIntArrayRegion(env, jtbo, 0, 6, t);
    (*env)->GetIntArrayRegion(env, jq, 0, 17, q);
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
    qcfg_from(q, f[1], &cfg.qtrim);
    cfg.entropy.k = e[0];
    cfg.entropy.window = e[1];
    cfg.entropy.high_pass = e[2];
    cfg.entropy.cutoff = f[2];
    if (stage_in(env, jbases, jquals, joffsets, nReads, &b, &qq, &o)) return 1;
    if (jid0) out.id0 = (int32_t *)need(B_ID0, 4 * n + 4);
    out.lo = (int32_t *)need(B_LO, 4 * n + 4);
    out.hi = (int32_t *)need(B_HI, 4 * n + 4);
    out.flags = (uint8_t *)need(B_FLAGS, n + 4);
    if (!out.lo || !out.hi || !out.flags) return 1;
    const jint rc = bbduk_b200_process_chain(H(handle), &cfg, b, qq, o, (int64_t)nReads, paired ? 1 : 0, &out, &st, extra, extra + 2,
                                             extra + 10);
    if (rc) return rc;
    if (jid0) (*env)->SetIntArrayRegion(env, jid0, 0, (jsize)n, (const jint *)out.id0);
    (*env)->SetIntArrayRegion(env, jlo, 0, (jsize)n, (const jint *)out.lo);
    (*env)->SetIntArrayRegion(env, jhi, 0, (jsize)n, (const jint *)out.hi);
    (*env)->SetByteArrayRegion(env, jflags, 0, (jsize)n, (const jbyte *)out.flags);
    if (jstats28) {
        (*env)->SetLongArrayRegion(env, jstats28, 0, 8, (const jlong *)&st);
        (*env)->SetLongArrayRegion(env, jstats28, 8, 12, (const jlong *)extra);
    }
    return 0;
}

JNIEXPORT jint JNICALL Java_bbduk_BBDukIndexGPU_scaffoldCountsNative(JNIEnv *env, jclass cls, jlong handle,
                                                                     jlongArray jreads, jlongArray jbases) {
    const jsize n = (*env)->GetArrayLength(env, jreads);
    int64_t *r = (int64_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int64_t));
    int64_t *b = (int64_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int64_t));
    jint rc = 1;
    if (r && b && (*env)->GetArrayLength(env, jbases) >= n) {
        rc = bbduk_b200_scaffold_counts(H(handle), r, b, n);
        if (!rc) {
            (*env)->SetLongArrayRegion(env, jreads, 0, n, (const jlong *)r);
            (*env)->SetLongArrayRegion(env, jbases, 0, n, (const jlong *)b);
        }
    }
    free(r);
    free(b);
    return rc;
}

JNIEXPORT jstring JNICALL Java_bbduk_BBDukIndexGPU_lastErrorNative(JNIEnv *env, jclass cls, jlong handle) {
    return (*env)->NewStringUTF(env, bbduk_b200_last_error(H(handle)));
}

JNIEXPORT void JNICALL Java_bbduk_BBDukIndexGPU_destroyNative(JNIEnv *env, jclass cls, jlong handle) {
    bbduk_b200_destroy(H(handle));
}
