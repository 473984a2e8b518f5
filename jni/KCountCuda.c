/*
 * KCountCuda.c -- JNI shim between a GPU-backed kmer.KmerTableSet (Java) and the counting entry points of
 * libbbduk_b200.so (include/kcount_b200.h). Pure marshalling, same conventions as jni/BBDukCuda.c (arrays copied
 * with Get / Set*ArrayRegion, no critical section across a CUDA call). NOT compiled against a real JDK in this repository's image.
 *
 * Java side: KmerTableSet.LoadThread (kmer/KmerTableSet.java:489-592) aggregates reads into one byte[] +
 * long[] offsets per >= 1 M reads and calls addReadsNative instead of addKmersToTable (:652-716);
 * AbstractKmerTableSet.fillHistogram (:370) calls khistNative; "Unique Kmers" comes from statsNative[3].
 */
#include <jni.h>
#include <stddef.h>
#include <stdlib.h>

#include "kcount_b200.h"

JNIEXPORT jlong JNICALL Java_kmer_KmerTableSetGPU_createNative(JNIEnv *env, jclass cls, jint k, jboolean rcomp, jlong initialKeys) {
    kcount_handle *h = NULL;
    if (kcount_b200_create(k, rcomp ? 1 : 0, initialKeys, -1, &h)) return 0;
    return (jlong)(intptr_t)h;
}

/* The arrays are COPIED (Get*ArrayRegion): kcount_b200_add_reads synchronises with the device and may grow the table,
 * far too long to hold a JNI critical section (the garbage collector would be blocked JVM-wide). */
JNIEXPORT jint JNICALL Java_kmer_KmerTableSetGPU_addReadsNative(JNIEnv *env, jclass cls, jlong handle, jbyteArray jbases,
                                                                jlongArray joffsets, jlong nReads) {
    const jsize nb = (*env)->GetArrayLength(env, jbases);
    jbyte *b = (jbyte *)malloc(nb > 0 ? (size_t)nb : 1);
    jlong *o = (jlong *)malloc((size_t)(nReads + 1) * sizeof(jlong));
    jint rc = 1;
    if (b && o && (*env)->GetArrayLength(env, joffsets) >= nReads + 1) {
        (*env)->GetByteArrayRegion(env, jbases, 0, nb, b);
        (*env)->GetLongArrayRegion(env, joffsets, 0, (jsize)(nReads + 1), o);
        rc = kcount_b200_add_reads((kcount_handle *)(intptr_t)handle, (const uint8_t *)b, (const int64_t *)o, nReads);
    }
    free(b);
    free(o);
    return rc;
}

/* stats4 = {readsIn, basesIn, kmersIn, unique k-mers} */
JNIEXPORT jint JNICALL Java_kmer_KmerTableSetGPU_statsNative(JNIEnv *env, jclass cls, jlong handle, jlongArray jstats4) {
    int64_t v[4];
    const jint rc = kcount_b200_stats((kcount_handle *)(intptr_t)handle, v);
    if (!rc) (*env)->SetLongArrayRegion(env, jstats4, 0, 4, (const jlong *)v);
    return rc;
}

JNIEXPORT jint JNICALL Java_kmer_KmerTableSetGPU_khistNative(JNIEnv *env, jclass cls, jlong handle, jint histMax, jlongArray jhist) {
    int64_t *hst = (int64_t *)malloc((size_t)(histMax + 1) * sizeof(int64_t));
    jint rc = 1;
    if (hst && (*env)->GetArrayLength(env, jhist) >= histMax + 1) {
        rc = kcount_b200_khist((kcount_handle *)(intptr_t)handle, histMax, hst);
        if (!rc) (*env)->SetLongArrayRegion(env, jhist, 0, histMax + 1, (const jlong *)hst);
    }
    free(hst);
    return rc;
}

JNIEXPORT jstring JNICALL Java_kmer_KmerTableSetGPU_lastErrorNative(JNIEnv *env, jclass cls, jlong handle) {
    return (*env)->NewStringUTF(env, kcount_b200_last_error((kcount_handle *)(intptr_t)handle));
}

JNIEXPORT void JNICALL Java_kmer_KmerTableSetGPU_destroyNative(JNIEnv *env, jclass cls, jlong handle) {
    kcount_b200_destroy((kcount_handle *)(intptr_t)handle);
}
