/*
 * kcount_b200.h -- C ABI (part of libbbduk_b200.so) of KmerCountExact's counting path on B200:
 * the shared 2-bit encode + HashArray insert/count loop (BASELINE.json configs[4]).
 *
 * Plain C only. Every entry point names the reference interface it replaces (paths relative to
 * /root/reference/current). Returns 0 on success, non-zero on failure (kcount_b200_last_error()).
 * No CPU fallback: kcount_b200_create fails without a CUDA device.
 *
 * A handle owns one device-resident open-addressed table of 16-byte slots {uint64 key, uint32 count, pad}
 * (kmer.HashArray1D's long[] array + int[] values, kmer/HashArray1D.java:41-89). It grows by
 * doubling when the next batch could push the load above 70 % (autoResize, kmer/HashArray1D.java:260-339).
 */
#ifndef KCOUNT_B200_H
#define KCOUNT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define KCOUNT_API __attribute__((visibility("default")))
#else
#define KCOUNT_API
#endif

typedef struct kcount_handle kcount_handle;

/* Replaces: KmerTableSet construction + allocateTables (kmer/KmerTableSet.java:273-303, :381-394).
 * k in [1,31]; rcomp!=0 -> canonical key max(kmer,rkmer) (kmer/KmerTableSet.java:1887-1895).
 * initial_keys: expected number of distinct k-mers (the reference's prealloc/initialSize); <=0 -> small. */
KCOUNT_API int kcount_b200_create(int32_t k, int32_t rcomp, int64_t initial_keys, int32_t device, kcount_handle **out);

/* Replaces: LoadThread.addKmersToTable for a batch of reads (kmer/KmerTableSet.java:652-716):
 * rolling 2-bit encode, an undefined base resets len, kmer and rkmer, every window of k defined bases
 * increments its key (saturating at INT32_MAX, kmer/HashArray1D.java:68-89). HOST buffers: concatenated
 * ASCII bases + offsets[n_reads+1]. */
KCOUNT_API int kcount_b200_add_reads(kcount_handle *h, const uint8_t *bases, const int64_t *offsets, int64_t n_reads);

/* Same on DEVICE buffers (32-bit offsets, total bases < 4 GiB), asynchronous on `stream` unless the table
 * has to grow first (then it synchronises the stream once). */
KCOUNT_API int kcount_b200_add_reads_device(kcount_handle *h, const uint8_t *d_bases, const uint32_t *d_offsets,
                                            int64_t n_reads, int64_t n_bases, void *stream);

/* v[4] = {readsIn, basesIn, kmersIn, unique k-mers}: the counters behind KmerCountExact's "Unique Kmers:"
 * line (jgi/KmerCountExact.java:355-367; kmer/KmerTableSet.java:509-513, :706). */
KCOUNT_API int kcount_b200_stats(kcount_handle *h, int64_t *v);

/* Replaces: fillHistogram (kmer/HashArray.java:577-588): hist[min(count,histmax)]++ over all keys;
 * hist has histmax+1 entries (host). */
KCOUNT_API int kcount_b200_khist(kcount_handle *h, int32_t histmax, int64_t *hist);

/* Replaces: dumpKmersAsBytes' traversal (kmer/AbstractKmerTable.java:490-516): every (key,count) with
 * mincount<=count<=maxcount, in table order (layout dependent: compare as a multiset). Host arrays of
 * capacity cap; *n_out = number of matching entries (may exceed cap; then only cap were written). */
KCOUNT_API int kcount_b200_dump(kcount_handle *h, int32_t mincount, int32_t maxcount, uint64_t *keys, int32_t *counts,
                                int64_t cap, int64_t *n_out);

/* Multi-GPU exchange: private tables per GPU, ONE exchange at the end.
 * _export_partitioned writes all entries to DEVICE arrays grouped by owner = mix(key) % n_parts
 * (part 0 first) and the per-part sizes to the HOST array part_sizes[n_parts]; d_keys/d_counts need
 * room for `unique` entries. _merge_device adds pre-counted DEVICE entries into this handle's table
 * (saturating), as received from the peers' all-to-all. */
KCOUNT_API int kcount_b200_export_partitioned(kcount_handle *h, int32_t n_parts, uint64_t *d_keys, int32_t *d_counts,
                                              int64_t *part_sizes, void *stream);
KCOUNT_API int kcount_b200_merge_device(kcount_handle *h, const uint64_t *d_keys, const int32_t *d_counts, int64_t n,
                                        void *stream);

/* Table geometry and kernel launches so far: v[3] = {n_slots, bytes, launches}. */
KCOUNT_API int kcount_b200_table_info(kcount_handle *h, int64_t *v);

/* Bench/test helper (no reference counterpart): fills DEVICE buffers with n_reads reads of read_len bases
 * sampled uniformly from a synthetic genome of genome_len bases (base i = f(seed_genome, i)), forward or
 * reverse-complement strand, sub_per_10k substitutions per 10 000 bases: the cfg-5 workload of SURVEY.md 8d.
 * Byte-identical to bbtools_b200/synth.py:genome_reads. */
KCOUNT_API int kcount_b200_synth_reads(uint8_t *d_bases, uint32_t *d_offsets, int64_t n_reads, int64_t first_read,
                                       int32_t read_len, int64_t genome_len, uint64_t seed, int32_t sub_per_10k,
                                       void *stream);

KCOUNT_API const char *kcount_b200_last_error(kcount_handle *h);
KCOUNT_API void kcount_b200_destroy(kcount_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* KCOUNT_B200_H */
