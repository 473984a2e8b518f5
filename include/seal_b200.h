/*
 * seal_b200.h -- C ABI (part of libbbduk_b200.so) of Seal's k-mer matching path on B200: the multi-value
 * table (one k-mer -> the set of reference sequences that contain it) and the per-pair assignment
 * ("Seal + multi-value tables").
 *
 * Plain C only. Every entry point names the reference interface it replaces (paths relative to
 * /root/reference/current). Returns 0 on success, non-zero on failure (seal_b200_last_error()).
 * No CPU fallback: seal_b200_create fails without a CUDA device.
 *
 * Covered: jgi.Seal's loader (jgi/Seal.java:1760-1990, Hamming neighbourhoods up to hdist=2) and its matching
 * block from "Do kmer matching" on (jgi/Seal.java:2186-2276): findBestMatch with the default count array
 * (:2864-2907), condenseLoose (:2654-2667), filterTopScaffolds with the clear zone (:2697-2708), the minimum
 * hit rule (:2223), assignTogether (:2386-2453) / assignIndependently (:2462-2606) with the four ambiguous
 * modes, the three match modes, restrictleft / restrictright, qskip, speed, rskip, middle masking and forbidn.
 * Not covered (rejected by seal_b200_create or absent from seal_cfg): edist > 0, qhdist > 0, k > 31,
 * processcontainedref, taxonomy, barcodes, rename, the quality / length preamble (jgi/Seal.java:2049-2183,
 * served by bbduk_b200_qtrim) and usecountvector=t (whose kpt=f branch reads the wrong list, :2241).
 */
#ifndef SEAL_B200_H
#define SEAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SEAL_API __attribute__((visibility("default")))
#else
#define SEAL_API
#endif

typedef struct seal_handle seal_handle;

/* jgi/Seal.java:3315-3316 */
enum { SEAL_AMBIG_ALL = 1, SEAL_AMBIG_FIRST = 2, SEAL_AMBIG_TOSS = 3, SEAL_AMBIG_RANDOM = 4 };
enum { SEAL_MATCH_ALL = 1, SEAL_MATCH_FIRST = 2, SEAL_MATCH_UNIQUE = 3 };

/* User flags of seal.sh, before the constructor's derivations (jgi/Seal.java:486-571 are applied inside). */
typedef struct seal_cfg {
    int32_t struct_size;          /* sizeof(seal_cfg) */
    int32_t k;                    /* k= (1..31), default 31 */
    int32_t rcomp;                /* rcomp=t */
    int32_t mask_middle;          /* mm=t (jgi/Seal.java:3088) */
    int32_t mid_mask_len;         /* mm=<int>; 0 = 2-(k&1) when mask_middle (jgi/Seal.java:548-552) */
    int32_t forbid_ns;            /* forbidn=f; forced on when hdist < 1 (jgi/Seal.java:492) */
    int32_t hdist;                /* hdist= 0..2 */
    int32_t speed;                /* speed= 0..15 (jgi/Seal.java:2983-2989) */
    int32_t qskip;                /* qskip= (jgi/Seal.java:2792) */
    int32_t rskip;                /* rskip= (jgi/Seal.java:1785-1801) */
    int32_t restrict_left;        /* restrictleft= */
    int32_t restrict_right;       /* restrictright= */
    int32_t ambig_mode;           /* ambig= : SEAL_AMBIG_*, default RANDOM */
    int32_t match_mode;           /* match= : SEAL_MATCH_*, default ALL */
    int32_t keep_pairs_together;  /* kpt=t */
    int32_t clearzone;            /* cz= */
    float clearzone_fraction;     /* czf= */
    int32_t min_kmer_hits;        /* mkh= (>= 1) */
    float min_kmer_fraction;      /* mkf= */
    int32_t device;               /* CUDA device ordinal */
    int32_t table_load_pct;       /* hash-array load in percent (layout only), default 50 */
    int32_t ids_stride;           /* assigned ids written per unit by seal_b200_process (0 = none) */
    int32_t reserved[6];
} seal_cfg;

/* Per-batch results. A UNIT is a pair when the batch is paired and keep_pairs_together is set, else a read. */
typedef struct seal_out {
    int32_t *n_assigned; /* [n_units] reference sequences the unit was assigned to (stop-start, jgi/Seal.java:2452) */
    int32_t *first_id;   /* [n_units] first assigned id (finalList.get(start)), 0 if none; ids start at 1 (:129-131) */
    int32_t *n_sites;    /* [n_units] finalList.size after the clear-zone filter (:2220, :2270) */
    int32_t *max_hits;   /* [n_units] highest per-reference hit count (condenseLoose's return value) */
    int32_t *ids;        /* [n_units * ids_stride] assigned ids in order, 0-padded; may be NULL */
} seal_out;

/* readsIn, basesIn, readsMatched, basesMatched, readsUnmatched, basesUnmatched (jgi/Seal.java:2044-2046,
 * :2226-2227, :2443-2449, :2523-2531); the last two words are reserved. */
typedef struct seal_stats {
    int64_t reads_in, bases_in, reads_matched, bases_matched, reads_unmatched, bases_unmatched, reserved[2];
} seal_stats;

SEAL_API void seal_b200_cfg_default(seal_cfg *cfg);

/* Replaces: the table part of Seal's constructor (jgi/Seal.java:486-571) -- validates and derives constants. */
SEAL_API int seal_b200_create(const seal_cfg *cfg, seal_handle **out);

/* Replaces: LoadThread.addToMap(Read, skip) for n_seqs reference sequences (jgi/Seal.java:1760-1829); HOST
 * buffers, concatenated ASCII bases + offsets[n_seqs+1]. Ids are assigned in call order starting at 1. */
SEAL_API int seal_b200_add_ref(seal_handle *h, const uint8_t *bases, const int64_t *offsets, int32_t n_seqs);

/* Replaces: spawnLoadThreads' join (jgi/Seal.java:1261-1553): builds the device table.
 * v[3] = {storedKmers ("Added N kmers", :744), (k-mer, id) entries, refKmers}. */
SEAL_API int seal_b200_finalize(seal_handle *h, int64_t *v);

/* Units of a batch: n_reads/2 when paired and kpt, else n_reads. */
SEAL_API int64_t seal_b200_n_units(const seal_handle *h, int64_t n_reads, int32_t paired);

/* Replaces: ProcessThread.run's matching block for a batch (jgi/Seal.java:2186-2276). HOST buffers: ASCII
 * bases + offsets[n_reads+1] (paired: r1,r2 interleaved); first_numeric_id = Read.numericID of the batch's
 * first pair / read (ambig=random picks finalList[numericID % sites], :2403). Results to host arrays.
 * A unit may hit at most 1152 distinct reference ids (128 in shared memory + 1024 in the warp's scratch); beyond that the call
 * fails with an error instead of truncating the list. One call at a time per handle (the scratch belongs to the handle). */
SEAL_API int seal_b200_process(seal_handle *h, const uint8_t *bases, const int64_t *offsets, int64_t n_reads,
                               int32_t paired, int64_t first_numeric_id, const seal_out *out, seal_stats *stats);

/* Same on DEVICE buffers (32-bit offsets), asynchronous on `stream`; d_stats = 8 device words that are ADDED to. */
SEAL_API int seal_b200_process_device(seal_handle *h, const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_reads,
                                      int32_t paired, int64_t first_numeric_id, const seal_out *d_out,
                                      unsigned long long *d_stats, void *stream);

/* Replaces: scaffoldReadCounts / BaseCounts / FragCounts / AmbigReadCounts (jgi/Seal.java:2431-2441): running
 * totals since finalize, host arrays of n = n_seqs+1 words each (index = id; any may be NULL). */
SEAL_API int seal_b200_scaffold_counts(seal_handle *h, int64_t *reads, int64_t *bases, int64_t *frags, int64_t *ambig,
                                       int32_t n);

/* Table export for tests: every (key, id) entry sorted by key then id; *n_out = entries (may exceed cap). */
SEAL_API int seal_b200_table_export(seal_handle *h, uint64_t *keys, int32_t *ids, int64_t cap, int64_t *n_out);

SEAL_API int64_t seal_b200_launch_count(const seal_handle *h);
SEAL_API const char *seal_b200_last_error(seal_handle *h);
SEAL_API void seal_b200_destroy(seal_handle *h);

#ifdef __cplusplus
}
#endif
#endif
