/*
 * bbduk_b200.h -- C ABI of libbbduk_b200.so: BBDuk's per-read k-mer match-and-trim hot path on B200.
 *
 * This is the drop-in boundary. Everything here is plain C (pointers, sizes, POD structs); no C++
 * or torch types cross it. A JNI shim (jni/BBDukCuda.c) or any other FFI binds exactly these
 * symbols. Each entry point names the reference interface it replaces (paths relative to the
 * BBTools tree, "jgi/" = current/jgi, "bbduk/" = current/bbduk).
 *
 * Thread safety: one handle owns one device table. bbduk_b200_process* may be called from several
 * host threads on the same handle concurrently (the reference calls its index from THREADS
 * ProcessThreads, bbduk/BBDukS.java:317-319); each call takes a private stream + staging slot.
 * create/add_ref/finalize/destroy are single-threaded per handle.
 *
 * Errors: every function returns 0 on success, non-zero on failure; bbduk_b200_last_error() gives
 * the message. Nothing throws, nothing falls back to a CPU path.
 */
#ifndef BBDUK_B200_H
#define BBDUK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BBDUK_B200_ABI_VERSION 2

#if defined(__GNUC__)
#define BBDUK_API __attribute__((visibility("default")))
#else
#define BBDUK_API
#endif

/* generation: which reference main class' derived-constant quirks to follow */
#define BBDUK_GEN_JGI 0   /* jgi.BBDuk   (bbdukOld.sh): unset mink -> 6   (jgi/BBDuk.java:804) */
#define BBDUK_GEN_S 1     /* bbduk.BBDukS (bbduk.sh)   : unset mink stays -1 (bbduk/BBDukParser.java:245) */

/*
 * User-level parameters with the meaning of the bbduk.sh flags of the same name. The derived
 * constants (mask, shift2, middleMask, minlen2, forbidNs, useShortKmers, kbig ...) are computed by
 * the library exactly as jgi/BBDuk.java:583-585, :672-877 does. Fill with bbduk_b200_cfg_default()
 * first, then override.
 */
typedef struct bbduk_cfg {
    int32_t struct_size;          /* = sizeof(bbduk_cfg); ABI guard */
    int32_t generation;           /* BBDUK_GEN_* */
    int32_t k;                    /* k=        ; <=0 means unset -> 27 (jgi/BBDuk.java:708); >31 -> kbig */
    int32_t mink;                 /* mink=     ; -1 unset */
    int32_t use_short_kmers;      /* usk=t     ; (jgi/BBDuk.java:255) */
    int32_t hdist;                /* hdist=    */
    int32_t hdist2;               /* hdist2=   ; -1 -> hdist  (jgi/BBDuk.java:583) */
    int32_t edist;                /* edist=    */
    int32_t edist2;               /* edist2=   ; -1 -> edist */
    int32_t qhdist;               /* qhdist=   */
    int32_t qhdist2;              /* qhdist2=  ; -1 -> qhdist */
    int32_t rcomp;                /* rcomp=    ; default 1 */
    int32_t mask_middle;          /* mm=t/f    ; default 1 */
    int32_t mid_mask_len;         /* mm=<n>    ; 0 = automatic 2-(k&1) */
    int32_t forbid_ns;            /* fn=       */
    int32_t ktrim_left;           /* ktrim=l   */
    int32_t ktrim_right;          /* ktrim=r   ; both = ktrim=rl / ktrimtips */
    int32_t ktrim_n;              /* ktrim=n / kmask= */
    int32_t ksplit;               /* ksplit=t  */
    int32_t ktrim_exclusive;      /* ktrimexclusive= */
    int32_t trim_pad;             /* tp=       */
    int32_t restrict_left;        /* restrictleft=  */
    int32_t restrict_right;       /* restrictright= */
    int32_t skip_r1;              /* skipr1=   */
    int32_t skip_r2;              /* skipr2=   */
    int32_t qskip;                /* qskip=    ; default 1 */
    int32_t speed;                /* speed=    ; 0..16 */
    int32_t min_skip;             /* minskip=  ; default 1 */
    int32_t max_skip;             /* maxskip=  ; default 1 */
    int32_t max_bad_kmers;        /* mbk= (mkh-1) ; default 0 */
    float   min_kmer_fraction;    /* mkf=      */
    float   min_covered_fraction; /* mcf=      */
    int32_t find_best_match;      /* fbm=      */
    int32_t kmask_fully_covered;  /* mfc=      */
    int32_t kmask_lowercase;      /* kmask=lc  */
    int32_t trim_symbol;          /* kmask=<c> ; default 'N' */
    int32_t min_read_length;      /* minlen=   ; default 10 */
    float   min_len_fraction;     /* mlf=      ; default 0 */
    int32_t require_both_bad;     /* rieb=f    ; default 0 */
    int32_t trim_pairs_evenly;    /* tpe=      */
    int32_t trim_failures_to_1bp; /* tossbrokenreads-style "trimfailuresto1bp" (jgi/BBDuk.java:3260-3266) */
    int32_t device;               /* CUDA device ordinal; -1 = current device */
    int32_t table_load_pct;       /* device table load factor in percent (layout only, no effect on results); 0 -> 50 */
    int32_t minlen2;              /* 0 = derive it. > 0: the caller has derived its constants already and hands in ITS minlen2:
                                     bbduk/BBDukParser.java:276 computes minlen2 from maskMiddle BEFORE useShortKmers / kbig>k
                                     switch maskMiddle off (:239-242, :290-294), so a host that marshals the parser's final
                                     maskMiddle / midMaskLen could not reproduce it (forbidNs and the distances are idempotent) */
    int32_t reserved[6];
} bbduk_cfg;

/* per-read flag bits written to bbduk_out.flags */
#define BBDUK_F_DISCARDED 0x01 /* read.discarded() after the k-mer block (jgi/BBDuk.java:2740-2777, :2848-2856) */
#define BBDUK_F_REMOVED   0x02 /* pair-level 'remove' (shouldRemove, jgi/BBDuk.java:2794, :2860, :3286-3289) */
#define BBDUK_F_KTRIMMED  0x04 /* the read's own ktrim/kmask call returned x>0 */
#define BBDUK_F_TPE       0x08 /* shortened by trimpairsevenly (jgi/BBDuk.java:2801-2811) */
#define BBDUK_F_SPLIT     0x10 /* ksplit produced a second segment [count, len-1) */
#define BBDUK_F_TBO       0x20 /* shortened by trim-by-overlap (bbduk_b200_tbo; jgi/BBDuk.java:2911-2924) */
#define BBDUK_F_QTRIMMED  0x40 /* shortened by quality trimming (bbduk_b200_qtrim; jgi/BBDuk.java:3077-3090) */
#define BBDUK_F_POLYTRIMMED 0x80 /* shortened by poly-A / poly-G / poly-C trimming (bbduk_b200_qtrim; jgi/BBDuk.java:2954-3052) */

/*
 * Per-read outputs, struct of arrays, each n_reads long; any pointer may be NULL (not wanted).
 * Reads of a pair are adjacent (2i, 2i+1). Host pointers for bbduk_b200_process, device pointers
 * for bbduk_b200_process_device.
 *   id0     scaffold id credited for this read (the id whose scaffoldReadCounts is incremented,
 *           e.g. jgi/BBDuk.java:3984-3992, :3439-3445), -1 if none. ktrim=rl: the right-tip credit.
 *   id0b    ktrim=rl only: the left-tip credit, else -1.
 *   lo, hi  the read keeps original bases [lo, hi) after the k-mer block (ktrim + tpe + ksplit).
 *   flags   BBDUK_F_*.
 *   count   the scan's return value: ktrim/ktrimTips bases trimmed; kmask BitSet cardinality;
 *           kfilter hit count as returned by countSetKmers / countCoveredBases / countSetKmersBig,
 *           or the id returned by findBestMatch; ksplit with BBDUK_F_SPLIT: start s of the new mate,
 *           which is original bases [s, len-1) -- the reference's subRead(rightmost+1, len-1) is
 *           end-exclusive and drops the last base (jgi/BBDuk.java:4370, stream/Read.java:3729-3731).
 *   maskbits  kmask only: bit (i&31) of word mask_off[r]+(i>>5) set <=> base i of read r is masked
 *           (jgi/BBDuk.java:4187-4196). mask_off has n_reads+1 entries, in words.
 */
typedef struct bbduk_out {
    int32_t  *id0;
    int32_t  *id0b;
    int32_t  *lo;
    int32_t  *hi;
    uint8_t  *flags;
    int32_t  *count;
    uint32_t *maskbits;
    const int64_t *mask_off;
} bbduk_out;

/* Aggregate counters of one process call; same meaning as the reference's per-thread sums
 * (jgi/BBDuk.java:2812-2813, :2863-2868; summed at :2085-2131). */
typedef struct bbduk_stats {
    int64_t reads_in, bases_in;
    int64_t reads_ktrimmed, bases_ktrimmed;
    int64_t reads_kfiltered, bases_kfiltered;
    int64_t reads_out, bases_out;          /* pairs not removed: reads and kept bases */
} bbduk_stats;

/* Device table blobs, for replication to other GPUs (one NCCL broadcast each at build time). */
typedef struct bbduk_table_desc {
    int64_t n_slots;        /* key slots (power of two) */
    int64_t n_filter_words; /* 32-bit words of the on-chip pre-filter image */
    int64_t stored_kmers;   /* distinct keys == the reference's "Added N kmers" (jgi/BBDuk.java:1973) */
    int32_t n_scaffolds;    /* ids are 1..n_scaffolds */
    int32_t reserved;
    void   *d_keys;         /* uint64[n_slots], device */
    void   *d_vals;         /* int32[n_slots], device */
    void   *d_filter;       /* uint32[n_filter_words], device */
    int64_t scalars[8];     /* hash seeds / geometry the kernels need; opaque, broadcast as-is */
} bbduk_table_desc;

typedef struct bbduk_handle bbduk_handle;

/* Library/ABI version (BBDUK_B200_ABI_VERSION). */
BBDUK_API int bbduk_b200_version(void);

/* Defaults of jgi.BBDuk's constructor (jgi/BBDuk.java:107-147, :4953-4977). */
BBDUK_API void bbduk_b200_cfg_default(bbduk_cfg *cfg);

/* The derived constants for a configuration, without touching the GPU (host logic; used by tests and by
 * the Java side to print the reference's "maskMiddle was disabled" style notices). v[16] =
 * {k, kbig, mink, useShortKmers, maskMiddle, midMaskLen, minlen, minlen2, minminlen, forbidNs,
 *  hammingDistance, hammingDistance2, middleMask, mask, kfilter, removePairsIfEitherBad}. */
BBDUK_API int bbduk_b200_describe_cfg(const bbduk_cfg *cfg, int64_t *v);

/* Replaces: BBDuk constructor's constant derivation + index allocation
 * (jgi/BBDuk.java:583-877, :1017-1021; bbduk/BBDukLoader.java:35-76). */
BBDUK_API int bbduk_b200_create(const bbduk_cfg *cfg, bbduk_handle **out);

/* Replaces: spawnLoadThreads' scaffold numbering + LoadThread.addToMap scan for a block of
 * reference sequences (jgi/BBDuk.java:1849-1863, :2210-2288). bases = concatenated ASCII,
 * offsets[n_seqs+1]. Scaffold ids continue from the previous call (first id 1). Host pointers. */
BBDUK_API int bbduk_b200_add_ref(bbduk_handle *h, const uint8_t *bases, const int64_t *offsets, int32_t n_seqs);

/* Replaces: the join of the LoadThreads ("Added N kmers", jgi/BBDuk.java:1945-1976): expands the
 * hdist/edist neighbourhoods and short-k-mer tails on the device and fills the hash array
 * (kmer.AbstractKmerTable.setIfNotPresent semantics: key -> smallest scaffold id). */
BBDUK_API int bbduk_b200_finalize(bbduk_handle *h, int64_t *stored_kmers);

/* Replaces: the k-mer block of ProcessThread's per-pair loop for one batch of reads
 * (jgi/BBDuk.java:2727-2873 == bbduk/BBDukProcessorS.java:947-1093), including ktrim / ktrimTips /
 * kmask / ksplit / countSetKmers / countCoveredBases / findBestMatch / countSetKmersBig and
 * TrimRead.trimToPosition's coordinate rule. HOST buffers; copies to/from the device inside.
 * paired!=0: reads 2i and 2i+1 are mates (pairnum 0 / 1). stats may be NULL. */
BBDUK_API int bbduk_b200_process(bbduk_handle *h, const uint8_t *bases, const int64_t *offsets, int64_t n_reads,
                       int32_t paired, const bbduk_out *out, bbduk_stats *stats);

/* The same call for a host that already holds its reads 2-bit packed (e.g. a FASTQ parser that packs while it scans, the
 * role of stream/FastqStreamer.java + dna/AminoAcid.java:269-285 fused): F = big-endian 2-bit codes, 16 bases per word, D =
 * defined bits (bit 15-b = base b of the group), both over the CONCATENATED bases exactly as bbduk_b200_pack_bases writes them
 * (group g = bases [16g, 16g+16)); offsets as in bbduk_b200_process. 0.375 B/base cross PCIe instead of 1 and no host
 * packing pass runs. Modes the tuned kernels do not serve are spelled out again on the device (undefined -> 'N'); kmask,
 * whose output depends on the bases' case, is refused. */
BBDUK_API int bbduk_b200_process_packed(bbduk_handle *h, const uint32_t *F, const uint16_t *D, const int64_t *offsets,
                                        int64_t n_reads, int32_t paired, const bbduk_out *out, bbduk_stats *stats);

/* Same, on DEVICE buffers (bases, 32-bit offsets[n_reads+1], outputs), asynchronous on `stream`
 * (a cudaStream_t, NULL = default stream). total bases < 4 GiB per call. d_stats: device
 * bbduk_stats to accumulate into, may be NULL. */
BBDUK_API int bbduk_b200_process_device(bbduk_handle *h, const uint8_t *d_bases, const uint32_t *d_offsets,
                              int64_t n_reads, int32_t paired, const bbduk_out *d_out,
                              bbduk_stats *d_stats, void *stream);

/* Optional hint for bbduk_b200_process_device: an upper bound on the read length of the batches to come
 * (reads longer than the hint are still handled, through the generic kernel). 0 = unknown: every call
 * then measures the batch with one extra reduction kernel and a stream synchronisation. */
BBDUK_API int bbduk_b200_set_max_read_len(bbduk_handle *h, int32_t max_read_len);

/* Per-scaffold hit accounting accumulated on the device by process calls, index 0..n_scaffolds
 * (replaces scaffoldReadCounts/scaffoldBaseCounts, jgi/BBDuk.java:1968-1969, :3984-3992). */
BBDUK_API int bbduk_b200_scaffold_counts(bbduk_handle *h, int64_t *read_counts, int64_t *base_counts, int32_t n);

/* Table replication across GPUs: rank 0 describes its table; other ranks allocate the same
 * geometry with _table_alloc, receive the three blobs (NCCL broadcast done by the caller on the
 * returned device pointers), then _table_commit. */
BBDUK_API int bbduk_b200_table_describe(bbduk_handle *h, bbduk_table_desc *desc);
BBDUK_API int bbduk_b200_table_alloc(bbduk_handle *h, bbduk_table_desc *desc /* in: geometry+scalars; out: pointers */);
BBDUK_API int bbduk_b200_table_commit(bbduk_handle *h);

/*
 * Single-process multi-GPU (SURVEY.md 8b `device_ids[] / n_devices`, 8e). The reference runs THREADS ProcessThreads of ONE
 * JVM against one shared index (bbduk/BBDukS.java:317-319, bbduk/BBDukProcessorS.java:768); a JNI caller has no torchrun.
 * bbduk_b200_replicate gives every listed GPU its own handle holding a copy of `src`'s finished table: one NCCL
 * broadcast per blob (keys, ids, filter images) over NVLink with ncclCommInitAll inside this process (libnccl.so.2 is
 * resolved with dlopen at run time), or cudaMemcpyPeerAsync for targets NCCL cannot serve (a target on src's own GPU,
 * duplicate ordinals, libnccl missing; BBDUK_B200_REPLICATE=peer forces it). out[n_devices] receives the new handles
 * (same bbduk_cfg as src, cfg.device = device_ids[i]); destroy them with bbduk_b200_destroy. On failure nothing is
 * left allocated. bbduk_b200_replica_transport: 0 = table built on this handle, 1 = received by NCCL broadcast,
 * 2 = received by peer copy.
 */
BBDUK_API int bbduk_b200_replicate(bbduk_handle *src, const int32_t *device_ids, int32_t n_devices, bbduk_handle **out);
BBDUK_API int bbduk_b200_replica_transport(bbduk_handle *h);

/* Replaces: the split of the input among the ProcessThreads (reads are independent units): the batch is
 * cut into n_handles contiguous slices (pairs never split), slice i runs bbduk_b200_process on handles[i] from its own
 * host thread, results land in input order in the caller's arrays and the counters are summed as the reference sums its
 * per-thread counters (jgi/BBDuk.java:2085-2131). HOST buffers, same contract as bbduk_b200_process. */
BBDUK_API int bbduk_b200_process_sharded(bbduk_handle **handles, int32_t n_handles, const uint8_t *bases,
                                         const int64_t *offsets, int64_t n_reads, int32_t paired, const bbduk_out *out,
                                         bbduk_stats *stats);
/* Per-scaffold hit counts summed over the handles of a replica set. */
BBDUK_API int bbduk_b200_scaffold_counts_sum(bbduk_handle **handles, int32_t n_handles, int64_t *read_counts,
                                             int64_t *base_counts, int32_t n);

/* Replaces: AbstractKmerTable.dumpKmersAsBytes behind BBDukIndex.dump (bbduk/BBDukIndex.java:77, bbduk/BBDukLoader.java:161-166):
 * copies the stored (key, scaffold id) pairs to HOST arrays of `cap` entries, in table order (compare as a set). The key
 * carries its length marker (toValue, bbduk/BBDukIndexMask2.java:533-545). *n_out = number of stored keys; entries beyond
 * cap are not written (call with cap = stored_kmers). */
BBDUK_API int bbduk_b200_table_export(bbduk_handle *h, uint64_t *keys, int32_t *ids, int64_t cap, int64_t *n_out);
/* refKmers of the finished table (the loader's count of reference k-mers seen, bbduk/BBDukLoader.java:461). */
BBDUK_API int64_t bbduk_b200_ref_kmers(bbduk_handle *h);

/* Bytes bbduk_b200_process / bbduk_b200_process_packed have copied host->device and device->host on this handle so far
 * (what really crossed PCIe: packed chunks count 0.375 B/base, ASCII chunks 1; bench.py's e2e.h2d_bytes_per_step). */
BBDUK_API int bbduk_b200_transfer_bytes(bbduk_handle *h, int64_t *h2d, int64_t *d2h);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
BBDUK_API int64_t bbduk_b200_launch_count(bbduk_handle *h);

/* Bench/test helper (no reference counterpart; the reference's generators need a JVM): fills DEVICE
 * buffers with n_pairs interleaved synthetic 2 x read_len bp pairs with adapter read-through, the
 * cfg-2 workload of SURVEY.md 8d. bases: 2*n_pairs*read_len bytes; offsets: 2*n_pairs+1 words.
 * Byte-identical to bbtools_b200/synth.py:paired_adapter_reads. */
BBDUK_API int bbduk_b200_synth_pairs(uint8_t *d_bases, uint32_t *d_offsets, int64_t n_pairs, int64_t first_pair,
                                     int32_t read_len, uint64_t seed, int32_t sub_per_10k, int32_t n_per_10k,
                                     void *stream);

/* Bench/test helpers for the cfg-3 / cfg-4 workloads: a uniform-ACGT reference (synth.py:random_reference, one
 * byte per base into a DEVICE buffer) and SE reads of which contam_pct % are copied from that reference
 * (synth.py:contaminant_reads). */
BBDUK_API int bbduk_b200_synth_reference(uint8_t *d_out, int64_t n, uint64_t seed, void *stream);
BBDUK_API int bbduk_b200_synth_contam(uint8_t *d_bases, uint32_t *d_offsets, int64_t n_reads, int64_t first_read,
                                      int32_t read_len, const uint8_t *d_ref, int64_t ref_len, uint64_t seed,
                                      int32_t contam_pct, int32_t sub_per_10k, int32_t n_per_10k, void *stream);

/*
 * Trim by overlap (tbo=t), the step that follows the k-mer block in the canonical adapter-trimming command
 * (`ktrim=r k=23 mink=11 hdist=1 tpe tbo`). Parameters with the meaning of the bbduk.sh flags; -1 / 0 = the reference's
 * default (jgi/BBDuk.java:5368-5371, :712-728).
 */
typedef struct bbduk_tbo_cfg {
    int32_t struct_size;     /* = sizeof(bbduk_tbo_cfg) */
    int32_t strict_overlap;  /* strictoverlap= ; default 1 */
    int32_t min_overlap0;    /* -1 -> 7  */
    int32_t min_overlap;     /* minoverlap= ; -1 -> 14 */
    int32_t min_insert0;     /* -1 -> 16 */
    int32_t min_insert;      /* mininsert= ; -1 -> 40 */
    int32_t qual_offset;     /* ASCII offset of the quality bytes; 0 -> 33 */
    float   mee_filter;      /* 0 -> 15 (strict) / off (loose) */
    int32_t reserved[4];
} bbduk_tbo_cfg;
BBDUK_API void bbduk_b200_tbo_cfg_default(bbduk_tbo_cfg *cfg);

/* Replaces: the tbo block of the per-pair loop (jgi/BBDuk.java:2878-2926) for one batch that has been through
 * bbduk_b200_process: the expectedErrors guard (needs quals; NULL = reads without qualities, guard passes),
 * r2.reverseComplementFast(), BBMergeOverlapper.mateByOverlapRatio (jgi/BBMergeOverlapper.java:98-136, :411-621,
 * :785-836; useq=f), the minInsert cut, and TrimRead.trimToPosition(r, 0, bestInsert-1, 1) of both mates.
 * Reads are paired (2i, 2i+1) and currently keep [lo,hi); pairs whose flags carry BBDUK_F_REMOVED are skipped.
 * hi[] and flags[] (|= BBDUK_F_TBO) are updated in place; insert[n_reads/2] (may be NULL) receives the insert size
 * used, -1 for none, -2 for ambiguous; stats2 += {readsTrimmedByOverlap, basesTrimmedByOverlap}.
 * HOST buffers; quals has the layout of bases. Reads longer than 1008 bases after the k-mer block are an error. */
BBDUK_API int bbduk_b200_tbo(bbduk_handle *h, const bbduk_tbo_cfg *cfg, const uint8_t *bases, const uint8_t *quals,
                             const int64_t *offsets, int64_t n_reads, const int32_t *lo, int32_t *hi, uint8_t *flags,
                             int32_t *insert, int64_t *stats2);

/* Same on DEVICE buffers (32-bit offsets), asynchronous on `stream`; max_read_len = upper bound on the read lengths
 * (<= 1008); d_stats2 = device int64[2] to accumulate into, may be NULL. */
BBDUK_API int bbduk_b200_tbo_device(bbduk_handle *h, const bbduk_tbo_cfg *cfg, const uint8_t *d_bases, const uint8_t *d_quals,
                                    const uint32_t *d_offsets, int64_t n_reads, int32_t max_read_len, const int32_t *d_lo,
                                    int32_t *d_hi, uint8_t *d_flags, int32_t *d_insert, int64_t *d_stats2, void *stream);

/*
 * Poly-X trimming, quality trimming and the per-read quality / length / N filters (first part): the
 * blocks that follow the k-mer block and tbo in the per-pair loop, jgi/BBDuk.java:2954-3052 and :3074-3170
 * (= bbduk/BBDukProcessorS.java's copy; entropy masking / trimming, which sits between them, must be off). Parameters carry the meaning of the bbduk.sh flags; minlen, minlenfraction, rieb and
 * trimfailuresto1bp come from the handle's bbduk_cfg, as in the reference.
 */
typedef struct bbduk_qtrim_cfg {
    int32_t struct_size;       /* = sizeof(bbduk_qtrim_cfg) */
    int32_t qtrim_left;        /* qtrim=l / rl */
    int32_t qtrim_right;       /* qtrim=r / rl */
    float   trimq;             /* trimq= ; default 6 (jgi/BBDuk.java:126); trimE = phredToProbError(trimq) */
    int32_t min_base_quality;  /* mbq= ; 0 = off */
    int32_t max_ns;            /* maxns= ; -1 = off */
    int32_t max_read_length;   /* maxlen= ; 0 = unlimited */
    int32_t qual_offset;       /* subtracted from every quality byte; 33 for FASTQ text, 0 for Read.quality */
    int32_t trim_poly_a;       /* trimpolya= ; 0 = off (parse/Parser.java:386-411: a bare flag means 2) */
    int32_t trim_poly_g_left, trim_poly_g_right, filter_poly_g; /* trimpolyg[left|right]=, filterpolyg= */
    int32_t trim_poly_c_left, trim_poly_c_right, filter_poly_c; /* trimpolyc[left|right]=, filterpolyc= */
    int32_t max_non_poly;      /* maxnonpoly= ; default 1 */
    float   min_avg_quality;   /* maq= / minavgquality= ; 0 = off (average by error probability, stream/Read.java:2181-2226) */
    int32_t min_avg_quality_bases; /* maq=Q,N / maqb= : only the first N bases; 0 = all */
    float   max_n_rate;        /* maxnrate= / maxnfraction= ; a read with more than rate * length undefined bases is discarded
                                  (jgi/BBDuk.java:3138-3149); >= 1 (the default, 1) = off */
    int32_t min_consecutive_bases; /* mcb= / minconsecutivebases= ; reads without a run of that many defined bases are discarded
                                  (jgi/BBDuk.java:3151-3154, stream/Read.java:2846-2858); 0 = off */
    float   min_base_frequency; /* minbasefrequency= ; reads whose rarest of A C G T (upper case, stream/Read.java:2864-2874)
                                  occurs fewer than frequency * length times are discarded (jgi/BBDuk.java:3156-3159); 0 = off */
    int32_t trim_mode;         /* 0 = optimal (TrimRead.optimalMode, the default), 1 = qtrim=window / w[,N] (right end only: the
                                  first window whose quality sum is below window * trimq, shared/TrimRead.java:438-455),
                                  2 = optitrim=f (testLeft / testRight: trim through the last base of quality <= trimq that is
                                  met before min_good_interval good ones in a row, :416-475) */
    int32_t window_length;     /* qtrim=w,N ; default 4 */
    int32_t min_good_interval; /* trimgoodinterval= ; default 2 */
} bbduk_qtrim_cfg;
BBDUK_API void bbduk_b200_qtrim_cfg_default(bbduk_qtrim_cfg *cfg);

/* Replaces, for one batch that has been through bbduk_b200_process (and bbduk_b200_tbo): the poly-A / poly-G / poly-C blocks
 * (trimPolyA, trimPoly, detectPolyLeft / Right, jgi/BBDuk.java:4721-4825, each followed by its minlen test and
 * shouldRemove; the reference's poly-C filter of r2 looks at r1, :3035, and so does this), then TrimRead.trimFast in its default
 * "optimal" mode (shared/TrimRead.java:113-169, :348-410: the maximum-sum run of avgErrorRate - probError, single
 * precision, then trimByAmount(r, a, b, 1)), the minlen / maxlen test, shouldRemove, then minavgquality, minbasequality and
 * maxns, maxnrate, minconsecutivebases and minbasefrequency with their shouldRemove (jgi/BBDuk.java:3074-3170; (the rest is at its
 * defaults = off). Reads WITHOUT qualities (quals = NULL): the trimming rules fall back to trimming N's
 * (testLeftN / testRightN, :348-353, :416-418, :438-440, :457-459, :477-503) and minavgquality / minbasequality do not apply.
 * minavgquality compares -10*log10(expectedErrors/bases) in double precision with the threshold; the
 * device compares the error probability with the smallest float for which that test holds (found on the host with the
 * C library's log10), which is the same predicate. paired != 0: reads (2i, 2i+1) are mates. Units whose flags carry BBDUK_F_REMOVED are skipped.
 * lo[], hi[] and flags[] (BBDUK_F_QTRIMMED, BBDUK_F_POLYTRIMMED, BBDUK_F_DISCARDED, BBDUK_F_REMOVED) are updated in place;
 * stats8 += {readsQTrimmed, basesQTrimmed, readsQFiltered, basesQFiltered, readsNFiltered, basesNFiltered, readsPolyTrimmed,
 * basesPolyTrimmed}.
 * HOST buffers; quals has the layout of bases and is required when qtrim or mbq is set. */
BBDUK_API int bbduk_b200_qtrim(bbduk_handle *h, const bbduk_qtrim_cfg *cfg, const uint8_t *bases, const uint8_t *quals,
                               const int64_t *offsets, int64_t n_reads, int32_t paired, int32_t *lo, int32_t *hi, uint8_t *flags,
                               int64_t *stats8);

/* Same on DEVICE buffers (32-bit offsets), asynchronous on `stream`; d_stats8 = device int64[8] to accumulate into, may
 * be NULL. */
BBDUK_API int bbduk_b200_qtrim_device(bbduk_handle *h, const bbduk_qtrim_cfg *cfg, const uint8_t *d_bases, const uint8_t *d_quals,
                                      const uint32_t *d_offsets, int64_t n_reads, int32_t paired, int32_t *d_lo, int32_t *d_hi,
                                      uint8_t *d_flags, int64_t *d_stats8, void *stream);

/*
 * Low-entropy read filter (entropy=<cutoff>), the "Test entropy" block that follows the quality filters in the per-pair
 * loop (jgi/BBDuk.java:3175-3186) with tracker/EntropyTracker.java underneath (average over all windows of `window`
 * bases of the Shannon entropy of the window's k-mers, scaled to 0..1; a read passes iff highpass XOR (entropy < cutoff)).
 * rieb and trimfailuresto1bp come from the handle's bbduk_cfg. entropymask / entropytrim / entropymark are not covered.
 */
typedef struct bbduk_entropy_cfg {
    int32_t struct_size;  /* = sizeof(bbduk_entropy_cfg) */
    float   cutoff;       /* entropy= / minentropy= ; the tracker uses max(0, cutoff) (jgi/BBDuk.java:2518) */
    int32_t k;            /* entropyk= ; 0 -> 5 (tracker/EntropyTracker.java:1206); device path: k <= 5 */
    int32_t window;       /* entropywindow= ; 0 -> 50; device path: window - k + 1 <= 254 */
    int32_t high_pass;    /* entropyHighpass, default 1 */
    int32_t reserved[4];
} bbduk_entropy_cfg;
BBDUK_API void bbduk_b200_entropy_cfg_default(bbduk_entropy_cfg *cfg);

/* For one batch that has been through bbduk_b200_process (and tbo / qtrim): units (reads, or pairs 2i / 2i+1 if paired)
 * whose flags carry BBDUK_F_REMOVED are skipped; reads that are not discarded are measured on their kept interval [lo,hi).
 * flags[] (BBDUK_F_DISCARDED, BBDUK_F_REMOVED) and, with trimfailuresto1bp, hi[] are updated in place;
 * stats2 += {readsEFiltered, basesEFiltered}. HOST buffers. */
BBDUK_API int bbduk_b200_entropy(bbduk_handle *h, const bbduk_entropy_cfg *cfg, const uint8_t *bases, const int64_t *offsets,
                                 int64_t n_reads, int32_t paired, const int32_t *lo, int32_t *hi, uint8_t *flags, int64_t *stats2);

/* Same on DEVICE buffers (32-bit offsets), asynchronous on `stream`; d_stats2 = device int64[2], may be NULL. */
BBDUK_API int bbduk_b200_entropy_device(bbduk_handle *h, const bbduk_entropy_cfg *cfg, const uint8_t *d_bases,
                                        const uint32_t *d_offsets, int64_t n_reads, int32_t paired, const int32_t *d_lo,
                                        int32_t *d_hi, uint8_t *d_flags, int64_t *d_stats2, void *stream);

/* Entropy masking / trimming (entropymask=t|lc, entropytrim=; jgi/BBDuk.java:3055-3067, :4432-4478, :4505-4526; the step sits
 * between the poly-X block and quality trimming): every full window of `window` bases without an undefined base whose entropy
 * fails the cutoff marks its bases. mode 1: the marked bases are to become 'N' (quality 0), mode 2: lower case -- the call
 * writes the mark bits (bit j of the read's mask words = base j of its kept interval [lo,hi)) to maskbits[mask_off[i] ..],
 * mask_off[n_reads+1] in 32-bit words with room for ceil(read length / 32) words per read, and the caller rewrites its
 * bases; mode 3: the marked runs at both ends are trimmed (TrimRead.trimByAmount(r, left, right, 1)): lo[] / hi[] are updated
 * and the mask words come back zero. Reads of removed units or discarded reads are skipped.
 * stats2 += {readsEFiltered, basesEFiltered} = reads with / number of bases changed (an N stays an N, lower case stays lower
 * case) or trimmed. HOST buffers. */
BBDUK_API int bbduk_b200_entropy_mask(bbduk_handle *h, const bbduk_entropy_cfg *cfg, int32_t mode, const uint8_t *bases,
                                      const int64_t *offsets, int64_t n_reads, int32_t paired, int32_t *lo, int32_t *hi,
                                      const uint8_t *flags, uint32_t *maskbits, const int64_t *mask_off, int64_t *stats2);
/* Same on DEVICE buffers (32-bit offsets, mask_off relative to d_maskbits), asynchronous on `stream`. */
BBDUK_API int bbduk_b200_entropy_mask_device(bbduk_handle *h, const bbduk_entropy_cfg *cfg, int32_t mode, const uint8_t *d_bases,
                                             const uint32_t *d_offsets, int64_t n_reads, int32_t paired, int32_t *d_lo,
                                             int32_t *d_hi, const uint8_t *d_flags, uint32_t *d_maskbits,
                                             const int64_t *d_mask_off, int64_t *d_stats2, void *stream);

/*
 * The whole device part of the per-pair loop in ONE call: the k-mer block, then (each optional) trim by overlap, poly-X /
 * quality trimming with the quality / length / N filters, and the low-entropy filter, in the reference's order
 * (jgi/BBDuk.java:2727-2873, :2878-2926, :2954-3052 + :3074-3170, :3175-3186). The batch crosses PCIe once (bases, and
 * qualities if a step needs them); the steps hand lo / hi / flags to each other on the device.
 */
typedef struct bbduk_chain_cfg {
    int32_t struct_size;        /* = sizeof(bbduk_chain_cfg) */
    int32_t do_tbo, do_qtrim, do_entropy;
    bbduk_tbo_cfg tbo;
    bbduk_qtrim_cfg qtrim;
    bbduk_entropy_cfg entropy;
} bbduk_chain_cfg;
BBDUK_API void bbduk_b200_chain_cfg_default(bbduk_chain_cfg *cfg);

/* HOST buffers, as bbduk_b200_process (ktrim / kfilter modes; kmask and ksplit rewrite bases and are not chained).
 * out->lo, out->hi and out->flags are required, out->id0 / out->count optional. stats = the k-mer block's counters;
 * tbo_stats2, qtrim_stats8, entropy_stats2 are added to as by the single-step entry points (any may be NULL). */
BBDUK_API int bbduk_b200_process_chain(bbduk_handle *h, const bbduk_chain_cfg *cfg, const uint8_t *bases, const uint8_t *quals,
                                       const int64_t *offsets, int64_t n_reads, int32_t paired, const bbduk_out *out,
                                       bbduk_stats *stats, int64_t *tbo_stats2, int64_t *qtrim_stats8, int64_t *entropy_stats2);

/* Host helper (no GPU needed): the 2-bit packing bbduk_b200_process applies to a chunk before it crosses PCIe when
 * the tuned kernel takes the whole chunk (set BBDUK_B200_PACK_HOST=0 to ship ASCII instead). F[i] = big-endian
 * 2-bit codes of bases 16i..16i+15 (A0 C1 G2 T/U3, anything else 0), D[i] = "defined" bits (bit 15-b = base 16i+b);
 * both have (n+15)/16 entries. Same tables as dna/AminoAcid.java:269-285, :1289-1320. */
BBDUK_API int bbduk_b200_pack_bases(const uint8_t *bases, int64_t n, uint32_t *F, uint16_t *D);

BBDUK_API const char *bbduk_b200_last_error(bbduk_handle *h);
BBDUK_API void bbduk_b200_destroy(bbduk_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* BBDUK_B200_H */
