/*
 * bbduk_oracle.c -- TEST INFRASTRUCTURE ONLY. CPU restatement (plain C) of BBDuk's k-mer
 * match-and-trim path, following the reference's Java statement by statement. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it; the
 * product (libbbduk_b200.so) never links, loads or calls anything in oracle/.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or unit tests for this
 * path and no JVM exists in this image, so this port is pinned only by
 * (a) an independently written closed-form restatement (oracle/closed_form.py) that must agree with
 * it, (b) the distinct-key counts of SURVEY.md Appendix B, and (c) the reference's own in-code
 * asserts, which are reproduced here as ORA_ASSERT.
 *
 * Reference files followed (paths under /root/reference/current):
 *   dna/AminoAcid.java:269-285,1289-1320   2-bit code tables
 *   dna/AminoAcid.java:585-603             reverseComplementBinaryFast
 *   jgi/BBDuk.java:583-585,672-877         derived constants
 *   jgi/BBDuk.java:2210-2452               loader: addToMap, left/right shift tails, mutate
 *   jgi/BBDuk.java:3335-3386               getValue / getValueInner
 *   jgi/BBDuk.java:3395-3677               countSetKmers, countCoveredBases, findBestMatch, countSetKmersBig
 *   jgi/BBDuk.java:3679-4013               ktrim, ktrimTips, ktrimTip
 *   jgi/BBDuk.java:4022-4377               kmask, ksplit
 *   jgi/BBDuk.java:2587-2593,2727-2873,3260-3289   per-pair k-mer block
 *   shared/TrimRead.java:273-276,299-346   trimToPosition / trimByAmount
 *   stream/Read.java:1673-1683             numValidKmers
 * The physical table layout (kmer/HashArray1D, 7 ways) is NOT followed: results only depend on the
 * map semantics key -> first writer, with scaffolds loaded in file order.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/bbduk_b200.h"

#define ORA_ASSERT(c)                                                                                  \
    do {                                                                                               \
        if (!(c)) {                                                                                    \
            fprintf(stderr, "oracle assertion failed: %s (%s:%d)\n", #c, __FILE__, __LINE__);          \
            abort();                                                                                   \
        }                                                                                              \
    } while (0)

typedef int64_t jlong; /* Java long */
typedef uint64_t ulong64;

/* ------------------------------------------------------------------------------------------ */
/* dna/AminoAcid.java:269-285, :1289-1320                                                       */
static int8_t baseToNumber[128], baseToNumber0[128], baseToComplementNumber0[128];
static pthread_once_t tables_once = PTHREAD_ONCE_INIT;
static void init_tables(void) {
    memset(baseToNumber, -1, sizeof baseToNumber);
    memset(baseToNumber0, 0, sizeof baseToNumber0);
    memset(baseToComplementNumber0, 0, sizeof baseToComplementNumber0);
    const char *acgt = "ACGT";
    for (int i = 0; i < 4; i++) {
        int x = acgt[i], x2 = x + 32;
        baseToNumber0[x] = baseToNumber[x] = (int8_t)i;
        baseToNumber0[x2] = baseToNumber[x2] = (int8_t)i;
    }
    baseToNumber0['U'] = baseToNumber['U'] = 3;
    baseToNumber0['u'] = baseToNumber['u'] = 3;
    baseToComplementNumber0['A'] = baseToComplementNumber0['a'] = 3;
    baseToComplementNumber0['C'] = baseToComplementNumber0['c'] = 2;
    baseToComplementNumber0['G'] = baseToComplementNumber0['g'] = 1;
    baseToComplementNumber0['T'] = baseToComplementNumber0['t'] = 0;
    baseToComplementNumber0['U'] = baseToComplementNumber0['u'] = 0;
}

/* dna/AminoAcid.java:585-603 */
static jlong reverseComplementBinaryFast(jlong kmer, int k) {
    ulong64 x = ~(ulong64)kmer;
    x = ((x & 0x3333333333333333ULL) << 2) | ((x & 0xCCCCCCCCCCCCCCCCULL) >> 2);
    x = ((x & 0x0F0F0F0F0F0F0F0FULL) << 4) | ((x & 0xF0F0F0F0F0F0F0F0ULL) >> 4);
    x = ((x & 0x00FF00FF00FF00FFULL) << 8) | ((x & 0xFF00FF00FF00FF00ULL) >> 8);
    x = ((x & 0x0000FFFF0000FFFFULL) << 16) | ((x & 0xFFFF0000FFFF0000ULL) >> 16);
    x = (x << 32) | (x >> 32);
    x = x >> (2 * (32 - k));
    return (jlong)x;
}

/* ------------------------------------------------------------------------------------------ */
/* key -> id map with setIfNotPresent / getValue semantics (kmer/AbstractKmerTable.java:54,:61)  */
typedef struct {
    ulong64 *keys;
    int32_t *vals;
    int64_t cap, size; /* cap power of two */
} kmap;
#define KMAP_EMPTY 0xFFFFFFFFFFFFFFFFULL
static inline ulong64 kmap_hash(ulong64 x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
static void kmap_init(kmap *m, int64_t cap) {
    m->cap = cap;
    m->size = 0;
    m->keys = (ulong64 *)malloc(sizeof(ulong64) * cap);
    m->vals = (int32_t *)malloc(sizeof(int32_t) * cap);
    memset(m->keys, 0xFF, sizeof(ulong64) * cap);
}
static void kmap_free(kmap *m) {
    free(m->keys);
    free(m->vals);
    m->keys = NULL;
    m->vals = NULL;
}
static int kmap_set_if_not_present(kmap *m, ulong64 key, int32_t id);
static void kmap_grow(kmap *m) {
    kmap n;
    kmap_init(&n, m->cap * 2);
    for (int64_t i = 0; i < m->cap; i++)
        if (m->keys[i] != KMAP_EMPTY) kmap_set_if_not_present(&n, m->keys[i], m->vals[i]);
    kmap_free(m);
    *m = n;
}
/* returns 1 if the key was added, 0 if it was already present (first writer wins) */
static int kmap_set_if_not_present(kmap *m, ulong64 key, int32_t id) {
    if ((m->size + 1) * 10 > m->cap * 6) kmap_grow(m);
    int64_t mask = m->cap - 1, i = (int64_t)(kmap_hash(key) & (ulong64)mask);
    while (m->keys[i] != KMAP_EMPTY) {
        if (m->keys[i] == key) return 0;
        i = (i + 1) & mask;
    }
    m->keys[i] = key;
    m->vals[i] = id;
    m->size++;
    return 1;
}
static inline int32_t kmap_get(const kmap *m, ulong64 key) {
    int64_t mask = m->cap - 1, i = (int64_t)(kmap_hash(key) & (ulong64)mask);
    while (m->keys[i] != KMAP_EMPTY) {
        if (m->keys[i] == key) return m->vals[i];
        i = (i + 1) & mask;
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------ */
/* the tool instance: fields named as in jgi/BBDuk.java                                          */
struct ora {
    bbduk_cfg cfg;
    /* derived, jgi/BBDuk.java:672-877 */
    int k, k2, kbig, keff, mink, midMaskLen, maskMiddle, useShortKmers;
    int hammingDistance, hammingDistance2, editDistance, editDistance2, qHammingDistance, qHammingDistance2;
    int minSkip, maxSkip, forbidNs, rcomp;
    int restrictLeft, restrictRight, speed, qSkip, skipR1, skipR2;
    int speedMask2; /* generation == BBDUK_GEN_S: the speed= rule of bbduk/BBDukIndexMask2.java */
    int ktrimLeft, ktrimRight, ktrimN, ksplit, ktrimExclusive, kfilter, trimPad;
    int findBestMatch, kmaskFullyCovered, kmaskLowercase, trimSymbol;
    int maxBadKmers0, removePairsIfEitherBad, trimPairsEvenly, trimFailuresTo1bp;
    int minReadLength;
    float minLenFraction, minKmerFraction, minCoveredFraction;
    int minlen, minminlen, minlen2, shift, shift2;
    jlong mask, kmask, middleMask;
    int bitsPerBase, symbols, maxSymbol, symbolArrayLen;
    jlong symbolMask;
    jlong clearMasks[32], leftMasks[32], rightMasks[32], lengthMasks[32], setMasks[4][32];
    /* state */
    kmap map;
    int32_t nScaffolds;  /* scaffoldNames.size()-1 */
    jlong storedKmers;   /* "Added N kmers" */
    jlong refKmers;
    int64_t *scaffoldReadCounts, *scaffoldBaseCounts; /* index by id, allocated at finalize */
    int finalized;
};

static __thread char ora_errbuf[512];
const char *ora_error(void) { return ora_errbuf; }

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline jlong lmax(jlong a, jlong b) { return a > b ? a : b; }
/* shared/Tools.java:5158 */
static inline int mid3(int x, int y, int z) { return imax(imin(x, y), imin(imax(x, y), z)); }

/* jgi/BBDuk.java:5355-5357 */
static inline int isFullyDefined(uint8_t symbol) { return symbol < 128 && baseToNumber[symbol] >= 0; }

void ora_cfg_default(bbduk_cfg *c) {
    memset(c, 0, sizeof *c);
    c->struct_size = (int32_t)sizeof *c;
    c->generation = BBDUK_GEN_JGI;
    c->k = 0;
    c->mink = -1;
    c->hdist2 = c->edist2 = c->qhdist2 = -1;
    c->rcomp = 1;
    c->mask_middle = 1;
    c->qskip = 1;
    c->min_skip = c->max_skip = 1;
    c->trim_symbol = 'N';
    c->min_read_length = 10;
    c->device = -1;
}

struct ora *ora_create(const bbduk_cfg *c) {
    pthread_once(&tables_once, init_tables);
    if (!c || c->struct_size != (int32_t)sizeof(bbduk_cfg)) {
        snprintf(ora_errbuf, sizeof ora_errbuf, "bad cfg struct_size");
        return NULL;
    }
    struct ora *o = (struct ora *)calloc(1, sizeof *o);
    o->cfg = *c;
    /* jgi/BBDuk.java:583-585 */
    o->hammingDistance = c->hdist;
    o->editDistance = c->edist;
    o->qHammingDistance = c->qhdist;
    o->hammingDistance2 = (c->hdist2 == -1 ? c->hdist : c->hdist2);
    o->qHammingDistance2 = (c->qhdist2 == -1 ? c->qhdist : c->qhdist2);
    o->editDistance2 = (c->edist2 == -1 ? c->edist : c->edist2);
    /* :672-701 */
    o->hammingDistance = imax(o->editDistance, o->hammingDistance);
    o->hammingDistance2 = imax(o->editDistance2, o->hammingDistance2);
    o->minSkip = imax(1, imin(c->min_skip, c->max_skip));
    o->maxSkip = imax(o->minSkip, c->max_skip);
    o->forbidNs = (c->forbid_ns || o->hammingDistance < 1);
    o->trimSymbol = c->trim_symbol;
    o->kmaskLowercase = c->kmask_lowercase;
    o->kmaskFullyCovered = c->kmask_fully_covered;
    o->trimPairsEvenly = c->trim_pairs_evenly;
    o->restrictLeft = imax(c->restrict_left, 0);
    o->restrictRight = imax(c->restrict_right, 0);
    o->findBestMatch = c->find_best_match; /* rename is host-side; (rename || findBestMatch_) */
    o->speed = c->speed;
    o->speedMask2 = (c->generation == BBDUK_GEN_S);
    o->qSkip = c->qskip;
    o->skipR1 = c->skip_r1;
    o->skipR2 = c->skip_r2;
    o->trimFailuresTo1bp = c->trim_failures_to_1bp;
    o->removePairsIfEitherBad = (!c->require_both_bad) && (!c->trim_failures_to_1bp); /* :631 */
    o->rcomp = c->rcomp;                                                             /* amino=false */
    o->trimPad = c->trim_pad;
    o->minReadLength = c->min_read_length;
    o->minLenFraction = c->min_len_fraction;
    /* :706-710 */
    int k_ = c->k, kbig_;
    const int maxSupportedK = 31;
    if (k_ <= 0) k_ = 27;
    kbig_ = (k_ > maxSupportedK ? k_ : -1);
    k_ = imin(k_, maxSupportedK);
    /* :764-780 */
    if (c->ktrim_left || c->ktrim_right || c->ktrim_n || c->ksplit) {
        if (kbig_ > k_) kbig_ = k_;
    }
    if ((o->speed > 0 || o->qSkip > 1) && kbig_ > k_) kbig_ = k_;
    /* :787-805 */
    o->k = k_;
    o->k2 = k_ - 1;
    o->kbig = kbig_;
    o->keff = imax(o->k, o->kbig);
    o->maskMiddle = c->mask_middle;
    o->midMaskLen = c->mid_mask_len;
    if (c->mid_mask_len > 0) o->maskMiddle = 1; /* mm=<n>: maskMiddle=midMaskLen>0 (:309-310) */
    if (o->maskMiddle) {
        o->midMaskLen = (o->midMaskLen > 0 ? o->midMaskLen : 2 - (o->k & 1));
    } else {
        o->midMaskLen = 0;
    }
    if (o->kbig > o->k) {
        o->minSkip = o->maxSkip = 0;
        if (o->maskMiddle) {
            o->maskMiddle = 0;
            o->midMaskLen = 0;
        }
    }
    if (c->generation == BBDUK_GEN_JGI) {
        o->mink = imin((c->mink < 1 ? 6 : c->mink), o->k); /* jgi/BBDuk.java:804 */
    } else {
        o->mink = imin(c->mink, o->k); /* bbduk/BBDukParser.java mink=Tools.min(mink,k) */
    }
    o->maxBadKmers0 = c->max_bad_kmers;
    /* :807-841 */
    o->bitsPerBase = 2;
    o->maxSymbol = 3;
    o->symbols = 4;
    o->symbolArrayLen = (64 + 2 - 1) / 2;
    o->symbolMask = 3;
    for (int i = 0; i < o->symbolArrayLen; i++) {
        o->clearMasks[i] = (jlong) ~((ulong64)o->symbolMask << (2 * i));
        o->leftMasks[i] = (jlong)(((ulong64)-1LL) << (2 * i));
        o->rightMasks[i] = ~(jlong)(((ulong64)-1LL) << (2 * i));
        o->lengthMasks[i] = (jlong)(1ULL << (2 * i));
        for (jlong j = 0; j < o->symbols; j++) o->setMasks[j][i] = (jlong)((ulong64)j << (2 * i));
    }
    o->minlen = o->k - 1;
    o->minminlen = o->mink - 1;
    o->minlen2 = (o->maskMiddle ? (o->k - o->midMaskLen) / 2 : o->k); /* note: before usk disables mm */
    if (c->minlen2 > 0 && c->minlen2 <= o->k) o->minlen2 = c->minlen2; /* include/bbduk_b200.h: a host's own derived value */
    o->shift = 2 * o->k;
    o->shift2 = o->shift - 2;
    o->mask = (o->shift > 63 ? -1LL : ~(jlong)(((ulong64)-1LL) << o->shift));
    o->kmask = o->lengthMasks[o->k];
    o->minKmerFraction = (c->min_kmer_fraction > 0 ? c->min_kmer_fraction : 0);
    o->minCoveredFraction = (c->min_covered_fraction > 0 ? c->min_covered_fraction : 0);
    /* :849-856 */
    o->useShortKmers = c->use_short_kmers;
    if (c->mink > 0 && c->mink < o->k) o->useShortKmers = 1;
    if (o->useShortKmers) {
        if (o->maskMiddle) {
            o->maskMiddle = 0;
            o->midMaskLen = 0;
        }
    }
    /* :858-866 */
    o->ktrimRight = c->ktrim_right;
    o->ktrimLeft = c->ktrim_left;
    o->ktrimN = c->ktrim_n;
    o->ksplit = c->ksplit;
    o->ktrimExclusive = c->ktrim_exclusive;
    o->kfilter = !(o->ktrimRight || o->ktrimLeft || o->ktrimN || o->ksplit);
    if (o->findBestMatch && o->kfilter && o->kbig > o->k) {
        snprintf(ora_errbuf, sizeof ora_errbuf, "K must be less than 32 in 'findBestMatch' mode");
        free(o);
        return NULL;
    }
    if (o->useShortKmers && !(o->ktrimRight || o->ktrimLeft || o->ktrimN || o->ksplit)) {
        snprintf(ora_errbuf, sizeof ora_errbuf, "Setting mink or useShortKmers also requires setting a ktrim mode");
        free(o);
        return NULL;
    }
    /* :868-877 */
    if (o->maskMiddle) {
        if (!(o->k > o->midMaskLen + 1)) {
            snprintf(ora_errbuf, sizeof ora_errbuf, "k must exceed midMaskLen+1");
            free(o);
            return NULL;
        }
        int bits = o->midMaskLen * 2;
        int shift = ((o->k - o->midMaskLen) / 2) * 2;
        o->middleMask = ~((~(jlong)(((ulong64)-1LL) << bits)) << shift);
    } else {
        o->middleMask = -1LL;
    }
    kmap_init(&o->map, 1 << 16);
    o->nScaffolds = 0;
    return o;
}

void ora_destroy(struct ora *o) {
    if (!o) return;
    kmap_free(&o->map);
    free(o->scaffoldReadCounts);
    free(o->scaffoldBaseCounts);
    free(o);
}

/* jgi/BBDuk.java:4673-4685 */
static inline jlong toValue(const struct ora *o, jlong kmer, jlong rkmer, jlong lengthMask) {
    ORA_ASSERT(lengthMask == 0 || (kmer < lengthMask && rkmer < lengthMask));
    const jlong value = (o->rcomp ? lmax(kmer, rkmer) : kmer);
    return (value & o->middleMask) | lengthMask;
}
/* :4693-4695 */
static inline jlong rcomp_(jlong kmer, int len) { return reverseComplementBinaryFast(kmer, len); }
/* shared/Tools.java:5482-5497 hash64plus2 (MurmurHash3 finalizer, sign bit cleared, the top values folded) */
static inline jlong hash64plus2(jlong key0) {
    ulong64 key = (ulong64)key0;
    key ^= key >> 33;
    key *= 0xff51afd7ed558ccdULL;
    key ^= key >> 33;
    key *= 0xc4ceb9fe1a85ec53ULL;
    key ^= key >> 33;
    key &= 0x7FFFFFFFFFFFFFFFULL;
    return (jlong)(key < 0x7FFFF800FFFFFFFFULL ? key : (key - 0x7FFFF800FFFFFFFFULL) * 64ULL);
}
/* jgi.BBDuk :4702-4713 (key%17); bbduk.BBDukS's default index bbduk/BBDukIndexMask2.java:566-577 (hash bits 16-19) */
static inline int passesSpeed(const struct ora *o, jlong key) {
    if (o->speedMask2) return o->speed < 2 || (((hash64plus2(key) >> 16) & 15) + 1) >= o->speed;
    return o->speed < 1 || ((key & INT64_MAX) % 17) >= o->speed;
}
static inline int failsSpeed(const struct ora *o, jlong key) {
    if (o->speedMask2) return o->speed > 1 && (((hash64plus2(key) >> 16) & 15) + 1) < o->speed;
    return o->speed > 0 && ((key & INT64_MAX) % 17) < o->speed;
}

/* ------------------------------------------------------------------------------------------ */
/* loader: jgi/BBDuk.java:2359-2452 (single map; key%WAYS routing dropped, see header)          */
static jlong mutate(struct ora *o, const jlong kmer, const jlong rkmer, const int len, const int id, const int dist,
                    const jlong extraBase) {
    jlong added = 0;
    const jlong key = toValue(o, kmer, rkmer, o->lengthMasks[len]);
    added += kmap_set_if_not_present(&o->map, (ulong64)key, id);
    if (dist > 0) {
        const int dist2 = dist - 1;
        /* Sub */
        for (int j = 0; j < o->symbols; j++) {
            for (int i = 0; i < len; i++) {
                const jlong temp = (kmer & o->clearMasks[i]) | o->setMasks[j][i];
                if (temp != kmer) {
                    jlong rtemp = rcomp_(temp, len);
                    added += mutate(o, temp, rtemp, len, id, dist2, extraBase);
                }
            }
        }
        if (o->editDistance > 0) {
            /* Del */
            if (extraBase >= 0 && extraBase <= o->maxSymbol) {
                for (int i = 1; i < len; i++) {
                    const jlong temp =
                        (kmer & o->leftMasks[i]) | ((jlong)((ulong64)kmer << 2) & o->rightMasks[i]) | extraBase;
                    if (temp != kmer) {
                        jlong rtemp = rcomp_(temp, len);
                        added += mutate(o, temp, rtemp, len, id, dist2, -1);
                    }
                }
            }
            /* Ins */
            const jlong eb2 = kmer & o->symbolMask;
            for (int i = 1; i < len; i++) {
                const jlong temp0 = (kmer & o->leftMasks[i]) | ((kmer & o->rightMasks[i]) >> 2);
                for (int j = 0; j < o->symbols; j++) {
                    const jlong temp = temp0 | o->setMasks[j][i - 1];
                    if (temp != kmer) {
                        jlong rtemp = rcomp_(temp, len);
                        added += mutate(o, temp, rtemp, len, id, dist2, eb2);
                    }
                }
            }
        }
    }
    return added;
}

static jlong addToMapK(struct ora *o, const jlong kmer, const jlong rkmer, const int len, const jlong extraBase,
                       const int id, const jlong kmask0, const int hdist, const int edist) {
    ORA_ASSERT(kmask0 == o->lengthMasks[len]);
    ORA_ASSERT((kmer & kmask0) == 0);
    jlong added;
    if (hdist == 0) {
        const jlong key = toValue(o, kmer, rkmer, kmask0);
        if (failsSpeed(o, key)) return 0;
        added = kmap_set_if_not_present(&o->map, (ulong64)key, id);
    } else if (edist > 0) {
        added = mutate(o, kmer, rkmer, len, id, edist, extraBase);
    } else {
        added = mutate(o, kmer, rkmer, len, id, hdist, -1);
    }
    return added;
}

/* :2299-2317 */
static jlong addToMapLeftShift(struct ora *o, jlong kmer, jlong rkmer, const jlong extraBase, const int id) {
    jlong added = 0;
    for (int i = o->k - 1; i >= o->mink; i--) {
        kmer = kmer & o->rightMasks[i];
        rkmer = (jlong)((ulong64)rkmer >> 2);
        added += addToMapK(o, kmer, rkmer, i, extraBase, id, o->lengthMasks[i], o->hammingDistance2, o->editDistance2);
    }
    return added;
}
/* :2327-2346 */
static jlong addToMapRightShift(struct ora *o, jlong kmer, jlong rkmer, const int id) {
    jlong added = 0;
    for (int i = o->k - 1; i >= o->mink; i--) {
        jlong extraBase = kmer & o->symbolMask;
        kmer = (jlong)((ulong64)kmer >> 2);
        rkmer = rkmer & o->rightMasks[i];
        added += addToMapK(o, kmer, rkmer, i, extraBase, id, o->lengthMasks[i], o->hammingDistance2, o->editDistance2);
    }
    return added;
}

/* :2210-2288 */
static jlong addToMapRead(struct ora *o, const uint8_t *bases, int64_t blen, int id, int skip) {
    skip = imax(o->minSkip, imin(o->maxSkip, skip));
    jlong kmer = 0, rkmer = 0, added = 0;
    int len = 0;
    const int k = o->k;
    if (bases == NULL || blen < k) return 0;
    for (int64_t i = 0; i < blen; i++) {
        const uint8_t b = bases[i];
        ORA_ASSERT(b < 128);
        const jlong x = baseToNumber0[b];
        const jlong x2 = baseToComplementNumber0[b];
        kmer = ((jlong)((ulong64)kmer << 2) | x) & o->mask;
        rkmer = ((jlong)((ulong64)rkmer >> 2) | (x2 << o->shift2)) & o->mask;
        if (isFullyDefined(b)) {
            len++;
        } else {
            len = 0;
            rkmer = 0;
        }
        if (len >= k) {
            o->refKmers++;
            if (skip > 1 && (len % skip != 0)) continue;
            const jlong extraBase = (i >= blen - 1 ? -1 : baseToNumber[bases[i + 1]]);
            added += addToMapK(o, kmer, rkmer, k, extraBase, id, o->kmask, o->hammingDistance, o->editDistance);
            if (o->useShortKmers) {
                if (i == o->k2) added += addToMapRightShift(o, kmer, rkmer, id);
                if (i == blen - 1) added += addToMapLeftShift(o, kmer, rkmer, extraBase, id);
            }
        }
    }
    return added;
}

/* scaffold numbering + per-scaffold skip rule: jgi/BBDuk.java:1849-1863, :2187-2193 */
int ora_add_ref(struct ora *o, const uint8_t *bases, const int64_t *offsets, int32_t n_seqs) {
    if (o->finalized) return 1;
    for (int32_t s = 0; s < n_seqs; s++) {
        const int id = ++o->nScaffolds;
        const int64_t rblen = offsets[s + 1] - offsets[s];
        o->storedKmers +=
            addToMapRead(o, bases + offsets[s], rblen, id,
                         rblen > 20000000 ? o->k : rblen > 5000000 ? 11 : rblen > 500000 ? 2 : 0);
    }
    return 0;
}

int64_t ora_finalize(struct ora *o) {
    if (!o->finalized) {
        o->scaffoldReadCounts = (int64_t *)calloc((size_t)o->nScaffolds + 1, sizeof(int64_t));
        o->scaffoldBaseCounts = (int64_t *)calloc((size_t)o->nScaffolds + 1, sizeof(int64_t));
        o->finalized = 1;
    }
    return o->storedKmers;
}

int32_t ora_n_scaffolds(struct ora *o) { return o->nScaffolds; }

/* dump the table (any order) for table-parity tests */
int64_t ora_dump_table(struct ora *o, uint64_t *keys, int32_t *vals, int64_t cap) {
    int64_t n = 0;
    for (int64_t i = 0; i < o->map.cap; i++) {
        if (o->map.keys[i] == KMAP_EMPTY) continue;
        if (n < cap) {
            keys[n] = o->map.keys[i];
            vals[n] = o->map.vals[i];
        }
        n++;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* per-thread scratch                                                                           */
typedef struct {
    struct ora *o;
    int64_t *scaffoldReadCountsT, *scaffoldBaseCountsT;
    int *countArray; /* findBestMatch */
    int *idList, idListSize, *countList;
    uint64_t *bs;    /* kmask BitSet words */
    int64_t bsWords;
} octx;

/* a read as the scan functions see it: current bases = orig[lo,hi) */
typedef struct {
    const uint8_t *orig;
    int origLen;
    int lo, hi;
    int pairnum;
    int discarded;
    int credit0, credit1; /* scaffold ids credited, in order */
    /* kmask */
    uint32_t *maskOut; /* may be NULL */
    /* ksplit */
    int splitStart; /* start (orig coords) of the new mate, or -1 */
} oread;

static inline int rlength(const oread *r) { return r->hi - r->lo; }
static inline const uint8_t *rbases(const oread *r) { return r->orig + r->lo; }

static void credit(octx *c, oread *r, int id, int blen) {
    c->scaffoldReadCountsT[id]++;
    c->scaffoldBaseCountsT[id] += blen;
    if (r->credit0 < 0)
        r->credit0 = id;
    else
        r->credit1 = id;
}

/* shared/TrimRead.java:299-346 (no match string / samline) */
static int trimByAmount(oread *r, int leftTrimAmount, int rightTrimAmount, int minResultingLength) {
    leftTrimAmount = imax(leftTrimAmount, 0);
    rightTrimAmount = imax(rightTrimAmount, 0);
    const int len = rlength(r);
    if (len < 1) return 0;
    minResultingLength = imin(len, imax(minResultingLength, 0));
    if (leftTrimAmount + rightTrimAmount + minResultingLength > len) {
        rightTrimAmount = imax(1, len - minResultingLength);
        leftTrimAmount = 0;
    }
    const int total = leftTrimAmount + rightTrimAmount;
    if (total > 0) {
        r->lo += leftTrimAmount;
        r->hi -= rightTrimAmount;
    }
    return total;
}
/* shared/TrimRead.java:273-276 */
static int trimToPosition(oread *r, int leftLoc, int rightLoc, int minResultingLength) {
    const int len = rlength(r);
    return trimByAmount(r, leftLoc, len - rightLoc - 1, minResultingLength);
}

/* jgi/BBDuk.java:3365-3386 */
static inline int getValueInner(const struct ora *o, const jlong kmer, const jlong rkmer, const jlong lengthMask,
                                const int len, const int qPos) {
    (void)len;
    ORA_ASSERT(lengthMask == 0 || (kmer < lengthMask && rkmer < lengthMask));
    if (o->qSkip > 1 && (qPos % o->qSkip != 0)) return -1;
    const jlong max = (o->rcomp ? lmax(kmer, rkmer) : kmer);
    const jlong key = (max & o->middleMask) | lengthMask;
    if (passesSpeed(o, key)) return kmap_get(&o->map, (ulong64)key);
    return -1;
}
/* jgi/BBDuk.java:3335-3354 */
static int getValue(const struct ora *o, const jlong kmer, const jlong rkmer, const jlong lengthMask, const int qPos,
                    const int len, const int qHDist) {
    int id = getValueInner(o, kmer, rkmer, lengthMask, len, qPos);
    if (id < 1 && qHDist > 0) {
        const int qHDist2 = qHDist - 1;
        for (int j = 0; j < o->symbols && id < 1; j++) {
            for (int i = 0; i < len && id < 1; i++) {
                const jlong temp = (kmer & o->clearMasks[i]) | o->setMasks[j][i];
                if (temp != kmer) {
                    jlong rtemp = rcomp_(temp, len);
                    id = getValue(o, temp, rtemp, lengthMask, qPos, len, qHDist2);
                }
            }
        }
    }
    return id;
}

#define ROLL(b)                                                                                        \
    do {                                                                                               \
        ORA_ASSERT((b) < 128);                                                                         \
        const jlong x = baseToNumber0[(b)];                                                            \
        const jlong x2 = baseToComplementNumber0[(b)];                                                 \
        kmer = ((jlong)((ulong64)kmer << 2) | x) & o->mask;                                                            \
        rkmer = ((jlong)((ulong64)rkmer >> 2) | (x2 << o->shift2)) & o->mask;                          \
        if (o->forbidNs && !isFullyDefined(b)) {                                                       \
            len = 0;                                                                                   \
            rkmer = 0;                                                                                 \
        } else {                                                                                       \
            len++;                                                                                     \
        }                                                                                              \
    } while (0)

/* jgi/BBDuk.java:3395-3457 (hitCounts==null: the deprecated duk= histogram is not modelled) */
static int countSetKmers(octx *c, oread *r, const int maxBadKmers) {
    struct ora *o = c->o;
    if (r == NULL || rlength(r) < o->k || o->storedKmers < 1) return 0;
    if ((o->skipR1 && r->pairnum == 0) || (o->skipR2 && r->pairnum == 1)) return 0;
    const uint8_t *bases = rbases(r);
    const int blen = rlength(r);
    jlong kmer = 0, rkmer = 0;
    int found = 0, len = 0;
    const int start = (o->restrictRight < 1 ? 0 : imax(0, blen - o->restrictRight));
    const int stop = (o->restrictLeft < 1 ? blen : imin(blen, o->restrictLeft));
    for (int i = start; i < stop; i++) {
        uint8_t b = bases[i];
        ROLL(b);
        if (len >= o->minlen2 && i >= o->minlen) {
            const int id = getValue(o, kmer, rkmer, o->kmask, i, o->k, o->qHammingDistance);
            if (id > 0) {
                if (found == maxBadKmers) {
                    credit(c, r, id, blen);
                    return (found = found + 1);
                }
                found++;
            }
        }
    }
    return found;
}

/* jgi/BBDuk.java:3466-3519 */
static int countCoveredBases(octx *c, oread *r, const int minCoveredBases) {
    struct ora *o = c->o;
    if (r == NULL || rlength(r) < o->k || o->storedKmers < 1) return 0;
    if ((o->skipR1 && r->pairnum == 0) || (o->skipR2 && r->pairnum == 1)) return 0;
    const uint8_t *bases = rbases(r);
    const int blen = rlength(r);
    jlong kmer = 0, rkmer = 0;
    int found = 0, len = 0, lastFound = -1;
    const int start = (o->restrictRight < 1 ? 0 : imax(0, blen - o->restrictRight));
    const int stop = (o->restrictLeft < 1 ? blen : imin(blen, o->restrictLeft));
    for (int i = start; i < stop; i++) {
        uint8_t b = bases[i];
        ROLL(b);
        if (len >= o->minlen2 && i >= o->minlen) {
            const int id = getValue(o, kmer, rkmer, o->kmask, i, o->k, o->qHammingDistance);
            if (id > 0) {
                int extra = imin(o->k, i - lastFound);
                found += extra;
                lastFound = i;
                if (found >= minCoveredBases) {
                    credit(c, r, id, blen); /* 'recorded' is never set to true in the reference */
                    return found;
                }
            }
        }
    }
    return found;
}

/* jgi/BBDuk.java:3527-3589. The reference leaves countArray dirty when found<=maxBadKmers
 * (condenseLoose is skipped); that cross-read leak is thread-schedule dependent and is NOT
 * reproduced: counts are per read here. Identical whenever maxBadKmers==0 (the default). */
static int findBestMatch(octx *c, oread *r, const int maxBadKmers) {
    struct ora *o = c->o;
    c->idListSize = 0;
    if (r == NULL || rlength(r) < o->k || o->storedKmers < 1) return -1;
    if ((o->skipR1 && r->pairnum == 0) || (o->skipR2 && r->pairnum == 1)) return -1;
    const uint8_t *bases = rbases(r);
    const int blen = rlength(r);
    jlong kmer = 0, rkmer = 0;
    int found = 0, len = 0;
    const int start = (o->restrictRight < 1 ? 0 : imax(0, blen - o->restrictRight));
    const int stop = (o->restrictLeft < 1 ? blen : imin(blen, o->restrictLeft));
    for (int i = start; i < stop; i++) {
        uint8_t b = bases[i];
        ROLL(b);
        if (len >= o->minlen2 && i >= o->minlen) {
            const int id = getValue(o, kmer, rkmer, o->kmask, i, o->k, o->qHammingDistance);
            if (id > 0) {
                c->countArray[id]++;
                if (c->countArray[id] == 1) c->idList[c->idListSize++] = id;
                found++;
            }
        }
    }
    int id;
    /* condenseLoose :4406-4419 */
    int max = 0;
    for (int i = 0; i < c->idListSize; i++) {
        const int p = c->idList[i];
        const int cc = c->countArray[p];
        c->countList[i] = cc;
        c->countArray[p] = 0;
        max = imax(max, cc);
    }
    if (found > maxBadKmers) {
        int id0 = -1;
        for (int i = 0; i < c->idListSize; i++) {
            if (c->countList[i] == max) {
                id0 = c->idList[i];
                break;
            }
        }
        id = id0;
    } else {
        id = -1;
    }
    if (found > maxBadKmers) credit(c, r, id, blen);
    return id;
}

/* jgi/BBDuk.java:3596-3677 */
static int countSetKmersBig(octx *c, oread *r, const int maxBadKmers) {
    struct ora *o = c->o;
    if (r == NULL || rlength(r) < o->kbig || o->storedKmers < 1) return 0;
    if ((o->skipR1 && r->pairnum == 0) || (o->skipR2 && r->pairnum == 1)) return 0;
    ORA_ASSERT(o->kbig > o->k);
    const int sub = o->kbig - o->k - 1;
    ORA_ASSERT(sub >= 0);
    const uint8_t *bases = rbases(r);
    const int blen = rlength(r);
    jlong kmer = 0, rkmer = 0;
    int found = 0, len = 0;
    int bkStart = -1, bkStop = -1;
    int id = -1, lastId = -1;
    const int start = (o->restrictRight < 1 ? 0 : imax(0, blen - o->restrictRight));
    const int stop = (o->restrictLeft < 1 ? blen : imin(blen, o->restrictLeft));
    for (int i = start; i < stop; i++) {
        uint8_t b = bases[i];
        ROLL(b);
        if (len >= o->minlen2 && i >= o->minlen) {
            id = getValue(o, kmer, rkmer, o->kmask, i, o->k, o->qHammingDistance);
            if (id > 0) {
                lastId = id;
                if (bkStart == -1) bkStart = i;
                bkStop = i;
            } else {
                if (bkStart > -1) {
                    int dif = bkStop - bkStart - sub;
                    bkStop = bkStart = -1;
                    if (dif > 0) {
                        int old = found;
                        found += dif;
                        if (found > maxBadKmers && old <= maxBadKmers) {
                            credit(c, r, lastId, blen);
                            return found;
                        }
                    }
                }
            }
        }
    }
    if (bkStart > -1) {
        int dif = bkStop - bkStart - sub;
        bkStop = bkStart = -1;
        if (dif > 0) {
            int old = found;
            found += dif;
            if (found > maxBadKmers && old <= maxBadKmers) credit(c, r, lastId, blen);
        }
    }
    return found;
}

/* jgi/BBDuk.java:3866-4013 (ktrim) and :3708-3858 (ktrimTip) are the same body; `left`/`right`
 * replace ktrimLeft/ktrimRight in the tip version. */
static int ktrimBody(octx *c, oread *r, const int start, const int stop, const int right, const int left) {
    struct ora *o = c->o;
    const int k = o->k;
    if (r == NULL || rlength(r) < imax(1, (o->useShortKmers ? imin(k, o->mink) : k)) || o->storedKmers < 1) return 0;
    if ((o->skipR1 && r->pairnum == 0) || (o->skipR2 && r->pairnum == 1)) return 0;
    const uint8_t *bases = rbases(r);
    const int blen = rlength(r);
    jlong kmer = 0, rkmer = 0;
    int found = 0, len = 0, id0 = -1;
    int minLoc = 999999999, minLocExclusive = 999999999;
    int maxLoc = -1, maxLocExclusive = -1;
    for (int i = start; i < stop; i++) {
        uint8_t b = bases[i];
        ROLL(b);
        if (len >= o->minlen2 && i >= o->minlen) {
            const int id = getValue(o, kmer, rkmer, o->kmask, i, k, o->qHammingDistance);
            if (id > 0) {
                if (id0 < 0) id0 = id;
                minLoc = imin(minLoc, i - k + 1);
                ORA_ASSERT(minLoc >= 0);
                maxLoc = i;
                found++;
            }
        }
    }
    if (minLoc != minLocExclusive) minLocExclusive = minLoc + k;
    if (maxLoc != maxLocExclusive) maxLocExclusive = maxLoc - k;
    if (o->useShortKmers && found == 0) {
        ORA_ASSERT(!o->maskMiddle && o->middleMask == -1);
        if (left) {
            kmer = 0;
            rkmer = 0;
            len = 0;
            const int lim = imin(k, stop);
            for (int i = start; i < lim; i++) {
                uint8_t b = bases[i];
                ORA_ASSERT(b < 128);
                jlong x = baseToNumber0[b];
                jlong x2 = baseToComplementNumber0[b];
                kmer = ((jlong)((ulong64)kmer << 2) | x) & o->mask;
                rkmer = rkmer | (x2 << (2 * len));
                len++;
                if (len >= o->mink) {
                    const int id = getValue(o, kmer, rkmer, o->lengthMasks[len], i, len, o->qHammingDistance2);
                    if (id > 0) {
                        if (id0 < 0) id0 = id;
                        minLoc = 0;
                        minLocExclusive = imin(minLocExclusive, i + 1);
                        maxLoc = imax(maxLoc, i);
                        maxLocExclusive = imax(maxLocExclusive, 0);
                        found++;
                    }
                }
            }
        }
        if (right) {
            kmer = 0;
            rkmer = 0;
            len = 0;
            const int lim = imax(-1, stop - k);
            for (int i = stop - 1; i > lim; i--) {
                uint8_t b = bases[i];
                ORA_ASSERT(b < 128);
                jlong x = baseToNumber0[b];
                jlong x2 = baseToComplementNumber0[b];
                kmer = kmer | (x << (2 * len));
                rkmer = ((jlong)((ulong64)rkmer << 2) | x2) & o->mask;
                len++;
                if (len >= o->mink) {
                    const int id = getValue(o, kmer, rkmer, o->lengthMasks[len], i, len, o->qHammingDistance2);
                    if (id > 0) {
                        if (id0 < 0) id0 = id;
                        minLoc = i;
                        minLocExclusive = imin(minLocExclusive, blen);
                        maxLoc = blen - 1;
                        maxLocExclusive = imax(maxLocExclusive, i - 1);
                        found++;
                    }
                }
            }
        }
    }
    if (found == 0) return 0;
    credit(c, r, id0, blen);
    if (o->trimPad != 0) {
        maxLoc = mid3(0, maxLoc + o->trimPad, blen);
        minLoc = mid3(0, minLoc - o->trimPad, blen);
        maxLocExclusive = mid3(0, maxLocExclusive + o->trimPad, blen);
        minLocExclusive = mid3(0, minLocExclusive - o->trimPad, blen);
    }
    if (left) {
        return trimToPosition(r, o->ktrimExclusive ? maxLocExclusive + 1 : maxLoc + 1, blen - 1, 1);
    } else {
        ORA_ASSERT(right);
        return trimToPosition(r, 0, o->ktrimExclusive ? minLocExclusive - 1 : minLoc - 1, 1);
    }
}

/* jgi/BBDuk.java:3679-3684 */
static int ktrim(octx *c, oread *r) {
    struct ora *o = c->o;
    const int len = rlength(r);
    const int start = (o->restrictRight < 1 ? 0 : imax(0, len - o->restrictRight));
    const int stop = (o->restrictLeft < 1 ? len : imin(len, o->restrictLeft));
    /* in ktrim() the tails are selected by the global flags; exactly one of them is set here */
    return ktrimBody(c, r, start, stop, o->ktrimRight, o->ktrimLeft);
}

/* jgi/BBDuk.java:3686-3699 */
static int ktrimTips(octx *c, oread *r) {
    struct ora *o = c->o;
    const int len = rlength(r);
    const int mid = len / 2 - (o->k - 1) / 2;
    int sum = 0;
    if (o->ktrimRight) {
        int start = imax(0, (o->restrictRight < 1 ? mid : len - o->restrictRight));
        sum += ktrimBody(c, r, start, len, 1, 0);
    }
    if (o->ktrimLeft) {
        int stop = imin(rlength(r), (o->restrictLeft < 1 ? mid + o->k - 1 : o->restrictLeft));
        sum += ktrimBody(c, r, 0, stop, 0, 1);
    }
    return sum;
}

/* java.util.BitSet subset */
static void bs_reserve(octx *c, int64_t bits) {
    int64_t w = (bits + 63) / 64 + 1;
    if (w > c->bsWords) {
        c->bs = (uint64_t *)realloc(c->bs, sizeof(uint64_t) * w);
        c->bsWords = w;
    }
}
static void bs_range(octx *c, int from, int to, int val) {
    ORA_ASSERT(from >= 0 && from <= to);
    for (int i = from; i < to; i++) {
        if (val)
            c->bs[i >> 6] |= (1ULL << (i & 63));
        else
            c->bs[i >> 6] &= ~(1ULL << (i & 63));
    }
}

/* jgi/BBDuk.java:4022-4199 */
static int kmask(octx *c, oread *r) {
    struct ora *o = c->o;
    const int k = o->k;
    if (r == NULL || rlength(r) < imax(1, (o->useShortKmers ? imin(k, o->mink) : k)) || o->storedKmers < 1) return 0;
    if ((o->skipR1 && r->pairnum == 0) || (o->skipR2 && r->pairnum == 1)) return 0;
    const uint8_t *bases = rbases(r);
    const int blen = rlength(r);
    if (blen < k) return 0;
    jlong kmer = 0, rkmer = 0;
    int found = 0, len = 0, id0 = -1;
    const int64_t nbits = (int64_t)blen + imax(o->trimPad, 0) + 2 + k;
    bs_reserve(c, nbits);
    memset(c->bs, 0, sizeof(uint64_t) * c->bsWords);
    if (o->kmaskFullyCovered) bs_range(c, 0, blen, 1);
    const int minus = k - 1 - o->trimPad;
    const int plus = o->trimPad + 1;
    const int start = (o->restrictRight < 1 ? 0 : imax(0, blen - o->restrictRight));
    const int stop = (o->restrictLeft < 1 ? blen : imin(blen, o->restrictLeft));
    for (int i = start; i < stop; i++) {
        uint8_t b = bases[i];
        ROLL(b);
        if (i >= o->minlen) {
            int id;
            if (len >= o->minlen2) {
                id = getValue(o, kmer, rkmer, o->kmask, i, k, o->qHammingDistance);
            } else {
                id = -1;
            }
            if (id > 0) {
                if (id0 < 0) id0 = id;
                if (!o->kmaskFullyCovered) bs_range(c, imax(0, i - minus), i + plus, 1);
                found++;
            } else if (o->kmaskFullyCovered) {
                bs_range(c, imax(0, i - minus), i + plus, 0);
            }
        }
    }
    if (o->useShortKmers) {
        ORA_ASSERT(!o->maskMiddle && o->middleMask == -1);
        {
            kmer = 0;
            rkmer = 0;
            len = 0;
            int len2 = 0;
            const int lim = imin(k, stop);
            for (int i = start; i < lim; i++) {
                uint8_t b = bases[i];
                ORA_ASSERT(b < 128);
                jlong x = baseToNumber0[b];
                jlong x2 = baseToComplementNumber0[b];
                kmer = ((jlong)((ulong64)kmer << 2) | x) & o->mask;
                rkmer = rkmer | (x2 << (2 * len));
                len++;
                len2++;
                if (len2 >= o->minminlen) {
                    int id;
                    if (len >= o->mink) {
                        id = getValue(o, kmer, rkmer, o->lengthMasks[len], i, len, o->qHammingDistance2);
                    } else {
                        id = -1;
                    }
                    if (id > 0) {
                        if (id0 < 0) id0 = id;
                        if (!o->kmaskFullyCovered) bs_range(c, 0, imin(blen, i + o->trimPad + 1), 1);
                        found++;
                    } else if (o->kmaskFullyCovered) {
                        bs_range(c, 0, imin(blen, i + o->trimPad + 1), 0);
                    }
                }
            }
        }
        {
            kmer = 0;
            rkmer = 0;
            len = 0;
            int len2 = 0;
            const int lim = imax(-1, stop - k);
            for (int i = stop - 1; i > lim; i--) {
                uint8_t b = bases[i];
                ORA_ASSERT(b < 128);
                jlong x = baseToNumber0[b];
                jlong x2 = baseToComplementNumber0[b];
                kmer = kmer | (x << (2 * len));
                rkmer = ((jlong)((ulong64)rkmer << 2) | x2) & o->mask;
                len++;
                len2++;
                if (len2 >= o->minminlen) {
                    int id;
                    if (len >= o->mink) {
                        id = getValue(o, kmer, rkmer, o->lengthMasks[len], i, len, o->qHammingDistance2);
                    } else {
                        id = -1;
                    }
                    if (id > 0) {
                        if (id0 < 0) id0 = id;
                        if (!o->kmaskFullyCovered) bs_range(c, imax(0, i - o->trimPad), blen, 1);
                        found++;
                    } else if (o->kmaskFullyCovered) {
                        bs_range(c, imax(0, i - o->trimPad), blen, 0);
                    }
                }
            }
        }
    }
    if (found == 0) return 0;
    credit(c, r, id0, blen);
    int cardinality = 0;
    for (int64_t w = 0; w < c->bsWords; w++) cardinality += __builtin_popcountll(c->bs[w]);
    if (r->maskOut) {
        for (int i = 0; i < blen; i++) {
            if ((c->bs[i >> 6] >> (i & 63)) & 1) {
                const int oi = r->lo + i;
                r->maskOut[oi >> 5] |= (1u << (oi & 31));
            }
        }
    }
    return cardinality;
}

/* jgi/BBDuk.java:4208-4377; returns 1 if the read was split into two */
static int ksplit(octx *c, oread *r) {
    struct ora *o = c->o;
    const int k = o->k;
    if (r == NULL || rlength(r) < imax(1, (o->useShortKmers ? imin(k, o->mink) : k)) || o->storedKmers < 1) return 0;
    const uint8_t *bases = rbases(r);
    const int blen = rlength(r);
    if (blen < k) return 0;
    jlong kmer = 0, rkmer = 0;
    jlong found = 0;
    int len = 0, id0 = -1;
    int leftmost = INT32_MAX, rightmost = -1;
    const int minus = k - 1 - o->trimPad;
    const int plus = o->trimPad;
    const int start = (o->restrictRight < 1 ? 0 : imax(0, blen - o->restrictRight));
    const int stop = (o->restrictLeft < 1 ? blen : imin(blen, o->restrictLeft));
    for (int i = start; i < stop; i++) {
        uint8_t b = bases[i];
        ROLL(b);
        if (i >= o->minlen) {
            int id;
            if (len >= o->minlen2) {
                id = getValue(o, kmer, rkmer, o->kmask, i, k, o->qHammingDistance);
            } else {
                id = -1;
            }
            if (id > 0) {
                if (id0 < 0) id0 = id;
                leftmost = imin(leftmost, imax(0, i - minus));
                rightmost = imax(rightmost, i + plus);
                found++;
            }
        }
    }
    if (o->useShortKmers && id0 == -1) {
        ORA_ASSERT(!o->maskMiddle && o->middleMask == -1);
        {
            kmer = 0;
            rkmer = 0;
            len = 0;
            int len2 = 0;
            const int lim = imax(-1, stop - k);
            for (int i = stop - 1; i > lim; i--) {
                uint8_t b = bases[i];
                ORA_ASSERT(b < 128);
                jlong x = baseToNumber0[b];
                jlong x2 = baseToComplementNumber0[b];
                kmer = kmer | (x << (2 * len));
                rkmer = ((jlong)((ulong64)rkmer << 2) | x2) & o->mask;
                len++;
                len2++;
                if (len2 >= o->minminlen) {
                    int id;
                    if (len >= o->mink) {
                        id = getValue(o, kmer, rkmer, o->lengthMasks[len], i, len, o->qHammingDistance2);
                    } else {
                        id = -1;
                    }
                    if (id > 0) {
                        if (id0 < 0) id0 = id;
                        leftmost = imin(leftmost, imax(0, i - o->trimPad));
                        rightmost = blen - 1;
                        found++;
                    }
                }
            }
        }
        if (id0 == -1) {
            kmer = 0;
            rkmer = 0;
            len = 0;
            int len2 = 0;
            const int lim = imin(k, stop);
            for (int i = start; i < lim; i++) {
                uint8_t b = bases[i];
                ORA_ASSERT(b < 128);
                jlong x = baseToNumber0[b];
                jlong x2 = baseToComplementNumber0[b];
                kmer = ((jlong)((ulong64)kmer << 2) | x) & o->mask;
                rkmer = rkmer | (x2 << (2 * len));
                len++;
                len2++;
                if (len2 >= o->minminlen) {
                    int id;
                    if (len >= o->mink) {
                        id = getValue(o, kmer, rkmer, o->lengthMasks[len], i, len, o->qHammingDistance2);
                    } else {
                        id = -1;
                    }
                    if (id > 0) {
                        if (id0 < 0) id0 = id;
                        leftmost = 0;
                        rightmost = imax(rightmost, i + o->trimPad);
                        found++;
                    }
                }
            }
        }
    }
    if (found == 0) return 0;
    credit(c, r, id0, blen);
    if (leftmost == 0) {
        trimToPosition(r, rightmost + 1, blen - 1, 1);
        return 0;
    } else if (rightmost == blen - 1) {
        trimToPosition(r, 0, leftmost - 1, 1);
        return 0;
    } else {
        /* Read r2=r.subRead(rightmost+1, bases.length-1): [rightmost+1, blen-1) of the current read */
        ORA_ASSERT(rightmost + 1 <= blen - 1); /* copyOfRange would throw otherwise (only with trimpad>0) */
        r->splitStart = r->lo + rightmost + 1;
        trimToPosition(r, 0, leftmost - 1, 1);
        return 1;
    }
}

/* stream/Read.java:1673-1683 */
static int numValidKmers(const oread *r, int k) {
    if (r == NULL) return 0;
    const uint8_t *bases = rbases(r);
    int len = 0, counted = 0;
    for (int i = 0; i < rlength(r); i++) {
        int x = bases[i] < 128 ? baseToNumber[bases[i]] : -1;
        if (x < 0) {
            len = 0;
        } else {
            len++;
        }
        if (len >= k) counted++;
    }
    return counted;
}

/* jgi/BBDuk.java:3260-3289 */
static void setDiscarded(const struct ora *o, oread *r) {
    if (o->trimFailuresTo1bp) {
        if (rlength(r) > 1) trimByAmount(r, 0, rlength(r) - 1, 1);
    } else {
        r->discarded = 1;
    }
}
static int isDiscarded(const struct ora *o, const oread *r) {
    if (r == NULL) return 0;
    if (r->discarded) return 1;
    return o->trimFailuresTo1bp && rlength(r) == 1;
}
static int isNullOrDiscarded(const struct ora *o, const oread *r) {
    if (r == NULL) return 1;
    if (r->discarded) return 1;
    return o->trimFailuresTo1bp && rlength(r) == 1;
}
static int isNotDiscarded(const struct ora *o, const oread *r) {
    if (r == NULL) return 0;
    if (r->discarded) return 0;
    return !(o->trimFailuresTo1bp && rlength(r) == 1);
}
static int shouldRemove(const struct ora *o, const oread *r1, const oread *r2) {
    return (o->removePairsIfEitherBad && (isDiscarded(o, r1) || isDiscarded(o, r2))) ||
           (isDiscarded(o, r1) && isNullOrDiscarded(o, r2));
}

typedef struct {
    int remove, ktrimmed1, ktrimmed2, tpe1, tpe2, split;
    int count1, count2;
    int64_t readsKTrimmed, basesKTrimmed, readsKFiltered, basesKFiltered;
} pair_result;

/* jgi/BBDuk.java:2587-2593 and :2727-2873 */
static void processPair(octx *c, oread *r1, oread *r2, pair_result *pr) {
    struct ora *o = c->o;
    memset(pr, 0, sizeof *pr);
    const int initialLength1 = rlength(r1);
    const int initialLength2 = (r2 ? rlength(r2) : 0);
    const int pairCount = (r2 ? 2 : 1);
    const int minlen1 = (int)fmaxf((float)initialLength1 * o->minLenFraction, (float)o->minReadLength);
    const int minlen2 = (int)fmaxf((float)initialLength2 * o->minLenFraction, (float)o->minReadLength);
    int remove = 0;
    const int ktrimLeftOrRight = o->ktrimLeft || o->ktrimRight;
    const int ktrimTipsMode = (o->ktrimLeft && o->ktrimRight);
    const int doKmerTrimming = o->storedKmers > 0 && (o->ktrimLeft || o->ktrimRight || o->ktrimN || o->ksplit);
    const int doKmerFiltering = o->storedKmers > 0 && !doKmerTrimming;

    if (doKmerTrimming) {
        int rlen1 = 0, rlen2 = 0, xsum = 0, rktsum = 0;
        if (ktrimTipsMode) {
            if (r1 != NULL) {
                int x = ktrimTips(c, r1);
                xsum += x;
                rktsum += (x > 0 ? 1 : 0);
                pr->count1 = x;
                rlen1 = rlength(r1);
                if (rlen1 < minlen1) setDiscarded(o, r1);
            }
            if (r2 != NULL) {
                int x = ktrimTips(c, r2);
                xsum += x;
                rktsum += (x > 0 ? 1 : 0);
                pr->count2 = x;
                rlen2 = rlength(r2);
                if (rlen2 < minlen2) setDiscarded(o, r2);
            }
        } else if (ktrimLeftOrRight) {
            if (r1 != NULL) {
                int x = ktrim(c, r1);
                xsum += x;
                rktsum += (x > 0 ? 1 : 0);
                pr->count1 = x;
                rlen1 = rlength(r1);
                if (rlen1 < minlen1) setDiscarded(o, r1);
            }
            if (r2 != NULL) {
                int x = ktrim(c, r2);
                xsum += x;
                rktsum += (x > 0 ? 1 : 0);
                pr->count2 = x;
                rlen2 = rlength(r2);
                if (rlen2 < minlen2) setDiscarded(o, r2);
            }
        } else if (o->ktrimN) {
            if (r1 != NULL) {
                int x = kmask(c, r1);
                xsum += x;
                rktsum += (x > 0 ? 1 : 0);
                pr->count1 = x;
                rlen1 = rlength(r1);
                if (rlen1 < minlen1) setDiscarded(o, r1);
            }
            if (r2 != NULL) {
                int x = kmask(c, r2);
                xsum += x;
                rktsum += (x > 0 ? 1 : 0);
                pr->count2 = x;
                rlen2 = rlength(r2);
                if (rlen2 < minlen2) setDiscarded(o, r2);
            }
        } else if (o->ksplit) {
            ORA_ASSERT(r2 == NULL);
            if (r1 != NULL) {
                int oldLen = rlength(r1);
                const int oldHi = r1->hi;
                int b = ksplit(c, r1);
                /* the new mate is subRead(rightmost+1, len-1), END-EXCLUSIVE (stream/Read.java:3729-3731):
                 * it loses the last base */
                int newPairLen = rlength(r1) + (b ? (oldHi - 1 - r1->splitStart) : 0);
                int trimmed = oldLen - newPairLen;
                xsum += trimmed;
                rktsum += (trimmed > 0 ? 1 : 0);
                pr->split = b;
                rlen1 = rlength(r1);
            }
        }
        pr->ktrimmed1 = (pr->count1 > 0);
        pr->ktrimmed2 = (pr->count2 > 0);
        if (o->ksplit) {
            remove = pr->split; /* remove=(r1.mate!=null) */
            pr->ktrimmed1 = (xsum > 0);
        } else if (shouldRemove(o, r1, r2)) {
            if (!o->ktrimN) {
                xsum += (rlen1 + rlen2);
                rktsum = pairCount;
            }
            remove = 1;
        } else if (o->ktrimRight && o->trimPairsEvenly && xsum > 0 && r2 != NULL && rlength(r1) != rlength(r2)) {
            int x;
            if (rlength(r1) > rlength(r2)) {
                x = trimToPosition(r1, 0, rlength(r2) - 1, 1);
                pr->tpe1 = 1;
            } else {
                x = trimToPosition(r2, 0, rlength(r1) - 1, 1);
                pr->tpe2 = 1;
            }
            if (rktsum < 2) rktsum++;
            xsum += x;
            /* the reference asserts r1.length()==r2.length() here (jgi/BBDuk.java:2810); it can only fail
             * when one mate is empty (trimByAmount never trims to 0), where the reference would die with
             * an AssertionError under its default -ea. Such input is outside the contract; carry on. */
        }
        pr->basesKTrimmed += xsum;
        pr->readsKTrimmed += rktsum;
    } else if (doKmerFiltering) {
        if (o->minCoveredFraction > 0) {
            if (isNotDiscarded(o, r1)) {
                const int minCoveredBases = (int)ceil((double)(o->minCoveredFraction * (float)rlength(r1)));
                const int covered = countCoveredBases(c, r1, minCoveredBases);
                pr->count1 = covered;
                if (covered >= minCoveredBases) setDiscarded(o, r1);
            }
            if (isNotDiscarded(o, r2)) {
                const int minCoveredBases = (int)ceil((double)(o->minCoveredFraction * (float)rlength(r2)));
                const int covered = countCoveredBases(c, r2, minCoveredBases);
                pr->count2 = covered;
                if (covered >= minCoveredBases) setDiscarded(o, r2);
            }
        } else {
            int maxBadKmersR1, maxBadKmersR2;
            if (o->minKmerFraction == 0) {
                maxBadKmersR1 = maxBadKmersR2 = o->maxBadKmers0;
            } else {
                const int vk1 = numValidKmers(r1, o->keff), vk2 = (r2 == NULL ? 0 : numValidKmers(r2, o->keff));
                maxBadKmersR1 = imax(o->maxBadKmers0, (int)((float)(vk1 - 1) * o->minKmerFraction));
                maxBadKmersR2 = imax(o->maxBadKmers0, (int)((float)(vk2 - 1) * o->minKmerFraction));
            }
            if (!o->findBestMatch) {
                const int a = (o->kbig <= o->k ? countSetKmers(c, r1, maxBadKmersR1) : countSetKmersBig(c, r1, maxBadKmersR1));
                const int b = (r2 == NULL ? 0
                                          : (o->kbig <= o->k ? countSetKmers(c, r2, maxBadKmersR2)
                                                             : countSetKmersBig(c, r2, maxBadKmersR2)));
                pr->count1 = a;
                pr->count2 = b;
                if (r1 != NULL && a > maxBadKmersR1) setDiscarded(o, r1);
                if (r2 != NULL && b > maxBadKmersR2) setDiscarded(o, r2);
            } else {
                const int a = findBestMatch(c, r1, maxBadKmersR1);
                const int b = (r2 == NULL ? -1 : findBestMatch(c, r2, maxBadKmersR2));
                pr->count1 = a;
                pr->count2 = b;
                if (r1 != NULL && a > 0) setDiscarded(o, r1);
                if (r2 != NULL && b > 0) setDiscarded(o, r2);
            }
        }
        if (shouldRemove(o, r1, r2)) {
            remove = 1;
            if (r1 != NULL) {
                pr->readsKFiltered++;
                pr->basesKFiltered += initialLength1;
            }
            if (r2 != NULL) {
                pr->readsKFiltered++;
                pr->basesKFiltered += initialLength2;
            }
        }
    }
    pr->remove = remove;
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    struct ora *o;
    const uint8_t *bases;
    const int64_t *offsets;
    int64_t unit0, unit1; /* pairs (or single reads) [unit0, unit1) */
    int paired;
    const bbduk_out *out;
    bbduk_stats st;
    int64_t *srcT, *sbcT;
} job;

static void fill_out(const bbduk_out *out, int64_t idx, const oread *r, int removed, int ktrimmed, int tpe, int split,
                     int count) {
    if (out->id0) out->id0[idx] = r->credit0;
    if (out->id0b) out->id0b[idx] = r->credit1;
    if (out->lo) out->lo[idx] = r->lo;
    if (out->hi) out->hi[idx] = r->hi;
    if (out->count) out->count[idx] = split ? r->splitStart : count;
    if (out->flags) {
        uint8_t f = 0;
        if (r->discarded) f |= BBDUK_F_DISCARDED;
        if (removed) f |= BBDUK_F_REMOVED;
        if (ktrimmed) f |= BBDUK_F_KTRIMMED;
        if (tpe) f |= BBDUK_F_TPE;
        if (split) f |= BBDUK_F_SPLIT;
        out->flags[idx] = f;
    }
}

static void *job_run(void *arg) {
    job *j = (job *)arg;
    struct ora *o = j->o;
    octx c;
    memset(&c, 0, sizeof c);
    c.o = o;
    c.scaffoldReadCountsT = j->srcT;
    c.scaffoldBaseCountsT = j->sbcT;
    c.countArray = (int *)calloc((size_t)o->nScaffolds + 2, sizeof(int));
    c.idList = (int *)calloc((size_t)o->nScaffolds + 2, sizeof(int));
    c.countList = (int *)calloc((size_t)o->nScaffolds + 2, sizeof(int));
    const int per = j->paired ? 2 : 1;
    for (int64_t u = j->unit0; u < j->unit1; u++) {
        oread rr[2];
        for (int p = 0; p < per; p++) {
            int64_t idx = u * per + p;
            oread *r = &rr[p];
            r->orig = j->bases + j->offsets[idx];
            r->origLen = (int)(j->offsets[idx + 1] - j->offsets[idx]);
            r->lo = 0;
            r->hi = r->origLen;
            r->pairnum = p;
            r->discarded = 0;
            r->credit0 = r->credit1 = -1;
            r->splitStart = -1;
            r->maskOut = NULL;
            if (j->out->maskbits && j->out->mask_off) {
                r->maskOut = j->out->maskbits + j->out->mask_off[idx];
                int64_t nw = j->out->mask_off[idx + 1] - j->out->mask_off[idx];
                memset(r->maskOut, 0, sizeof(uint32_t) * (size_t)nw);
            }
        }
        pair_result pr;
        processPair(&c, &rr[0], per == 2 ? &rr[1] : NULL, &pr);
        fill_out(j->out, u * per, &rr[0], pr.remove, pr.ktrimmed1, pr.tpe1, pr.split, pr.count1);
        if (per == 2) fill_out(j->out, u * per + 1, &rr[1], pr.remove, pr.ktrimmed2, pr.tpe2, 0, pr.count2);
        j->st.reads_in += per;
        j->st.bases_in += rr[0].origLen + (per == 2 ? rr[1].origLen : 0);
        j->st.reads_ktrimmed += pr.readsKTrimmed;
        j->st.bases_ktrimmed += pr.basesKTrimmed;
        j->st.reads_kfiltered += pr.readsKFiltered;
        j->st.bases_kfiltered += pr.basesKFiltered;
        if (!pr.remove) {
            j->st.reads_out += per;
            j->st.bases_out += rlength(&rr[0]) + (per == 2 ? rlength(&rr[1]) : 0);
        }
    }
    free(c.countArray);
    free(c.idList);
    free(c.countList);
    free(c.bs);
    return NULL;
}

int ora_process(struct ora *o, const uint8_t *bases, const int64_t *offsets, int64_t n_reads, int32_t paired,
                const bbduk_out *out, bbduk_stats *stats, int nthreads) {
    if (!o->finalized) ora_finalize(o);
    if (paired && (n_reads & 1)) return 1;
    if (nthreads < 1) nthreads = 1;
    const int64_t units = paired ? n_reads / 2 : n_reads;
    if (units < nthreads) nthreads = (int)(units > 0 ? units : 1);
    job *jobs = (job *)calloc((size_t)nthreads, sizeof(job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    const size_t ns = (size_t)o->nScaffolds + 1;
    for (int t = 0; t < nthreads; t++) {
        jobs[t].o = o;
        jobs[t].bases = bases;
        jobs[t].offsets = offsets;
        jobs[t].unit0 = units * t / nthreads;
        jobs[t].unit1 = units * (t + 1) / nthreads;
        jobs[t].paired = paired;
        jobs[t].out = out;
        jobs[t].srcT = (int64_t *)calloc(ns, sizeof(int64_t));
        jobs[t].sbcT = (int64_t *)calloc(ns, sizeof(int64_t));
        if (nthreads > 1) pthread_create(&th[t], NULL, job_run, &jobs[t]);
    }
    if (nthreads == 1) job_run(&jobs[0]);
    bbduk_stats st;
    memset(&st, 0, sizeof st);
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        st.reads_in += jobs[t].st.reads_in;
        st.bases_in += jobs[t].st.bases_in;
        st.reads_ktrimmed += jobs[t].st.reads_ktrimmed;
        st.bases_ktrimmed += jobs[t].st.bases_ktrimmed;
        st.reads_kfiltered += jobs[t].st.reads_kfiltered;
        st.bases_kfiltered += jobs[t].st.bases_kfiltered;
        st.reads_out += jobs[t].st.reads_out;
        st.bases_out += jobs[t].st.bases_out;
        for (size_t s = 0; s < ns; s++) {
            o->scaffoldReadCounts[s] += jobs[t].srcT[s];
            o->scaffoldBaseCounts[s] += jobs[t].sbcT[s];
        }
        free(jobs[t].srcT);
        free(jobs[t].sbcT);
    }
    if (stats) *stats = st;
    free(jobs);
    free(th);
    return 0;
}

void ora_scaffold_counts(struct ora *o, int64_t *read_counts, int64_t *base_counts, int32_t n) {
    for (int32_t i = 0; i < n && i <= o->nScaffolds; i++) {
        if (read_counts) read_counts[i] = o->scaffoldReadCounts ? o->scaffoldReadCounts[i] : 0;
        if (base_counts) base_counts[i] = o->scaffoldBaseCounts ? o->scaffoldBaseCounts[i] : 0;
    }
}

/* derived constants, for tests of the host-side derivation (product vs oracle) */
void ora_derived(struct ora *o, int64_t *v /* [16] */) {
    v[0] = o->k;
    v[1] = o->kbig;
    v[2] = o->mink;
    v[3] = o->useShortKmers;
    v[4] = o->maskMiddle;
    v[5] = o->midMaskLen;
    v[6] = o->minlen;
    v[7] = o->minlen2;
    v[8] = o->minminlen;
    v[9] = o->forbidNs;
    v[10] = o->hammingDistance;
    v[11] = o->hammingDistance2;
    v[12] = o->middleMask;
    v[13] = o->mask;
    v[14] = o->kfilter;
    v[15] = o->removePairsIfEitherBad;
}

/* direct probe for unit tests: id stored for a canonical key, or -1 */
int32_t ora_lookup_key(struct ora *o, uint64_t key) { return kmap_get(&o->map, key); }
