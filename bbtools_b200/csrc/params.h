// params.h -- derived BBDuk constants shared by host code and kernels (POD, passed by value to kernels).
//
// Field names follow the reference (jgi/BBDuk.java:5121-5347); derive_params() in derive.cpp computes
// them exactly as the reference's constructor does (jgi/BBDuk.java:583-585, :672-877).
#pragma once
#include <stdint.h>

enum BBMode : int32_t {
    MODE_KTRIM = 0,    // ktrim=r or ktrim=l          (jgi/BBDuk.java:3866-4013)
    MODE_KTRIM_TIPS,   // ktrim=rl                    (jgi/BBDuk.java:3686-3858)
    MODE_KMASK,        // ktrim=n / kmask             (jgi/BBDuk.java:4022-4199)
    MODE_KSPLIT,       // ksplit                      (jgi/BBDuk.java:4208-4377)
    MODE_KFILTER,      // countSetKmers               (jgi/BBDuk.java:3395-3457)
    MODE_KFILTER_BIG,  // countSetKmersBig, k>31      (jgi/BBDuk.java:3596-3677)
    MODE_KCOVER,       // countCoveredBases, mcf>0    (jgi/BBDuk.java:3466-3519)
    MODE_KBEST         // findBestMatch               (jgi/BBDuk.java:3527-3589)
};

struct BBParams {
    // k-mer geometry
    int32_t k, k2, kbig, keff, mink;
    int32_t minlen, minlen2, minminlen, shift2;
    int32_t midMaskLen, maskMiddle, useShortKmers;
    uint64_t mask, kmask, middleMask;
    // distances
    int32_t hammingDistance, hammingDistance2, editDistance, editDistance2;
    int32_t qHammingDistance, qHammingDistance2;
    int32_t minSkip, maxSkip;
    // query behaviour
    int32_t forbidNs, rcomp, speed, qSkip;
    int32_t speedMask2;  // speed= follows bbduk.BBDukS's default index (hash bits 16-19, bbduk/BBDukIndexMask2.java:566-577) instead of jgi.BBDuk's key%17
    int32_t restrictLeft, restrictRight, skipR1, skipR2;
    int32_t mode;  // BBMode
    int32_t ktrimLeft, ktrimRight, ktrimExclusive, trimPad;
    int32_t kmaskFullyCovered;
    int32_t maxBadKmers0;
    float minKmerFraction, minCoveredFraction;
    // pair logic
    int32_t minReadLength;
    float minLenFraction;
    int32_t removePairsIfEitherBad, trimPairsEvenly, trimFailuresTo1bp;
};

// Device hash-array geometry (layout only; results do not depend on it -- SURVEY.md section 0.2).
struct BBTable {
    const uint64_t *keys;   // [n_slots]; EMPTY = ~0
    const int32_t *vals;    // [n_slots]; scaffold id (min over writers)
    uint64_t slot_mask;     // n_slots-1 (n_slots power of two, >= 1024)
    uint32_t bucket_shift;  // 34 - log2(n_slots): bucket = hash32 >> bucket_shift
    // filter images, one device buffer: [canonical bloom of all keys | part filter | short-key bloom | 8-mer byte map |
    // tail bitmaps | L2 filter]
    const uint32_t *filter;
    uint32_t n_filter_words;  // canonical bloom words
    uint32_t part_words;      // part filter words (0 = not available for this configuration)
    uint32_t short_words;     // bloom over the short (len<k) keys only
    uint32_t samp_words;      // byte map over all 8-mers that occur inside a reference part (sampled scan of probe_fast2.cu; 0 = none)
    uint32_t tail_words;      // two direct bitmaps over tail_q-mers in front of the short-k-mer tails (0 = none), followed (tail_q >= 9)
                              // by their two 8-mer level-0 images of BB_TAIL0_WORDS words each
    int32_t tail_q;           // bases the tail bitmaps are indexed by = min(mink, 12)
    uint32_t big_words;       // L2-resident one-bit-per-key filter in front of an HBM-resident array (0 = none)
    int32_t n_parts;          // pigeonhole parts = hdist+1
    int32_t part_w;           // bases per part (<=16)
    int32_t part_lag[4];      // part j is the part_w-mer that ends part_lag[j] bases before the window end
    int32_t n_scaffolds;
    int64_t stored;         // distinct keys
};

#define BB_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define BB_TAIL0_WORDS 2048u  // 4^8 bits

// returns 0 on success, else writes a message (reference's assertion texts) into err
int derive_params(const struct bbduk_cfg *cfg, BBParams *p, char *err, int errlen);
