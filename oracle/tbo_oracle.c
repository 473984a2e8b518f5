/*
 * tbo_oracle.c -- CPU restatement of BBDuk's trim-by-overlap step (tbo=t).
 *
 * TEST INFRASTRUCTURE ONLY: linked by tests/, never by bbtools_b200/. PARITY UNPINNED: the reference ships no golden
 * vectors for this step and there is no JVM here (the C file jni/BBMergeOverlapper.c of the reference is an older
 * variant that the Java no longer calls, jgi/BBMergeOverlapper.java:114, and differs in its limits). Pinned only by an
 * independent Python restatement in tests/test_tbo_oracle.py.
 *
 * Follows, statement by statement (paths relative to /root/reference/current):
 *   jgi/BBDuk.java:2878-2926          the tbo block of the per-pair loop (guard, reverse complement of r2,
 *                                     mateByOverlapRatio, minInsert cut, ambig, trimToPosition of both mates, counters)
 *   jgi/BBDuk.java:712-728, :5368-5371 strict / loose constants, minOverlap0=7 minOverlap=14 minInsert0=16 minInsert=40
 *   jgi/BBMergeOverlapper.java:98-136  mateByOverlapRatio dispatch (useQuality=false -> mateByOverlapRatioJava)
 *   jgi/BBMergeOverlapper.java:411-621 mateByOverlapRatioJava (float accumulators, extraBadlimit=20 :1464)
 *   jgi/BBMergeOverlapper.java:785-836 findBestRatio
 *   stream/Read.java:2985-3003         expectedErrors; align2/QualityTools.java:688-698 PROB_ERROR
 *   dna/AminoAcid.java:1315-1332, :206-231 baseToComplementExtended; :468-490 reverseComplementBasesInPlace
 *   shared/TrimRead.java:273-276, :299-346 trimToPosition / trimByAmount
 * All arithmetic is IEEE single precision in the reference's evaluation order (compile with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct tbo_params {
    int32_t min_overlap0, min_overlap, min_insert0, min_insert;
    float max_ratio, min_second_ratio, ratio_margin, ratio_offset, g_incr, b_incr, mee_filter;
    int32_t qual_offset; /* ASCII offset of the quality bytes (33) */
} tbo_params;

#define EXTRA_BADLIMIT 20 /* jgi/BBMergeOverlapper.java:1464 */

static uint8_t g_comp[128];
static float g_prob_error[128];
static int g_ready = 0;

static void init_tables(void) {
    if (g_ready) return;
    /* dna/AminoAcid.java:206-231 */
    const char *nb = " ACMGRSVTWYHKDBNX       ";
    const char *nc = " TGKCYSBAWRDMHVNX       ";
    for (int i = 0; i < 128; i++) g_comp[i] = (uint8_t)i;
    for (int i = 0; i < 24; i++) {
        const unsigned char x = (unsigned char)nb[i], x2 = (unsigned char)nc[i];
        g_comp[x] = x2;
        const unsigned char xl = (x >= 'A' && x <= 'Z') ? (unsigned char)(x + 32) : x;
        const unsigned char x2l = (x2 >= 'A' && x2 <= 'Z') ? (unsigned char)(x2 + 32) : x2;
        g_comp[xl] = x2l;
    }
    g_comp['U'] = 'A';
    g_comp['u'] = 'a';
    g_comp['?'] = '?';
    g_comp[' '] = ' ';
    g_comp['-'] = '-';
    g_comp['*'] = '*';
    g_comp['.'] = '.';
    /* align2/QualityTools.java:688-698 */
    for (int i = 0; i < 128; i++) g_prob_error[i] = (float)pow(10.0, 0 - .1 * i);
    g_prob_error[0] = .75f;
    g_prob_error[1] = .7f;
    g_ready = 1;
}

void tbo_ora_tables(uint8_t *comp128, float *prob_error128) {
    init_tables();
    if (comp128) memcpy(comp128, g_comp, 128);
    if (prob_error128) memcpy(prob_error128, g_prob_error, sizeof g_prob_error);
}

static int is_fully_defined(uint8_t b) { /* jgi/BBDuk.java:5355-5357 == AminoAcid.isFullyDefined */
    const uint8_t y = (uint8_t)(b | 0x20);
    return b < 128 && (y == 'a' || y == 'c' || y == 'g' || y == 't' || y == 'u');
}

/* stream/Read.java:2985-3003 with countUndefined=false, maxBases=-1 */
static float expected_errors(const uint8_t *bases, const uint8_t *quals, int len, int qoff) {
    if (!quals) return 0;
    float sum = 0;
    for (int i = 0; i < len; i++) {
        if (is_fully_defined(bases[i])) sum += g_prob_error[(quals[i] - qoff) & 127];
    }
    return sum;
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }
static int imid(int a, int b, int c) { /* shared/Tools.mid: the median */
    return a < b ? (b < c ? b : imax(a, c)) : (a < c ? a : imax(b, c));
}

/* jgi/BBMergeOverlapper.java:785-836 */
static float find_best_ratio(const uint8_t *abases, int alen, const uint8_t *bbases, int blen, int minOverlap0, int minOverlap,
                             int minInsert, float maxRatio, float offset, float gIncr, float bIncr) {
    float bestRatio = maxRatio + 0.0001f;
    const float halfmax = maxRatio * 0.5f;
    const int largestInsertToTest = alen + blen - minOverlap;
    const int smallestInsertToTest = minInsert;
    for (int insert = largestInsertToTest; insert >= smallestInsertToTest; insert--) {
        const int istart = (insert <= blen ? 0 : insert - blen);
        const int jstart = (insert >= blen ? 0 : blen - insert);
        const int overlapLength = imin(alen - istart, imin(blen - jstart, insert));
        const float badlimit = bestRatio * overlapLength + EXTRA_BADLIMIT;
        float good = 0, bad = 0;
        const int imax_ = istart + overlapLength;
        for (int i = istart, j = jstart; i < imax_ && bad <= badlimit; i++, j++) {
            const uint8_t ca = abases[i], cb = bbases[j];
            if (ca == cb) {
                if (ca != 'N') good += gIncr;
            } else {
                bad += bIncr;
            }
        }
        if (bad <= badlimit) {
            if (bad == 0 && good > minOverlap0 && good < minOverlap) return 100.f;
            const float ratio = (bad + offset) / overlapLength;
            if (ratio < bestRatio) {
                bestRatio = ratio;
                if (good >= minOverlap && ratio < halfmax) return bestRatio;
            }
        }
    }
    return bestRatio;
}

/* jgi/BBMergeOverlapper.java:411-621 with TAG_CUSTOM = MAKE_VECTOR = false; rvector[2] = bestBadInt, rvector[4] = ambig */
static int mate_by_overlap_ratio(const uint8_t *abases, int alen, const uint8_t *bbases, int blen, int *rvector, int minOverlap0,
                                 int minOverlap, int minInsert0, int minInsert, float maxRatio, float minSecondRatio, float margin,
                                 float offset, float gIncr, float bIncr) {
    minOverlap = imax(4, imax(minOverlap0, minOverlap));
    minOverlap0 = imid(4, minOverlap0, minOverlap);
    const int minLength = imin(alen, blen);
    {
        const float x = find_best_ratio(abases, alen, bbases, blen, minOverlap0, minOverlap, minInsert, maxRatio, offset, gIncr, bIncr);
        if (x > maxRatio) {
            rvector[2] = minLength;
            rvector[4] = 0;
            return -1;
        }
        maxRatio = (maxRatio < x ? maxRatio : x);
    }
    const float margin2 = (margin + offset) / minLength;
    int bestInsert = -1, bestBadInt = -1;
    float bestRatio = 1;
    int ambig = 0;
    float secondBestRatio = 1;
    const float extraMult = 1.2f;
    const int largestInsertToTest = alen + blen - minOverlap0;
    const int smallestInsertToTest = minInsert0;
    for (int insert = largestInsertToTest; insert >= smallestInsertToTest; insert--) {
        const int istart = (insert <= blen ? 0 : insert - blen);
        const int jstart = (insert >= blen ? 0 : blen - insert);
        const int overlapLength = imin(alen - istart, imin(blen - jstart, insert));
        const float badlimit = extraMult * ((bestRatio < maxRatio ? bestRatio : maxRatio) * margin * overlapLength) + 1.f + EXTRA_BADLIMIT;
        float good = 0, bad = 0;
        int badInt = 0;
        const int imax_ = istart + overlapLength;
        for (int i = istart, j = jstart; i < imax_ && bad <= badlimit; i++, j++) {
            const uint8_t ca = abases[i], cb = bbases[j];
            if (ca == cb) {
                if (ca != 'N') good += gIncr;
            } else {
                bad += bIncr;
                badInt++;
            }
        }
        if (bad <= badlimit) {
            if (bad == 0 && good > minOverlap0 && good < minOverlap) {
                rvector[2] = bestBadInt;
                rvector[4] = 1;
                return -1;
            }
            const float ratio = (bad + offset) / overlapLength;
            if (ratio < bestRatio * margin) {
                ambig = (ratio * margin >= bestRatio || good < minOverlap);
                if (ratio < bestRatio) {
                    secondBestRatio = bestRatio;
                    bestInsert = insert;
                    bestRatio = ratio;
                    bestBadInt = badInt;
                } else if (ratio < secondBestRatio) {
                    secondBestRatio = ratio;
                }
                if ((ambig && bestRatio < margin2) || secondBestRatio < minSecondRatio) {
                    rvector[2] = bestBadInt;
                    rvector[4] = 1;
                    return -1;
                }
            }
        }
    }
    if (!ambig && bestRatio > maxRatio) bestInsert = -1;
    rvector[2] = bestBadInt;
    rvector[4] = (ambig ? 1 : 0);
    return (bestInsert < 0 ? -1 : bestInsert);
}

/* jgi/BBDuk.java:712-728: strictoverlap=t (default) / f */
void tbo_ora_default_params(tbo_params *p, int strict) {
    p->min_overlap0 = 7;
    p->min_overlap = 14;
    p->min_insert0 = 16;
    p->min_insert = 40;
    p->min_second_ratio = 0.12f;
    p->g_incr = p->b_incr = 0.95f;
    p->qual_offset = 33;
    if (strict) {
        p->max_ratio = 0.05f;
        p->ratio_margin = 9.f;
        p->ratio_offset = 0.5f;
        p->mee_filter = 15.f;
    } else {
        p->max_ratio = 0.10f;
        p->ratio_margin = 5.f;
        p->ratio_offset = 0.4f;
        p->mee_filter = 999999999.f;
    }
}

/*
 * jgi/BBDuk.java:2878-2926 for every pair (2i, 2i+1) of a batch that went through the k-mer block:
 * reads keep original bases [lo,hi); flags bit 0x02 = pair removed. quals may be NULL (then expectedErrors = 0).
 * hi[] is updated in place; insert_out[pair] = bestInsert after the minInsert cut (-1 none), ambig_out[pair];
 * stats[0] += reads trimmed, stats[1] += bases trimmed (readsTrimmedByOverlapT / basesTrimmedByOverlapT).
 */
void tbo_ora_process(const uint8_t *bases, const uint8_t *quals, const int64_t *offsets, int64_t n_reads, const int32_t *lo,
                     int32_t *hi, const uint8_t *flags, const tbo_params *p, int32_t *insert_out, uint8_t *ambig_out,
                     int64_t *stats) {
    init_tables();
    uint8_t *rc = NULL;
    int rc_cap = 0;
    for (int64_t u = 0; u + 1 < n_reads; u += 2) {
        const int64_t i1 = u, i2 = u + 1;
        if (insert_out) insert_out[u / 2] = -1;
        if (ambig_out) ambig_out[u / 2] = 0;
        if (flags[i1] & 0x02) continue; /* remove */
        const uint8_t *a = bases + offsets[i1] + lo[i1];
        const int alen = hi[i1] - lo[i1];
        const uint8_t *b0 = bases + offsets[i2] + lo[i2];
        const int blen = hi[i2] - lo[i2];
        const float ea = expected_errors(a, quals ? quals + offsets[i1] + lo[i1] : NULL, alen, p->qual_offset);
        const float eb = expected_errors(b0, quals ? quals + offsets[i2] + lo[i2] : NULL, blen, p->qual_offset);
        if (!((ea > eb ? ea : eb) < p->mee_filter)) continue;
        if (blen > rc_cap) {
            rc_cap = blen + 64;
            rc = (uint8_t *)realloc(rc, (size_t)rc_cap);
        }
        for (int j = 0; j < blen; j++) rc[j] = g_comp[b0[blen - 1 - j] & 127]; /* r2.reverseComplementFast() */
        int rvector[5] = {0, 0, 0, 0, 0};
        int bestInsert = mate_by_overlap_ratio(a, alen, rc, blen, rvector, p->min_overlap0, p->min_overlap, p->min_insert0,
                                               p->min_insert, p->max_ratio, p->min_second_ratio, p->ratio_margin, p->ratio_offset,
                                               p->g_incr, p->b_incr);
        if (bestInsert < p->min_insert) bestInsert = -1;
        const int ambig = (rvector[4] == 1);
        if (insert_out) insert_out[u / 2] = bestInsert;
        if (ambig_out) ambig_out[u / 2] = (uint8_t)ambig;
        if (bestInsert > 0 && !ambig) {
            /* TrimRead.trimToPosition(r, 0, bestInsert-1, 1) = trimByAmount(r, 0, len-bestInsert, 1): bestInsert >= 1 bases stay */
            if (bestInsert < alen) {
                hi[i1] = lo[i1] + bestInsert;
                stats[0] += 1;
                stats[1] += alen - bestInsert;
            }
            if (bestInsert < blen) {
                hi[i2] = lo[i2] + bestInsert;
                stats[0] += 1;
                stats[1] += blen - bestInsert;
            }
        }
    }
    free(rc);
}
