/*
 * qtrim_oracle.c -- CPU restatement of BBDuk's quality-trimming block and the per-read quality / length / N filters
 * (first part).
 *
 * TEST INFRASTRUCTURE ONLY: linked by tests/, never by bbtools_b200/. PARITY UNPINNED: the reference ships no golden
 * vectors for this step and there is no JVM here. Pinned only by an independent Python restatement in
 * tests/test_qtrim_oracle.py.
 *
 * Follows, statement by statement (paths relative to /root/reference/current):
 *   jgi/BBDuk.java:2954-3052, :4721-4825 poly-A / poly-G / poly-C trimming and filtering (trimPolyA, trimPoly, detectPolyLeft,
 *                                       detectPolyRight; stream/Read.java:3387-3401 countLeft / countRight), each with its
 *                                       minlen test and shouldRemove -- including the reference's use of r1 in the
 *                                       filterpolyc test of r2 (:3035)
 *   jgi/BBDuk.java:3074-3108            "Do quality trimming": trimFast of r1 and r2, minlen / maxlen, shouldRemove
 *   jgi/BBDuk.java:3110-3170            "Do quality filtering": minavgquality (stream/Read.java:2181-2226, :2985-3001,
 *                                       align2/QualityTools.java:674-680), minbasequality, maxns, shouldRemove (maxnrate,
 *                                       minconsecutivebases, minbasefrequency at their defaults = off)
 *   jgi/BBDuk.java:3260-3289            setDiscarded, isDiscarded, isNullOrDiscarded, isNotDiscarded, shouldRemove
 *   shared/TrimRead.java:140-169        trimFast (optimalMode=true :954, discardUnder=0)
 *   shared/TrimRead.java:348-410        testOptimal (NPROB=0.75f :964)
 *   shared/TrimRead.java:299-346        trimByAmount
 *   parse/Parser.java:1757-1759, align2/QualityTools.java:650-654, :688-698   trimE, PROB_ERROR
 *   stream/Read.java:2281-2289, :2835-2844   minQuality, countUndefined
 * All arithmetic is IEEE single precision in the reference's evaluation order (compile with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct qtrim_params {
    int32_t qtrim_left, qtrim_right;
    float trimq;
    int32_t min_base_quality, max_ns, max_read_length, qual_offset;
    /* from the BBDuk configuration */
    int32_t min_read_length;
    float min_len_fraction;
    int32_t remove_pairs_if_either_bad, trim_failures_to_1bp;
    /* poly-X (parse/Parser.java:386-411, :1815-1831) */
    int32_t trim_poly_a, trim_poly_g_left, trim_poly_g_right, filter_poly_g, trim_poly_c_left, trim_poly_c_right, filter_poly_c,
        max_non_poly;
    /* minavgquality= (parse/Parser.java:516-526) */
    float min_avg_quality;
    int32_t min_avg_quality_bases;
    /* maxnrate= (>= 1 = off), minconsecutivebases=, minbasefrequency= (jgi/BBDuk.java:3138-3159) */
    float max_n_rate;
    int32_t min_consecutive_bases;
    float min_base_frequency;
    /* TrimRead's static modes: 0 = optimalMode (default), 1 = windowMode (qtrim=w[,N]), 2 = neither (optitrim=f) */
    int32_t trim_mode, window_length, min_good_interval;
} qtrim_params;

static float g_pe[128];
static int g_pe_ready = 0;

static void init_pe(void) {
    if (g_pe_ready) return;
    for (int i = 0; i < 128; i++) g_pe[i] = (float)pow(10.0, 0 - .1 * i); /* align2/QualityTools.java:688-698 */
    g_pe[0] = .75f;
    g_pe[1] = .7f;
    g_pe_ready = 1;
}

/* align2/QualityTools.java:650-654 */
static double phred_to_prob_error(double q) {
    if (q <= 0) return 0.75;
    if (q <= 1) return 0.75 - q * 0.05;
    const double p = pow(10, -0.1 * q);
    return p < 0.7 ? p : 0.7;
}

typedef struct qread {
    const uint8_t *bases, *quals; /* the read's first base / quality byte (untrimmed) */
    int lo, hi;                   /* the interval it currently keeps */
    int discarded;
} qread;

static int is_defined(uint8_t b) { /* dna/AminoAcid.java:1289-1320 baseToNumber >= 0 */
    const uint8_t y = (uint8_t)(b | 0x20);
    return b < 128 && (y == 'a' || y == 'c' || y == 'g' || y == 't' || y == 'u');
}

/* shared/TrimRead.java:299-346 on the kept interval */
static int trim_by_amount(qread *r, int left, int right, int min_len) {
    if (left < 0) left = 0;
    if (right < 0) right = 0;
    const int len = r->hi - r->lo;
    if (len < 1) return 0;
    if (min_len < 0) min_len = 0;
    if (min_len > len) min_len = len;
    if (left + right + min_len > len) {
        right = (len - min_len) > 1 ? (len - min_len) : 1;
        left = 0;
    }
    r->lo += left;
    r->hi -= right;
    return left + right;
}

/* shared/TrimRead.java:348-410; quality bytes are (byte)(ascii - qual_offset) */
static void test_optimal(const qread *r, float avg_error_rate, int qual_offset, int *left_out, int *right_out) {
    const int n = r->hi - r->lo;
    const uint8_t *bases = r->bases + r->lo, *qual = r->quals + r->lo;
    float maxScore = 0, score = 0;
    int maxLoc = -1, maxCount = -1, count = 0;
    float nprob = avg_error_rate * 1.1f;
    if (nprob > 1) nprob = 1;
    if (nprob < 0.75f) nprob = 0.75f;
    for (int i = 0; i < n; i++) {
        const uint8_t b = bases[i];
        const int8_t q = (int8_t)(qual[i] - qual_offset);
        const float probError = (b == 'N' || q < 1) ? nprob : g_pe[q];
        const float delta = avg_error_rate - probError;
        score = score + delta;
        if (score > 0) {
            count++;
            if (score > maxScore || (score == maxScore && count > maxCount)) {
                maxScore = score;
                maxCount = count;
                maxLoc = i;
            }
        } else {
            score = 0;
            count = 0;
        }
    }
    if (maxScore > 0) {
        *left_out = maxLoc - maxCount + 1;
        *right_out = n - maxLoc - 1;
    } else {
        *left_out = 0;
        *right_out = n;
    }
}

/* shared/TrimRead.java:477-489 */
static int test_left_n(const uint8_t *bases, int n, int minGoodInterval) {
    if (n == 0) return 0;
    int good = 0, lastBad = -1;
    for (int i = 0; i < n && good < minGoodInterval; i++) {
        if (bases[i] != 'N') good++;
        else { good = 0; lastBad = i; }
    }
    return lastBad + 1;
}
/* shared/TrimRead.java:491-503 */
static int test_right_n(const uint8_t *bases, int n, int minGoodInterval) {
    if (n == 0) return 0;
    int good = 0, lastBad = n;
    for (int i = n - 1; i >= 0 && good < minGoodInterval; i--) {
        if (bases[i] != 'N') good++;
        else { good = 0; lastBad = i; }
    }
    return n - lastBad;
}
/* shared/TrimRead.java:416-436 */
static int test_left(const uint8_t *bases, const uint8_t *qual, int n, int8_t trimq, int qual_offset, int minGoodInterval) {
    if (n == 0) return 0;
    if (!qual) return trimq < 0 ? 0 : test_left_n(bases, n, minGoodInterval);
    int good = 0, lastBad = -1;
    for (int i = 0; i < n && good < minGoodInterval; i++) {
        const int8_t q = (int8_t)(qual[i] - qual_offset);
        if (q > trimq) good++;
        else { good = 0; lastBad = i; }
    }
    return lastBad + 1;
}
/* shared/TrimRead.java:457-475 */
static int test_right(const uint8_t *bases, const uint8_t *qual, int n, int8_t trimq, int qual_offset, int minGoodInterval) {
    if (n == 0) return 0;
    if (!qual) return trimq < 0 ? 0 : test_right_n(bases, n, minGoodInterval);
    int good = 0, lastBad = n;
    for (int i = n - 1; i >= 0 && good < minGoodInterval; i--) {
        const int8_t q = (int8_t)(qual[i] - qual_offset);
        if (q > trimq) good++;
        else { good = 0; lastBad = i; }
    }
    return n - lastBad;
}
/* shared/TrimRead.java:438-455 */
static int test_right_window(const uint8_t *bases, const uint8_t *qual, int n, int8_t trimq, int qual_offset, int window,
                             int minGoodInterval) {
    if (n == 0) return 0;
    if (!qual || n < window) return trimq > 0 ? 0 : test_right_n(bases, n, minGoodInterval);
    int thresh = window * trimq;
    if (thresh < 1) thresh = 1;
    int sum = 0;
    for (int i = 0, j = -window; i < n; i++, j++) {
        sum += (int8_t)(qual[i] - qual_offset);
        if (j >= -1) {
            if (j >= 0) sum -= (int8_t)(qual[j] - qual_offset);
            if (sum < thresh) return n - j - 1;
        }
    }
    return 0;
}

/* shared/TrimRead.java:140-171 with discardUnder = 0, trimClip = false */
static int trim_fast(qread *r, const qtrim_params *p, float trimE) {
    const int n = r->hi - r->lo;
    if (n < 1) return 0;
    const uint8_t *bases = r->bases + r->lo, *qual = r->quals ? r->quals + r->lo : NULL;
    const int8_t tq = (int8_t)(int)p->trimq; /* (byte)trimq */
    int a, b;
    if (p->trim_mode == 0) {
        int a0, b0;
        if (!qual) { /* :352 */
            a0 = trimE >= 1 ? 0 : test_left_n(bases, n, p->min_good_interval);
            b0 = trimE >= 1 ? 0 : test_right_n(bases, n, p->min_good_interval);
        } else {
            test_optimal(r, trimE, p->qual_offset, &a0, &b0);
        }
        a = p->qtrim_left ? a0 : 0;
        b = p->qtrim_right ? b0 : 0;
    } else if (p->trim_mode == 1) {
        a = 0;
        b = p->qtrim_right ? test_right_window(bases, qual, n, tq, p->qual_offset, p->window_length, p->min_good_interval) : 0;
    } else {
        a = p->qtrim_left ? test_left(bases, qual, n, tq, p->qual_offset, p->min_good_interval) : 0;
        b = p->qtrim_right ? test_right(bases, qual, n, tq, p->qual_offset, p->min_good_interval) : 0;
    }
    return trim_by_amount(r, a, b, 1);
}

/* stream/Read.java:3387-3401 on the kept interval */
static int count_left(const qread *r, uint8_t c) {
    const int n = r->hi - r->lo;
    for (int i = 0; i < n; i++)
        if (r->bases[r->lo + i] != c) return i;
    return n;
}
static int count_right(const qread *r, uint8_t c) {
    const int n = r->hi - r->lo;
    for (int i = n - 1; i >= 0; i--)
        if (r->bases[r->lo + i] != c) return n - i - 1;
    return n;
}

/* jgi/BBDuk.java:4721-4736 */
static int trim_poly_a(qread *r, int minPoly) {
    if (r->hi - r->lo < minPoly) return 0;
    int la = count_left(r, 'A'), lt = count_left(r, 'T'), ra = count_right(r, 'A'), rt = count_right(r, 'T');
    int left = la > lt ? la : lt, right = ra > rt ? ra : rt;
    if (left < minPoly) left = 0;
    if (right < minPoly) right = 0;
    int trimmed = 0;
    if (left > 0 || right > 0) trimmed = trim_by_amount(r, left, right, 1);
    return trimmed;
}

/* jgi/BBDuk.java:4771-4791 */
static int detect_poly_left(const qread *r, int minPoly, int maxNonPoly, uint8_t c) {
    const int n = r->hi - r->lo;
    if (n < minPoly) return 0;
    int trimTo = -1;
    for (int i = 0, polymer = 0, nonpoly = 0; i < n && nonpoly <= maxNonPoly; i++) {
        if (r->bases[r->lo + i] == c) {
            polymer++;
            if (polymer >= minPoly) {
                nonpoly = 0;
                trimTo = i;
            }
        } else {
            polymer = 0;
            nonpoly++;
        }
    }
    return trimTo + 1;
}

/* jgi/BBDuk.java:4802-4822 */
static int detect_poly_right(const qread *r, int minPoly, int maxNonPoly, uint8_t c) {
    const int n = r->hi - r->lo;
    if (n < minPoly) return 0;
    int trimTo = n;
    for (int i = n - 1, polymer = 0, nonpoly = 0; i >= 0 && nonpoly <= maxNonPoly; i--) {
        if (r->bases[r->lo + i] == c) {
            polymer++;
            if (polymer >= minPoly) {
                nonpoly = 0;
                trimTo = i;
            }
        } else {
            polymer = 0;
            nonpoly++;
        }
    }
    return n - trimTo;
}

/* jgi/BBDuk.java:4747-4760 */
static int trim_poly(qread *r, int minLeft, int minRight, int maxNonPoly, uint8_t c) {
    const int left = minLeft > 0 ? detect_poly_left(r, minLeft, maxNonPoly, c) : 0;
    const int right = minRight > 0 ? detect_poly_right(r, minRight, maxNonPoly, c) : 0;
    int trimmed = 0;
    if (left > 0 || right > 0) trimmed = trim_by_amount(r, left, right, 1);
    return trimmed;
}

/* Read.avgQuality(false, maxBases) with AVERAGE_QUALITY_BY_PROBABILITY (stream/Read.java:2181-2226): expectedErrors over
 * the defined bases (:2985-3001), divided by the number of bases looked at, as a phred score in double precision */
static double avg_quality(const qread *r, int qual_offset, int max_bases) {
    const int n = r->hi - r->lo;
    if (n == 0) return 0;
    const int limit = max_bases < 1 ? n : (max_bases < n ? max_bases : n);
    float sum = 0;
    for (int i = 0; i < limit; i++) {
        if (is_defined(r->bases[r->lo + i])) {
            const int8_t q = (int8_t)(r->quals[r->lo + i] - qual_offset);
            sum += g_pe[q < 0 ? 0 : q];
        }
    }
    const float pr = sum / limit;
    const double prob = pr; /* align2/QualityTools.java:674-680 */
    if (prob >= 1) return 0;
    if (prob <= 0.000001) return 60;
    return -10 * log10(prob);
}

static void set_discarded(const qtrim_params *p, qread *r) { /* jgi/BBDuk.java:3260-3266 */
    if (p->trim_failures_to_1bp) {
        if (r->hi - r->lo > 1) trim_by_amount(r, 0, r->hi - r->lo - 1, 1);
    } else {
        r->discarded = 1;
    }
}
static int is_discarded(const qtrim_params *p, const qread *r) {
    if (!r) return 0;
    if (r->discarded) return 1;
    return p->trim_failures_to_1bp && (r->hi - r->lo) == 1;
}
static int is_null_or_discarded(const qtrim_params *p, const qread *r) { return !r || is_discarded(p, r); }
static int is_not_discarded(const qtrim_params *p, const qread *r) { return r && !is_discarded(p, r); }
static int should_remove(const qtrim_params *p, const qread *r1, const qread *r2) {
    return (p->remove_pairs_if_either_bad && (is_discarded(p, r1) || is_discarded(p, r2))) ||
           (is_discarded(p, r1) && is_null_or_discarded(p, r2));
}

/*
 * For every unit (read, or pair 2i / 2i+1) of a batch that went through the k-mer block: reads keep [lo,hi) of their
 * original bases; flags bit 0x01 = discarded, 0x02 = unit removed. lo / hi / flags are updated in place (0x40 = quality
 * trimmed, 0x80 = poly-X trimmed); stats[0..7] += readsQTrimmed, basesQTrimmed, readsQFiltered, basesQFiltered,
 * readsNFiltered, basesNFiltered, readsPolyTrimmed, basesPolyTrimmed.
 */
void qtrim_ora_process(const uint8_t *bases, const uint8_t *quals, const int64_t *offsets, int64_t n_reads, int paired,
                       int32_t *lo, int32_t *hi, uint8_t *flags, const qtrim_params *p, int64_t *stats) {
    init_pe();
    const float trimE = (float)phred_to_prob_error(p->trimq);
    const int per = paired ? 2 : 1;
    const int max_len = p->max_read_length > 0 ? p->max_read_length : 0x7FFFFFFF;
    for (int64_t u = 0; u + per <= n_reads; u += per) {
        if (flags[u] & 0x02) continue; /* remove */
        qread rr[2];
        int minlen[2];
        for (int q = 0; q < per; q++) {
            const int64_t i = u + q;
            rr[q].bases = bases + offsets[i];
            rr[q].quals = quals ? quals + offsets[i] : NULL;
            rr[q].lo = lo[i];
            rr[q].hi = hi[i];
            rr[q].discarded = (flags[i] & 0x01) != 0;
            const float initial = (float)(offsets[i + 1] - offsets[i]);
            const float a = initial * p->min_len_fraction, b = (float)p->min_read_length;
            minlen[q] = (int)(a > b ? a : b); /* jgi/BBDuk.java:2587-2593 */
        }
        qread *r1 = &rr[0], *r2 = per == 2 ? &rr[1] : NULL;
        int remove = 0;
        int qtrimmed[2] = {0, 0}, ptrimmed[2] = {0, 0};
        /* :2954-2979 poly-A */
        if (!remove && p->trim_poly_a > 0) {
            for (int q = 0; q < per; q++) {
                const int x = trim_poly_a(&rr[q], p->trim_poly_a);
                stats[7] += x;
                stats[6] += (x > 0 ? 1 : 0);
                ptrimmed[q] |= x > 0;
                if (rr[q].hi - rr[q].lo < minlen[q]) set_discarded(p, &rr[q]);
            }
            if (should_remove(p, r1, r2)) {
                stats[7] += (r1->hi - r1->lo) + (r2 ? r2->hi - r2->lo : 0);
                remove = 1;
            }
        }
        /* :2981-3016 poly-G, :3018-3052 poly-C */
        for (int which = 0; which < 2; which++) {
            const uint8_t c = which == 0 ? 'G' : 'C';
            const int tl = which == 0 ? p->trim_poly_g_left : p->trim_poly_c_left;
            const int tr = which == 0 ? p->trim_poly_g_right : p->trim_poly_c_right;
            const int fp = which == 0 ? p->filter_poly_g : p->filter_poly_c;
            if (remove || !(tl > 0 || tr > 0 || fp > 0)) continue;
            for (int q = 0; q < per; q++) {
                /* the poly-C filter of r2 looks at r1 (:3035) */
                const qread *probe = (which == 1 && q == 1) ? r1 : &rr[q];
                if (fp > 0 && detect_poly_left(probe, fp, p->max_non_poly, c) >= fp) {
                    set_discarded(p, &rr[q]);
                    stats[6] += 1;
                } else if (tl > 0 || tr > 0) {
                    const int x = trim_poly(&rr[q], tl, tr, p->max_non_poly, c);
                    stats[7] += x;
                    stats[6] += (x > 0 ? 1 : 0);
                    ptrimmed[q] |= x > 0;
                    if (rr[q].hi - rr[q].lo < minlen[q]) set_discarded(p, &rr[q]);
                }
            }
            if (should_remove(p, r1, r2)) {
                stats[7] += (r1->hi - r1->lo) + (r2 ? r2->hi - r2->lo : 0);
                remove = 1;
            }
        }
        /* :3074-3108 */
        if (!remove) {
        if (p->qtrim_left || p->qtrim_right) {
            for (int q = 0; q < per; q++) {
                const int x = trim_fast(&rr[q], p, trimE);
                stats[1] += x;
                stats[0] += (x > 0 ? 1 : 0);
                qtrimmed[q] = x > 0;
            }
        }
        for (int q = 0; q < per; q++) {
            if (is_not_discarded(p, &rr[q])) {
                const int len = rr[q].hi - rr[q].lo;
                if (len < minlen[q] || len > max_len) set_discarded(p, &rr[q]);
            }
        }
        if (should_remove(p, r1, r2)) {
            stats[1] += (r1->hi - r1->lo) + (r2 ? r2->hi - r2->lo : 0); /* basesQTrimmedT+=r1.pairLength() */
            remove = 1;
        }
        }
        /* :3110-3170 */
        if (!remove) {
            if (p->min_avg_quality > 0) {
                for (int q = 0; q < per; q++)
                    if (rr[q].quals && avg_quality(&rr[q], p->qual_offset, p->min_avg_quality_bases) < p->min_avg_quality)
                        set_discarded(p, &rr[q]);
            }
            if (p->min_base_quality > 0) {
                for (int q = 0; q < per; q++) {
                    if (!rr[q].quals) continue;
                    int mn = 41;
                    for (int i = rr[q].lo; i < rr[q].hi; i++) {
                        const int8_t v = (int8_t)(rr[q].quals[i] - p->qual_offset);
                        if (v < mn) mn = v;
                    }
                    if (mn < p->min_base_quality) set_discarded(p, &rr[q]);
                }
            }
            if (p->max_ns >= 0) {
                for (int q = 0; q < per; q++) {
                    int n = 0;
                    for (int i = rr[q].lo; i < rr[q].hi; i++) n += is_defined(rr[q].bases[i]) ? 0 : 1;
                    if (n > p->max_ns) {
                        stats[4] += 1;
                        stats[5] += rr[q].hi - rr[q].lo;
                        set_discarded(p, &rr[q]);
                    }
                }
            }
            /* jgi/BBDuk.java:3136-3149: the same as a fraction of the read length; r.discarded() is the raw flag */
            if (p->max_n_rate < 1) {
                for (int q = 0; q < per; q++) {
                    if (rr[q].discarded) continue;
                    int n = 0;
                    for (int i = rr[q].lo; i < rr[q].hi; i++) n += is_defined(rr[q].bases[i]) ? 0 : 1;
                    if ((float)n > p->max_n_rate * (float)(rr[q].hi - rr[q].lo)) {
                        stats[4] += 1;
                        stats[5] += rr[q].hi - rr[q].lo;
                        set_discarded(p, &rr[q]);
                    }
                }
            }
            /* :3150-3154, stream/Read.java:2846-2858 hasMinConsecutiveBases */
            if (p->min_consecutive_bases > 0) {
                for (int q = 0; q < per; q++) {
                    if (!is_not_discarded(p, &rr[q])) continue;
                    int len = 0, ok = 0;
                    for (int i = rr[q].lo; i < rr[q].hi; i++) {
                        if (!is_defined(rr[q].bases[i])) {
                            len = 0;
                        } else {
                            len++;
                            if (len >= p->min_consecutive_bases) {
                                ok = 1;
                                break;
                            }
                        }
                    }
                    if (!ok) set_discarded(p, &rr[q]);
                }
            }
            /* :3155-3159, stream/Read.java:2864-2874 minBaseCount: upper-case A C G T only */
            if (p->min_base_frequency > 0) {
                for (int q = 0; q < per; q++) {
                    int a = 0, c = 0, g = 0, t = 0;
                    for (int i = rr[q].lo; i < rr[q].hi; i++) {
                        const uint8_t b = rr[q].bases[i];
                        if (b == 'A') a++;
                        else if (b == 'C') c++;
                        else if (b == 'G') g++;
                        else if (b == 'T') t++;
                    }
                    int mn = a < c ? a : c;
                    if (g < mn) mn = g;
                    if (t < mn) mn = t;
                    if ((float)mn < p->min_base_frequency * (float)(rr[q].hi - rr[q].lo)) set_discarded(p, &rr[q]);
                }
            }
            if (should_remove(p, r1, r2)) {
                stats[3] += (r1->hi - r1->lo) + (r2 ? r2->hi - r2->lo : 0);
                stats[2] += per;
                remove = 1;
            }
        }
        for (int q = 0; q < per; q++) {
            const int64_t i = u + q;
            lo[i] = rr[q].lo;
            hi[i] = rr[q].hi;
            uint8_t f = (uint8_t)(flags[i] & ~0x03);
            if (rr[q].discarded) f |= 0x01;
            if (remove) f |= 0x02;
            if (qtrimmed[q]) f |= 0x40;
            if (ptrimmed[q]) f |= 0x80;
            flags[i] = f;
        }
    }
}
