/*
 * entropy_oracle.c -- CPU restatement of BBDuk's low-entropy read filter (entropy=<cutoff>).
 *
 * TEST INFRASTRUCTURE ONLY: linked by tests/, never by bbtools_b200/. PARITY UNPINNED: the reference ships no golden
 * vectors for this step and there is no JVM here. One more caveat: the 48-entry table of pk*log(pk) is built with the C
 * library's log(); Java's Math.log may differ from it in the last bit for some arguments, which could move a read that
 * sits exactly on the cutoff. Pinned only by an independent Python restatement in tests/test_entropy_oracle.py.
 *
 * Follows, statement by statement (paths relative to /root/reference/current):
 *   jgi/BBDuk.java:3175-3186, :2517-2518  the "Test entropy" block (passes(r.bases, true), setDiscarded, shouldRemove,
 *                                         basesEFilteredT / readsEFilteredT); tracker built with max(0, cutoff), highpass
 *   jgi/BBDuk.java:3260-3289              setDiscarded, isDiscarded, isNullOrDiscarded, isNotDiscarded, shouldRemove
 *   tracker/EntropyTracker.java:62-118    constructor, makeEntropyArray, entropyMult
 *   tracker/EntropyTracker.java:194-201   calcEntropyFast (speed=FAST :1215)
 *   tracker/EntropyTracker.java:657-703   averageEntropy
 *   tracker/EntropyTracker.java:798-806   passes(sequence, allowNs)
 *   tracker/EntropyTracker.java:815-946   add (sliding window: incoming k-mer, outgoing k-mer, running sum in double)
 *   tracker/EntropyTracker.java:950-975   clear
 * Doubles and floats are evaluated in the reference's order (compile with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct entropy_params {
    float cutoff;                /* entropy= ; the tracker gets max(0, cutoff) */
    int32_t k, window;           /* entropyk= (5), entropywindow= (50) */
    int32_t high_pass;           /* entropyHighpass, default true */
    int32_t remove_pairs_if_either_bad, trim_failures_to_1bp;
} entropy_params;

typedef struct etracker {
    int k, windowBases, windowKmers, mask, kmerSpace;
    double *entropy, entropyMult;
    short *counts;
    uint8_t *ring;
    /* mutable state */
    int kmer, kmer2, len, pos, pos2, unique, ns;
    double currentEsum;
} etracker;

static int sym0(uint8_t b) { /* dna/AminoAcid.java symbolToNumber0: A0 C1 G2 T/U3, anything else 0 */
    const uint8_t y = (uint8_t)(b | 0x20);
    if (b >= 128) return 0;
    return y == 'c' ? 1 : y == 'g' ? 2 : (y == 't' || y == 'u') ? 3 : 0;
}

static int is_def(uint8_t b) { /* dna/AminoAcid.java isFullyDefined: A C G T U, either case */
    const uint8_t y = (uint8_t)(b | 0x20);
    return b < 128 && (y == 'a' || y == 'c' || y == 'g' || y == 't' || y == 'u');
}

static void et_init(etracker *t, int k, int window) {
    t->k = k;
    t->windowBases = window;
    t->windowKmers = window - k + 1;
    t->mask = (k > 15 ? -1 : (int)~(~0u << (2 * k)));
    t->kmerSpace = 1 << (2 * k);
    t->entropy = (double *)calloc((size_t)t->windowKmers + 2, sizeof(double));
    const double mult = 1.0 / t->windowKmers;
    for (int i = 1; i < t->windowKmers + 2; i++) {
        const double pk = i * mult;
        t->entropy[i] = pk * log(pk);
    }
    t->entropyMult = -1 / log((double)t->windowKmers);
    t->counts = (short *)calloc((size_t)t->kmerSpace, sizeof(short));
    t->ring = (uint8_t *)calloc((size_t)window, 1);
}

static void et_clear(etracker *t) { /* :950-975 */
    t->kmer = t->kmer2 = t->len = t->pos = t->unique = t->ns = 0;
    t->pos2 = 0 - t->windowBases + t->k - 1;
    t->currentEsum = 0;
    memset(t->counts, 0, sizeof(short) * (size_t)t->kmerSpace);
}

static void et_add(etracker *t, uint8_t b) { /* :815-946 */
    const uint8_t oldBase = t->ring[t->pos];
    t->len++;
    {
        t->ring[t->pos] = b;
        const int n = sym0(b);
        t->kmer = ((t->kmer << 2) | n) & t->mask;
        if (!is_def(b)) t->ns++;
        if (t->len >= t->k) {
            const short oldCount = t->counts[t->kmer];
            if (oldCount < 1) t->unique++;
            const short newCount = t->counts[t->kmer] = (short)(oldCount + 1);
            t->currentEsum = t->currentEsum + t->entropy[newCount] - t->entropy[oldCount];
        }
    }
    if (t->pos2 >= 0) {
        const uint8_t b2 = (t->k > 1 ? t->ring[t->pos2] : oldBase);
        const int n2 = sym0(b2);
        t->kmer2 = ((t->kmer2 << 2) | n2) & t->mask;
        if (t->len > t->windowBases) {
            if (!is_def(oldBase)) t->ns--;
            const short oldCount = t->counts[t->kmer2];
            const short newCount = t->counts[t->kmer2] = (short)(oldCount - 1);
            if (newCount < 1) t->unique--;
            t->currentEsum = t->currentEsum + t->entropy[newCount] - t->entropy[oldCount];
        }
    }
    t->pos++;
    t->pos2++;
    if (t->pos >= t->windowBases) t->pos = 0;
    if (t->pos2 >= t->windowBases) t->pos2 = 0;
}

static float et_calc(const etracker *t) { /* calcEntropyFast :194-201 */
    const float f = (float)(t->currentEsum * t->entropyMult);
    return f > 0 ? f : 0;
}

/* averageEntropy(bases, allowNs = true, 0, n-1) :657-703 */
static float et_average(etracker *t, const uint8_t *bases, int n) {
    et_clear(t);
    int i = 0;
    double sum = 0;
    int divisor = 0;
    const int lim = n < t->windowBases ? n : t->windowBases;
    for (; i < lim; i++) et_add(t, bases[i]);
    sum += et_calc(t);
    divisor++;
    for (; i <= n - 1; i++) {
        et_add(t, bases[i]);
        sum += et_calc(t);
        divisor++;
    }
    const double avg = sum / (divisor > 1 ? divisor : 1);
    return (float)avg;
}

/* entropy of every read's kept interval, for the tests */
void entropy_ora_values(const uint8_t *bases, const int64_t *offsets, int64_t n_reads, const int32_t *lo, const int32_t *hi, int k,
                        int window, float *out) {
    etracker t;
    et_init(&t, k, window);
    for (int64_t i = 0; i < n_reads; i++) out[i] = et_average(&t, bases + offsets[i] + lo[i], hi[i] - lo[i]);
    free(t.entropy);
    free(t.counts);
    free(t.ring);
}

/*
 * jgi/BBDuk.java:3175-3186 for every unit (read, or pair 2i / 2i+1): reads keep [lo,hi); flags bit 0x01 = discarded,
 * 0x02 = unit removed. flags (and, with trimfailuresto1bp, hi) are updated in place;
 * stats[0..1] += readsEFiltered, basesEFiltered.
 */
void entropy_ora_process(const uint8_t *bases, const int64_t *offsets, int64_t n_reads, int paired, const int32_t *lo, int32_t *hi,
                         uint8_t *flags, const entropy_params *p, int64_t *stats) {
    etracker t;
    et_init(&t, p->k, p->window);
    const float cutoff = p->cutoff > 0 ? p->cutoff : 0;
    const int per = paired ? 2 : 1;
    for (int64_t u = 0; u + per <= n_reads; u += per) {
        if (flags[u] & 0x02) continue;
        int disc[2] = {0, 0}, len[2] = {0, 0};
        for (int q = 0; q < per; q++) {
            disc[q] = (flags[u + q] & 0x01) != 0;
            len[q] = hi[u + q] - lo[u + q];
        }
        for (int q = 0; q < per; q++) {
            const int is_disc = disc[q] || (p->trim_failures_to_1bp && len[q] == 1);
            if (is_disc) continue; /* isNotDiscarded */
            const float e = et_average(&t, bases + offsets[u + q] + lo[u + q], len[q]);
            const int passes = (p->high_pass != 0) ^ (e < cutoff);
            if (!passes) { /* setDiscarded */
                if (p->trim_failures_to_1bp) {
                    if (len[q] > 1) { /* trimByAmount(r, 0, len-1, 1) keeps the first base */
                        hi[u + q] = lo[u + q] + 1;
                        len[q] = 1;
                    }
                } else {
                    disc[q] = 1;
                }
            }
        }
        int d[2];
        for (int q = 0; q < per; q++) d[q] = disc[q] || (p->trim_failures_to_1bp && len[q] == 1);
        const int remove = per == 2 ? ((p->remove_pairs_if_either_bad && (d[0] || d[1])) || (d[0] && d[1])) : d[0];
        if (remove) {
            stats[1] += len[0] + (per == 2 ? len[1] : 0);
            stats[0] += per;
        }
        for (int q = 0; q < per; q++) {
            uint8_t f = (uint8_t)(flags[u + q] & ~0x03);
            if (disc[q]) f |= 0x01;
            if (remove) f |= 0x02;
            flags[u + q] = f;
        }
    }
    free(t.entropy);
    free(t.counts);
    free(t.ring);
}

/*
 * jgi/BBDuk.java:3055-3067 with maskLowEntropy (:4432-4446) / trimLowEntropy (:4448-4478) and maskFromBitset (:4505-4526) for every
 * read that is not discarded and whose unit is not removed. mode 1: mask to N, 2: mask to lower case, 3: trim. The BitSet of a
 * read goes to maskbits + mask_off[i] (bit j = base j of the kept interval; zero in mode 3, where lo / hi change instead);
 * stats[0..1] += readsEFiltered, basesEFiltered.
 */
void entropy_ora_mask(const uint8_t *bases, const int64_t *offsets, int64_t n_reads, int paired, int32_t *lo, int32_t *hi,
                      const uint8_t *flags, const entropy_params *p, int mode, uint32_t *maskbits, const int64_t *mask_off,
                      int64_t *stats) {
    etracker t;
    et_init(&t, p->k, p->window);
    const float cutoff = p->cutoff > 0 ? p->cutoff : 0;
    const int per = paired ? 2 : 1;
    for (int64_t i = 0; i < n_reads; i++) {
        const int64_t u = i - (i % per);
        uint32_t *bs = maskbits + mask_off[i];
        const int full = (int)(offsets[i + 1] - offsets[i]);
        for (int w = 0; w < (full + 31) / 32; w++) bs[w] = 0;
        if (flags[u] & 0x02) continue;
        const int n = hi[i] - lo[i];
        const int is_disc = (flags[i] & 0x01) || (p->trim_failures_to_1bp && n == 1);
        if (is_disc) continue; /* isNotDiscarded */
        if (n < t.windowBases) continue;
        const uint8_t *b = bases + offsets[i] + lo[i];
        et_clear(&t);
        for (int j = 0, min = t.windowBases - 1; j < n; j++) {
            et_add(&t, b[j]);
            if (j >= min && t.ns < 1) {
                const float e = et_calc(&t);
                const int passes = (p->high_pass != 0) ^ (e < cutoff);
                if (!passes)
                    for (int q = t.len - t.windowBases; q <= t.len - 1; q++) bs[q >> 5] |= 1u << (q & 31); /* leftPos..rightPos */
            }
        }
        int masked = 0;
        if (mode == 3) {
            int left = 0, right = 0;
            for (int j = 0; j < n; j++) {
                if ((bs[j >> 5] >> (j & 31)) & 1u) left++;
                else break;
            }
            for (int j = n - 1; j >= 0; j--) {
                if ((bs[j >> 5] >> (j & 31)) & 1u) right++;
                else break;
            }
            for (int w = 0; w < (full + 31) / 32; w++) bs[w] = 0;
            if (left || right) { /* TrimRead.trimByAmount(r, left, right, 1), shared/TrimRead.java:299-346 */
                const int minLen = n < 1 ? n : 1;
                if (left + right + minLen > n) {
                    right = n - minLen > 1 ? n - minLen : 1;
                    left = 0;
                }
                lo[i] += left;
                hi[i] -= right;
                masked = left + right;
            }
        } else {
            for (int j = 0; j < n; j++) {
                if (!((bs[j >> 5] >> (j & 31)) & 1u)) continue;
                if (mode == 1) masked += b[j] != 'N';
                else masked += !(b[j] >= 'a' && b[j] <= 'z') && b[j] != 'N';
            }
        }
        stats[1] += masked;
        stats[0] += masked > 0;
    }
    free(t.entropy);
    free(t.counts);
    free(t.ring);
}
