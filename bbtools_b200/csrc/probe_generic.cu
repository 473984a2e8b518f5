// probe_generic.cu -- the complete per-read k-mer block, every mode, one thread per read pair.
//
// This kernel covers every mode and flag of the path (ktrim l/r/tips, kmask, ksplit, the four kfilter
// scorers, qhdist, qskip, speed, restrictleft/right, k>31, reads of any length, N/IUPAC) with the
// reference's exact scan order; probe_fast.cu is the tuned kernel for the common configurations and
// hands anything it does not cover to this one. Rolling state per base follows
// jgi/BBDuk.java:3882-3888 (== bbduk/BBDukProcessorS.java:2009-2016); see each function for its lines.
#include <algorithm>

#include "bbduk_dev.cuh"
#include "probe.h"

namespace {

struct RState {  // a read as the scans see it: current bases = b[lo..hi)
    const uint8_t *b;
    int lo, hi;
    int pairnum;
    int discarded;
    int credit0, credit1;
    int splitStart;
    int64_t base_off;  // offset of b[0] in the batch (diagnostics only)
    uint32_t *mask;    // kmask bit words of this read, or null
};

struct Ctx {
    const BBParams &p;
    const BBTable &t;
    unsigned long long *scaf_reads, *scaf_bases;
};

__device__ __forceinline__ int rlen(const RState &r) { return r.hi - r.lo; }

__device__ __forceinline__ void credit(const Ctx &c, RState &r, int id, int blen) {
    if (c.scaf_reads) {
        atomicAdd(c.scaf_reads + id, 1ull);
        atomicAdd(c.scaf_bases + id, (unsigned long long)blen);
    }
    if (r.credit0 < 0)
        r.credit0 = id;
    else
        r.credit1 = id;
}

// shared/TrimRead.java:299-346 trimByAmount (plain reads: no match string, no cigar)
__device__ __forceinline__ int trim_by_amount(RState &r, int left, int right, int minLen) {
    left = max(left, 0);
    right = max(right, 0);
    const int len = rlen(r);
    if (len < 1) return 0;
    minLen = min(len, max(minLen, 0));
    if (left + right + minLen > len) {
        right = max(1, len - minLen);
        left = 0;
    }
    const int total = left + right;
    if (total > 0) {
        r.lo += left;
        r.hi -= right;
    }
    return total;
}
// shared/TrimRead.java:273-276
__device__ __forceinline__ int trim_to_position(RState &r, int leftLoc, int rightLoc, int minLen) {
    return trim_by_amount(r, leftLoc, rlen(r) - rightLoc - 1, minLen);
}

// jgi/BBDuk.java:3365-3386 getValueInner
__device__ __forceinline__ int get_value_inner(const Ctx &c, uint64_t kmer, uint64_t rkmer, uint64_t lengthMask, int qPos) {
    if (c.p.qSkip > 1 && (qPos % c.p.qSkip != 0)) return -1;
    const uint64_t key = bb_to_value(c.p, kmer, rkmer, lengthMask);
    if (bb_passes_speed(c.p, key)) return bb_table_get(c.t, key);
    return -1;
}

// jgi/BBDuk.java:3335-3354 getValue: on a miss, try every single substitution of the QUERY k-mer,
// symbol-major then position, recursively qHDist deep; the first hit wins.
template <int DEPTH>
__device__ int get_value_rec(const Ctx &c, uint64_t kmer, uint64_t rkmer, uint64_t lengthMask, int qPos, int len, int qHDist) {
    int id = get_value_inner(c, kmer, rkmer, lengthMask, qPos);
    if constexpr (DEPTH > 0) {
        if (id < 1 && qHDist > 0) {
            for (int j = 0; j < 4 && id < 1; j++) {
                for (int i = 0; i < len && id < 1; i++) {
                    const uint64_t temp = (kmer & ~(3ull << (2 * i))) | ((uint64_t)j << (2 * i));
                    if (temp != kmer) id = get_value_rec<DEPTH - 1>(c, temp, bb_rcomp(temp, len), lengthMask, qPos, len, qHDist - 1);
                }
            }
        }
    }
    return id;
}
__device__ __forceinline__ int get_value(const Ctx &c, uint64_t kmer, uint64_t rkmer, uint64_t lengthMask, int qPos, int len, int qHDist) {
    if (qHDist <= 0) return get_value_inner(c, kmer, rkmer, lengthMask, qPos);
    return get_value_rec<3>(c, kmer, rkmer, lengthMask, qPos, len, qHDist);
}

#define BB_ROLL(ch)                                                            \
    do {                                                                       \
        const uint32_t c_ = (ch);                                              \
        const bool def_ = bb_defined(c_);                                      \
        const uint64_t x_ = def_ ? bb_code_raw(c_) : 0u;                       \
        const uint64_t x2_ = def_ ? (3u - bb_code_raw(c_)) : 0u;               \
        kmer = ((kmer << 2) | x_) & p.mask;                                    \
        rkmer = ((rkmer >> 2) | (x2_ << p.shift2)) & p.mask;                   \
        if (p.forbidNs && !def_) {                                             \
            len = 0;                                                           \
            rkmer = 0;                                                         \
        } else {                                                               \
            len++;                                                             \
        }                                                                      \
    } while (0)

#define BB_SCAN_WINDOW()                                                                         \
    const int start = (p.restrictRight < 1 ? 0 : max(0, blen - p.restrictRight));                \
    const int stop = (p.restrictLeft < 1 ? blen : min(blen, p.restrictLeft))

__device__ __forceinline__ bool skipped(const BBParams &p, const RState &r) {
    return (p.skipR1 && r.pairnum == 0) || (p.skipR2 && r.pairnum == 1);
}

// jgi/BBDuk.java:3395-3457 countSetKmers
__device__ int count_set_kmers(const Ctx &c, RState &r, int maxBadKmers) {
    const BBParams &p = c.p;
    const int blen = rlen(r);
    if (blen < p.k || c.t.stored < 1 || skipped(p, r)) return 0;
    const uint8_t *bases = r.b + r.lo;
    uint64_t kmer = 0, rkmer = 0;
    int found = 0, len = 0;
    BB_SCAN_WINDOW();
    for (int i = start; i < stop; i++) {
        BB_ROLL(bases[i]);
        if (len >= p.minlen2 && i >= p.minlen) {
            const int id = get_value(c, kmer, rkmer, p.kmask, i, p.k, p.qHammingDistance);
            if (id > 0) {
                if (found == maxBadKmers) {
                    credit(c, r, id, blen);
                    return found + 1;
                }
                found++;
            }
        }
    }
    return found;
}

// jgi/BBDuk.java:3466-3519 countCoveredBases
__device__ int count_covered_bases(const Ctx &c, RState &r, int minCoveredBases) {
    const BBParams &p = c.p;
    const int blen = rlen(r);
    if (blen < p.k || c.t.stored < 1 || skipped(p, r)) return 0;
    const uint8_t *bases = r.b + r.lo;
    uint64_t kmer = 0, rkmer = 0;
    int found = 0, len = 0, lastFound = -1;
    BB_SCAN_WINDOW();
    for (int i = start; i < stop; i++) {
        BB_ROLL(bases[i]);
        if (len >= p.minlen2 && i >= p.minlen) {
            const int id = get_value(c, kmer, rkmer, p.kmask, i, p.k, p.qHammingDistance);
            if (id > 0) {
                found += min(p.k, i - lastFound);
                lastFound = i;
                if (found >= minCoveredBases) {
                    credit(c, r, id, blen);
                    return found;
                }
            }
        }
    }
    return found;
}

// jgi/BBDuk.java:3527-3589 findBestMatch. The winner is the first-seen id with the maximal hit
// count. Per-thread storage for per-id counters is not available, so ties and counts are resolved by
// re-scanning: pass 1 collects hits, pass 2..n evaluate candidates in first-seen order. To keep it
// O(hits^2) at worst only the distinct ids seen are compared (reads carry few distinct ids).
__device__ int find_best_match(const Ctx &c, RState &r, int maxBadKmers) {
    const BBParams &p = c.p;
    const int blen = rlen(r);
    if (blen < p.k || c.t.stored < 1 || skipped(p, r)) return -1;
    const uint8_t *bases = r.b + r.lo;
    BB_SCAN_WINDOW();
    int found = 0, best_id = -1, best_cnt = 0;
    // outer: candidate = id of the n-th distinct hit in first-seen order
    constexpr int MAXD = 16;  // distinct ids tracked in registers; more falls back to rescans
    int ids[MAXD], cnts[MAXD], nd = 0;
    bool overflow = false;
    {
        uint64_t kmer = 0, rkmer = 0;
        int len = 0;
        for (int i = start; i < stop; i++) {
            BB_ROLL(bases[i]);
            if (len >= p.minlen2 && i >= p.minlen) {
                const int id = get_value(c, kmer, rkmer, p.kmask, i, p.k, p.qHammingDistance);
                if (id > 0) {
                    found++;
                    int q = 0;
                    for (; q < nd; q++)
                        if (ids[q] == id) break;
                    if (q < nd)
                        cnts[q]++;
                    else if (nd < MAXD) {
                        ids[nd] = id;
                        cnts[nd] = 1;
                        nd++;
                    } else
                        overflow = true;
                }
            }
        }
    }
    if (found <= maxBadKmers) return -1;
    if (!overflow) {
        for (int q = 0; q < nd; q++)
            if (cnts[q] > best_cnt) {
                best_cnt = cnts[q];
                best_id = ids[q];
            }
    } else {
        // many distinct ids: for each hit position in order, count that id's total by a rescan
        uint64_t kmer = 0, rkmer = 0;
        int len = 0;
        for (int i = start; i < stop; i++) {
            BB_ROLL(bases[i]);
            if (len >= p.minlen2 && i >= p.minlen) {
                const int id = get_value(c, kmer, rkmer, p.kmask, i, p.k, p.qHammingDistance);
                if (id > 0) {
                    uint64_t kmer2 = 0, rkmer2 = 0;
                    int len2 = 0, cnt = 0;
                    bool seen_before = false;
                    for (int i2 = start; i2 < stop; i2++) {
                        {
                            const uint32_t c_ = bases[i2];
                            const bool def_ = bb_defined(c_);
                            const uint64_t x_ = def_ ? bb_code_raw(c_) : 0u, x2_ = def_ ? (3u - bb_code_raw(c_)) : 0u;
                            kmer2 = ((kmer2 << 2) | x_) & p.mask;
                            rkmer2 = ((rkmer2 >> 2) | (x2_ << p.shift2)) & p.mask;
                            if (p.forbidNs && !def_) {
                                len2 = 0;
                                rkmer2 = 0;
                            } else
                                len2++;
                        }
                        if (len2 >= p.minlen2 && i2 >= p.minlen) {
                            const int id2 = get_value(c, kmer2, rkmer2, p.kmask, i2, p.k, p.qHammingDistance);
                            if (id2 == id) {
                                if (i2 < i) seen_before = true;
                                cnt++;
                            }
                        }
                    }
                    if (!seen_before && cnt > best_cnt) {
                        best_cnt = cnt;
                        best_id = id;
                    }
                }
            }
        }
    }
    credit(c, r, best_id, blen);
    return best_id;
}

// jgi/BBDuk.java:3596-3677 countSetKmersBig (k>31 emulated by runs of consecutive 31-mer hits)
__device__ int count_set_kmers_big(const Ctx &c, RState &r, int maxBadKmers) {
    const BBParams &p = c.p;
    const int blen = rlen(r);
    if (blen < p.kbig || c.t.stored < 1 || skipped(p, r)) return 0;
    const int sub = p.kbig - p.k - 1;
    const uint8_t *bases = r.b + r.lo;
    uint64_t kmer = 0, rkmer = 0;
    int found = 0, len = 0, bkStart = -1, bkStop = -1, lastId = -1;
    BB_SCAN_WINDOW();
    for (int i = start; i < stop; i++) {
        BB_ROLL(bases[i]);
        if (len >= p.minlen2 && i >= p.minlen) {
            const int id = get_value(c, kmer, rkmer, p.kmask, i, p.k, p.qHammingDistance);
            if (id > 0) {
                lastId = id;
                if (bkStart == -1) bkStart = i;
                bkStop = i;
            } else if (bkStart > -1) {
                const int dif = bkStop - bkStart - sub;
                bkStop = bkStart = -1;
                if (dif > 0) {
                    const int old = found;
                    found += dif;
                    if (found > maxBadKmers && old <= maxBadKmers) {
                        credit(c, r, lastId, blen);
                        return found;
                    }
                }
            }
        }
    }
    if (bkStart > -1) {
        const int dif = bkStop - bkStart - sub;
        if (dif > 0) {
            const int old = found;
            found += dif;
            if (found > maxBadKmers && old <= maxBadKmers) credit(c, r, lastId, blen);
        }
    }
    return found;
}

// jgi/BBDuk.java:3866-4013 ktrim and :3708-3858 ktrimTip (same body; left/right select the tails)
__device__ int ktrim_body(const Ctx &c, RState &r, int start, int stop, bool right, bool left) {
    const BBParams &p = c.p;
    const int k = p.k;
    const int blen = rlen(r);
    if (blen < max(1, (p.useShortKmers ? min(k, p.mink) : k)) || c.t.stored < 1 || skipped(p, r)) return 0;
    const uint8_t *bases = r.b + r.lo;
    uint64_t kmer = 0, rkmer = 0;
    int found = 0, len = 0, id0 = -1;
    int minLoc = 999999999, minLocExclusive = 999999999, maxLoc = -1, maxLocExclusive = -1;
    for (int i = start; i < stop; i++) {
        BB_ROLL(bases[i]);
        if (len >= p.minlen2 && i >= p.minlen) {
            const int id = get_value(c, kmer, rkmer, p.kmask, i, k, p.qHammingDistance);
            if (id > 0) {
                if (id0 < 0) id0 = id;
                minLoc = min(minLoc, i - k + 1);
                maxLoc = i;
                found++;
            }
        }
    }
    if (minLoc != minLocExclusive) minLocExclusive = minLoc + k;
    if (maxLoc != maxLocExclusive) maxLocExclusive = maxLoc - k;
    if (p.useShortKmers && found == 0) {
        if (left) {  // :3910-3942 prefixes of the read, growing
            kmer = 0;
            rkmer = 0;
            len = 0;
            const int lim = min(k, stop);
            for (int i = start; i < lim; i++) {
                const uint32_t ch = bases[i];
                kmer = ((kmer << 2) | bb_code0(ch)) & p.mask;
                rkmer = rkmer | ((uint64_t)bb_comp0(ch) << (2 * len));
                len++;
                if (len >= p.mink) {
                    const int id = get_value(c, kmer, rkmer, 1ull << (2 * len), i, len, p.qHammingDistance2);
                    if (id > 0) {
                        if (id0 < 0) id0 = id;
                        minLoc = 0;
                        minLocExclusive = min(minLocExclusive, i + 1);
                        maxLoc = max(maxLoc, i);
                        maxLocExclusive = max(maxLocExclusive, 0);
                        found++;
                    }
                }
            }
        }
        if (right) {  // :3945-3975 suffixes of the read, growing leftwards
            kmer = 0;
            rkmer = 0;
            len = 0;
            const int lim = max(-1, stop - k);
            for (int i = stop - 1; i > lim; i--) {
                const uint32_t ch = bases[i];
                kmer = kmer | ((uint64_t)bb_code0(ch) << (2 * len));
                rkmer = ((rkmer << 2) | bb_comp0(ch)) & p.mask;
                len++;
                if (len >= p.mink) {
                    const int id = get_value(c, kmer, rkmer, 1ull << (2 * len), i, len, p.qHammingDistance2);
                    if (id > 0) {
                        if (id0 < 0) id0 = id;
                        minLoc = i;
                        minLocExclusive = min(minLocExclusive, blen);
                        maxLoc = blen - 1;
                        maxLocExclusive = max(maxLocExclusive, i - 1);
                        found++;
                    }
                }
            }
        }
    }
    if (found == 0) return 0;
    credit(c, r, id0, blen);
    if (p.trimPad != 0) {  // Tools.mid(0, x, len), shared/Tools.java:5158
        auto mid3 = [](int x, int y, int z) { return max(min(x, y), min(max(x, y), z)); };
        maxLoc = mid3(0, maxLoc + p.trimPad, blen);
        minLoc = mid3(0, minLoc - p.trimPad, blen);
        maxLocExclusive = mid3(0, maxLocExclusive + p.trimPad, blen);
        minLocExclusive = mid3(0, minLocExclusive - p.trimPad, blen);
    }
    if (left) return trim_to_position(r, p.ktrimExclusive ? maxLocExclusive + 1 : maxLoc + 1, blen - 1, 1);
    return trim_to_position(r, 0, p.ktrimExclusive ? minLocExclusive - 1 : minLoc - 1, 1);
}

// jgi/BBDuk.java:3679-3684
__device__ int ktrim(const Ctx &c, RState &r) {
    const BBParams &p = c.p;
    const int blen = rlen(r);
    BB_SCAN_WINDOW();
    return ktrim_body(c, r, start, stop, p.ktrimRight, p.ktrimLeft);
}
// jgi/BBDuk.java:3686-3699
__device__ int ktrim_tips(const Ctx &c, RState &r) {
    const BBParams &p = c.p;
    const int len = rlen(r);
    const int mid = len / 2 - (p.k - 1) / 2;
    int sum = 0;
    if (p.ktrimRight) sum += ktrim_body(c, r, max(0, (p.restrictRight < 1 ? mid : len - p.restrictRight)), len, true, false);
    if (p.ktrimLeft) sum += ktrim_body(c, r, 0, min(rlen(r), (p.restrictLeft < 1 ? mid + p.k - 1 : p.restrictLeft)), false, true);
    return sum;
}

// BitSet.set/clear(from,to) on the read's own mask words, bits clipped to [0, nbits)
__device__ __forceinline__ void bits_range(uint32_t *w, int from, int to, int nbits, bool val) {
    from = max(from, 0);
    to = min(to, nbits);
    if (from >= to) return;
    const int w0 = from >> 5, w1 = (to - 1) >> 5;
    for (int i = w0; i <= w1; i++) {
        uint32_t m = 0xFFFFFFFFu;
        if (i == w0) m &= 0xFFFFFFFFu << (from & 31);
        if (i == w1) m &= 0xFFFFFFFFu >> (31 - ((to - 1) & 31));
        if (val)
            w[i] |= m;
        else
            w[i] &= ~m;
    }
}

// jgi/BBDuk.java:4022-4199 kmask. The BitSet lives in the read's output mask words; bits at or
// beyond the read length (possible only with trimpad>0, from full-length hits) are tracked as a
// count because they only contribute to the returned cardinality.
__device__ int kmask_read(const Ctx &c, RState &r) {
    const BBParams &p = c.p;
    const int k = p.k;
    const int blen = rlen(r);
    if (blen < max(1, (p.useShortKmers ? min(k, p.mink) : k)) || c.t.stored < 1 || skipped(p, r)) return 0;
    if (blen < k) return 0;
    const uint8_t *bases = r.b + r.lo;
    uint32_t *bs = r.mask;
    uint64_t kmer = 0, rkmer = 0;
    int found = 0, len = 0, id0 = -1, maxEnd = 0;
    const bool mfc = p.kmaskFullyCovered;
    if (mfc) bits_range(bs, 0, blen, blen, true);
    const int minus = k - 1 - p.trimPad, plus = p.trimPad + 1;
    BB_SCAN_WINDOW();
    for (int i = start; i < stop; i++) {
        BB_ROLL(bases[i]);
        if (i >= p.minlen) {
            const int id = (len >= p.minlen2) ? get_value(c, kmer, rkmer, p.kmask, i, k, p.qHammingDistance) : -1;
            if (id > 0) {
                if (id0 < 0) id0 = id;
                if (!mfc) {
                    bits_range(bs, max(0, i - minus), i + plus, blen, true);
                    maxEnd = max(maxEnd, i + plus);
                }
                found++;
            } else if (mfc) {
                bits_range(bs, max(0, i - minus), i + plus, blen, false);
            }
        }
    }
    if (p.useShortKmers) {
        {  // :4081-4121 left side
            kmer = 0;
            rkmer = 0;
            len = 0;
            int len2 = 0;
            const int lim = min(k, stop);
            for (int i = start; i < lim; i++) {
                const uint32_t ch = bases[i];
                kmer = ((kmer << 2) | bb_code0(ch)) & p.mask;
                rkmer = rkmer | ((uint64_t)bb_comp0(ch) << (2 * len));
                len++;
                len2++;
                if (len2 >= p.minminlen) {
                    const int id = (len >= p.mink) ? get_value(c, kmer, rkmer, 1ull << (2 * len), i, len, p.qHammingDistance2) : -1;
                    if (id > 0) {
                        if (id0 < 0) id0 = id;
                        if (!mfc) bits_range(bs, 0, min(blen, i + p.trimPad + 1), blen, true);
                        found++;
                    } else if (mfc) {
                        bits_range(bs, 0, min(blen, i + p.trimPad + 1), blen, false);
                    }
                }
            }
        }
        {  // :4124-4164 right side
            kmer = 0;
            rkmer = 0;
            len = 0;
            int len2 = 0;
            const int lim = max(-1, stop - k);
            for (int i = stop - 1; i > lim; i--) {
                const uint32_t ch = bases[i];
                kmer = kmer | ((uint64_t)bb_code0(ch) << (2 * len));
                rkmer = ((rkmer << 2) | bb_comp0(ch)) & p.mask;
                len++;
                len2++;
                if (len2 >= p.minminlen) {
                    const int id = (len >= p.mink) ? get_value(c, kmer, rkmer, 1ull << (2 * len), i, len, p.qHammingDistance2) : -1;
                    if (id > 0) {
                        if (id0 < 0) id0 = id;
                        if (!mfc) bits_range(bs, max(0, i - p.trimPad), blen, blen, true);
                        found++;
                    } else if (mfc) {
                        bits_range(bs, max(0, i - p.trimPad), blen, blen, false);
                    }
                }
            }
        }
    }
    const int nw = (blen + 31) >> 5;
    if (found == 0) {
        for (int i = 0; i < nw; i++) bs[i] = 0;  // the read is left untouched (:4170)
        return 0;
    }
    credit(c, r, id0, blen);
    int card = max(0, maxEnd - blen);
    for (int i = 0; i < nw; i++) card += __popc(bs[i]);
    return card;
}

// jgi/BBDuk.java:4208-4377 ksplit; returns true when the read is split in two
__device__ bool ksplit_read(const Ctx &c, RState &r) {
    const BBParams &p = c.p;
    const int k = p.k;
    const int blen = rlen(r);
    if (blen < max(1, (p.useShortKmers ? min(k, p.mink) : k)) || c.t.stored < 1) return false;
    if (blen < k) return false;
    const uint8_t *bases = r.b + r.lo;
    uint64_t kmer = 0, rkmer = 0;
    int found = 0, len = 0, id0 = -1;
    int leftmost = 0x7FFFFFFF, rightmost = -1;
    const int minus = k - 1 - p.trimPad, plus = p.trimPad;
    BB_SCAN_WINDOW();
    for (int i = start; i < stop; i++) {
        BB_ROLL(bases[i]);
        if (i >= p.minlen) {
            const int id = (len >= p.minlen2) ? get_value(c, kmer, rkmer, p.kmask, i, k, p.qHammingDistance) : -1;
            if (id > 0) {
                if (id0 < 0) id0 = id;
                leftmost = min(leftmost, max(0, i - minus));
                rightmost = max(rightmost, i + plus);
                found++;
            }
        }
    }
    if (p.useShortKmers && id0 == -1) {
        {  // right side first (:4264-4303)
            kmer = 0;
            rkmer = 0;
            len = 0;
            int len2 = 0;
            const int lim = max(-1, stop - k);
            for (int i = stop - 1; i > lim; i--) {
                const uint32_t ch = bases[i];
                kmer = kmer | ((uint64_t)bb_code0(ch) << (2 * len));
                rkmer = ((rkmer << 2) | bb_comp0(ch)) & p.mask;
                len++;
                len2++;
                if (len2 >= p.minminlen) {
                    const int id = (len >= p.mink) ? get_value(c, kmer, rkmer, 1ull << (2 * len), i, len, p.qHammingDistance2) : -1;
                    if (id > 0) {
                        if (id0 < 0) id0 = id;
                        leftmost = min(leftmost, max(0, i - p.trimPad));
                        rightmost = blen - 1;
                        found++;
                    }
                }
            }
        }
        if (id0 == -1) {  // then the left side (:4306-4345)
            kmer = 0;
            rkmer = 0;
            len = 0;
            int len2 = 0;
            const int lim = min(k, stop);
            for (int i = start; i < lim; i++) {
                const uint32_t ch = bases[i];
                kmer = ((kmer << 2) | bb_code0(ch)) & p.mask;
                rkmer = rkmer | ((uint64_t)bb_comp0(ch) << (2 * len));
                len++;
                len2++;
                if (len2 >= p.minminlen) {
                    const int id = (len >= p.mink) ? get_value(c, kmer, rkmer, 1ull << (2 * len), i, len, p.qHammingDistance2) : -1;
                    if (id > 0) {
                        if (id0 < 0) id0 = id;
                        leftmost = 0;
                        rightmost = max(rightmost, i + p.trimPad);
                        found++;
                    }
                }
            }
        }
    }
    if (found == 0) return false;
    credit(c, r, id0, blen);
    if (leftmost == 0) {
        trim_to_position(r, rightmost + 1, blen - 1, 1);
        return false;
    } else if (rightmost == blen - 1) {
        trim_to_position(r, 0, leftmost - 1, 1);
        return false;
    }
    // new mate = subRead(rightmost+1, blen-1), end-exclusive (stream/Read.java:3729-3731). With
    // trimpad>0 rightmost+1 can pass blen-1, where the reference throws; clamp to an empty mate.
    r.splitStart = r.lo + min(rightmost + 1, blen - 1);
    trim_to_position(r, 0, leftmost - 1, 1);
    return true;
}

// stream/Read.java:1673-1683
__device__ int num_valid_kmers(const RState &r, int k) {
    const uint8_t *bases = r.b + r.lo;
    int len = 0, counted = 0;
    for (int i = 0; i < rlen(r); i++) {
        if (bb_defined(bases[i]))
            len++;
        else
            len = 0;
        if (len >= k) counted++;
    }
    return counted;
}

// jgi/BBDuk.java:3260-3289
__device__ __forceinline__ void set_discarded(const BBParams &p, RState &r) {
    if (p.trimFailuresTo1bp) {
        if (rlen(r) > 1) trim_by_amount(r, 0, rlen(r) - 1, 1);
    } else
        r.discarded = 1;
}
__device__ __forceinline__ bool is_discarded(const BBParams &p, const RState *r) {
    if (!r) return false;
    if (r->discarded) return true;
    return p.trimFailuresTo1bp && rlen(*r) == 1;
}
__device__ __forceinline__ bool is_null_or_discarded(const BBParams &p, const RState *r) {
    if (!r) return true;
    return is_discarded(p, r);
}
__device__ __forceinline__ bool should_remove(const BBParams &p, const RState *r1, const RState *r2) {
    return (p.removePairsIfEitherBad && (is_discarded(p, r1) || is_discarded(p, r2))) ||
           (is_discarded(p, r1) && is_null_or_discarded(p, r2));
}

__device__ __forceinline__ void write_out(const bbduk_out &o, int64_t idx, const RState &r, bool removed, bool ktrimmed,
                                          bool tpe, bool split, int count) {
    if (o.id0) o.id0[idx] = r.credit0;
    if (o.id0b) o.id0b[idx] = r.credit1;
    if (o.lo) o.lo[idx] = r.lo;
    if (o.hi) o.hi[idx] = r.hi;
    if (o.count) o.count[idx] = split ? r.splitStart : count;
    if (o.flags) {
        uint8_t f = 0;
        if (r.discarded) f |= BBDUK_F_DISCARDED;
        if (removed) f |= BBDUK_F_REMOVED;
        if (ktrimmed) f |= BBDUK_F_KTRIMMED;
        if (tpe) f |= BBDUK_F_TPE;
        if (split) f |= BBDUK_F_SPLIT;
        o.flags[idx] = f;
    }
}

}  // namespace

// One thread per unit (a pair, or a single read). `units` optionally lists the unit indices to
// process (the fast kernel's hand-offs); null = all units [0, n_units).
__global__ void __launch_bounds__(128)
bbduk_generic_kernel(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ offsets, int64_t n_units_arg, int paired,
                     const int32_t *__restrict__ units, const unsigned int *__restrict__ n_units_dev, BBParams p, BBTable t,
                     bbduk_out out, bbduk_stats *stats, unsigned long long *scaf_reads, unsigned long long *scaf_bases) {
    // n_units_dev: the number of listed units lives on the device (written by the fast kernel), so the
    // host never has to synchronise to learn it; the grid then strides over the list.
    const int64_t n_units = n_units_dev ? (int64_t)*n_units_dev : n_units_arg;
    long long s_rk = 0, s_bk = 0, s_rf = 0, s_bf = 0, s_ro = 0, s_bo = 0, s_ri = 0, s_bi = 0;
    for (int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tid < n_units; tid += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = units ? (int64_t)units[tid] : tid;
        const Ctx c{p, t, scaf_reads, scaf_bases};
        const int per = paired ? 2 : 1;
        RState rr[2];
        for (int q = 0; q < per; q++) {
            const int64_t idx = u * per + q;
            RState &r = rr[q];
            const uint32_t o0 = offsets[idx], o1 = offsets[idx + 1];
            r.b = bases + o0;
            r.base_off = o0;
            r.lo = 0;
            r.hi = (int)(o1 - o0);
            r.pairnum = q;
            r.discarded = 0;
            r.credit0 = r.credit1 = -1;
            r.splitStart = -1;
            r.mask = (out.maskbits && out.mask_off) ? out.maskbits + out.mask_off[idx] : nullptr;
            if (r.mask) {
                const int nw = (r.hi + 31) >> 5;
                for (int i = 0; i < nw; i++) r.mask[i] = 0;
            }
        }
        RState *r1 = &rr[0], *r2 = paired ? &rr[1] : nullptr;
        const int initialLength1 = rlen(*r1), initialLength2 = r2 ? rlen(*r2) : 0;
        const int pairCount = per;
        // (int)Tools.max(initialLength*minLenFraction, minReadLength), float arithmetic (jgi/BBDuk.java:2592-2593)
        const int minlen1 = (int)fmaxf(__fmul_rn((float)initialLength1, p.minLenFraction), (float)p.minReadLength);
        const int minlen2 = (int)fmaxf(__fmul_rn((float)initialLength2, p.minLenFraction), (float)p.minReadLength);
        bool remove = false, split = false, tpe1 = false, tpe2 = false;
        int count1 = 0, count2 = 0;
        bool kt1 = false, kt2 = false;
        const bool trimming = t.stored > 0 && (p.mode == MODE_KTRIM || p.mode == MODE_KTRIM_TIPS || p.mode == MODE_KMASK || p.mode == MODE_KSPLIT);
        const bool filtering = t.stored > 0 && !trimming;
        if (trimming) {  // jgi/BBDuk.java:2728-2813
            int rlen1 = 0, rlen2 = 0, xsum = 0, rktsum = 0;
            if (p.mode == MODE_KSPLIT) {
                const int oldLen = rlen(*r1), oldHi = r1->hi;
                split = ksplit_read(c, *r1);
                const int newPair = rlen(*r1) + (split ? (oldHi - 1 - r1->splitStart) : 0);
                const int trimmed = oldLen - newPair;
                xsum += trimmed;
                rktsum += (trimmed > 0 ? 1 : 0);
                kt1 = trimmed > 0;
                rlen1 = rlen(*r1);
            } else {
                for (int q = 0; q < per; q++) {
                    RState &r = rr[q];
                    const int x = (p.mode == MODE_KTRIM_TIPS) ? ktrim_tips(c, r) : (p.mode == MODE_KTRIM) ? ktrim(c, r) : kmask_read(c, r);
                    xsum += x;
                    rktsum += (x > 0 ? 1 : 0);
                    if (q == 0) {
                        count1 = x;
                        kt1 = x > 0;
                        rlen1 = rlen(r);
                        if (rlen1 < minlen1) set_discarded(p, r);
                    } else {
                        count2 = x;
                        kt2 = x > 0;
                        rlen2 = rlen(r);
                        if (rlen2 < minlen2) set_discarded(p, r);
                    }
                }
            }
            if (p.mode == MODE_KSPLIT) {
                remove = split;
            } else if (should_remove(p, r1, r2)) {
                if (p.mode != MODE_KMASK) {
                    xsum += (rlen1 + rlen2);
                    rktsum = pairCount;
                }
                remove = true;
            } else if (p.ktrimRight && p.trimPairsEvenly && xsum > 0 && r2 && rlen(*r1) != rlen(*r2)) {
                int x;
                if (rlen(*r1) > rlen(*r2)) {
                    x = trim_to_position(*r1, 0, rlen(*r2) - 1, 1);
                    tpe1 = true;
                } else {
                    x = trim_to_position(*r2, 0, rlen(*r1) - 1, 1);
                    tpe2 = true;
                }
                if (rktsum < 2) rktsum++;
                xsum += x;
            }
            s_bk += xsum;
            s_rk += rktsum;
        } else if (filtering) {  // jgi/BBDuk.java:2815-2873
            if (p.mode == MODE_KCOVER) {
                for (int q = 0; q < per; q++) {
                    RState &r = rr[q];
                    if (!is_discarded(p, &r)) {
                        const int minCov = (int)ceil((double)__fmul_rn(p.minCoveredFraction, (float)rlen(r)));
                        const int covered = count_covered_bases(c, r, minCov);
                        (q == 0 ? count1 : count2) = covered;
                        if (covered >= minCov) set_discarded(p, r);
                    }
                }
            } else {
                int mb1, mb2;
                if (p.minKmerFraction == 0) {
                    mb1 = mb2 = p.maxBadKmers0;
                } else {
                    const int vk1 = num_valid_kmers(*r1, p.keff), vk2 = r2 ? num_valid_kmers(*r2, p.keff) : 0;
                    mb1 = max(p.maxBadKmers0, (int)__fmul_rn((float)(vk1 - 1), p.minKmerFraction));
                    mb2 = max(p.maxBadKmers0, (int)__fmul_rn((float)(vk2 - 1), p.minKmerFraction));
                }
                if (p.mode == MODE_KBEST) {
                    count1 = find_best_match(c, *r1, mb1);
                    count2 = r2 ? find_best_match(c, *r2, mb2) : -1;
                    if (count1 > 0) set_discarded(p, *r1);
                    if (r2 && count2 > 0) set_discarded(p, *r2);
                } else {
                    const bool big = p.mode == MODE_KFILTER_BIG;
                    count1 = big ? count_set_kmers_big(c, *r1, mb1) : count_set_kmers(c, *r1, mb1);
                    count2 = r2 ? (big ? count_set_kmers_big(c, *r2, mb2) : count_set_kmers(c, *r2, mb2)) : 0;
                    if (count1 > mb1) set_discarded(p, *r1);
                    if (r2 && count2 > mb2) set_discarded(p, *r2);
                }
            }
            if (should_remove(p, r1, r2)) {
                remove = true;
                s_rf += per;
                s_bf += initialLength1 + initialLength2;
            }
        }
        write_out(out, u * per, *r1, remove, kt1, tpe1, split, count1);
        if (r2) write_out(out, u * per + 1, *r2, remove, kt2, tpe2, false, count2);
        s_ri += per;
        s_bi += initialLength1 + initialLength2;
        if (!remove) {
            s_ro += per;
            s_bo += rlen(*r1) + (r2 ? rlen(*r2) : 0);
        }
    }
    if (stats) {  // warp-level sums, one atomic per warp and counter
        long long v[8] = {s_ri, s_bi, s_rk, s_bk, s_rf, s_bf, s_ro, s_bo};
#pragma unroll
        for (int q = 0; q < 8; q++) {
            long long x = v[q];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
            if ((threadIdx.x & 31) == 0 && x) atomicAdd((unsigned long long *)stats + q, (unsigned long long)x);
        }
    }
}

int launch_generic(const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_units, int paired, const int32_t *d_units,
                   const unsigned int *d_n_units, const BBParams &p, const BBTable &t, const bbduk_out &out,
                   bbduk_stats *d_stats, unsigned long long *scaf_reads, unsigned long long *scaf_bases, int sm_count,
                   cudaStream_t st) {
    if (n_units <= 0) return 0;
    // with a device-side count the real number of units is unknown here (usually 0): a small grid strides
    const int64_t nb = d_n_units ? std::min<int64_t>((n_units + 127) / 128, 8 * (int64_t)sm_count) : (n_units + 127) / 128;
    bbduk_generic_kernel<<<(unsigned)nb, 128, 0, st>>>(d_bases, d_offsets, n_units, paired, d_units, d_n_units, p, t, out,
                                                       d_stats, scaf_reads, scaf_bases);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
