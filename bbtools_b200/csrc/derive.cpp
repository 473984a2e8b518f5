// derive.cpp -- user flags -> derived constants, in the order the reference's constructor applies them
// (jgi/BBDuk.java:583-585 inherit *2 distances; :672-710 finals + K; :764-780 kbig clamps;
//  :787-877 masks, minlen2, usk-disables-maskmiddle, mode selection, middleMask).
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "../../include/bbduk_b200.h"
#include "params.h"

static int fail(char *err, int errlen, const char *msg) {
    if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", msg);
    return 1;
}

int derive_params(const bbduk_cfg *c, BBParams *p, char *err, int errlen) {
    if (!c || c->struct_size != (int32_t)sizeof(bbduk_cfg)) return fail(err, errlen, "bbduk_cfg.struct_size mismatch");
    memset(p, 0, sizeof *p);

    // the reference asserts these ranges at parse time (jgi/BBDuk.java:254, :261, :264, :267-270, :292)
    if (c->hdist < 0 || c->hdist > 3 || c->qhdist < 0 || c->qhdist > 3)
        return fail(err, errlen, "hamming distance must be between 0 and 3; default is 0.");
    if (c->edist < 0 || c->edist > 2) return fail(err, errlen, "edit distance must be between 0 and 2; default is 0.");
    if (c->speed < 0 || c->speed > 16) return fail(err, errlen, "Speed range is 0 to 16.");
    if (!(c->mink < 0 || (c->mink > 0 && c->mink < 32))) return fail(err, errlen, "kmin must be between 1 and 31");
    if (c->qskip < 1) return fail(err, errlen, "qskip must be at least 1");

    int hd = c->hdist, ed = c->edist, qhd = c->qhdist;
    int hd2 = (c->hdist2 == -1 ? hd : c->hdist2);
    int qhd2 = (c->qhdist2 == -1 ? qhd : c->qhdist2);
    int ed2 = (c->edist2 == -1 ? ed : c->edist2);
    if (hd2 < 0 || hd2 > 3 || qhd2 < 0 || qhd2 > 3 || ed2 < 0 || ed2 > 2)
        return fail(err, errlen, "hdist2/qhdist2/edist2 out of range");
    hd = std::max(ed, hd);
    hd2 = std::max(ed2, hd2);
    p->hammingDistance = hd;
    p->hammingDistance2 = hd2;
    p->editDistance = ed;
    p->editDistance2 = ed2;
    p->qHammingDistance = qhd;
    p->qHammingDistance2 = qhd2;
    p->minSkip = std::max(1, std::min(c->min_skip, c->max_skip));
    p->maxSkip = std::max(p->minSkip, c->max_skip);
    p->forbidNs = (c->forbid_ns || hd < 1) ? 1 : 0;
    p->restrictLeft = std::max(c->restrict_left, 0);
    p->restrictRight = std::max(c->restrict_right, 0);
    p->speed = c->speed;
    p->speedMask2 = (c->generation == BBDUK_GEN_S) ? 1 : 0;
    p->qSkip = c->qskip;
    p->skipR1 = c->skip_r1 != 0;
    p->skipR2 = c->skip_r2 != 0;
    p->rcomp = c->rcomp != 0;
    p->trimPad = c->trim_pad;
    p->ktrimExclusive = c->ktrim_exclusive != 0;
    p->kmaskFullyCovered = c->kmask_fully_covered != 0;
    p->trimFailuresTo1bp = c->trim_failures_to_1bp != 0;
    p->removePairsIfEitherBad = (!c->require_both_bad && !c->trim_failures_to_1bp) ? 1 : 0;
    p->trimPairsEvenly = c->trim_pairs_evenly != 0;
    p->minReadLength = c->min_read_length;
    p->minLenFraction = c->min_len_fraction;
    p->maxBadKmers0 = c->max_bad_kmers;
    p->minKmerFraction = std::max(c->min_kmer_fraction, 0.0f);
    p->minCoveredFraction = std::max(c->min_covered_fraction, 0.0f);
    if (p->minKmerFraction > 1.0f) return fail(err, errlen, "minKmerFraction must range from 0 to 1");
    if (p->minCoveredFraction > 1.0f) return fail(err, errlen, "minCoveredFraction must range from 0 to 1");

    const bool trimmode = c->ktrim_left || c->ktrim_right || c->ktrim_n || c->ksplit;
    const int maxSupportedK = 31;
    int k = c->k > 0 ? c->k : 27;
    int kbig = (k > maxSupportedK ? k : -1);
    k = std::min(k, maxSupportedK);
    if (trimmode && kbig > k) kbig = k;
    if ((c->speed > 0 || c->qskip > 1) && kbig > k) kbig = k;
    p->k = k;
    p->k2 = k - 1;
    p->kbig = kbig;
    p->keff = std::max(k, kbig);

    bool mm = c->mask_middle != 0;
    int mml = c->mid_mask_len;
    if (mml > 0) mm = true;
    mml = mm ? (mml > 0 ? mml : 2 - (k & 1)) : 0;
    if (kbig > k) {
        p->minSkip = p->maxSkip = 0;
        if (mm) {
            mm = false;
            mml = 0;
        }
    }
    p->mink = (c->generation == BBDUK_GEN_JGI) ? std::min((c->mink < 1 ? 6 : c->mink), k) : std::min(c->mink, k);

    p->minlen = k - 1;
    p->minminlen = p->mink - 1;
    p->minlen2 = mm ? (k - mml) / 2 : k;  // computed BEFORE useShortKmers switches maskMiddle off (:836 vs :849-856)
    if (c->minlen2 > 0) {  // a host that marshals already-derived parser fields (java/bbduk/BBDukIndexGPU.java) hands in its own
        if (c->minlen2 > k) return fail(err, errlen, "cfg.minlen2 must not exceed k");
        p->minlen2 = c->minlen2;
    }
    p->shift2 = 2 * k - 2;
    p->mask = (2 * k > 63) ? ~0ull : ~((~0ull) << (2 * k));
    p->kmask = 1ull << (2 * k);

    bool usk = c->use_short_kmers != 0;
    if (c->mink > 0 && c->mink < k) usk = true;
    if (usk && mm) {
        mm = false;
        mml = 0;
    }
    p->useShortKmers = usk;
    p->maskMiddle = mm;
    p->midMaskLen = mml;
    if (usk && !trimmode)
        return fail(err, errlen,
                    "Setting mink or useShortKmers also requires setting a ktrim mode, such as 'r', 'l', or 'n'");
    if (usk && p->mink < 1) return fail(err, errlen, "useShortKmers needs mink>=1");

    p->ktrimLeft = c->ktrim_left != 0;
    p->ktrimRight = c->ktrim_right != 0;
    if (c->ksplit)
        p->mode = MODE_KSPLIT;  // order of the dispatch in the per-pair block: tips, l|r, n, split (:2734-2789)
    if (c->ktrim_n) p->mode = MODE_KMASK;
    if (c->ktrim_left || c->ktrim_right) p->mode = (c->ktrim_left && c->ktrim_right) ? MODE_KTRIM_TIPS : MODE_KTRIM;
    if (!trimmode) {
        if (p->minCoveredFraction > 0)
            p->mode = MODE_KCOVER;
        else if (c->find_best_match) {
            if (kbig > k) return fail(err, errlen, "K must be less than 32 in 'findBestMatch' mode");
            p->mode = MODE_KBEST;
        } else
            p->mode = (kbig > k) ? MODE_KFILTER_BIG : MODE_KFILTER;
    }
    if ((p->mode == MODE_KMASK || p->mode == MODE_KSPLIT) && p->trimPad < 0)
        return fail(err, errlen, "kmask/ksplit need trimpad>=0 (the reference's BitSet ranges throw otherwise)");

    if (mm) {
        if (!(k > mml + 1)) return fail(err, errlen, "k must be greater than midMaskLen+1");
        const int bits = mml * 2;
        const int shift = ((k - mml) / 2) * 2;
        p->middleMask = ~((~((~0ull) << bits)) << shift);
    } else {
        p->middleMask = ~0ull;
    }
    return 0;
}
