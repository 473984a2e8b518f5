// qtrim.cu -- BBDuk's poly-X trimming, quality-trimming block and the per-read quality / length / N filters on the device:
// bbduk_b200_qtrim / bbduk_b200_qtrim_device (first part).
//
// Replaces jgi/BBDuk.java:2954-3052 (poly-A, poly-G, poly-C: trimPolyA / trimPoly / detectPolyLeft / detectPolyRight
// :4721-4825, each with its minlen test and shouldRemove) and :3074-3170: TrimRead.trimFast in its default "optimal" mode (shared/TrimRead.java:140-169,
// :348-410: the maximum-sum run of avgErrorRate - probError in single precision, ties to the longer run, then
// trimByAmount(r, a, b, 1) :299-346), the minlen / maxlen test with shouldRemove, then minbasequality and maxns with
// their shouldRemove (:3260-3289 for setDiscarded / isDiscarded / shouldRemove).
//
// One lane per read, mates on neighbouring lanes (pair decisions by lane shuffles, as in stage D of probe_fast.cu).
// The running score is a chain of dependent single-precision adds in the reference's order, so a read is scanned by
// one lane; what is vectorised is the memory side: 16 quality bytes + 16 bases per load, 'N' bases folded into the
// quality word with byte-lane logic, and ONE shared-memory lookup per base: D[raw quality byte] = trimE - probError
// (the table is built on the host with the same libm calls as the tables of tbo.cu and handed over as a kernel
// parameter; undefined qualities and N map to trimE - nprob). The kernel is bound by HBM when the reads are long
// enough for the 16-byte loads to dominate: 2 bytes per base in, 9 bytes per read in/out.
#include <algorithm>
#include <cmath>
#include <cstring>

#include <cuda_runtime.h>

#include "../../include/bbduk_b200.h"
#include "probe.h"

namespace {

constexpr int QT_THREADS = 256;
constexpr int QS_BYTES = 32 * 160 + 32;  // staged quality bytes per warp (static shared memory: 8 warps -> 41 KB per block)

struct QtrimDev {
    int qtrim_left, qtrim_right, mbq, max_ns, max_len, qual_offset;
    int minReadLength;
    float minLenFraction;
    int rieb, tf1;
    int poly_a, poly_g_left, poly_g_right, filter_g, poly_c_left, poly_c_right, filter_c, max_non_poly;
    int maq_on, maq_bases;
    float max_n_rate, min_base_freq;  // maxnrate (>= 1 = off), minbasefrequency (0 = off)
    int mcb;                          // minconsecutivebases (0 = off)
    int alt_trim;                     // quality trimming by another rule than the staged optimal scan: 1 = window, 2 = optitrim=f,
                                      // 3 = optimal without qualities (N's only); 0 = none
    int trimq_byte, window, good_interval, n_only_off;  // (byte)trimq; n_only_off: the no-quality rule trims nothing
    float maq_prob;    // discard iff expectedErrors / bases >= maq_prob  (<=> phred average < minavgquality, see launch_qtrim)
    float delta[256];  // per raw quality byte: trimE - probError (trimE - nprob for q < 1)
    float pe[256];     // per raw quality byte: PROB_ERROR[max(q, 0)] (only staged when maq is on)
};

// shared/TrimRead.java:299-346 on a kept interval
__device__ __forceinline__ int trim_amounts(int &lo, int &hi, int left, int right, int minLen) {
    left = max(left, 0);
    right = max(right, 0);
    const int len = hi - lo;
    if (len < 1) return 0;
    minLen = min(len, max(minLen, 0));
    if (left + right + minLen > len) {
        right = max(1, len - minLen);
        left = 0;
    }
    lo += left;
    hi -= right;
    return left + right;
}

__device__ __forceinline__ bool defined_base(uint8_t b) {
    const uint8_t y = b | 0x20;
    return b < 128 && (y == 'a' || y == 'c' || y == 'g' || y == 't' || y == 'u');
}

// stream/Read.java:3387-3401 on the kept interval [l,h) of the read whose first base is b[0]
__device__ __forceinline__ int count_left(const uint8_t *b, int l, int h, uint8_t c) {
    int i = l;
    while (i < h && b[i] == c) i++;
    return i - l;
}
__device__ __forceinline__ int count_right(const uint8_t *b, int l, int h, uint8_t c) {
    int i = h - 1;
    while (i >= l && b[i] == c) i--;
    return h - 1 - i;
}
// jgi/BBDuk.java:4771-4791
__device__ __forceinline__ int detect_poly_left(const uint8_t *b, int l, int h, int minPoly, int maxNonPoly, uint8_t c) {
    const int n = h - l;
    if (n < minPoly) return 0;
    int trimTo = -1;
    for (int i = 0, polymer = 0, nonpoly = 0; i < n && nonpoly <= maxNonPoly; i++) {
        if (b[l + i] == c) {
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
// jgi/BBDuk.java:4802-4822
__device__ __forceinline__ int detect_poly_right(const uint8_t *b, int l, int h, int minPoly, int maxNonPoly, uint8_t c) {
    const int n = h - l;
    if (n < minPoly) return 0;
    int trimTo = n;
    for (int i = n - 1, polymer = 0, nonpoly = 0; i >= 0 && nonpoly <= maxNonPoly; i--) {
        if (b[l + i] == c) {
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

// 0xFF in every byte lane of w that holds 'N'
__device__ __forceinline__ uint32_t n_lanes(uint32_t w) {
    const uint32_t y = w ^ 0x4E4E4E4Eu;
    const uint32_t z = ((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y;  // bit 7 of a lane <=> byte != 0
    return ((~z & 0x80808080u) >> 7) * 0xFFu;
}

template <bool QT, bool POLY>
__global__ void __launch_bounds__(QT_THREADS)
qtrim_kernel(const uint8_t *__restrict__ bases, const uint8_t *__restrict__ quals, const uint32_t *__restrict__ offsets,
             int64_t n_reads, int paired, int32_t *lo_io, int32_t *hi_io, uint8_t *flags_io, const QtrimDev p,
             unsigned long long *stats) {
    __shared__ float D[256];
    __shared__ float PEs[256];
    if (p.maq_on)
        for (int i = threadIdx.x; i < 256; i += QT_THREADS) PEs[i] = p.pe[i];
    __shared__ __align__(16) uint8_t Qs_all[QT ? (QT_THREADS / 32) * QS_BYTES : 16];
    uint8_t *Qs = Qs_all + (QT ? (threadIdx.x >> 5) * QS_BYTES : 0);
    for (int i = threadIdx.x; i < 256; i += QT_THREADS) D[i] = p.delta[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (QT_THREADS / 32);
    const int64_t n_tiles = (n_reads + 31) >> 5;
    const uint32_t off_word = (uint32_t)(p.qual_offset & 0xFF) * 0x01010101u;
    unsigned int s_rq = 0, s_bq = 0, s_rf = 0, s_bf = 0, s_rn = 0, s_bn = 0, s_rp = 0, s_bp = 0;
    for (int64_t tile = (int64_t)blockIdx.x * (QT_THREADS / 32) + (threadIdx.x >> 5); tile < n_tiles; tile += warps_total) {
        const int64_t r = tile * 32 + lane;
        const bool live = r < n_reads;
        const uint32_t o0 = live ? offsets[r] : 0u;
        const int L = live ? (int)(offsets[r + 1] - o0) : 0;
        int l = live ? lo_io[r] : 0, h = live ? hi_io[r] : 0;
        const int f = live ? (int)flags_io[r] : BBDUK_F_REMOVED;
        const int f_first = paired ? __shfl_sync(0xFFFFFFFFu, f, lane & ~1) : f;
        const bool removed = !live || (f_first & BBDUK_F_REMOVED) != 0;
        bool discarded = (f & BBDUK_F_DISCARDED) != 0;
        bool gone = removed;  // the unit has been removed (on input, or by an earlier step of this kernel)
        int x = 0;
        bool ptrim = false;
        const uint8_t *rb = bases + o0;
        const int minlenR = (int)fmaxf(__fmul_rn((float)L, p.minLenFraction), (float)p.minReadLength);
        const bool first = !paired || !(lane & 1);
        auto is_disc = [&]() { return discarded || (p.tf1 && h - l == 1); };
        auto set_disc = [&]() {  // jgi/BBDuk.java:3260-3266
            if (p.tf1) {
                if (h - l > 1) trim_amounts(l, h, 0, h - l - 1, 1);
            } else {
                discarded = true;
            }
        };
        // shouldRemove (jgi/BBDuk.java:3286-3289) for the unit, then basesPolyTrimmedT += r1.pairLength(); executed by all lanes
        auto poly_close = [&](int xp) {
            const bool d = is_disc();
            const bool dm = __shfl_xor_sync(0xFFFFFFFFu, (int)d, 1) != 0;
            const bool rem = !gone && (paired ? (p.rieb ? (d || dm) : (d && dm)) : d);
            const int len = h - l;
            const int lenm = __shfl_xor_sync(0xFFFFFFFFu, len, 1);
            if (rem && first) s_bp += (unsigned int)(len + (paired ? lenm : 0));
            gone = gone || rem;
            ptrim = ptrim || xp > 0;
        };
        if (POLY && p.poly_a > 0) {  // :2954-2979
            int xp = 0;
            if (!gone) {
                if (h - l >= p.poly_a) {  // trimPolyA :4721-4736
                    int left = max(count_left(rb, l, h, 'A'), count_left(rb, l, h, 'T'));
                    int right = max(count_right(rb, l, h, 'A'), count_right(rb, l, h, 'T'));
                    if (left < p.poly_a) left = 0;
                    if (right < p.poly_a) right = 0;
                    if (left > 0 || right > 0) xp = trim_amounts(l, h, left, right, 1);
                }
                s_bp += (unsigned int)xp;
                s_rp += xp > 0;
                if (h - l < minlenR) set_disc();
            }
            poly_close(xp);
        }
#pragma unroll 1
        for (int which = 0; POLY && which < 2; which++) {  // :2981-3016 poly-G, :3018-3052 poly-C
            const uint8_t c = which == 0 ? 'G' : 'C';
            const int tl = which == 0 ? p.poly_g_left : p.poly_c_left, tr = which == 0 ? p.poly_g_right : p.poly_c_right;
            const int fp = which == 0 ? p.filter_g : p.filter_c;
            if (!(tl > 0 || tr > 0 || fp > 0)) continue;
            int xp = 0;
            // the poly-C filter of r2 looks at r1 as the r1 step of this block left it (:3035): second mates run after the first
#pragma unroll 1
            for (int turn = 0; turn < 2; turn++) {
                const uint32_t o0m = __shfl_xor_sync(0xFFFFFFFFu, o0, 1);
                const int lm = __shfl_xor_sync(0xFFFFFFFFu, l, 1), hm = __shfl_xor_sync(0xFFFFFFFFu, h, 1);
                const bool second = paired && (lane & 1);
                if (gone || (turn == 0) == second) continue;
                const bool probe_mate = which == 1 && second;
                const uint8_t *pb = probe_mate ? bases + o0m : rb;
                const int pl = probe_mate ? lm : l, ph = probe_mate ? hm : h;
                if (fp > 0 && detect_poly_left(pb, pl, ph, fp, p.max_non_poly, c) >= fp) {
                    set_disc();
                    s_rp += 1;
                } else if (tl > 0 || tr > 0) {  // trimPoly :4747-4760
                    const int left = tl > 0 ? detect_poly_left(rb, l, h, tl, p.max_non_poly, c) : 0;
                    const int right = tr > 0 ? detect_poly_right(rb, l, h, tr, p.max_non_poly, c) : 0;
                    if (left > 0 || right > 0) xp = trim_amounts(l, h, left, right, 1);
                    s_bp += (unsigned int)xp;
                    s_rp += xp > 0;
                    if (h - l < minlenR) set_disc();
                }
            }
            poly_close(xp);
        }
        // Stage the tile's quality bytes in shared memory: the 32 reads of a warp are contiguous in the batch, so the warp
        // reads them with coalesced 16-byte loads (every DRAM sector fetched once, nothing evicted from L1 in between) and
        // folds the 'N' bases in on the way; each lane then scans its own read out of shared memory. Tiles too long for the
        // buffer (reads beyond ~160 bases on average) scan global memory directly.
        bool staged = false;
        uint32_t a0t = 0;
        if (QT) {
            const int last_lane = (int)min((long long)31, (long long)(n_reads - 1 - tile * 32));
            const uint32_t t_lo = __shfl_sync(0xFFFFFFFFu, o0, 0);
            const uint32_t t_hi = __shfl_sync(0xFFFFFFFFu, o0 + (uint32_t)L, last_lane);
            a0t = t_lo & ~15u;
            const uint32_t nchunks = (t_hi - a0t + 15u) >> 4;
            staged = nchunks * 16u + 16u <= (uint32_t)QS_BYTES;
            if (staged) {
                uint4 *dst = reinterpret_cast<uint4 *>(Qs);
                for (uint32_t c = lane; c < nchunks; c += 32) {
                    const uint4 qv = __ldg(reinterpret_cast<const uint4 *>(quals + a0t) + c);
                    const uint4 bv = __ldg(reinterpret_cast<const uint4 *>(bases + a0t) + c);
                    uint4 o;
                    uint32_t nm = n_lanes(bv.x);
                    o.x = (qv.x & ~nm) | (off_word & nm);  // an N base reads as quality 0
                    nm = n_lanes(bv.y);
                    o.y = (qv.y & ~nm) | (off_word & nm);
                    nm = n_lanes(bv.z);
                    o.z = (qv.z & ~nm) | (off_word & nm);
                    nm = n_lanes(bv.w);
                    o.w = (qv.w & ~nm) | (off_word & nm);
                    dst[c] = o;
                }
            }
            __syncwarp();
        }
        if (!QT && p.alt_trim && !gone && h - l >= 1) {
            // the other rules of TrimRead.trimFast (shared/TrimRead.java:153-171): rare modes, plain per-lane loops
            const int n = h - l;
            const uint8_t *bb = bases + o0 + l;
            const uint8_t *qq = quals ? quals + o0 + l : nullptr;
            auto qv = [&](int i) { return (int)(int8_t)(uint8_t)(qq[i] - p.qual_offset); };
            auto left_n = [&]() {  // testLeftN :477-489
                int good = 0, lastBad = -1;
                for (int i = 0; i < n && good < p.good_interval; i++) {
                    if (bb[i] != 'N') good++;
                    else { good = 0; lastBad = i; }
                }
                return lastBad + 1;
            };
            auto right_n = [&]() {  // testRightN :491-503
                int good = 0, lastBad = n;
                for (int i = n - 1; i >= 0 && good < p.good_interval; i--) {
                    if (bb[i] != 'N') good++;
                    else { good = 0; lastBad = i; }
                }
                return n - lastBad;
            };
            int a = 0, b = 0;
            if (p.alt_trim == 3) {  // testOptimal without qualities (:352): avgErrorRate >= 1 trims nothing
                if (!p.n_only_off) {
                    a = left_n();
                    b = right_n();
                }
                a = p.qtrim_left ? a : 0;
                b = p.qtrim_right ? b : 0;
            } else if (p.alt_trim == 1) {  // testRightWindow :438-455
                if (p.qtrim_right) {
                    if (!qq || n < p.window) {
                        b = p.trimq_byte > 0 ? 0 : right_n();
                    } else {
                        const int thresh = max(p.window * p.trimq_byte, 1);
                        int sum = 0;
                        for (int i = 0, j = -p.window; i < n; i++, j++) {
                            sum += qv(i);
                            if (j >= -1) {
                                if (j >= 0) sum -= qv(j);
                                if (sum < thresh) {
                                    b = n - j - 1;
                                    break;
                                }
                            }
                        }
                    }
                }
            } else {  // testLeft / testRight :416-475
                if (p.qtrim_left) {
                    if (!qq) {
                        a = p.trimq_byte < 0 ? 0 : left_n();
                    } else {
                        int good = 0, lastBad = -1;
                        for (int i = 0; i < n && good < p.good_interval; i++) {
                            if (qv(i) > p.trimq_byte) good++;
                            else { good = 0; lastBad = i; }
                        }
                        a = lastBad + 1;
                    }
                }
                if (p.qtrim_right) {
                    if (!qq) {
                        b = p.trimq_byte < 0 ? 0 : right_n();
                    } else {
                        int good = 0, lastBad = n;
                        for (int i = n - 1; i >= 0 && good < p.good_interval; i--) {
                            if (qv(i) > p.trimq_byte) good++;
                            else { good = 0; lastBad = i; }
                        }
                        b = n - lastBad;
                    }
                }
            }
            x = trim_amounts(l, h, a, b, 1);  // trimFast -> trimByAmount(r, a, b, 1)
        }
        if (QT && !gone && h - l >= 1) {
            // testOptimal (shared/TrimRead.java:348-410) over [l,h)
            // Branch-free: the reference's "score>maxScore || (score==maxScore && count>maxCount)" is ONE unsigned 64-bit
            // compare of (score bits : count) -- scores that reach it are positive floats, whose order is the order of their
            // bit patterns; a reset step carries the key (0 : 0), which never beats the initial (0 : 0xFFFFFFFF).
            float score = 0.0f;
            uint32_t count = 0, best_hi = 0, best_lo = 0xFFFFFFFFu;
            int maxLoc = -1, i = 0;
            auto step = [&](uint32_t qbyte) {
                const float s2 = __fadd_rn(score, D[qbyte]);
                const bool pos = s2 > 0.0f;
                score = pos ? s2 : 0.0f;
                count = pos ? count + 1u : 0u;
                const uint32_t k_hi = __float_as_uint(score);
                const bool better = k_hi > best_hi || (k_hi == best_hi && count > best_lo);
                best_hi = better ? k_hi : best_hi;
                best_lo = better ? count : best_lo;
                maxLoc = better ? i : maxLoc;
                i++;
            };
            if (staged) {
                const uint32_t sa = o0 + (uint32_t)l - a0t;  // byte offset of the read's first kept base in Qs
                const uint32_t *W = reinterpret_cast<const uint32_t *>(Qs) + (sa >> 2);
                const uint32_t sh = (sa & 3u) * 8u;
                const int ns = h - l;
                uint32_t cur = W[0];
                int g = 0;
                for (; 4 * g + 4 <= ns; g++) {
                    const uint32_t nxt = W[g + 1];
                    const uint32_t w = __funnelshift_r(cur, nxt, sh);
                    cur = nxt;
                    step(w & 0xFFu);
                    step((w >> 8) & 0xFFu);
                    step((w >> 16) & 0xFFu);
                    step(w >> 24);
                }
                uint32_t w = __funnelshift_r(cur, W[g + 1], sh);
                for (int t = 4 * g; t < ns; t++, w >>= 8) step(w & 0xFFu);
            } else {
            uint32_t pos = o0 + (uint32_t)l;
            const uint32_t end = o0 + (uint32_t)h;
            for (; pos < end && (pos & 15u); pos++) step(bases[pos] == 'N' ? (uint32_t)(p.qual_offset & 0xFF) : (uint32_t)quals[pos]);
            for (; pos + 16u <= end; pos += 16u) {
                const uint4 qv = __ldg(reinterpret_cast<const uint4 *>(quals + pos));
                const uint4 bv = __ldg(reinterpret_cast<const uint4 *>(bases + pos));
                const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t nm = n_lanes(bw[k]);
                    const uint32_t w = (qw[k] & ~nm) | (off_word & nm);  // an N base reads as quality 0
                    step(w & 0xFFu);
                    step((w >> 8) & 0xFFu);
                    step((w >> 16) & 0xFFu);
                    step(w >> 24);
                }
            }
            for (; pos < end; pos++) step(bases[pos] == 'N' ? (uint32_t)(p.qual_offset & 0xFF) : (uint32_t)quals[pos]);
            }
            const float maxScore = __uint_as_float(best_hi);
            const int maxCount = (int)best_lo;
            const int n = h - l;
            int a0 = 0, b0 = n;
            if (maxScore > 0.0f) {
                a0 = maxLoc - maxCount + 1;
                b0 = n - maxLoc - 1;
            }
            x = trim_amounts(l, h, p.qtrim_left ? a0 : 0, p.qtrim_right ? b0 : 0, 1);  // trimFast -> trimByAmount(r, a, b, 1)
        }
        // :3092-3099 minlen / maxlen
        if (!gone && !is_disc()) {
            const int len = h - l;
            if (len < minlenR || len > p.max_len) set_disc();
        }
        // :3102-3106 shouldRemove after quality trimming
        const bool d1 = is_disc();
        const bool d1m = __shfl_xor_sync(0xFFFFFFFFu, (int)d1, 1) != 0;
        const bool rem1 = !gone && (paired ? (p.rieb ? (d1 || d1m) : (d1 && d1m)) : d1);
        const int len1 = h - l;
        const int len1m = __shfl_xor_sync(0xFFFFFFFFu, len1, 1);
        const int xm = __shfl_xor_sync(0xFFFFFFFFu, x, 1);
        if (!gone && first) {
            s_bq += (unsigned int)(x + (paired ? xm : 0));
            s_rq += (x > 0) + ((paired && xm > 0) ? 1 : 0);
            if (rem1) s_bq += (unsigned int)(len1 + (paired ? len1m : 0));
        }
        // :3110-3148 minbasequality, maxns
        if (!gone && !rem1) {
            if (p.maq_on && quals) {  // Read.avgQuality(false, maxBases) < minAvgQuality (stream/Read.java:2181-2226, :2985-3001)
                const int n = h - l;
                bool low = true;  // an empty read averages 0
                if (n > 0) {
                    const int limit = p.maq_bases < 1 ? n : min(p.maq_bases, n);
                    float sum = 0.0f;
                    for (int i = 0; i < limit; i++)
                        if (defined_base(bases[o0 + l + i])) sum = __fadd_rn(sum, PEs[quals[o0 + l + i]]);
                    low = __fdiv_rn(sum, (float)limit) >= p.maq_prob;
                }
                if (low) set_disc();
            }
            if (p.mbq > 0 && quals) {
                int mn = 41;
                for (int i = l; i < h; i++) mn = min(mn, (int)(int8_t)(quals[o0 + i] - p.qual_offset));
                if (mn < p.mbq) set_disc();
            }
            if (p.max_ns >= 0) {
                int nu = 0;
                for (int i = l; i < h; i++) nu += defined_base(bases[o0 + i]) ? 0 : 1;
                if (nu > p.max_ns) {
                    s_rn += 1;
                    s_bn += (unsigned int)(h - l);
                    set_disc();
                }
            }
            // :3138-3149 the same as a fraction of the read length; r.discarded() is the raw flag (no double count after maxns)
            if (p.max_n_rate < 1.0f && !discarded) {
                int nu = 0;
                for (int i = l; i < h; i++) nu += defined_base(bases[o0 + i]) ? 0 : 1;
                if ((float)nu > __fmul_rn(p.max_n_rate, (float)(h - l))) {
                    s_rn += 1;
                    s_bn += (unsigned int)(h - l);
                    set_disc();
                }
            }
            // :3151-3154 Read.hasMinConsecutiveBases (stream/Read.java:2846-2858): some run of mcb defined bases
            if (p.mcb > 0 && !is_disc()) {
                int run = 0;
                bool ok = false;
                for (int i = l; i < h && !ok; i++) {
                    run = defined_base(bases[o0 + i]) ? run + 1 : 0;
                    ok = run >= p.mcb;
                }
                if (!ok) set_disc();
            }
            // :3156-3159 Read.minBaseCount (stream/Read.java:2864-2874): upper-case A, C, G, T only
            if (p.min_base_freq > 0.0f) {
                int na = 0, nc = 0, ng = 0, nt = 0;
                for (int i = l; i < h; i++) {
                    const uint8_t b = bases[o0 + i];
                    na += b == 'A';
                    nc += b == 'C';
                    ng += b == 'G';
                    nt += b == 'T';
                }
                if ((float)min(min(na, nc), min(ng, nt)) < __fmul_rn(p.min_base_freq, (float)(h - l))) set_disc();
            }
        }
        // :3162-3167 shouldRemove after quality filtering
        const bool d2 = is_disc();
        const bool d2m = __shfl_xor_sync(0xFFFFFFFFu, (int)d2, 1) != 0;
        const bool rem2 = !gone && !rem1 && (paired ? (p.rieb ? (d2 || d2m) : (d2 && d2m)) : d2);
        const int len2 = h - l;
        const int len2m = __shfl_xor_sync(0xFFFFFFFFu, len2, 1);
        if (rem2 && first) {
            s_bf += (unsigned int)(len2 + (paired ? len2m : 0));
            s_rf += paired ? 2 : 1;
        }
        if (live && !removed) {
            lo_io[r] = l;
            hi_io[r] = h;
            flags_io[r] = (uint8_t)((f & ~(BBDUK_F_DISCARDED | BBDUK_F_REMOVED)) | (discarded ? BBDUK_F_DISCARDED : 0) |
                                    ((gone || rem1 || rem2) ? BBDUK_F_REMOVED : 0) | (x > 0 ? BBDUK_F_QTRIMMED : 0) |
                                    (ptrim ? BBDUK_F_POLYTRIMMED : 0));
        }
    }
    if (stats) {
        const unsigned int v[8] = {s_rq, s_bq, s_rf, s_bf, s_rn, s_bn, s_rp, s_bp};
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const unsigned int t = __reduce_add_sync(0xFFFFFFFFu, v[q]);
            if (lane == 0 && t) atomicAdd(stats + q, (unsigned long long)t);
        }
    }
}

// align2/QualityTools.java:650-654
double phred_to_prob_error(double q) {
    if (q <= 0) return 0.75;
    if (q <= 1) return 0.75 - q * 0.05;
    return std::min(0.7, std::pow(10.0, -0.1 * q));
}

}  // namespace

// launcher used by abi.cu; returns 0 on success, 1 on CUDA failure
int launch_qtrim(int sm_count, const bbduk_qtrim_cfg *cfg, const BBParams &bp, const uint8_t *d_bases, const uint8_t *d_quals,
                 const uint32_t *d_offsets, int64_t n_reads, int paired, int32_t *d_lo, int32_t *d_hi, uint8_t *d_flags,
                 unsigned long long *d_stats, cudaStream_t st) {
    if (n_reads < 1) return 0;
    QtrimDev p;
    p.qtrim_left = cfg->qtrim_left != 0;
    p.qtrim_right = cfg->qtrim_right != 0;
    p.mbq = cfg->min_base_quality;
    p.max_ns = cfg->max_ns;
    p.max_n_rate = cfg->max_n_rate;
    p.mcb = cfg->min_consecutive_bases;
    p.min_base_freq = cfg->min_base_frequency;
    p.max_len = cfg->max_read_length > 0 ? cfg->max_read_length : 0x7FFFFFFF;
    p.qual_offset = cfg->qual_offset;
    p.minReadLength = bp.minReadLength;
    p.minLenFraction = bp.minLenFraction;
    p.rieb = bp.removePairsIfEitherBad;
    p.tf1 = bp.trimFailuresTo1bp;
    p.poly_a = cfg->trim_poly_a;
    p.poly_g_left = cfg->trim_poly_g_left;
    p.poly_g_right = cfg->trim_poly_g_right;
    p.filter_g = cfg->filter_poly_g;
    p.poly_c_left = cfg->trim_poly_c_left;
    p.poly_c_right = cfg->trim_poly_c_right;
    p.filter_c = cfg->filter_poly_c;
    p.max_non_poly = cfg->max_non_poly;
    // parse/Parser.java:1757-1759 trimE; shared/TrimRead.java:364 nprob; align2/QualityTools.java:688-698 PROB_ERROR
    const float e = (float)phred_to_prob_error((double)cfg->trimq);
    volatile float n11 = e * 1.1f;
    const float nprob = std::max(std::min((float)n11, 1.0f), 0.75f);
    for (int raw = 0; raw < 256; raw++) {
        const int8_t q = (int8_t)(uint8_t)(raw - cfg->qual_offset);
        float pe = nprob;
        if (q >= 1) {
            pe = (float)std::pow(10.0, 0 - .1 * q);
            if (q == 1) pe = .7f;
        }
        volatile float d = e - pe;
        p.delta[raw] = d;
    }
    // minavgquality: the reference tests phred(p) < maq with phred(p) = p >= 1 ? 0 : p <= 1e-6 ? 60 : -10 * log10((double)p)
    // (align2/QualityTools.java:674-680), a predicate that only ever switches from false to true as the float p grows.
    // maq_prob = the smallest non-negative float for which it holds, found by bisection over the float bit patterns.
    p.maq_on = cfg->min_avg_quality > 0.0f;
    p.maq_bases = cfg->min_avg_quality_bases;
    p.maq_prob = 0.0f;
    if (p.maq_on) {
        const double maq = (double)cfg->min_avg_quality;
        auto low = [&](uint32_t bits) {
            float pr;
            memcpy(&pr, &bits, 4);
            const double prob = pr;
            const double phred = prob >= 1 ? 0.0 : (prob <= 0.000001 ? 60.0 : -10 * std::log10(prob));
            return phred < maq;
        };
        uint32_t lo_b = 0, hi_b = 0x7F800000u;  // +0 .. +inf; low(+inf) holds because maq > 0
        if (low(lo_b)) hi_b = lo_b;
        while (hi_b - lo_b > 1 && !low(lo_b)) {
            const uint32_t mid = lo_b + (hi_b - lo_b) / 2;
            if (low(mid)) hi_b = mid;
            else lo_b = mid;
        }
        memcpy(&p.maq_prob, &hi_b, 4);
        for (int raw = 0; raw < 256; raw++) {
            const int8_t q = (int8_t)(uint8_t)(raw - cfg->qual_offset);
            float pe = .75f;
            if (q >= 1) {
                pe = (float)std::pow(10.0, 0 - .1 * q);
                if (q == 1) pe = .7f;
            }
            p.pe[raw] = pe;
        }
    }
    const int64_t n_tiles = (n_reads + 31) / 32;
    const int blocks = (int)std::min<int64_t>((n_tiles + QT_THREADS / 32 - 1) / (QT_THREADS / 32), (int64_t)sm_count * 8);
    // which trimming rule runs: the staged optimal scan needs qualities and the default mode; the other rules (and reads
    // without qualities) take the plain per-lane path of the non-QT kernels
    const bool want_trim = p.qtrim_left || p.qtrim_right;
    p.trimq_byte = (int)(int8_t)(int)cfg->trimq;  // Java's (byte)trimq
    p.window = cfg->window_length;
    p.good_interval = cfg->min_good_interval;
    p.n_only_off = e >= 1.0f;
    p.alt_trim = !want_trim ? 0 : cfg->trim_mode == 1 ? 1 : cfg->trim_mode == 2 ? 2 : (d_quals ? 0 : 3);
    if (cfg->trim_mode == 1) p.qtrim_left = 0;  // window mode trims the right end only (parse/Parser.java:352-357)
    const bool qt = want_trim && !p.alt_trim;
    const bool poly = p.poly_a > 0 || p.poly_g_left > 0 || p.poly_g_right > 0 || p.filter_g > 0 || p.poly_c_left > 0 ||
                      p.poly_c_right > 0 || p.filter_c > 0;
#define QT_GO(A, B) qtrim_kernel<A, B><<<blocks, QT_THREADS, 0, st>>>(d_bases, d_quals, d_offsets, n_reads, paired, d_lo, d_hi, d_flags, p, d_stats)
    if (qt && poly) QT_GO(true, true);
    else if (qt) QT_GO(true, false);
    else if (poly) QT_GO(false, true);
    else QT_GO(false, false);
#undef QT_GO
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
