// qtrim.cu -- BBDuk's quality-trimming block and the per-read quality / length / N filters on the device:
// bbduk_b200_qtrim / bbduk_b200_qtrim_device (first part).
//
// Replaces jgi/BBDuk.java:3074-3170: TrimRead.trimFast in its default "optimal" mode (shared/TrimRead.java:140-169,
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

struct QtrimDev {
    int qtrim_left, qtrim_right, mbq, max_ns, max_len, qual_offset;
    int minReadLength;
    float minLenFraction;
    int rieb, tf1;
    float delta[256];  // per raw quality byte: trimE - probError (trimE - nprob for q < 1)
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

// 0xFF in every byte lane of w that holds 'N'
__device__ __forceinline__ uint32_t n_lanes(uint32_t w) {
    const uint32_t y = w ^ 0x4E4E4E4Eu;
    const uint32_t z = ((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y;  // bit 7 of a lane <=> byte != 0
    return ((~z & 0x80808080u) >> 7) * 0xFFu;
}

template <bool QT>
__global__ void __launch_bounds__(QT_THREADS)
qtrim_kernel(const uint8_t *__restrict__ bases, const uint8_t *__restrict__ quals, const uint32_t *__restrict__ offsets,
             int64_t n_reads, int paired, int32_t *lo_io, int32_t *hi_io, uint8_t *flags_io, const QtrimDev p,
             unsigned long long *stats) {
    __shared__ float D[256];
    for (int i = threadIdx.x; i < 256; i += QT_THREADS) D[i] = p.delta[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (QT_THREADS / 32);
    const int64_t n_tiles = (n_reads + 31) >> 5;
    const uint32_t off_word = (uint32_t)(p.qual_offset & 0xFF) * 0x01010101u;
    unsigned int s_rq = 0, s_bq = 0, s_rf = 0, s_bf = 0, s_rn = 0, s_bn = 0;
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
        int x = 0;
        if (QT && !removed && h - l >= 1) {
            // testOptimal (shared/TrimRead.java:348-410) over [l,h)
            float score = 0.0f, maxScore = 0.0f;
            int count = 0, maxLoc = -1, maxCount = -1, i = 0;
            auto step = [&](uint32_t qbyte) {
                score = __fadd_rn(score, D[qbyte]);
                if (score > 0.0f) {
                    count++;
                    if (score > maxScore || (score == maxScore && count > maxCount)) {
                        maxScore = score;
                        maxCount = count;
                        maxLoc = i;
                    }
                } else {
                    score = 0.0f;
                    count = 0;
                }
                i++;
            };
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
            const int n = h - l;
            int a0 = 0, b0 = n;
            if (maxScore > 0.0f) {
                a0 = maxLoc - maxCount + 1;
                b0 = n - maxLoc - 1;
            }
            x = trim_amounts(l, h, p.qtrim_left ? a0 : 0, p.qtrim_right ? b0 : 0, 1);  // trimFast -> trimByAmount(r, a, b, 1)
        }
        auto is_disc = [&]() { return discarded || (p.tf1 && h - l == 1); };
        auto set_disc = [&]() {  // jgi/BBDuk.java:3260-3266
            if (p.tf1) {
                if (h - l > 1) trim_amounts(l, h, 0, h - l - 1, 1);
            } else {
                discarded = true;
            }
        };
        // :3092-3099 minlen / maxlen
        const int minlenR = (int)fmaxf(__fmul_rn((float)L, p.minLenFraction), (float)p.minReadLength);
        if (!removed && !is_disc()) {
            const int len = h - l;
            if (len < minlenR || len > p.max_len) set_disc();
        }
        // :3102-3106 shouldRemove after quality trimming
        const bool d1 = is_disc();
        const bool d1m = __shfl_xor_sync(0xFFFFFFFFu, (int)d1, 1) != 0;
        const bool rem1 = !removed && (paired ? (p.rieb ? (d1 || d1m) : (d1 && d1m)) : d1);
        const int len1 = h - l;
        const int len1m = __shfl_xor_sync(0xFFFFFFFFu, len1, 1);
        const int xm = __shfl_xor_sync(0xFFFFFFFFu, x, 1);
        const bool first = !paired || !(lane & 1);
        if (!removed && first) {
            s_bq += (unsigned int)(x + (paired ? xm : 0));
            s_rq += (x > 0) + ((paired && xm > 0) ? 1 : 0);
            if (rem1) s_bq += (unsigned int)(len1 + (paired ? len1m : 0));
        }
        // :3110-3148 minbasequality, maxns
        if (!removed && !rem1) {
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
        }
        // :3162-3167 shouldRemove after quality filtering
        const bool d2 = is_disc();
        const bool d2m = __shfl_xor_sync(0xFFFFFFFFu, (int)d2, 1) != 0;
        const bool rem2 = !removed && !rem1 && (paired ? (p.rieb ? (d2 || d2m) : (d2 && d2m)) : d2);
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
                                    ((rem1 || rem2) ? BBDUK_F_REMOVED : 0) | (x > 0 ? BBDUK_F_QTRIMMED : 0));
        }
    }
    if (stats) {
        const unsigned int v[6] = {s_rq, s_bq, s_rf, s_bf, s_rn, s_bn};
#pragma unroll
        for (int q = 0; q < 6; q++) {
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
    p.max_len = cfg->max_read_length > 0 ? cfg->max_read_length : 0x7FFFFFFF;
    p.qual_offset = cfg->qual_offset;
    p.minReadLength = bp.minReadLength;
    p.minLenFraction = bp.minLenFraction;
    p.rieb = bp.removePairsIfEitherBad;
    p.tf1 = bp.trimFailuresTo1bp;
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
    const int64_t n_tiles = (n_reads + 31) / 32;
    const int blocks = (int)std::min<int64_t>((n_tiles + QT_THREADS / 32 - 1) / (QT_THREADS / 32), (int64_t)sm_count * 8);
    if (p.qtrim_left || p.qtrim_right)
        qtrim_kernel<true><<<blocks, QT_THREADS, 0, st>>>(d_bases, d_quals, d_offsets, n_reads, paired, d_lo, d_hi, d_flags, p, d_stats);
    else
        qtrim_kernel<false><<<blocks, QT_THREADS, 0, st>>>(d_bases, d_quals, d_offsets, n_reads, paired, d_lo, d_hi, d_flags, p, d_stats);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
