// fast_common.cuh -- device helpers shared by the tuned per-read kernels (probe_fast.cu, probe_fast2.cu): SIMD-in-register
// base classification, the staged 2-bit stream, the exact evaluator of one window, TrimRead arithmetic.
#pragma once
#include "bbduk_dev.cuh"

namespace bbfast {

constexpr int PAD = 2;           // zero chunks in front of the staged stream (windows reach back 32 bases)
constexpr int TAIL = 4;          // zero chunks behind it (the reverse-strand words run two steps ahead)
constexpr int MAX_FAST_LEN = 1008;
constexpr int FAST_SMEM_LIMIT = 227 * 1024;

__device__ __forceinline__ uint32_t pair_reverse_complement(uint32_t x) {
    // big-endian 16 bases -> little-endian complemented 16 bases
    uint32_t r = __brev(~x);
    return ((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u);
}

// 4 ASCII bases (byte 0 first) -> raw 2-bit codes in each byte's low bits, and "bad" (non-zero byte
// <=> the base is not one of ACGTUacgtu). With d = (c|0x20)^0x61 a base is valid iff bits 7,6,5,3 of d
// are 0, bit 4 equals q, and (q or bit 0 is 0), where q = "bits (2,1) are 10" marks the t/u class.
__device__ __forceinline__ void classify4(uint32_t w, uint32_t &codes, uint32_t &bad) {
    codes = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    const uint32_t d = (w | 0x20202020u) ^ 0x61616161u;
    const uint32_t q = (d >> 2) & ~(d >> 1) & 0x01010101u;
    bad = (d & 0xE8E8E8E8u) | (((d >> 4) ^ q) & 0x01010101u) | (d & ~q & 0x01010101u);
}
__device__ __forceinline__ uint32_t pack4(uint32_t codes) { return (codes * 0x40100401u) >> 24; }  // big-endian 8 bits
__device__ __forceinline__ uint32_t valid4(uint32_t bad) {                                           // 4 bits, base 0 in bit 3
    const uint32_t nz = (((bad & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | bad) & 0x80808080u;
    return (((nz ^ 0x80808080u) >> 7) * 0x08040201u) >> 24;
}

struct Stream {
    const uint32_t *F;  // big-endian 2-bit codes, 16 bases per word, index PAD = first chunk
    const uint16_t *D;  // defined bits, bit 15-b = base b of the chunk
    // 16 bases starting at stream base g, big-endian
    __device__ __forceinline__ uint32_t f16(int g) const {
        const int w = (g >> 4) + PAD;
        return __funnelshift_l(F[w + 1], F[w], (g & 15) * 2);
    }
    __device__ __forceinline__ uint32_t d16(int g) const {  // bit 15-b = base g+b
        const int w = (g >> 4) + PAD;
        const uint32_t x = ((uint32_t)D[w] << 16) | D[w + 1];
        return (x >> (16 - (g & 15))) & 0xFFFFu;
    }
    // defined bits of the 32 bases ending at stream base e: bit t = base e-t
    __device__ __forceinline__ uint32_t dwin(int e) const { return (d16(e - 31) << 16) | d16(e - 15); }
    // 2-bit codes of the 32 bases ending at stream base e: slot t (bits 2t+1,2t) = base e-t
    __device__ __forceinline__ uint64_t win(int e) const { return ((uint64_t)f16(e - 31) << 32) | f16(e - 15); }
};

__device__ __forceinline__ uint64_t spread2(uint32_t m) {  // bit t -> bits (2t+1, 2t)
    uint64_t x = m;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x | (x << 1);
}
__device__ __forceinline__ uint32_t spread16(uint32_t m) { return (uint32_t)spread2(m & 0xFFFFu); }  // 16 bits -> 16 slots
// reverse the order of the low `n` 2-bit slots (no complement)
__device__ __forceinline__ uint64_t rev2(uint64_t x, int n) { return bb_rcomp(~x, n); }

// exact id of the full-length probe whose window ends at stream base e, -1 if none / no probe.
// clean = the owning read has no undefined base. Otherwise undefined bases are handled exactly
// (SURVEY.md A.2): kmer keeps code 0 for them and is never reset; with forbidNs the reverse k-mer only
// holds the bases after the last undefined one and the probe needs len >= minlen2.
__device__ __forceinline__ int exact_full(const Stream &st, int e, bool clean, const BBParams &p, const BBTable &t) {
    const int k = p.k;
    uint64_t kmer = st.win(e) & p.mask;
    uint64_t rkmer;
    bool plain = clean;
    uint32_t dw = 0;
    const uint32_t kbits = (1u << k) - 1u;  // k <= 31
    if (!clean) {
        dw = st.dwin(e);
        plain = (dw & kbits) == kbits;
    }
    if (plain) {
        rkmer = bb_rcomp(kmer, k);
    } else {
        const uint64_t E = spread2(dw & kbits);
        kmer &= E;
        if (p.forbidNs) {
            const int len = __ffs(~dw) - 1;  // bases after the last undefined one (the window has one)
            if (len < p.minlen2) return -1;
            rkmer = bb_rcomp(kmer, k) & ~((1ull << (2 * (k - len))) - 1ull) & p.mask;
        } else {
            rkmer = bb_rcomp(kmer, k) & rev2(E, k);
        }
    }
    return bb_table_get(t, bb_to_value(p, kmer, rkmer, p.kmask));
}

__device__ __forceinline__ bool filter_pass(const uint32_t *filt, uint32_t nfw, uint64_t key) {
    const uint32_t tt = bb_fhash64(key);
    const uint32_t pat = bb_filter_bits(tt);
    return (filt[bb_filter_word(tt, nfw)] & pat) == pat;
}

__device__ __forceinline__ int mid3(int x, int y, int z) { return max(min(x, y), min(max(x, y), z)); }

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

// OR of x >> d for d in [0, n): every set bit also covers the n-1 lower bit positions
__device__ __forceinline__ uint32_t smear_right(uint32_t x, int n) {
    int have = 1;
    while (have < n) {
        const int s = min(have, n - have);
        x |= x >> s;
        have += s;
    }
    return (uint32_t)x;
}

enum { FM_KTRIM_R = 0, FM_KTRIM_L = 1, FM_KFILTER = 2, FM_KMASK = 3 };


}  // namespace bbfast
