// entropy.cu -- BBDuk's low-entropy read filter (entropy=<cutoff>) on the device: bbduk_b200_entropy / _entropy_device
// (SURVEY.md 8f row 4).
//
// Replaces the "Test entropy" block of the per-pair loop (jgi/BBDuk.java:3175-3186: passes(r.bases, true), setDiscarded,
// shouldRemove, basesEFilteredT / readsEFilteredT) with tracker/EntropyTracker.java underneath: averageEntropy (:657-703)
// over a sliding window of `window` bases, the k-mer counts of the window kept incrementally (add :815-946) and the
// window's entropy read off a running double-precision sum of pk*log(pk) terms (calcEntropyFast :194-201).
//
// One lane per read, mates on neighbouring lanes. The window's k-mer counts (4^k bytes, k <= 5: 1 KB) live in shared
// memory, one private table per lane; a read leaves its table zeroed by walking its last window backwards instead of
// clearing 4^k entries. The table of pk*log(pk) is built on the host with the C library's log() and handed over as a
// kernel parameter; every double and float operation keeps the reference's order.
#include <algorithm>
#include <cmath>
#include <cstring>

#include <cuda_runtime.h>

#include "../../include/bbduk_b200.h"
#include "probe.h"

namespace {

constexpr int EN_THREADS = 64;
constexpr int EN_MAX_WK = 254;  // counts are bytes: windowKmers + 1 must fit
constexpr int ES_BYTES = 32 * 150 + 32;  // staged base codes per warp (with k = 5 three blocks of 64 lanes just fit one SM)

struct EntropyDev {
    float cutoff;
    int k, window, mask, space, high_pass, rieb, tf1, n_e;  // n_e = entries of E staged in shared memory (even)
    double mult;                // entropyMult = -1 / log(windowKmers)
    double E[EN_MAX_WK + 2];    // E[c] = (c / windowKmers) * log(c / windowKmers)
};

__device__ __forceinline__ uint32_t sym0(uint8_t b) {  // dna/AminoAcid.java symbolToNumber0: A0 C1 G2 T/U3, anything else 0
    const uint8_t y = b | 0x20;
    if (b >= 128) return 0;
    return y == 'c' ? 1u : y == 'g' ? 2u : (y == 't' || y == 'u') ? 3u : 0u;
}

// 4 ASCII bases -> their codes (A0 C1 G2 T/U3, either case) in the byte lanes, 0 for anything else.
// With d = (c|0x20)^0x61 a base is valid iff bits 7,6,5,3 of d are 0, bit 4 equals q, and (q or bit 0 is 0), where
// q = "bits (2,1) are 10" marks the t/u class (the same classification as probe_fast.cu's stage A).
__device__ __forceinline__ uint32_t codes4(uint32_t w) {
    const uint32_t codes = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    const uint32_t d = (w | 0x20202020u) ^ 0x61616161u;
    const uint32_t q = (d >> 2) & ~(d >> 1) & 0x01010101u;
    const uint32_t bad = (d & 0xE8E8E8E8u) | (((d >> 4) ^ q) & 0x01010101u) | (d & ~q & 0x01010101u);
    const uint32_t nz = (((bad & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | bad) & 0x80808080u;  // bit 7 of a lane <=> byte of bad != 0
    return codes & ~((nz >> 7) * 0xFFu);
}

// The tile's bases -> codes in shared memory. Four 16-byte loads per lane are in flight at a time: with six warps per SM
// a loop that waits for every load on its own spends a third of the kernel here.
__device__ __forceinline__ void stage_codes(const uint8_t *__restrict__ src16, uint32_t nchunks, uint8_t *Bs, int lane) {
    const uint4 *src = reinterpret_cast<const uint4 *>(src16);
    uint4 *dst = reinterpret_cast<uint4 *>(Bs);
    uint32_t c = lane;
    for (; c + 96 < nchunks; c += 128) {
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = __ldg(src + c + 32 * j);
#pragma unroll
        for (int j = 0; j < 4; j++) dst[c + 32 * j] = make_uint4(codes4(v[j].x), codes4(v[j].y), codes4(v[j].z), codes4(v[j].w));
    }
    for (; c < nchunks; c += 32) {
        const uint4 v = __ldg(src + c);
        dst[c] = make_uint4(codes4(v.x), codes4(v.y), codes4(v.z), codes4(v.w));
    }
}
// pull the warp's next tile into L2 while this one is being worked on (one 128-byte line per lane and round)
__device__ __forceinline__ void prefetch_tile(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ offsets, int64_t n_reads,
                                              int64_t tile, int lane) {
    if (tile * 32 >= n_reads) return;
    const int64_t r1 = tile * 32 + 32 < n_reads ? tile * 32 + 32 : n_reads;
    const uint32_t b0 = __ldg(offsets + tile * 32) & ~127u, b1 = __ldg(offsets + r1);
    for (uint32_t a = b0 + 128u * (uint32_t)lane; a < b1; a += 128u * 32u) asm volatile("prefetch.global.L2 [%0];" ::"l"(bases + a));
}

// EntropyTracker.averageEntropy(bases, allowNs = true) (tracker/EntropyTracker.java:657-703) of one read. src = the read's
// staged codes (STAGED) or its ASCII bases. Cl = this lane's slice of the warp's count table: the count of k-mer x is the
// byte (x & 3) of word (x >> 2) * 32 + lane, so the 32 lanes of a warp always hit 32 different banks. The window loop is
// split where the reference's conditions change (first k-1 bases: no k-mer yet; up to `window` bases: k-mers only enter;
// afterwards one enters and one leaves per base), so the loops carry no position tests.
template <bool STAGED>
__device__ __forceinline__ float average_entropy(const uint8_t *__restrict__ src, int n, uint8_t *__restrict__ Cl,
                                                 const double *__restrict__ E, int k, int W, uint32_t mask,
                                                 double mult) {
    auto code = [&](int i) -> uint32_t { return STAGED ? (uint32_t)src[i] : sym0(src[i]); };
    auto slot = [&](uint32_t x) -> uint8_t * { return Cl + (((x & ~3u) << 5) | (x & 3u)); };
    double esum = 0.0, sum = 0.0;
    int div = 0;
    auto enter = [&](uint32_t x) {
        uint8_t *c = slot(x);
        const uint32_t oc = *c;
        *c = (uint8_t)(oc + 1);
        esum = __dsub_rn(__dadd_rn(esum, E[oc + 1]), E[oc]);
    };
    auto leave = [&](uint32_t x) {
        uint8_t *c = slot(x);
        const uint32_t oc = *c;
        *c = (uint8_t)(oc - 1);
        esum = __dsub_rn(__dadd_rn(esum, E[oc - 1]), E[oc]);
    };
    auto measure = [&]() {  // calcEntropyFast :194-201
        const float e1 = (float)__dmul_rn(esum, mult);
        sum = __dadd_rn(sum, (double)(e1 > 0.0f ? e1 : 0.0f));
        div++;
    };
    uint32_t kmer = 0;
    const int lim = min(n, W);
    int i = 0;
    for (; i < min(lim, k - 1); i++) kmer = ((kmer << 2) | code(i)) & mask;
    // Both loops are software-pipelined by hand: the only serial part of a base is the reference's chain of double-precision
    // adds (two while the first window fills, four afterwards). The counts and table terms of base i+1 are fetched while
    // the adds of base i are in flight; with six warps per SM nothing else would cover those latencies.
    if (i < lim) {
        kmer = ((kmer << 2) | code(i)) & mask;
        uint8_t *s = slot(kmer);
        uint32_t oc = *s;
        *s = (uint8_t)(oc + 1);
        double ea = E[oc + 1], eb = E[oc];
        for (i++; i < lim; i++) {
            kmer = ((kmer << 2) | code(i)) & mask;
            s = slot(kmer);
            oc = *s;
            esum = __dsub_rn(__dadd_rn(esum, ea), eb);
            *s = (uint8_t)(oc + 1);
            ea = E[oc + 1];
            eb = E[oc];
        }
        esum = __dsub_rn(__dadd_rn(esum, ea), eb);
    }
    measure();  // the first window (or the whole read if it is shorter)
    if (n > W) {
        uint32_t kmer2 = 0;  // the reference has rolled bases 0..k-2 into kmer2 by now
        for (int t = 0; t < k - 1; t++) kmer2 = ((kmer2 << 2) | code(t)) & mask;
        // One k-mer enters and one leaves per base. Both count loads are issued before either store (the two slots differ
        // unless the same k-mer enters and leaves, which is patched up in registers) and the next base's codes are read
        // ahead of the stores they could alias with.
        uint32_t c_in = code(i), c_out = code(i - W + k - 1);
        double e1a, e1b, e2a, e2b;
        {
            kmer = ((kmer << 2) | c_in) & mask;
            kmer2 = ((kmer2 << 2) | c_out) & mask;
            if (i + 1 < n) {
                c_in = code(i + 1);
                c_out = code(i + 1 - W + k - 1);
            }
            uint8_t *s1 = slot(kmer), *s2 = slot(kmer2);
            const uint32_t oc1 = *s1;
            uint32_t oc2 = *s2;
            if (s1 == s2) oc2 = oc1 + 1;  // it was just counted
            *s1 = (uint8_t)(oc1 + 1);
            *s2 = (uint8_t)(oc2 - 1);  // same slot: ends at oc1 again
            e1a = E[oc1 + 1], e1b = E[oc1], e2a = E[oc2 - 1], e2b = E[oc2];
        }
        for (i++; i < n; i++) {
            kmer = ((kmer << 2) | c_in) & mask;
            kmer2 = ((kmer2 << 2) | c_out) & mask;
            if (i + 1 < n) {
                c_in = code(i + 1);
                c_out = code(i + 1 - W + k - 1);
            }
            uint8_t *s1 = slot(kmer), *s2 = slot(kmer2);
            const uint32_t oc1 = *s1;
            uint32_t oc2 = *s2;
            // the previous base's adds run while this base's counts arrive
            esum = __dsub_rn(__dadd_rn(esum, e1a), e1b);
            esum = __dsub_rn(__dadd_rn(esum, e2a), e2b);
            measure();
            if (s1 == s2) oc2 = oc1 + 1;
            *s1 = (uint8_t)(oc1 + 1);
            *s2 = (uint8_t)(oc2 - 1);
            e1a = E[oc1 + 1], e1b = E[oc1], e2a = E[oc2 - 1], e2b = E[oc2];
        }
        esum = __dsub_rn(__dadd_rn(esum, e1a), e1b);
        esum = __dsub_rn(__dadd_rn(esum, e2a), e2b);
        measure();
    }
    {  // leave the table zeroed: the k-mers still inside the last window are the only non-zero counts
        const int start = max(0, n - W);
        uint32_t km = 0;
        int t = start;
        for (; t < min(n, start + k - 1); t++) km = ((km << 2) | code(t)) & mask;
        for (; t < n; t++) {
            km = ((km << 2) | code(t)) & mask;
            *slot(km) = 0;
        }
    }
    return (float)__ddiv_rn(sum, (double)max(1, div));
}

__global__ void __launch_bounds__(EN_THREADS)
entropy_kernel(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ offsets, int64_t n_reads, int paired,
               const int32_t *__restrict__ lo_in, int32_t *hi_io, uint8_t *flags_io, const EntropyDev p, unsigned long long *stats) {
    extern __shared__ __align__(16) uint8_t smem[];
    double *E = reinterpret_cast<double *>(smem);
    // this lane's slice of its warp's count table (all zero between reads)
    uint8_t *Cl = smem + sizeof(double) * p.n_e + (size_t)(threadIdx.x >> 5) * 32 * p.space + (size_t)(threadIdx.x & 31) * 4;
    uint8_t *Bs = smem + sizeof(double) * p.n_e + (size_t)EN_THREADS * p.space + (size_t)(threadIdx.x >> 5) * ES_BYTES;
    for (int i = threadIdx.x; i < p.n_e; i += EN_THREADS) E[i] = p.E[i];
    {
        uint32_t *z = reinterpret_cast<uint32_t *>(smem + sizeof(double) * p.n_e);
        for (int i = threadIdx.x; i < EN_THREADS * p.space / 4; i += EN_THREADS) z[i] = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (EN_THREADS / 32);
    const int64_t n_tiles = (n_reads + 31) >> 5;
    const int k = p.k, W = p.window;
    unsigned int s_r = 0, s_b = 0;
    for (int64_t tile = (int64_t)blockIdx.x * (EN_THREADS / 32) + (threadIdx.x >> 5); tile < n_tiles; tile += warps_total) {
        const int64_t r = tile * 32 + lane;
        const bool live = r < n_reads;
        const uint32_t o0 = live ? offsets[r] : 0u;
        const int l = live ? lo_in[r] : 0;
        int h = live ? hi_io[r] : 0;
        const int f = live ? (int)flags_io[r] : BBDUK_F_REMOVED;
        const int f_first = paired ? __shfl_sync(0xFFFFFFFFu, f, lane & ~1) : f;
        const bool removed = !live || (f_first & BBDUK_F_REMOVED) != 0;
        bool discarded = (f & BBDUK_F_DISCARDED) != 0;
        const bool was_disc = discarded || (p.tf1 && h - l == 1);
        // Stage the tile's base CODES in shared memory: the 32 reads of a warp are contiguous in the batch, so the warp reads
        // them once with coalesced 16-byte loads (per-lane byte loads thrash the few KB of L1 this kernel leaves and pull
        // every sector through L2 many times). Tiles too long for the buffer read global memory directly.
        const int L = live ? (int)(offsets[r + 1] - o0) : 0;
        const int last_lane = (int)min((long long)31, (long long)(n_reads - 1 - tile * 32));
        const uint32_t t_lo = __shfl_sync(0xFFFFFFFFu, o0, 0);
        const uint32_t t_hi = __shfl_sync(0xFFFFFFFFu, o0 + (uint32_t)L, last_lane);
        const uint32_t a0t = t_lo & ~15u;
        const uint32_t nchunks = (t_hi - a0t + 15u) >> 4;
        const bool staged = nchunks * 16u <= (uint32_t)ES_BYTES && (reinterpret_cast<uintptr_t>(bases) & 15) == 0;
        if (staged) stage_codes(bases + a0t, nchunks, Bs, lane);
        prefetch_tile(bases, offsets, n_reads, tile + warps_total, lane);
        __syncwarp();
        if (!removed && !was_disc) {  // isNotDiscarded(r) && !passes(r.bases, true)
            const int n = h - l;
            const float e = staged ? average_entropy<true>(Bs + (o0 + (uint32_t)l - a0t), n, Cl, E, k, W, (uint32_t)p.mask, p.mult)
                                   : average_entropy<false>(bases + o0 + l, n, Cl, E, k, W, (uint32_t)p.mask, p.mult);
            const bool passes = (p.high_pass != 0) != (e < p.cutoff);
            if (!passes) {  // setDiscarded (jgi/BBDuk.java:3260-3266)
                if (p.tf1) {
                    if (h - l > 1) h = l + 1;
                } else {
                    discarded = true;
                }
            }
        }
        const bool d = discarded || (p.tf1 && h - l == 1);
        const bool dm = __shfl_xor_sync(0xFFFFFFFFu, (int)d, 1) != 0;
        const bool rem = !removed && (paired ? (p.rieb ? (d || dm) : (d && dm)) : d);
        const int len = h - l;
        const int lenm = __shfl_xor_sync(0xFFFFFFFFu, len, 1);
        if (rem && (!paired || !(lane & 1))) {
            s_b += (unsigned int)(len + (paired ? lenm : 0));
            s_r += paired ? 2 : 1;
        }
        if (live && !removed) {
            hi_io[r] = h;
            flags_io[r] = (uint8_t)((f & ~(BBDUK_F_DISCARDED | BBDUK_F_REMOVED)) | (discarded ? BBDUK_F_DISCARDED : 0) |
                                    (rem ? BBDUK_F_REMOVED : 0));
        }
    }
    if (stats) {
        const unsigned int tr = __reduce_add_sync(0xFFFFFFFFu, s_r), tb = __reduce_add_sync(0xFFFFFFFFu, s_b);
        if (lane == 0 && tr) {
            atomicAdd(stats, (unsigned long long)tr);
            atomicAdd(stats + 1, (unsigned long long)tb);
        }
    }
}

// ---- entropy masking / trimming (jgi/BBDuk.java:3055-3067, :4432-4478) -----------------------------------------------------
// as codes4, plus bit 2 of a lane set when the base is undefined (the window's N count decides whether it is tested at all)
__device__ __forceinline__ uint32_t codes4u(uint32_t w) {
    const uint32_t codes = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    const uint32_t d = (w | 0x20202020u) ^ 0x61616161u;
    const uint32_t q = (d >> 2) & ~(d >> 1) & 0x01010101u;
    const uint32_t bad = (d & 0xE8E8E8E8u) | (((d >> 4) ^ q) & 0x01010101u) | (d & ~q & 0x01010101u);
    const uint32_t nz = (((bad & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | bad) & 0x80808080u;
    return (codes & ~((nz >> 7) * 0xFFu)) | (nz >> 5);
}
__device__ __forceinline__ bool defined_ascii(uint8_t b) {
    const uint8_t y = b | 0x20;
    return b < 128 && (y == 'a' || y == 'c' || y == 'g' || y == 't' || y == 'u');
}

// maskLowEntropy's BitSet (jgi/BBDuk.java:4432-4446) of one read of n >= W bases: every full window without an undefined base
// whose entropy fails the cutoff sets its W positions. bits = the read's zeroed mask words (bit j = base j of the kept interval).
// Same incremental counts and the same order of double-precision updates as average_entropy / EntropyTracker.add (:815-946).
template <bool STAGED>
__device__ __forceinline__ void low_entropy_bits(const uint8_t *__restrict__ src, int n, uint8_t *__restrict__ Cl,
                                                 const double *__restrict__ E, int k, int W, uint32_t mask, double mult, float cutoff,
                                                 bool high_pass, uint32_t *__restrict__ bits) {
    auto code = [&](int i) -> uint32_t { return STAGED ? (uint32_t)(src[i] & 3u) : sym0(src[i]); };
    auto undef = [&](int i) -> int { return STAGED ? (int)(src[i] >> 2) : (defined_ascii(src[i]) ? 0 : 1); };
    auto slot = [&](uint32_t x) -> uint8_t * { return Cl + (((x & ~3u) << 5) | (x & 3u)); };
    double esum = 0.0;
    int ns = 0;
    auto enter = [&](uint32_t x) {
        uint8_t *c = slot(x);
        const uint32_t oc = *c;
        *c = (uint8_t)(oc + 1);
        esum = __dsub_rn(__dadd_rn(esum, E[oc + 1]), E[oc]);
    };
    auto leave = [&](uint32_t x) {
        uint8_t *c = slot(x);
        const uint32_t oc = *c;
        *c = (uint8_t)(oc - 1);
        esum = __dsub_rn(__dadd_rn(esum, E[oc - 1]), E[oc]);
    };
    auto test = [&](int i) {  // the window [i-W+1, i]: et.ns() < 1 && !et.passes()
        if (ns > 0) return;
        const float e1 = (float)__dmul_rn(esum, mult);
        const float e = e1 > 0.0f ? e1 : 0.0f;
        if (high_pass != (e < cutoff)) return;  // passes
        const int a = i - W + 1;
        for (int w = a >> 5; w <= (i >> 5); w++) {
            uint32_t m = 0xFFFFFFFFu;
            if (w == (a >> 5)) m &= 0xFFFFFFFFu << (a & 31);
            if (w == (i >> 5)) m &= 0xFFFFFFFFu >> (31 - (i & 31));
            bits[w] |= m;
        }
    };
    uint32_t kmer = 0;
    int i = 0;
    for (; i < W; i++) {
        kmer = ((kmer << 2) | code(i)) & mask;
        ns += undef(i);
        if (i >= k - 1) enter(kmer);
    }
    test(W - 1);
    uint32_t kmer2 = 0;
    for (int t = 0; t < k - 1; t++) kmer2 = ((kmer2 << 2) | code(t)) & mask;
    for (; i < n; i++) {
        kmer = ((kmer << 2) | code(i)) & mask;
        ns += undef(i);
        enter(kmer);
        kmer2 = ((kmer2 << 2) | code(i - W + k - 1)) & mask;
        ns -= undef(i - W);
        leave(kmer2);
        test(i);
    }
    {  // leave the table zeroed
        uint32_t km = 0;
        for (int t = n - W; t < n; t++) {
            km = ((km << 2) | code(t)) & mask;
            if (t >= n - W + k - 1) {
                uint8_t *c = slot(km);
                *c = (uint8_t)(*c - 1);
            }
        }
    }
}

// mode 1: mask to N, 2: mask to lower case (the returned count differs, jgi/BBDuk.java:4505-4526), 3: trim the masked ends
__global__ void __launch_bounds__(EN_THREADS)
entropy_mask_kernel(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ offsets, int64_t n_reads, int paired,
                    int32_t *lo_io, int32_t *hi_io, const uint8_t *__restrict__ flags, const EntropyDev p, int mode,
                    uint32_t *__restrict__ maskbits, const int64_t *__restrict__ mask_off, unsigned long long *stats) {
    extern __shared__ __align__(16) uint8_t smem[];
    double *E = reinterpret_cast<double *>(smem);
    uint8_t *Cl = smem + sizeof(double) * p.n_e + (size_t)(threadIdx.x >> 5) * 32 * p.space + (size_t)(threadIdx.x & 31) * 4;
    uint8_t *Bs = smem + sizeof(double) * p.n_e + (size_t)EN_THREADS * p.space + (size_t)(threadIdx.x >> 5) * ES_BYTES;
    for (int i = threadIdx.x; i < p.n_e; i += EN_THREADS) E[i] = p.E[i];
    {
        uint32_t *z = reinterpret_cast<uint32_t *>(smem + sizeof(double) * p.n_e);
        for (int i = threadIdx.x; i < EN_THREADS * p.space / 4; i += EN_THREADS) z[i] = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (EN_THREADS / 32);
    const int64_t n_tiles = (n_reads + 31) >> 5;
    const int k = p.k, W = p.window;
    unsigned int s_r = 0, s_b = 0;
    for (int64_t tile = (int64_t)blockIdx.x * (EN_THREADS / 32) + (threadIdx.x >> 5); tile < n_tiles; tile += warps_total) {
        const int64_t r = tile * 32 + lane;
        const bool live = r < n_reads;
        const uint32_t o0 = live ? offsets[r] : 0u;
        int l = live ? lo_io[r] : 0, h = live ? hi_io[r] : 0;
        const int f = live ? (int)flags[r] : BBDUK_F_REMOVED;
        const int f_first = paired ? __shfl_sync(0xFFFFFFFFu, f, lane & ~1) : f;
        const bool removed = !live || (f_first & BBDUK_F_REMOVED) != 0;
        const bool was_disc = (f & BBDUK_F_DISCARDED) != 0 || (p.tf1 && h - l == 1);
        const int L = live ? (int)(offsets[r + 1] - o0) : 0;
        const int last_lane = (int)min((long long)31, (long long)(n_reads - 1 - tile * 32));
        const uint32_t t_lo = __shfl_sync(0xFFFFFFFFu, o0, 0);
        const uint32_t t_hi = __shfl_sync(0xFFFFFFFFu, o0 + (uint32_t)L, last_lane);
        const uint32_t a0t = t_lo & ~15u;
        const uint32_t nchunks = (t_hi - a0t + 15u) >> 4;
        const bool staged = nchunks * 16u <= (uint32_t)ES_BYTES && (reinterpret_cast<uintptr_t>(bases) & 15) == 0;
        if (staged) {
            uint4 *dst = reinterpret_cast<uint4 *>(Bs);
            for (uint32_t c = lane; c < nchunks; c += 32) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(bases + a0t) + c);
                dst[c] = make_uint4(codes4u(v.x), codes4u(v.y), codes4u(v.z), codes4u(v.w));
            }
        }
        __syncwarp();
        if (!live) continue;
        uint32_t *bits = (maskbits && mask_off) ? maskbits + mask_off[r] : nullptr;
        const int nw_all = (L + 31) >> 5;
        if (bits)
            for (int w = 0; w < nw_all; w++) bits[w] = 0u;
        const int n = h - l;
        if (removed || was_disc || n < W || !bits) continue;  // isNotDiscarded(r), r.length() >= window
        if (staged) low_entropy_bits<true>(Bs + (o0 + (uint32_t)l - a0t), n, Cl, E, k, W, (uint32_t)p.mask, p.mult, p.cutoff, p.high_pass != 0, bits);
        else low_entropy_bits<false>(bases + o0 + l, n, Cl, E, k, W, (uint32_t)p.mask, p.mult, p.cutoff, p.high_pass != 0, bits);
        const int nw = (n + 31) >> 5;
        int x = 0;
        if (mode == 3) {  // trimLowEntropy :4448-4478: the masked runs at both ends go, trimByAmount(r, left, right, 1)
            int left = 0, right = 0;
            while (left < n && ((bits[left >> 5] >> (left & 31)) & 1u)) left++;
            while (right < n && ((bits[(n - 1 - right) >> 5] >> ((n - 1 - right) & 31)) & 1u)) right++;
            if (left || right) {
                // shared/TrimRead.java:299-346 on the kept interval
                int lt = left, rt = right;
                const int minLen = min(n, 1);
                if (lt + rt + minLen > n) {
                    rt = max(1, n - minLen);
                    lt = 0;
                }
                l += lt;
                h -= rt;
                x = lt + rt;
                lo_io[r] = l;
                hi_io[r] = h;
            }
            for (int w = 0; w < nw; w++) bits[w] = 0u;  // nothing is masked in this mode
        } else {  // maskFromBitset :4505-4526: bases that change (an N stays an N; lower case stays lower case)
            for (int w = 0; w < nw; w++) {
                uint32_t m = bits[w];
                while (m) {
                    const int j = 32 * w + __ffs(m) - 1;
                    m &= m - 1;
                    const uint8_t b = bases[o0 + l + j];
                    if (mode == 1) x += b != 'N';
                    else x += !(b >= 'a' && b <= 'z') && b != 'N';
                }
            }
        }
        s_b += (unsigned int)x;
        s_r += x > 0;
    }
    if (stats) {
        const unsigned int tr = __reduce_add_sync(0xFFFFFFFFu, s_r), tb = __reduce_add_sync(0xFFFFFFFFu, s_b);
        if (lane == 0 && (tr | tb)) {
            atomicAdd(stats, (unsigned long long)tr);
            atomicAdd(stats + 1, (unsigned long long)tb);
        }
    }
}

}  // namespace

// launcher used by abi.cu; 0 ok, 1 CUDA failure, 2 unsupported k / window
static int launch_entropy_any(int sm_count, const bbduk_entropy_cfg *cfg, const BBParams &bp, const uint8_t *d_bases,
                              const uint32_t *d_offsets, int64_t n_reads, int paired, const int32_t *d_lo, int32_t *d_lo_io, int32_t *d_hi,
                              uint8_t *d_flags, unsigned long long *d_stats, int mode, uint32_t *d_maskbits, const int64_t *d_mask_off,
                              cudaStream_t st) {
    if (n_reads < 1) return 0;
    const int k = cfg->k > 0 ? cfg->k : 5, W = cfg->window > 0 ? cfg->window : 50;  // tracker/EntropyTracker.java:1206-1209
    const int wk = W - k + 1;
    if (k < 1 || k > 5 || wk < 1 || wk > EN_MAX_WK) return 2;
    EntropyDev p;
    memset(&p, 0, sizeof p);
    p.cutoff = std::max(0.0f, cfg->cutoff);  // jgi/BBDuk.java:2518
    p.k = k;
    p.window = W;
    p.mask = (int)~(~0u << (2 * k));
    p.space = 1 << (2 * k);
    p.high_pass = cfg->high_pass != 0;
    p.rieb = bp.removePairsIfEitherBad;
    p.tf1 = bp.trimFailuresTo1bp;
    const double mult = 1.0 / wk;  // tracker/EntropyTracker.java:106-118
    for (int i = 1; i < wk + 2; i++) {
        const double pk = i * mult;
        p.E[i] = pk * std::log(pk);
    }
    p.mult = -1 / std::log((double)wk);
    p.n_e = (wk + 2 + 1) & ~1;
    const size_t smem = sizeof(double) * p.n_e + (size_t)EN_THREADS * p.space + (size_t)(EN_THREADS / 32) * ES_BYTES;
    const int64_t n_tiles = (n_reads + 31) / 32;
    const int blocks = (int)std::min<int64_t>((n_tiles + EN_THREADS / 32 - 1) / (EN_THREADS / 32), (int64_t)sm_count * 3);
    if (mode != 0) {
        if (cudaFuncSetAttribute(entropy_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
        entropy_mask_kernel<<<blocks, EN_THREADS, smem, st>>>(d_bases, d_offsets, n_reads, paired, d_lo_io, d_hi, d_flags, p, mode,
                                                              d_maskbits, d_mask_off, d_stats);
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
    }
    if (cudaFuncSetAttribute(entropy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
    entropy_kernel<<<blocks, EN_THREADS, smem, st>>>(d_bases, d_offsets, n_reads, paired, d_lo, d_hi, d_flags, p, d_stats);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_entropy(int sm_count, const bbduk_entropy_cfg *cfg, const BBParams &bp, const uint8_t *d_bases, const uint32_t *d_offsets,
                   int64_t n_reads, int paired, const int32_t *d_lo, int32_t *d_hi, uint8_t *d_flags, unsigned long long *d_stats,
                   cudaStream_t st) {
    return launch_entropy_any(sm_count, cfg, bp, d_bases, d_offsets, n_reads, paired, d_lo, nullptr, d_hi, d_flags, d_stats, 0, nullptr,
                              nullptr, st);
}

// entropymask= / entropytrim= : mode 1 mask to N, 2 mask to lower case, 3 trim
int launch_entropy_mask(int sm_count, const bbduk_entropy_cfg *cfg, const BBParams &bp, const uint8_t *d_bases,
                        const uint32_t *d_offsets, int64_t n_reads, int paired, int32_t *d_lo, int32_t *d_hi, uint8_t *d_flags,
                        int mode, uint32_t *d_maskbits, const int64_t *d_mask_off, unsigned long long *d_stats, cudaStream_t st) {
    return launch_entropy_any(sm_count, cfg, bp, d_bases, d_offsets, n_reads, paired, d_lo, d_lo, d_hi, d_flags, d_stats, mode, d_maskbits,
                              d_mask_off, st);
}
