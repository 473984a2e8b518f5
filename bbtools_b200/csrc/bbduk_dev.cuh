// bbduk_dev.cuh -- device-side codec, key formula, hashing and hash-array access shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "params.h"

// ---- 2-bit codec (replaces dna.AminoAcid tables, dna/AminoAcid.java:269-285, :1289-1320) ----------
// A/a 0, C/c 1, G/g 2, T/t/U/u 3 from bits 1..2 of the ASCII code; "defined" is an exact membership
// test, everything else (N, IUPAC, '.', '-', ...) is undefined -> code0 = comp0 = 0.
__device__ __forceinline__ uint32_t bb_code_raw(uint32_t c) { return ((c >> 1) ^ (c >> 2)) & 3u; }
__device__ __forceinline__ bool bb_defined(uint32_t c) {
    // jgi/BBDuk.java:5355-5357 isFullyDefined: symbol>=0 && baseToNumber[symbol]>=0
    const uint32_t y = c | 0x20u;
    return (c < 128u) && (y == 'a' || y == 'c' || y == 'g' || y == 't' || y == 'u');
}
__device__ __forceinline__ uint32_t bb_code0(uint32_t c) { return bb_defined(c) ? bb_code_raw(c) : 0u; }
__device__ __forceinline__ uint32_t bb_comp0(uint32_t c) { return bb_defined(c) ? (3u - bb_code_raw(c)) : 0u; }
// baseToNumber (undefined -> -1), used for extraBase in the loader (jgi/BBDuk.java:2277)
__device__ __forceinline__ int bb_code_m1(uint32_t c) { return bb_defined(c) ? (int)bb_code_raw(c) : -1; }

// ---- reverse complement of a 2-bit packed k-mer (dna/AminoAcid.java:585-603) -----------------------
__device__ __forceinline__ uint64_t bb_rcomp(uint64_t kmer, int k) {
    uint32_t lo = ~(uint32_t)kmer, hi = ~(uint32_t)(kmer >> 32);
    lo = __brev(lo);
    hi = __brev(hi);
    // brev reversed the bits inside each base pair too: swap them back
    lo = ((lo & 0x55555555u) << 1) | ((lo >> 1) & 0x55555555u);
    hi = ((hi & 0x55555555u) << 1) | ((hi >> 1) & 0x55555555u);
    const uint64_t x = ((uint64_t)lo << 32) | hi;  // word swap completes the 64-bit reversal
    return x >> (2 * (32 - k));
}

// ---- key formula (jgi/BBDuk.java:4673-4685 toValue / :3373-3375) -----------------------------------
__device__ __forceinline__ uint64_t bb_to_value(const BBParams &p, uint64_t kmer, uint64_t rkmer, uint64_t lengthMask) {
    const uint64_t v = p.rcomp ? (kmer > rkmer ? kmer : rkmer) : kmer;
    return (v & p.middleMask) | lengthMask;
}
// speed= filter. jgi.BBDuk (bbdukOld.sh): (key & Long.MAX_VALUE) % 17 >= speed (jgi/BBDuk.java:4702-4713). bbduk.BBDukS
// (bbduk.sh) with its default index: speed < 2 || ((hash64plus2(key) >> 16) & 15) + 1 >= speed
// (bbduk/BBDukIndexMask2.java:566-577, shared/Tools.java:5482-5497: MurmurHash3 finalizer, sign bit cleared, top values folded).
__host__ __device__ __forceinline__ uint64_t bb_hash64plus2(uint64_t key) {
    key ^= key >> 33;
    key *= 0xff51afd7ed558ccdull;
    key ^= key >> 33;
    key *= 0xc4ceb9fe1a85ec53ull;
    key ^= key >> 33;
    key &= 0x7FFFFFFFFFFFFFFFull;
    return key < 0x7FFFF800FFFFFFFFull ? key : (key - 0x7FFFF800FFFFFFFFull) * 64ull;
}
__device__ __forceinline__ bool bb_passes_speed(const BBParams &p, uint64_t key) {
    if (p.speedMask2) return p.speed < 2 || (int)((bb_hash64plus2(key) >> 16) & 15ull) + 1 >= p.speed;
    return p.speed < 1 || (int)((key & 0x7FFFFFFFFFFFFFFFull) % 17ull) >= p.speed;
}

// ---- hashing: layout only, any mixing function gives the same results ------------------------------
// One 32-bit multiplicative hash of the key serves everything: its HIGH bits pick the hash-array
// bucket and the filter word, its LOW bits pick the filter bit pattern.
__device__ __forceinline__ uint32_t bb_fhash(uint32_t klo, uint32_t khi) { return klo * 0x9E3779B1u + khi * 0x85EBCA77u; }
__device__ __forceinline__ uint32_t bb_fhash64(uint64_t key) { return bb_fhash((uint32_t)key, (uint32_t)(key >> 32)); }

// Hash array: buckets of 4 consecutive slots = one 32-byte sector; linear probing over buckets. Slots
// of a bucket fill in order and nothing is ever deleted, so an EMPTY slot ends the search.
// n_slots is a power of two in [1024, 2^34]; bucket = hash >> (34 - log2(n_slots)).
#define BB_MAX_PROBE 8192
__device__ __forceinline__ uint32_t bb_bucket(uint32_t h, uint32_t bucket_shift) { return h >> bucket_shift; }

// Lookup = kmer.AbstractKmerTable.getValue (kmer/AbstractKmerTable.java:61): id or -1.
__device__ __forceinline__ int bb_table_get(const BBTable &t, uint64_t key) {
    uint64_t b = bb_bucket(bb_fhash64(key), t.bucket_shift);
    const uint64_t bmask = t.slot_mask >> 2;
#pragma unroll 1  // one bucket is the normal case; unrolled copies only cost instruction-cache space
    for (int probe = 0; probe < BB_MAX_PROBE / 4; probe++) {
        const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(t.keys + 4 * b);
        const ulonglong2 k01 = __ldg(q), k23 = __ldg(q + 1);  // one sector, both halves in flight
        int j = -1;
        if (k01.x == key) j = 0;
        else if (k01.y == key) j = 1;
        else if (k23.x == key) j = 2;
        else if (k23.y == key) j = 3;
        if (j >= 0) return __ldg(t.vals + 4 * b + j);
        if (k23.y == BB_EMPTY_KEY) return -1;  // slots fill in order: last one empty <=> bucket not full
        b = (b + 1) & bmask;
    }
    return -1;
}

// Insert = setIfNotPresent with "first writer wins" realised as min id.
// Returns 1 if this call created the key. The probe length is bounded so that an over-full array
// raises *overflow instead of spinning forever.
__device__ __forceinline__ int bb_table_put(uint64_t *keys, int32_t *vals, uint64_t slot_mask, uint32_t bucket_shift,
                                            uint64_t key, int32_t id, int *overflow) {
    uint64_t slot = (uint64_t)bb_bucket(bb_fhash64(key), bucket_shift) << 2;
    for (int probe = 0; probe < BB_MAX_PROBE; probe++) {
        uint64_t kk = keys[slot];
        if (kk == BB_EMPTY_KEY) {
            kk = atomicCAS((unsigned long long *)(keys + slot), (unsigned long long)BB_EMPTY_KEY, (unsigned long long)key);
            if (kk == BB_EMPTY_KEY) {
                atomicMin(vals + slot, id);
                return 1;
            }
        }
        if (kk == key) {
            atomicMin(vals + slot, id);
            return 0;
        }
        slot = (slot + 1) & slot_mask;
    }
    *overflow = 1;
    return 0;
}

// ---- blocked bloom pre-filter over all keys (no false negatives) ----------------------------------
// One 32-bit hash t of the key picks the filter word (high bits, multiply-shift range reduction) and
// a 4-bit pattern: two 2-bit stencils rotated by two independent 5-bit fields of t (funnel shifts use
// the shift amount mod 32, so no masking is needed). The same functions build and query the filter.
#define BB_FPAT1 0x00000081u
#define BB_FPAT2 0x00002001u
__device__ __forceinline__ uint32_t bb_filter_word(uint32_t t, uint32_t n_words) { return __umulhi(t, n_words); }
__device__ __forceinline__ uint32_t bb_filter_bits(uint32_t t) {
    return __funnelshift_l(BB_FPAT1, BB_FPAT1, t) | __funnelshift_l(BB_FPAT2, BB_FPAT2, t >> 5);
}

__device__ __forceinline__ uint32_t bb_big_word(uint32_t t, uint32_t n_words) { return __umulhi(t, n_words); }

// ---- pigeonhole part filter -------------------------------------------------------------------------
// A window can only hit the table if, on one strand, it agrees with an UNMUTATED reference k-mer on at
// least one of hdist+1 disjoint parts (substitutions only). The filter knows the parts of every
// reference k-mer and of its reverse complement, so the query needs the forward window only.
// It is a DIRECT bitmap over all BB_PART_WD-mers (4^9 bits = 32 KB): a part of w >= 9 bases is in the
// filter iff each of its w-8 overlapping 9-mers is, so the scan does ONE word load per read position
// (the 9-mer ending there: no hash, the word index and the bit are fields of the window itself) and
// rebuilds the w-mer answer bit-parallel as the AND of w-8 consecutive position bits.
// Bit layout: 9-mer value x (18 bits, newest base lowest) lives in word x>>5 at bit (x-1)&31, so that
// rotating the word right by x (the funnel shift takes x mod 32) leaves the answer in bit 31, from where
// one more funnel shift pushes it into the per-step accumulator: 6 instructions per read position.
#define BB_PART_WD 9
#define BB_PART_WORDS (1u << (2 * BB_PART_WD - 5))
__host__ __device__ __forceinline__ uint32_t bb_part_word(uint32_t x) { return (x >> 5) & (BB_PART_WORDS - 1u); }
__host__ __device__ __forceinline__ uint32_t bb_part_bit(uint32_t x) { return 1u << ((x - 1u) & 31u); }
// v: 2-bit codes of (at least) the w bases ending at a position, newest base in the low bits
__device__ __forceinline__ bool bb_part_test(const uint32_t *filt, uint32_t v, int w) {
    bool ok = true;
#pragma unroll 1
    for (int d = 0; d <= w - BB_PART_WD; d++) {
        const uint32_t x = v >> (2 * d);
        ok = ok && (filt[bb_part_word(x)] & bb_part_bit(x)) != 0;
    }
    return ok;
}

// ---- exact key of a window that may contain undefined bases (SURVEY.md A.2 closed form) -------------
// win: 2-bit codes of the 32 bases ending at the probe position, slot t = base i-t, undefined bases read as 0;
// dw: their "defined" bits, bit t = base i-t. kmer keeps code 0 for undefined bases and is never reset; with
// forbidNs the reverse k-mer only holds the bases after the last undefined one and the probe needs
// len >= minlen2 (jgi/BBDuk.java:3882-3900). Returns false if the reference would not probe here.
__device__ __forceinline__ uint64_t bb_spread2(uint32_t m) {  // bit t -> bits (2t+1, 2t)
    uint64_t x = m;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x | (x << 1);
}
__device__ __forceinline__ bool bb_window_key(const BBParams &p, uint64_t win, uint32_t dw, uint64_t *key) {
    const int k = p.k;
    const uint32_t kbits = (k >= 32) ? 0xFFFFFFFFu : ((1u << k) - 1u);
    uint64_t kmer = win & p.mask, rkmer;
    if ((dw & kbits) == kbits) {
        rkmer = bb_rcomp(kmer, k);
    } else {
        const uint64_t E = bb_spread2(dw & kbits);
        kmer &= E;
        if (p.forbidNs) {
            const int len = __ffs(~dw) - 1;  // bases after the last undefined one (the window has one)
            if (len < p.minlen2) return false;
            rkmer = bb_rcomp(kmer, k) & ~((1ull << (2 * (k - len))) - 1ull) & p.mask;
        } else {
            rkmer = bb_rcomp(kmer, k) & bb_rcomp(~E, k);  // E reversed slot-wise
        }
    }
    *key = bb_to_value(p, kmer, rkmer, p.kmask);
    return true;
}

// lookup continuing from a bucket whose 32 bytes the caller has already loaded
__device__ __forceinline__ int bb_table_get_from(const BBTable &t, uint64_t b, ulonglong2 k01, ulonglong2 k23, uint64_t key) {
    const uint64_t bmask = t.slot_mask >> 2;
#pragma unroll 1  // one bucket is the normal case; unrolled copies only cost instruction-cache space
    for (int probe = 0; probe < BB_MAX_PROBE / 4; probe++) {
        int j = -1;
        if (k01.x == key) j = 0;
        else if (k01.y == key) j = 1;
        else if (k23.x == key) j = 2;
        else if (k23.y == key) j = 3;
        if (j >= 0) return __ldg(t.vals + 4 * b + j);
        if (k23.y == BB_EMPTY_KEY) return -1;
        b = (b + 1) & bmask;
        const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(t.keys + 4 * b);
        k01 = __ldg(q);
        k23 = __ldg(q + 1);
    }
    return -1;
}
