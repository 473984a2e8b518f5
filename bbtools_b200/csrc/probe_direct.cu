// probe_direct.cu -- full-length k-mer scan against an HBM-resident hash array (BASELINE.json configs 3 and 4).
//
// When the reference k-mer set is too large for the on-chip filters of probe_fast.cu (contaminant genomes,
// hdist=2 neighbourhoods of anything but adapter-sized references), every read position is one probe of the
// hash array in HBM: one 32-byte bucket = one DRAM sector. The scan is position-parallel over the flat,
// concatenated base array -- the closed form of jgi/BBDuk.java:3882-3900 makes every position
// independent -- so the work is balanced for any read-length mix and each thread keeps 4 bucket loads in flight:
//   * pre-pass: one "a read starts here" bit per read (2 bytes per 16 bases);
//   * a warp stages a sub-span of 512 positions (+32 bases look-back) as 2-bit stream F, defined bits D, start
//     bits S in shared memory; windows crossing a read start do not exist (i < k-1); windows with an undefined
//     base take the exact slow path (N read as A / reverse k-mer truncated, forbidNs len rule);
//   * hits (rare) find their read by binary search over the offsets and fold into per-read
//     first hit (position,id) / last hit position with 64-bit atomicMin / atomicMax;
//   * the per-read epilogue (ktrim arithmetic, minlen, rieb/tpe pair logic, counters, outputs) is
//     probe_fast.cu's stage D, launched with the precomputed first/last hits.
// Covers what plan_direct() accepts: ktrim=r / ktrim=l without short k-mers and kfilter (countSetKmers,
// maxbadkmers=0), qhdist=0, speed=0, qskip=1, no restrictleft/right, k<=31.
#include <algorithm>

#include "bbduk_dev.cuh"
#include "probe.h"

namespace {

constexpr int DP_THREADS = 256;
constexpr int DP_SPAN = DP_THREADS * 16;
constexpr int DP_LOOK = 2;

__device__ __forceinline__ void dp_classify4(uint32_t w, uint32_t &codes, uint32_t &bad) {
    codes = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    const uint32_t d = (w | 0x20202020u) ^ 0x61616161u;
    const uint32_t q = (d >> 2) & ~(d >> 1) & 0x01010101u;
    bad = (d & 0xE8E8E8E8u) | (((d >> 4) ^ q) & 0x01010101u) | (d & ~q & 0x01010101u);
}
__device__ __forceinline__ uint32_t dp_pack4(uint32_t codes) { return (codes * 0x40100401u) >> 24; }
__device__ __forceinline__ uint32_t dp_valid4(uint32_t bad) {
    const uint32_t nz = (((bad & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | bad) & 0x80808080u;
    return (((nz ^ 0x80808080u) >> 7) * 0x08040201u) >> 24;
}
__device__ __forceinline__ uint64_t dp_smear(uint64_t x, int n) {  // OR of x >> d, d in [0,n)
    int have = 1;
    while (have < n) {
        const int s = min(have, n - have);
        x |= x >> s;
        have += s;
    }
    return x;
}

// one 32-byte bucket in one instruction; table lines are use-once, so they must not push the filter out of L2
__device__ __forceinline__ void dp_load_bucket(const uint64_t *p, ulonglong2 &k01, ulonglong2 &k23) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(k01.x), "=l"(k01.y), "=l"(k23.x), "=l"(k23.y)
                 : "l"(p));
}
__device__ __forceinline__ uint64_t dp_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint32_t dp_load_filter(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

// out-of-line slow paths keep the probe loop small enough for the instruction cache
__device__ __noinline__ bool dp_dirty_key(const BBParams &p, uint64_t win, uint32_t dw, uint64_t *key) {
    return bb_window_key(p, win, dw, key);
}
__device__ __noinline__ int dp_chain(const BBTable &t, uint64_t b, uint64_t key) {
    // the first bucket was full and did not hold the key: continue along the chain
    const uint64_t bmask = t.slot_mask >> 2;
    b = (b + 1) & bmask;
    const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(t.keys + 4 * b);
    return bb_table_get_from(t, b, __ldg(q), __ldg(q + 1), key);
}
// id of `key` given its first bucket's 32 bytes; -1 if absent
__device__ __forceinline__ int dp_resolve(const BBTable &t, uint64_t b, const ulonglong2 &k01, const ulonglong2 &k23, uint64_t key) {
    int j = -1;
    if (k01.x == key) j = 0;
    else if (k01.y == key) j = 1;
    else if (k23.x == key) j = 2;
    else if (k23.y == key) j = 3;
    if (j >= 0) return __ldg(t.vals + 4 * b + j);
    if (k23.y == BB_EMPTY_KEY) return -1;
    return dp_chain(t, b, key);
}

__global__ void dp_starts_kernel(const uint32_t *__restrict__ offsets, int64_t n_reads, uint32_t *sbits) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t o = offsets[r];
        atomicOr(&sbits[o >> 5], 1u << ((((o >> 4) & 1u) << 4) + 15u - (o & 15u)));
    }
}

__global__ void dp_init_kernel(unsigned long long *first64, int *lastpos, int64_t n_reads) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (int64_t)gridDim.x * blockDim.x) {
        first64[r] = ~0ull;
        if (lastpos) lastpos[r] = -1;
    }
}

// a hit at flat position g: credit the read that owns it
__device__ __noinline__ void dp_hit(const uint32_t *__restrict__ offsets, int64_t n_reads, int64_t g, int id, int paired,
                                    const BBParams &p, unsigned long long *first64, int *lastpos) {
    int64_t a = 0, b = n_reads + 1;  // first index with offsets[idx] > g
    while (a < b) {
        const int64_t m = (a + b) >> 1;
        if ((int64_t)offsets[m] <= g) a = m + 1;
        else b = m;
    }
    const int64_t r = a - 1;
    if (r < 0 || r >= n_reads) return;
    const int pairnum = (paired && (r & 1)) ? 1 : 0;
    if ((p.skipR1 && pairnum == 0) || (p.skipR2 && pairnum == 1)) return;
    const unsigned int pos = (unsigned int)(g - (int64_t)offsets[r]);
    atomicMin(first64 + r, ((unsigned long long)pos << 32) | (unsigned int)id);
    if (lastpos) atomicMax(lastpos + r, (int)pos);
}

// a hit at flat position g whose read index was derived from the start bits: verify, else search
__device__ __noinline__ void dp_credit(const uint32_t *__restrict__ offsets, int64_t n_reads, int64_t g, int64_t r, int id,
                                       int paired, const BBParams &p, unsigned long long *first64, int *lastpos) {
    bool ok = r >= 0 && r < n_reads;
    if (ok) ok = (int64_t)offsets[r] <= g && g < (int64_t)offsets[r + 1];
    if (!ok) {  // empty reads share a start bit: fall back to the search
        dp_hit(offsets, n_reads, g, id, paired, p, first64, lastpos);
        return;
    }
    const int pairnum = (paired && (r & 1)) ? 1 : 0;
    if ((p.skipR1 && pairnum == 0) || (p.skipR2 && pairnum == 1)) return;
    const unsigned int pos = (unsigned int)(g - (int64_t)offsets[r]);
    atomicMin(first64 + r, ((unsigned long long)pos << 32) | (unsigned int)id);
    if (lastpos) atomicMax(lastpos + r, (int)pos);
}

__global__ void __launch_bounds__(DP_THREADS, 3)
bbduk_direct_kernel(const uint8_t *__restrict__ bases, const uint16_t *__restrict__ sbits,
                    const uint32_t *__restrict__ offsets, int64_t n_reads, int64_t n_bases, int paired, BBParams p,
                    BBTable t, unsigned long long *first64, int *lastpos) {
    __shared__ uint32_t Fs_all[DP_THREADS / 32][32 + DP_LOOK];
    __shared__ uint32_t DSs_all[DP_THREADS / 32][32 + DP_LOOK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *Fs = Fs_all[warp], *DSs = DSs_all[warp];
    const int k = p.k;
    const int64_t n_spans = (n_bases + DP_SPAN - 1) / DP_SPAN;
    const uint64_t bmask = t.slot_mask >> 2;
    // L2-resident one-bit-per-key filter in front of the HBM probes (none for small arrays)
    const uint32_t *bigf = t.filter + t.n_filter_words + t.part_words + t.short_words + t.samp_words + t.tail_words;
    const uint32_t big_words = t.big_words;
    const uint64_t pol_keep = dp_policy_evict_last();

    for (int64_t span = blockIdx.x; span < n_spans; span += gridDim.x) {
        const int64_t g_lo = span * DP_SPAN + warp * 512;
        const int64_t c_lo = (g_lo >> 4) - DP_LOOK;
        __syncwarp();
        for (int i = lane; i < 32 + DP_LOOK; i += 32) {
            const int64_t c = c_lo + i;
            uint32_t f = 0, dbits = 0, sb = 0xFFFFu;  // outside the batch: every base "starts a read"
            if (c >= 0 && c * 16 < n_bases) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(bases) + c);
                sb = __ldg(sbits + c);
                uint32_t cw[4], bw[4];
                dp_classify4(v.x, cw[0], bw[0]);
                dp_classify4(v.y, cw[1], bw[1]);
                dp_classify4(v.z, cw[2], bw[2]);
                dp_classify4(v.w, cw[3], bw[3]);
                f = (dp_pack4(cw[0]) << 24) | (dp_pack4(cw[1]) << 16) | (dp_pack4(cw[2]) << 8) | dp_pack4(cw[3]);
                dbits = 0xFFFFu;
                if ((bw[0] | bw[1] | bw[2] | bw[3]) != 0)
                    dbits = (dp_valid4(bw[0]) << 12) | (dp_valid4(bw[1]) << 8) | (dp_valid4(bw[2]) << 4) | dp_valid4(bw[3]);
                const int64_t rem = n_bases - c * 16;
                if (rem < 16) {  // tail of the batch: bases past the end are start markers, so no window reaches them
                    dbits &= 0xFFFFu << (16 - rem);
                    sb |= 0xFFFFu >> rem;
                }
            }
            Fs[i] = f;
            DSs[i] = (dbits << 16) | sb;
        }
        __syncwarp();

        const int64_t g0 = g_lo + 16 * lane;
        const uint32_t f_m2 = Fs[lane], f_m1 = Fs[lane + 1], f_0 = Fs[lane + 2];
        const uint32_t w0 = DSs[lane], w1 = DSs[lane + 1], w2 = DSs[lane + 2];
        // 48-base strings, base i of the string at bit 47-i; this thread's positions are bases 32..47
        const uint64_t Dall = ((uint64_t)(w0 >> 16) << 32) | ((uint64_t)(w1 >> 16) << 16) | (w2 >> 16);
        const uint64_t Sall = ((uint64_t)(w0 & 0xFFFFu) << 32) | ((uint64_t)(w1 & 0xFFFFu) << 16) | (w2 & 0xFFFFu);
        const uint64_t U = (~Dall) & 0xFFFFFFFFFFFFull;
        // a read start at base i removes the windows ending in [i, i+k-2] (they would reach before the read);
        // positions past the end of the batch carry a start marker themselves
        const uint32_t cross = (uint32_t)(k > 1 ? dp_smear(Sall, k - 1) : 0ull) & 0xFFFFu;
        uint32_t exists = ~cross & 0xFFFFu;
        const int64_t left = n_bases - g0;
        if (left < 16) exists &= (left <= 0) ? 0u : (0xFFFFu << (16 - left));
        const uint32_t dirty = (uint32_t)dp_smear(U, k) & exists;  // window holds an undefined base: exact path
        const uint32_t clean = exists & ~dirty;
        // the read of a hit: index of the read holding the sub-span's first position (one binary search per warp,
        // only if the warp has a hit at all) + the read starts seen since then
        bool have_base = false;
        int64_t r_base = 0;
        int starts_before = 0;
        const uint32_t s16 = w2 & 0xFFFFu;  // bit 15-b = a read starts at position g0+b

#pragma unroll 1
        for (int b0 = 0; b0 < 16; b0 += 4) {
            const uint32_t cm = (clean >> (12 - b0)) & 0xFu, dm = (dirty >> (12 - b0)) & 0xFu;
            if (!__any_sync(0xFFFFFFFFu, (cm | dm) != 0)) continue;
            uint64_t keys[4], bk[4];
            uint32_t hs[4], fw[4];
            ulonglong2 k01[4], k23[4];
            bool probe[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int b = b0 + q;
                const int sh = 2 * (15 - b);
                const uint32_t klo = __funnelshift_r(f_0, f_m1, sh);
                const uint32_t khi = __funnelshift_r(f_m1, f_m2, sh);
                const uint64_t win = ((uint64_t)khi << 32) | klo;
                probe[q] = (cm >> (3 - q)) & 1u;
                uint64_t key = 0;
                if (probe[q]) {
                    const uint64_t kmer = win & p.mask;
                    key = bb_to_value(p, kmer, bb_rcomp(kmer, k), p.kmask);
                } else if ((dm >> (3 - q)) & 1u) {
                    const uint32_t dw = (uint32_t)(Dall >> (15 - b));  // bit t = base g-t
                    probe[q] = dp_dirty_key(p, win, dw, &key);
                }
                keys[q] = key;
                hs[q] = bb_fhash64(key);
                bk[q] = bb_bucket(hs[q], t.bucket_shift);
                if (probe[q] && big_words) fw[q] = dp_load_filter(bigf + bb_big_word(hs[q], big_words), pol_keep);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (big_words) probe[q] = probe[q] && ((fw[q] >> (hs[q] & 31u)) & 1u);
                if (probe[q]) dp_load_bucket(t.keys + 4 * bk[q], k01[q], k23[q]);
            }
            int ids[4];
            bool any_hit = false;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                ids[q] = probe[q] ? dp_resolve(t, bk[q] & bmask, k01[q], k23[q], keys[q]) : -1;
                any_hit |= ids[q] > 0;
            }
            if (!__any_sync(0xFFFFFFFFu, any_hit)) continue;
            if (!have_base) {
                int64_t rb = 0;
                if (lane == 0) {
                    int64_t a = 0, b = n_reads + 1;  // first index with offsets[idx] > g_lo
                    while (a < b) {
                        const int64_t m = (a + b) >> 1;
                        if ((int64_t)offsets[m] <= g_lo) a = m + 1;
                        else b = m;
                    }
                    rb = a - 1;
                }
                // the start bit of g_lo itself belongs to read r_base
                r_base = __shfl_sync(0xFFFFFFFFu, rb, 0) - (int64_t)((__shfl_sync(0xFFFFFFFFu, s16, 0) >> 15) & 1u);
                const int cnt = __popc(s16);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += v;
                }
                starts_before = incl - cnt;
                have_base = true;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (ids[q] <= 0) continue;
                const int b = b0 + q;
                dp_credit(offsets, n_reads, g0 + b, r_base + starts_before + __popc(s16 >> (15 - b)), ids[q], paired, p, first64,
                          lastpos);
            }
        }
    }
}

}  // namespace

bool plan_direct(const BBParams &p, const BBTable &t) {
    const bool mode_ok = (p.mode == MODE_KTRIM && !p.useShortKmers) || (p.mode == MODE_KFILTER && p.maxBadKmers0 == 0);
    if (!mode_ok) return false;
    if (p.qHammingDistance != 0 || p.speed != 0 || p.qSkip != 1 || p.restrictLeft != 0 || p.restrictRight != 0) return false;
    if (p.kbig > p.k || p.minKmerFraction != 0.0f || p.k < 2 || p.k > 31) return false;
    return t.stored > 0;
}

size_t direct_sbits_bytes(int64_t n_bases) { return (size_t)((n_bases + 31) / 32 + 2) * 4; }

int launch_direct(const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_reads, int64_t n_bases, int paired,
                  const BBParams &p, const BBTable &t, unsigned long long *d_first64, int *d_lastpos, uint16_t *d_sbits,
                  int sm_count, cudaStream_t st) {
    if (n_reads <= 0) return 0;
    int nl = 0;
    int *lastpos = (p.mode == MODE_KTRIM && p.ktrimLeft) ? d_lastpos : nullptr;
    dp_init_kernel<<<sm_count * 4, 256, 0, st>>>(d_first64, lastpos, n_reads);
    nl++;
    if (n_bases > 0) {
        if (cudaMemsetAsync(d_sbits, 0, direct_sbits_bytes(n_bases), st) != cudaSuccess) return -1;
        dp_starts_kernel<<<(unsigned)std::min<int64_t>((n_reads + 255) / 256, (int64_t)sm_count * 16), 256, 0, st>>>(
            d_offsets, n_reads, reinterpret_cast<uint32_t *>(d_sbits));
        nl++;
        const int64_t n_spans = (n_bases + DP_SPAN - 1) / DP_SPAN;
        const int blocks = (int)std::min<int64_t>(n_spans, (int64_t)sm_count * 3);
        bbduk_direct_kernel<<<blocks, DP_THREADS, 0, st>>>(d_bases, d_sbits, d_offsets, n_reads, n_bases, paired, p, t,
                                                          d_first64, lastpos);
        nl++;
    }
    return cudaGetLastError() == cudaSuccess ? nl : -1;
}
