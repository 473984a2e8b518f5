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
        if (g0 >= n_bases) continue;
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
        if (left < 16) exists &= 0xFFFFu << (16 - left);
        const uint32_t dirty = (uint32_t)dp_smear(U, k) & exists;  // window holds an undefined base: exact path
        const uint32_t clean = exists & ~dirty;

#pragma unroll
        for (int b0 = 0; b0 < 16; b0 += 4) {
            const uint32_t cm = (clean >> (12 - b0)) & 0xFu, dm = (dirty >> (12 - b0)) & 0xFu;
            if ((cm | dm) == 0) continue;
            uint64_t keys[4], bk[4];
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
                    probe[q] = bb_window_key(p, win, dw, &key);
                }
                keys[q] = key;
                bk[q] = bb_bucket(bb_fhash64(key), t.bucket_shift);
                if (probe[q]) {
                    const ulonglong2 *ptr = reinterpret_cast<const ulonglong2 *>(t.keys + 4 * bk[q]);
                    k01[q] = __ldg(ptr);
                    k23[q] = __ldg(ptr + 1);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (!probe[q]) continue;
                const int id = bb_table_get_from(t, bk[q] & bmask, k01[q], k23[q], keys[q]);
                if (id > 0) dp_hit(offsets, n_reads, g0 + b0 + q, id, paired, p, first64, lastpos);
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
