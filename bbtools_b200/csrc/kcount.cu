// kcount.cu -- KmerCountExact's counting path (include/kcount_b200.h) on sm_100a.
//
// Replaces kmer/KmerTableSet.java:652-716 (addKmersToTable) + kmer/HashArray1D.java:68-89 (increment).
// The reference walks each read serially with a rolling (kmer, rkmer, len) state; an undefined base resets
// all three. In closed form: position g of the concatenated base array ends a counted k-mer iff the k bases
// [g-k+1, g] are all defined and belong to one read. That makes every position independent, so the kernel is
// position-parallel over the flat base array (balanced for any read-length mix, including Mbp contigs):
//   * a pre-pass scatters one "a read starts here" bit per read into a bit stream S (2 bytes per 16 bases);
//   * a warp takes a sub-span of 512 positions, stages it (+32 bases of look-back) as a big-endian 2-bit stream
//     F, a "defined" bit stream D and S in shared memory (16-byte coalesced loads, SIMD-in-register
//     classification as in probe_fast.cu); warps never wait for each other;
//   * a thread owns 16 consecutive positions: validity of all 16 is decided bit-parallel (smears of ~D and S),
//     each window is two funnel shifts, the reverse complement a bit-reverse;
//   * insert = one 16-byte slot {key, count}: a k-mer costs one 32-byte sector read + one L2 atomic.
// The table is HBM-resident (config 5: ~1e9 distinct keys = 32 GB at load 0.5); a thread keeps the first-probe
// sectors of 4 windows in flight before the dependent compare/CAS/RED steps run.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/kcount_b200.h"
#include "bbduk_dev.cuh"

namespace {

constexpr uint64_t KC_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr int KC_THREADS = 256;
constexpr int KC_SPAN = KC_THREADS * 16;  // positions per CTA iteration
constexpr int KC_LOOK = 2;                // look-back chunks (32 bases >= k-1)
constexpr int KC_MAX_PROBE = 1 << 16;

struct __align__(16) KSlot {
    uint64_t key;
    uint32_t count;
    uint32_t pad;
};

struct KTable {
    KSlot *slots;
    uint64_t mask;   // n_slots - 1
    int shift;       // 64 - log2(n_slots)
};

// counters kept on the device: [0] kmers_in, [1] unique, [2] overflow flag
struct KCounters {
    unsigned long long kmers_in, unique, overflow, pad;
};

__device__ __forceinline__ uint64_t kc_slot_of(uint64_t key, int shift) {
    key ^= key >> 29;
    return (key * 0x9E3779B97F4A7C15ull) >> shift;
}
__host__ __device__ __forceinline__ uint64_t kc_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint32_t kc_owner(uint64_t key, uint32_t n_parts) {
    return (uint32_t)((kc_mix64(key) >> 32) % n_parts);
}

// count += incr with saturation at INT32_MAX (kmer/HashArray1D.java:74-75)
__device__ __forceinline__ void kc_add(uint32_t *c, uint32_t seen, uint32_t incr) {
    if (incr == 1u) {
        // the snapshot may lag by at most the number of threads in flight (< 2^31), so the 32-bit counter
        // cannot wrap; readers clamp to INT32_MAX
        if (seen < 0x7FFFFFFFu) atomicAdd(c, 1u);
        return;
    }
    uint32_t old = seen;
    while (true) {
        const uint64_t nv64 = (uint64_t)min(old, 0x7FFFFFFFu) + incr;
        const uint32_t nv = nv64 > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)nv64;
        const uint32_t got = atomicCAS(c, old, nv);
        if (got == old) return;
        old = got;
    }
}

// returns 1 if the key was created. Lookup/insert = incrementAndReturnNumCreated (kmer/HashArray1D.java:68-89).
// `v` is the caller's (possibly stale) 16-byte snapshot of slot s -- stale is fine: keys never change once set
// and the count is only a saturation hint.
__device__ __forceinline__ int kc_insert_from(const KTable &t, uint64_t s, ulonglong2 v, uint64_t key, uint32_t incr,
                                              unsigned long long *overflow) {
    for (int probe = 0; probe < KC_MAX_PROBE; probe++) {
        KSlot *p = t.slots + s;
        uint64_t kk = v.x;
        uint32_t seen = (uint32_t)v.y;
        int created = 0;
        if (kk == KC_EMPTY) {
            kk = atomicCAS(reinterpret_cast<unsigned long long *>(&p->key), (unsigned long long)KC_EMPTY,
                           (unsigned long long)key);
            if (kk == KC_EMPTY) {
                kk = key;
                created = 1;
            }
            seen = 0;  // a fresh slot: the count is far from saturation whoever created it
        }
        if (kk == key) {
            kc_add(&p->count, seen, incr);
            return created;
        }
        s = (s + 1) & t.mask;
        v = __ldcg(reinterpret_cast<const ulonglong2 *>(t.slots + s));
    }
    *overflow = 1ull;
    return 0;
}
__device__ __forceinline__ int kc_insert(const KTable &t, uint64_t key, uint32_t incr, unsigned long long *overflow) {
    const uint64_t s = kc_slot_of(key, t.shift);
    return kc_insert_from(t, s, __ldcg(reinterpret_cast<const ulonglong2 *>(t.slots + s)), key, incr, overflow);
}

__device__ __forceinline__ void kc_classify4(uint32_t w, uint32_t &codes, uint32_t &bad) {
    // same bit tests as probe_fast.cu classify4: exact ACGTUacgtu membership (dna/AminoAcid.java:1289-1320)
    codes = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    const uint32_t d = (w | 0x20202020u) ^ 0x61616161u;
    const uint32_t q = (d >> 2) & ~(d >> 1) & 0x01010101u;
    bad = (d & 0xE8E8E8E8u) | (((d >> 4) ^ q) & 0x01010101u) | (d & ~q & 0x01010101u);
}
__device__ __forceinline__ uint32_t kc_pack4(uint32_t codes) { return (codes * 0x40100401u) >> 24; }
__device__ __forceinline__ uint32_t kc_valid4(uint32_t bad) {
    const uint32_t nz = (((bad & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | bad) & 0x80808080u;
    return (((nz ^ 0x80808080u) >> 7) * 0x08040201u) >> 24;
}

// OR of x >> d for d in [0, n)
__device__ __forceinline__ uint64_t kc_smear(uint64_t x, int n) {
    int have = 1;
    while (have < n) {
        const int s = min(have, n - have);
        x |= x >> s;
        have += s;
    }
    return x;
}

// read-start bit stream: bit 15-b of the 16-bit word c set <=> some read starts at position 16c+b
__global__ void kc_starts_kernel(const uint32_t *__restrict__ offsets, int64_t n_reads, uint32_t *sbits) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t o = offsets[r];
        atomicOr(&sbits[o >> 5], 1u << ((((o >> 4) & 1u) << 4) + 15u - (o & 15u)));
    }
}

template <bool RCOMP>
__global__ void __launch_bounds__(KC_THREADS, 3)
kcount_kernel(const uint8_t *__restrict__ bases, const uint16_t *__restrict__ sbits, int64_t n_bases, int64_t span_lo,
              int64_t span_hi, int k, KTable t, KCounters *ctr) {
    // every warp works alone on 512-position sub-spans: no CTA-wide barrier ever waits on a DRAM chain
    __shared__ uint32_t Fs_all[KC_THREADS / 32][32 + KC_LOOK];
    __shared__ uint32_t DSs_all[KC_THREADS / 32][32 + KC_LOOK];  // D in the high half, S in the low half; bit 15-b = base b
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *Fs = Fs_all[warp], *DSs = DSs_all[warp];
    const int tid = lane;
    const uint64_t kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1ull);
    unsigned long long n_valid = 0, n_created = 0;

    for (int64_t span = span_lo + blockIdx.x; span < span_hi; span += gridDim.x) {
        const int64_t g_lo = span * KC_SPAN + warp * 512;  // first position of this warp's sub-span
        const int64_t c_lo = (g_lo >> 4) - KC_LOOK;        // first staged chunk (may be negative)
        __syncwarp();
        // ---- stage: chunk c_lo + i -> Fs[i], D and S bits ----------------------------------------
        for (int i = lane; i < 32 + KC_LOOK; i += 32) {
            const int64_t c = c_lo + i;
            uint32_t f = 0, dbits = 0, sb = 0;
            if (c >= 0 && c * 16 < n_bases) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(bases) + c);
                sb = __ldg(sbits + c);
                uint32_t cw[4], bw[4];
                kc_classify4(v.x, cw[0], bw[0]);
                kc_classify4(v.y, cw[1], bw[1]);
                kc_classify4(v.z, cw[2], bw[2]);
                kc_classify4(v.w, cw[3], bw[3]);
                f = (kc_pack4(cw[0]) << 24) | (kc_pack4(cw[1]) << 16) | (kc_pack4(cw[2]) << 8) | kc_pack4(cw[3]);
                dbits = 0xFFFFu;
                if ((bw[0] | bw[1] | bw[2] | bw[3]) != 0)
                    dbits = (kc_valid4(bw[0]) << 12) | (kc_valid4(bw[1]) << 8) | (kc_valid4(bw[2]) << 4) | kc_valid4(bw[3]);
                const int64_t rem = n_bases - c * 16;  // bases of this chunk inside the batch
                if (rem < 16) dbits &= 0xFFFFu << (16 - rem);
            }
            Fs[i] = f;
            DSs[i] = (dbits << 16) | sb;
        }
        __syncwarp();

        // ---- 16 positions per thread -----------------------------------------------------------
        const int64_t g0 = g_lo + 16 * tid;
        if (g0 < n_bases) {
            const uint32_t f_m2 = Fs[tid], f_m1 = Fs[tid + 1], f_0 = Fs[tid + 2];
            const uint32_t w0 = DSs[tid], w1 = DSs[tid + 1], w2 = DSs[tid + 2];
            // 48-base strings, base i of the string at bit 47-i
            const uint64_t Dall = ((uint64_t)(w0 >> 16) << 32) | ((uint64_t)(w1 >> 16) << 16) | (w2 >> 16);
            const uint64_t Sall = ((uint64_t)(w0 & 0xFFFFu) << 32) | ((uint64_t)(w1 & 0xFFFFu) << 16) | (w2 & 0xFFFFu);
            const uint64_t U = (~Dall) & 0xFFFFFFFFFFFFull;
            // an undefined base at i spoils windows ending in [i, i+k-1]; a read start at i those in [i, i+k-2]
            const uint64_t inval = kc_smear(U, k) | (k > 1 ? kc_smear(Sall, k - 1) : 0ull);
            uint32_t ok = (~(uint32_t)inval) & 0xFFFFu;  // bit 15-b = position g0+b ends a counted k-mer
            // the look-back of the very first chunks is "undefined", positions >= n_bases have D = 0
            n_valid += __popc(ok);
            // groups of 4 windows: keys, then the 4 first-probe slots are loaded together (4 sectors in
            // flight per thread; a deeper look-ahead would push the in-flight footprint past the L2), then the
            // dependent compare / CAS / RED steps run
#pragma unroll
            for (int b0 = 0; b0 < 16; b0 += 4) {
                if (((ok >> (12 - b0)) & 0xFu) == 0) continue;
                uint64_t keys[4], sl[4];
                ulonglong2 snap[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int sh = 2 * (15 - (b0 + q));
                    const uint32_t klo = __funnelshift_r(f_0, f_m1, sh);
                    const uint32_t khi = __funnelshift_r(f_m1, f_m2, sh);
                    const uint64_t kmer = (((uint64_t)khi << 32) | klo) & kmask;
                    uint64_t key = kmer;
                    if (RCOMP) {
                        const uint64_t rk = bb_rcomp(kmer, k);
                        key = rk > kmer ? rk : kmer;
                    }
                    keys[q] = key;
                    sl[q] = kc_slot_of(key, t.shift);
                    if ((ok >> (15 - (b0 + q))) & 1u) snap[q] = __ldcg(reinterpret_cast<const ulonglong2 *>(t.slots + sl[q]));
                }
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if ((ok >> (15 - (b0 + q))) & 1u) n_created += kc_insert_from(t, sl[q], snap[q], keys[q], 1u, &ctr->overflow);
            }
        }
    }
    // per-warp sums
    for (int o = 16; o > 0; o >>= 1) {
        n_valid += __shfl_xor_sync(0xFFFFFFFFu, n_valid, o);
        n_created += __shfl_xor_sync(0xFFFFFFFFu, n_created, o);
    }
    if ((tid & 31) == 0) {
        if (n_valid) atomicAdd(&ctr->kmers_in, n_valid);
        if (n_created) atomicAdd(&ctr->unique, n_created);
    }
}

__global__ void kc_fill_kernel(KSlot *slots, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        KSlot s;
        s.key = KC_EMPTY;
        s.count = 0;
        s.pad = 0;
        slots[i] = s;
    }
}

// re-insert every entry of `old` into t (resize, kmer/HashArray1D.java:260-339)
__global__ void kc_rehash_kernel(const KSlot *__restrict__ old, uint64_t n_old, KTable t, KCounters *ctr) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_old; i += (uint64_t)gridDim.x * blockDim.x) {
        const KSlot s = old[i];
        if (s.key != KC_EMPTY) kc_insert(t, s.key, min(s.count, 0x7FFFFFFFu), &ctr->overflow);
    }
}

__global__ void kc_merge_kernel(const uint64_t *__restrict__ keys, const int32_t *__restrict__ counts, int64_t n, KTable t,
                                KCounters *ctr) {
    unsigned long long created = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = counts[i];
        if (c > 0) created += kc_insert(t, keys[i], (uint32_t)c, &ctr->overflow);
    }
    for (int o = 16; o > 0; o >>= 1) created += __shfl_xor_sync(0xFFFFFFFFu, created, o);
    if ((threadIdx.x & 31) == 0 && created) atomicAdd(&ctr->unique, created);
}

constexpr int KC_HIST_SMEM = 2048;
__global__ void kc_hist_kernel(const KSlot *__restrict__ slots, uint64_t n, int histmax, unsigned long long *hist) {
    __shared__ unsigned int sh[KC_HIST_SMEM];
    for (int i = threadIdx.x; i < KC_HIST_SMEM; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const KSlot s = slots[i];
        if (s.key == KC_EMPTY) continue;
        const int c = (int)min(s.count, 0x7FFFFFFFu);
        const int bin = min(c, histmax);
        if (bin < KC_HIST_SMEM) atomicAdd(&sh[bin], 1u);
        else atomicAdd(&hist[bin], 1ull);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < KC_HIST_SMEM && i <= histmax; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

__global__ void kc_dump_kernel(const KSlot *__restrict__ slots, uint64_t n, int mincount, int maxcount, uint64_t *keys,
                               int32_t *counts, long long cap, unsigned long long *cursor) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const KSlot s = slots[i];
        if (s.key == KC_EMPTY) continue;
        const int c = (int)min(s.count, 0x7FFFFFFFu);
        if (c < mincount || c > maxcount) continue;
        const unsigned long long w = atomicAdd(cursor, 1ull);
        if ((long long)w < cap) {
            keys[w] = s.key;
            counts[w] = c;
        }
    }
}

// pass 0: sizes per owner; pass 1: scatter through per-owner cursors
__global__ void kc_export_kernel(const KSlot *__restrict__ slots, uint64_t n, uint32_t n_parts, int pass,
                                 unsigned long long *part_cursor, uint64_t *keys, int32_t *counts) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const KSlot s = slots[i];
        if (s.key == KC_EMPTY) continue;
        const uint32_t o = kc_owner(s.key, n_parts);
        const unsigned long long w = atomicAdd(&part_cursor[o], 1ull);
        if (pass == 1) {
            keys[w] = s.key;
            counts[w] = (int32_t)min(s.count, 0x7FFFFFFFu);
        }
    }
}

// ---- synthetic cfg-5 reads; same formulas as bbtools_b200/synth.py:genome_reads -----
__device__ __forceinline__ uint64_t kc_rnd(uint64_t seed, uint64_t stream, uint64_t idx) {
    return kc_mix64(seed + stream * 0x9E3779B97F4A7C15ull + idx * 0xD1B54A32D192ED03ull);
}
__global__ void kc_synth_kernel(uint8_t *bases, uint32_t *offsets, int64_t n_reads, int64_t first_read, int L,
                                int64_t genome_len, uint64_t seed, int sub_per_10k) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= n_reads) offsets[t] = (uint32_t)(t * L);
    if (t >= n_reads * L) return;
    const int64_t lr = t / L;
    const int j = (int)(t - lr * L);
    const uint64_t r = (uint64_t)(first_read + lr);
    const uint64_t c = kc_rnd(seed, 0, r);
    const bool rev = (c & 1ull) != 0;
    const uint64_t start = (c >> 8) % (uint64_t)(genome_len - L + 1);
    const uint64_t gi = start + (uint64_t)(rev ? (L - 1 - j) : j);
    uint32_t code = (uint32_t)(kc_rnd(seed, 5, gi) & 3ull);
    if (rev) code = 3u - code;
    const uint64_t e = kc_rnd(seed, 3, r * (uint64_t)L + (uint64_t)j);
    if ((int)(e % 10000ull) < sub_per_10k) code = (code + 1u + (uint32_t)((e >> 16) % 3ull)) & 3u;
    bases[t] = "ACGT"[code];
}

}  // namespace

struct kcount_handle {
    int k = 31, rcomp = 1, device = 0, sm_count = 148;
    KSlot *slots = nullptr;
    uint64_t n_slots = 0;
    KCounters *d_ctr = nullptr;
    int64_t reads_in = 0, bases_in = 0;
    int64_t unique_known = 0;    // last value read back from the device
    int64_t added_since = 0;     // positions processed since then (upper bound on new keys)
    std::atomic<int64_t> launches{0};
    std::mutex mu;
    std::string err;
    // read-start bit stream of the batch being counted (device entry point)
    uint16_t *d_sbits = nullptr;
    int64_t cap_sbits = 0;
    // host path: two staging sets so that the copy of chunk i+1 overlaps the counting of chunk i
    struct Stage {
        uint8_t *d_bases = nullptr;
        uint32_t *d_off = nullptr, *h_off = nullptr;  // h_off pinned
        uint16_t *d_sbits = nullptr;
        int64_t cap_bases = 0, cap_reads = 0, cap_sbits = 0;
        cudaStream_t st = nullptr;
        cudaEvent_t done = nullptr;
    } stage[2];
    cudaStream_t st = nullptr;
};

namespace {

thread_local std::string g_kerr;

int kerr(kcount_handle *h, const std::string &m) {
    if (h) h->err = m;
    g_kerr = m;
    return 1;
}

#define KCK(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            char b_[512];                                                                                  \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return kerr(h, b_);                                                                            \
        }                                                                                                  \
    } while (0)

KTable view(const kcount_handle *h) {
    KTable t;
    t.slots = h->slots;
    t.mask = h->n_slots - 1;
    int lg = 0;
    while ((1ull << lg) < h->n_slots) lg++;
    t.shift = 64 - lg;
    return t;
}

int alloc_table(kcount_handle *h, uint64_t n_slots, KSlot **out) {
    KSlot *p = nullptr;
    KCK(cudaMalloc(&p, sizeof(KSlot) * n_slots));
    kc_fill_kernel<<<h->sm_count * 8, 256>>>(p, n_slots);
    h->launches += 1;
    KCK(cudaGetLastError());
    *out = p;
    return 0;
}

int read_counters(kcount_handle *h, KCounters *c, cudaStream_t st) {
    KCK(cudaMemcpyAsync(c, h->d_ctr, sizeof *c, cudaMemcpyDeviceToHost, st));
    KCK(cudaStreamSynchronize(st));
    if (c->overflow) return kerr(h, "k-mer table overflow (probe bound exceeded)");
    h->unique_known = (int64_t)c->unique;
    h->added_since = 0;
    return 0;
}

int grow(kcount_handle *h, uint64_t new_slots, cudaStream_t st) {
    (void)st;
    KCK(cudaDeviceSynchronize());  // every stream that counts into the old table
    KSlot *ns = nullptr;
    if (alloc_table(h, new_slots, &ns)) return 1;
    KSlot *old = h->slots;
    const uint64_t n_old = h->n_slots;
    h->slots = ns;
    h->n_slots = new_slots;
    kc_rehash_kernel<<<h->sm_count * 8, 256>>>(old, n_old, view(h), h->d_ctr);
    h->launches += 1;
    KCK(cudaGetLastError());
    KCK(cudaDeviceSynchronize());
    KCK(cudaFree(old));
    return 0;
}

// make room for up to `incoming` new keys; returns the number of positions that may be processed now
int reserve(kcount_handle *h, int64_t incoming, int64_t *allowed, cudaStream_t st) {
    auto limit = [&]() { return (int64_t)(h->n_slots / 10 * 7); };
    if (h->unique_known + h->added_since + incoming <= limit()) {
        *allowed = incoming;
        return 0;
    }
    KCounters c;
    KCK(cudaDeviceSynchronize());  // kernels of the other staging stream also create keys
    if (read_counters(h, &c, st)) return 1;
    int64_t room = limit() - h->unique_known;
    // grow while less than 1/8 of the table is free for this launch
    while (room < incoming && room < (int64_t)(h->n_slots / 8)) {
        if (grow(h, h->n_slots * 2, st)) return 1;
        room = limit() - h->unique_known;
    }
    *allowed = std::min(incoming, room);
    return 0;
}

int count_device(kcount_handle *h, const uint8_t *d_bases, const uint32_t *d_off, int64_t n_reads, int64_t n_bases,
                 cudaStream_t st, uint16_t **sbits, int64_t *cap_sbits) {
    if (n_bases <= 0) return 0;
    if (reinterpret_cast<uintptr_t>(d_bases) & 15) return kerr(h, "d_bases must be 16-byte aligned");
    const int64_t n_spans = (n_bases + KC_SPAN - 1) / KC_SPAN;
    // read-start bit stream for this batch (2 bytes per 16 bases)
    const int64_t sb_bytes = ((n_bases + 31) / 32 + 2) * 4;
    if (sb_bytes > *cap_sbits) {
        KCK(cudaStreamSynchronize(st));
        cudaFree(*sbits);
        *sbits = nullptr;
        *cap_sbits = sb_bytes + sb_bytes / 8 + 4096;
        KCK(cudaMalloc(sbits, (size_t)*cap_sbits));
    }
    KCK(cudaMemsetAsync(*sbits, 0, (size_t)sb_bytes, st));
    kc_starts_kernel<<<(unsigned)std::min<int64_t>((n_reads + 255) / 256, (int64_t)h->sm_count * 16), 256, 0, st>>>(
        d_off, n_reads, reinterpret_cast<uint32_t *>(*sbits));
    h->launches += 1;
    int64_t span = 0;
    while (span < n_spans) {
        int64_t allowed = 0;
        const int64_t want = std::min(n_bases, (n_spans - span) * (int64_t)KC_SPAN);
        if (reserve(h, want, &allowed, st)) return 1;
        int64_t take = std::max<int64_t>(1, allowed / KC_SPAN);
        take = std::min(take, n_spans - span);
        const int blocks = (int)std::min<int64_t>(take, (int64_t)h->sm_count * 8);
        if (h->rcomp)
            kcount_kernel<true><<<blocks, KC_THREADS, 0, st>>>(d_bases, *sbits, n_bases, span, span + take, h->k, view(h), h->d_ctr);
        else
            kcount_kernel<false><<<blocks, KC_THREADS, 0, st>>>(d_bases, *sbits, n_bases, span, span + take, h->k, view(h), h->d_ctr);
        h->launches += 1;
        KCK(cudaGetLastError());
        h->added_since += take * (int64_t)KC_SPAN;
        span += take;
    }
    return 0;
}

}  // namespace

extern "C" {

const char *kcount_b200_last_error(kcount_handle *h) {
    if (h) g_kerr = h->err;
    return g_kerr.c_str();
}

int kcount_b200_create(int32_t k, int32_t rcomp, int64_t initial_keys, int32_t device, kcount_handle **out) {
    if (!out) return kerr(nullptr, "out is NULL");
    *out = nullptr;
    if (k < 1 || k > 31) return kerr(nullptr, "k must be in [1,31] (longer k-mers use the reference's KmerTableSetU path)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1)
        return kerr(nullptr, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libbbduk_b200 has no CPU fallback)");
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) device = 0;
    if (device >= ndev) return kerr(nullptr, "device out of range");
    kcount_handle *h = new kcount_handle();
    h->k = k;
    h->rcomp = rcomp ? 1 : 0;
    h->device = device;
    auto fail = [&](const char *m) {
        std::string msg = m;
        kcount_b200_destroy(h);
        return kerr(nullptr, msg);
    };
    if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice failed");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    // random 16-byte slot accesses: ask the L2 not to fetch more than the touched sector (a hint; may be ignored)
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    uint64_t n = 1 << 16;
    while (initial_keys > 0 && (int64_t)(n / 2) < initial_keys) n <<= 1;
    if (cudaMalloc(&h->d_ctr, sizeof(KCounters)) != cudaSuccess) return fail("cudaMalloc failed");
    cudaMemset(h->d_ctr, 0, sizeof(KCounters));
    if (alloc_table(h, n, &h->slots)) return fail("table allocation failed");
    h->n_slots = n;
    if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess) return fail("stream creation failed");
    if (cudaDeviceSynchronize() != cudaSuccess) return fail("table initialisation failed");
    *out = h;
    return 0;
}

int kcount_b200_add_reads_device(kcount_handle *h, const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_reads,
                                 int64_t n_bases, void *stream) {
    if (!h) return kerr(nullptr, "handle is NULL");
    if (n_reads < 0 || n_bases < 0 || n_bases >= (1ll << 32)) return kerr(h, "bad sizes (total bases must be < 4 GiB per call)");
    if (n_reads == 0) return 0;
    std::lock_guard<std::mutex> g(h->mu);
    KCK(cudaSetDevice(h->device));
    h->reads_in += n_reads;
    h->bases_in += n_bases;
    return count_device(h, d_bases, d_offsets, n_reads, n_bases, (cudaStream_t)stream, &h->d_sbits, &h->cap_sbits);
}

int kcount_b200_add_reads(kcount_handle *h, const uint8_t *bases, const int64_t *offsets, int64_t n_reads) {
    if (!h) return kerr(nullptr, "handle is NULL");
    if (n_reads < 0 || (n_reads > 0 && (!bases || !offsets))) return kerr(h, "bad add_reads arguments");
    if (n_reads == 0) return 0;
    std::lock_guard<std::mutex> g(h->mu);
    KCK(cudaSetDevice(h->device));
    const int64_t CH_READS = 1 << 21, CH_BYTES = 1ll << 29;
    int64_t r0 = 0;
    int chunk_no = 0;
    while (r0 < n_reads) {
        int64_t r1 = std::min(n_reads, r0 + CH_READS);
        while (r1 > r0 + 1 && offsets[r1] - offsets[r0] > CH_BYTES) r1 = r0 + (r1 - r0) / 2;
        const int64_t nb = offsets[r1] - offsets[r0], nr = r1 - r0;
        if (nb < 0 || nb >= (1ll << 32) - 64) return kerr(h, "a single read exceeds 4 GiB (or offsets decrease)");
        kcount_handle::Stage &sg = h->stage[chunk_no++ & 1];
        if (!sg.st) {
            KCK(cudaStreamCreateWithFlags(&sg.st, cudaStreamNonBlocking));
            KCK(cudaEventCreateWithFlags(&sg.done, cudaEventDisableTiming));
        }
        KCK(cudaEventSynchronize(sg.done));  // the previous chunk staged here has been counted
        if (nb + 64 > sg.cap_bases) {
            cudaFree(sg.d_bases);
            sg.d_bases = nullptr;
            sg.cap_bases = nb + nb / 8 + 4096;
            KCK(cudaMalloc(&sg.d_bases, (size_t)sg.cap_bases));
        }
        if (nr + 1 > sg.cap_reads) {
            cudaFree(sg.d_off);
            cudaFreeHost(sg.h_off);
            sg.d_off = sg.h_off = nullptr;
            sg.cap_reads = nr + nr / 8 + 1024;
            KCK(cudaMalloc(&sg.d_off, sizeof(uint32_t) * sg.cap_reads));
            KCK(cudaHostAlloc(&sg.h_off, sizeof(uint32_t) * sg.cap_reads, cudaHostAllocDefault));
        }
        for (int64_t i = 0; i <= nr; i++) {
            if (i > 0 && offsets[r0 + i] < offsets[r0 + i - 1]) return kerr(h, "offsets must be non-decreasing");
            sg.h_off[i] = (uint32_t)(offsets[r0 + i] - offsets[r0]);
        }
        KCK(cudaMemcpyAsync(sg.d_bases, bases + offsets[r0], (size_t)nb, cudaMemcpyHostToDevice, sg.st));
        KCK(cudaMemcpyAsync(sg.d_off, sg.h_off, sizeof(uint32_t) * (nr + 1), cudaMemcpyHostToDevice, sg.st));
        h->reads_in += nr;
        h->bases_in += nb;
        if (count_device(h, sg.d_bases, sg.d_off, nr, nb, sg.st, &sg.d_sbits, &sg.cap_sbits)) return 1;
        KCK(cudaEventRecord(sg.done, sg.st));
        r0 = r1;
    }
    for (auto &sg : h->stage)
        if (sg.st) KCK(cudaStreamSynchronize(sg.st));
    return 0;
}

int kcount_b200_stats(kcount_handle *h, int64_t *v) {
    if (!h || !v) return kerr(h, "NULL argument");
    std::lock_guard<std::mutex> g(h->mu);
    KCK(cudaSetDevice(h->device));
    KCK(cudaDeviceSynchronize());
    KCounters c;
    if (read_counters(h, &c, h->st)) return 1;
    v[0] = h->reads_in;
    v[1] = h->bases_in;
    v[2] = (int64_t)c.kmers_in;
    v[3] = (int64_t)c.unique;
    return 0;
}

int kcount_b200_khist(kcount_handle *h, int32_t histmax, int64_t *hist) {
    if (!h || !hist) return kerr(h, "NULL argument");
    if (histmax < 1) return kerr(h, "histmax must be >= 1");
    std::lock_guard<std::mutex> g(h->mu);
    KCK(cudaSetDevice(h->device));
    KCK(cudaDeviceSynchronize());
    unsigned long long *d_hist = nullptr;
    const size_t nb = sizeof(unsigned long long) * ((size_t)histmax + 1);
    KCK(cudaMalloc(&d_hist, nb));
    cudaMemset(d_hist, 0, nb);
    kc_hist_kernel<<<h->sm_count * 8, 256>>>(h->slots, h->n_slots, histmax, d_hist);
    h->launches += 1;
    cudaError_t e = cudaMemcpy(hist, d_hist, nb, cudaMemcpyDeviceToHost);
    cudaFree(d_hist);
    if (e != cudaSuccess) return kerr(h, std::string("khist failed: ") + cudaGetErrorString(e));
    return 0;
}

int kcount_b200_dump(kcount_handle *h, int32_t mincount, int32_t maxcount, uint64_t *keys, int32_t *counts, int64_t cap,
                     int64_t *n_out) {
    if (!h || !n_out || cap < 0 || (cap > 0 && (!keys || !counts))) return kerr(h, "bad dump arguments");
    std::lock_guard<std::mutex> g(h->mu);
    KCK(cudaSetDevice(h->device));
    KCK(cudaDeviceSynchronize());
    uint64_t *dk = nullptr;
    int32_t *dc = nullptr;
    unsigned long long *cur = nullptr;
    KCK(cudaMalloc(&dk, sizeof(uint64_t) * std::max<int64_t>(cap, 1)));
    KCK(cudaMalloc(&dc, sizeof(int32_t) * std::max<int64_t>(cap, 1)));
    KCK(cudaMalloc(&cur, sizeof(unsigned long long)));
    cudaMemset(cur, 0, sizeof(unsigned long long));
    kc_dump_kernel<<<h->sm_count * 8, 256>>>(h->slots, h->n_slots, mincount, maxcount, dk, dc, cap, cur);
    h->launches += 1;
    unsigned long long n = 0;
    cudaError_t e = cudaMemcpy(&n, cur, sizeof n, cudaMemcpyDeviceToHost);
    const int64_t m = std::min<int64_t>((int64_t)n, cap);
    if (e == cudaSuccess && m > 0) e = cudaMemcpy(keys, dk, sizeof(uint64_t) * m, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && m > 0) e = cudaMemcpy(counts, dc, sizeof(int32_t) * m, cudaMemcpyDeviceToHost);
    cudaFree(dk);
    cudaFree(dc);
    cudaFree(cur);
    if (e != cudaSuccess) return kerr(h, std::string("dump failed: ") + cudaGetErrorString(e));
    *n_out = (int64_t)n;
    return 0;
}

int kcount_b200_export_partitioned(kcount_handle *h, int32_t n_parts, uint64_t *d_keys, int32_t *d_counts,
                                   int64_t *part_sizes, void *stream) {
    if (!h || !part_sizes || n_parts < 1 || n_parts > 4096) return kerr(h, "bad export arguments");
    std::lock_guard<std::mutex> g(h->mu);
    KCK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    KCK(cudaDeviceSynchronize());
    unsigned long long *cur = nullptr;
    KCK(cudaMalloc(&cur, sizeof(unsigned long long) * n_parts));
    cudaMemsetAsync(cur, 0, sizeof(unsigned long long) * n_parts, st);
    kc_export_kernel<<<h->sm_count * 8, 256, 0, st>>>(h->slots, h->n_slots, (uint32_t)n_parts, 0, cur, nullptr, nullptr);
    h->launches += 1;
    std::vector<unsigned long long> sizes(n_parts), starts(n_parts);
    cudaError_t e = cudaMemcpyAsync(sizes.data(), cur, sizeof(unsigned long long) * n_parts, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) {
        unsigned long long acc = 0;
        for (int i = 0; i < n_parts; i++) {
            part_sizes[i] = (int64_t)sizes[i];
            starts[i] = acc;
            acc += sizes[i];
        }
        if (acc > 0 && d_keys && d_counts) {
            e = cudaMemcpyAsync(cur, starts.data(), sizeof(unsigned long long) * n_parts, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) {
                kc_export_kernel<<<h->sm_count * 8, 256, 0, st>>>(h->slots, h->n_slots, (uint32_t)n_parts, 1, cur, d_keys, d_counts);
                h->launches += 1;
                e = cudaStreamSynchronize(st);
            }
        }
    }
    cudaFree(cur);
    if (e != cudaSuccess) return kerr(h, std::string("export failed: ") + cudaGetErrorString(e));
    return 0;
}

int kcount_b200_merge_device(kcount_handle *h, const uint64_t *d_keys, const int32_t *d_counts, int64_t n, void *stream) {
    if (!h || n < 0 || (n > 0 && (!d_keys || !d_counts))) return kerr(h, "bad merge arguments");
    if (n == 0) return 0;
    std::lock_guard<std::mutex> g(h->mu);
    KCK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    int64_t done = 0;
    while (done < n) {
        int64_t allowed = 0;
        if (reserve(h, n - done, &allowed, st)) return 1;
        const int64_t take = std::max<int64_t>(1, allowed);
        kc_merge_kernel<<<h->sm_count * 8, 256, 0, st>>>(d_keys + done, d_counts + done, take, view(h), h->d_ctr);
        h->launches += 1;
        KCK(cudaGetLastError());
        h->added_since += take;
        done += take;
    }
    return 0;
}

int kcount_b200_table_info(kcount_handle *h, int64_t *v) {
    if (!h || !v) return kerr(h, "NULL argument");
    v[0] = (int64_t)h->n_slots;
    v[1] = (int64_t)(h->n_slots * sizeof(KSlot));
    v[2] = h->launches.load();
    return 0;
}

int kcount_b200_synth_reads(uint8_t *d_bases, uint32_t *d_offsets, int64_t n_reads, int64_t first_read, int32_t read_len,
                            int64_t genome_len, uint64_t seed, int32_t sub_per_10k, void *stream) {
    if (n_reads <= 0 || read_len <= 0 || genome_len < read_len) return 1;
    if (n_reads * (int64_t)read_len >= (1ll << 32)) return 1;
    const int64_t threads = std::max<int64_t>(n_reads * read_len, n_reads + 1);
    kc_synth_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_bases, d_offsets, n_reads, first_read, read_len, genome_len, seed, sub_per_10k);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

void kcount_b200_destroy(kcount_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->slots);
    cudaFree(h->d_ctr);
    for (auto &sg : h->stage) {
        cudaFree(sg.d_bases);
        cudaFree(sg.d_off);
        cudaFree(sg.d_sbits);
        cudaFreeHost(sg.h_off);
        if (sg.done) cudaEventDestroy(sg.done);
        if (sg.st) cudaStreamDestroy(sg.st);
    }
    cudaFree(h->d_sbits);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

}  // extern "C"
