// seal.cu -- Seal's multi-value k-mer table and per-pair assignment on B200 (include/seal_b200.h).
//
// Table (replaces kmer.HashArrayHybridFast behind jgi/Seal.java:1760-1946): the loader emits every (key, id) entry
// of every reference k-mer (with its Hamming ball) position-parallel, two stable radix sorts order them by key then
// id, duplicates are dropped, and the distinct keys go into a bucketed hash array (four slots per bucket, keys and
// values in one 64-byte line). A key with one id stores the id itself (one line per lookup, the common case); a key with
// several stores -(p+2), p = its first entry in the sorted id array whose last entry of a list carries bit 31.
// A key's ids are ascending, which is the order the reference's per-way loader appends them in.
//
// Matching (replaces ProcessThread's matching block, jgi/Seal.java:2186-2276): ONE WARP PER PAIR. The read is staged
// in shared memory, every lane rolls the reference's kmer / rkmer / len registers over its own short run of positions
// after a k-base warm-up (k steps flush the registers, so the state is the reference's), the 32-byte probes of a warp
// are in flight together, and the lookups land in a per-warp hit buffer in position order. The warp then folds the
// buffer into the pair's list of distinct ids in first-seen order with counts (the reference's idList + countArray),
// batching equal single-id hits of 32 positions into one update, and applies the clear zone, the minimum hit rule and
// the ambiguous mode collectively. Lists live in shared memory (128 ids behind a per-warp hash; a key's value list is
// added 32 ids at a time, one per lane) and spill to a per-warp global scratch.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <atomic>
#include <cub/cub.cuh>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/seal_b200.h"
#include "bbduk_dev.cuh"

namespace {

namespace cg = cooperative_groups;

struct SealParams {
    int32_t k, rcomp, forbidNs, minlen2, shift2;
    int32_t speed, qskip, restrictLeft, restrictRight;
    int32_t ambig, match, kpt, clearzone, minKmerHits, ids_stride;
    float czf, mkf;
    uint64_t mask, middleMask, kmask;
};

// Hash array of Seal's table: buckets of four slots, keys and values side by side in 64 bytes (a probe of this table is a random
// DRAM access, and an L2 miss fetches a whole 128-byte line, profiles/r01_e_random_access_microbench.txt: with the BBDuk layout
// -- keys and values in separate arrays -- every hit cost two lines). Linear probing over buckets; slots fill in order and nothing is deleted, so an empty last slot ends
// the search (same rule as bb_table_get).
struct __align__(64) SlBucket {
    uint64_t keys[4];
    int32_t vals[4];
    int32_t pad[4];
};
struct SealTable {
    const SlBucket *bk;
    uint64_t bmask;           // buckets - 1 (power of two)
    uint32_t bucket_shift;    // bucket = hash32 >> bucket_shift
    const int32_t *ent_ids;   // sorted entries' ids, bit 31 = last id of its key
};
__device__ __forceinline__ int32_t sl_table_get(const SealTable &t, uint64_t key) {
    uint64_t b = bb_bucket(bb_fhash64(key), t.bucket_shift);
#pragma unroll 1
    for (int probe = 0; probe < BB_MAX_PROBE / 4; probe++) {
        const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(t.bk[b].keys);
        const ulonglong2 k01 = __ldg(q), k23 = __ldg(q + 1);
        const int4 v = __ldg(reinterpret_cast<const int4 *>(t.bk[b].vals));  // same line: in flight with the keys
        if (k01.x == key) return v.x;
        if (k01.y == key) return v.y;
        if (k23.x == key) return v.z;
        if (k23.y == key) return v.w;
        if (k23.y == BB_EMPTY_KEY) return 0;
        b = (b + 1) & t.bmask;
    }
    return 0;
}

constexpr int SL_WARPS = 8;          // warps per block
constexpr int SL_CH = 160;           // positions per chunk (five per lane for 150 bp reads)
constexpr int SL_LCAP = 128;         // list entries in shared memory
constexpr int SL_HASH_BITS = 8;
constexpr int SL_HASH = 1 << SL_HASH_BITS;  // slots of the per-warp hash over them
constexpr int SL_SPILL = 1024;       // list entries in the per-warp global scratch
constexpr int SL_GHASH_BITS = 11;
constexpr int SL_GHASH = 1 << SL_GHASH_BITS;  // slots of the per-warp global hash over them
constexpr int SL_SCRATCH = 4 * SL_SPILL + 2 * SL_GHASH;  // int32 words of global scratch per warp
#ifndef SL_BLOCKS
#define SL_BLOCKS 3
#endif
constexpr int SL_BLOCKS_PER_SM = SL_BLOCKS;  // 80 registers x 256 threads: three blocks are resident

__device__ __forceinline__ uint64_t sl_to_value(const SealParams &p, uint64_t kmer, uint64_t rkmer) {
    const uint64_t v = p.rcomp ? (kmer > rkmer ? kmer : rkmer) : kmer;
    return (v & p.middleMask) | p.kmask;
}
__device__ __forceinline__ bool sl_passes_speed(int speed, uint64_t key) {
    return speed < 1 || (int)((key & 0x7FFFFFFFFFFFFFFFull) % 17ull) >= speed;  // jgi/Seal.java:2983-2985
}

// ---- loader ------------------------------------------------------------------------------------------------------
// One thread per reference base i: the k-mer ending at i exists iff the k bases are defined (x<0 resets len,
// jgi/Seal.java:1790, :1809); rskip keeps it iff the run of defined bases ending at i is a multiple of skip (:1794).
__global__ void sl_seed_kernel(const uint8_t *__restrict__ bases, const int64_t *__restrict__ offsets, int32_t n_seqs,
                               int32_t first_id, int64_t total, int k, int skip, uint64_t *seed_kmer, int32_t *seed_id,
                               unsigned long long *ctr /* [0] seeds, [1] refKmers */) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    int32_t lo = 0, hi = n_seqs;  // offsets[lo] <= g < offsets[hi]
    while (hi - lo > 1) {
        const int32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= g) lo = mid;
        else hi = mid;
    }
    const int64_t s0 = offsets[lo], s1 = offsets[lo + 1];
    const int64_t i = g - s0, L = s1 - s0;
    if (L < k || i < k - 1) return;
    uint64_t kmer = 0;
    for (int j = 0; j < k; j++) {
        const uint32_t c = bases[g - (k - 1) + j];
        if (!bb_defined(c)) return;
        kmer = (kmer << 2) | bb_code_raw(c);
    }
    {
        cg::coalesced_group grp = cg::coalesced_threads();
        if (grp.thread_rank() == 0) atomicAdd(ctr + 1, (unsigned long long)grp.size());
    }
    if (skip > 1) {
        int64_t len = k, q = g - k;
        while (q >= s0 && bb_defined(bases[q])) {
            len++;
            q--;
        }
        if (len % skip != 0) return;
    }
    cg::coalesced_group grp = cg::coalesced_threads();
    unsigned long long w = 0;
    if (grp.thread_rank() == 0) w = atomicAdd(ctr, (unsigned long long)grp.size());
    w = grp.shfl(w, 0) + grp.thread_rank();
    seed_kmer[w] = kmer;
    seed_id[w] = first_id + lo;
}

// Hamming ball of every seed (jgi/Seal.java:1890-1918 mutate, substitutions only): thread = (seed, first substitution);
// the second level enumerates later positions only -- the set of keys is what the recursion's visits add up to.
__global__ void sl_expand_kernel(const uint64_t *__restrict__ seed_kmer, const int32_t *__restrict__ seed_id, int64_t n_seeds,
                                 int hdist, SealParams p, uint64_t *keys, uint32_t *ids, unsigned long long *cursor) {
    const int k = p.k;
    const int n1 = hdist > 0 ? 1 + 3 * k : 1;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_seeds * n1) return;
    const int64_t s = t / n1;
    const int o = (int)(t - s * n1);
    const uint64_t kmer = seed_kmer[s];
    const uint32_t id = (uint32_t)seed_id[s];
    if (o == 0) {
        const uint64_t key = sl_to_value(p, kmer, bb_rcomp(kmer, k));
        if (hdist == 0 && !sl_passes_speed(p.speed, key)) return;  // failsSpeed only on the hdist==0 branch (:1848)
        const unsigned long long w = atomicAdd(cursor, 1ull);
        keys[w] = key;
        ids[w] = id;
        return;
    }
    const int pos = (o - 1) / 3, alt = (o - 1) % 3 + 1;
    const uint64_t x = kmer ^ ((uint64_t)alt << (2 * pos));
    const int cnt = 1 + (hdist > 1 ? 3 * (k - 1 - pos) : 0);
    unsigned long long w = atomicAdd(cursor, (unsigned long long)cnt);
    keys[w] = sl_to_value(p, x, bb_rcomp(x, k));
    ids[w] = id;
    w++;
    if (hdist > 1) {
        for (int pos2 = pos + 1; pos2 < k; pos2++) {
            for (int a2 = 1; a2 < 4; a2++) {
                const uint64_t y = x ^ ((uint64_t)a2 << (2 * pos2));
                keys[w] = sl_to_value(p, y, bb_rcomp(y, k));
                ids[w] = id;
                w++;
            }
        }
    }
}

__global__ void sl_keep_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ ids, int64_t n, uint32_t *keep) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keep[i] = (i == 0 || keys[i] != keys[i - 1] || ids[i] != ids[i - 1]) ? 1u : 0u;
}
__global__ void sl_scatter_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ ids, const uint32_t *__restrict__ keep,
                                  const uint32_t *__restrict__ pos, int64_t n, uint64_t *okeys, int32_t *oids) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (keep[i]) {
            okeys[pos[i]] = keys[i];
            oids[pos[i]] = (int32_t)ids[i];
        }
}
__global__ void sl_count_heads_kernel(const uint64_t *__restrict__ keys, int64_t n, unsigned long long *heads) {
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        c += (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
    if (c) atomicAdd(heads, c);
}
__global__ void sl_fill_kernel(SlBucket *bk, int64_t n_buckets) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 4 * n_buckets; i += (int64_t)gridDim.x * blockDim.x) {
        bk[i >> 2].keys[i & 3] = BB_EMPTY_KEY;
        bk[i >> 2].vals[i & 3] = 0;
        bk[i >> 2].pad[i & 3] = 0;
    }
}
// one thread per sorted entry; the first entry of a key inserts it. Distinct keys only, so every put creates its slot.
__global__ void sl_insert_kernel(const uint64_t *__restrict__ ekeys, int32_t *eids, int64_t n, SlBucket *bk, uint64_t bmask,
                                 uint32_t bucket_shift, int *overflow) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = ekeys[i];
        if (i > 0 && ekeys[i - 1] == key) continue;
        int64_t m = 1;
        while (i + m < n && ekeys[i + m] == key) m++;
        int32_t val;
        if (m == 1) val = eids[i];
        else {
            val = (int32_t)(-(i + 2));
            eids[i + m - 1] |= (int32_t)0x80000000;
        }
        uint64_t b = bb_bucket(bb_fhash64(key), bucket_shift);
        bool placed = false;
        for (int probe = 0; probe < BB_MAX_PROBE / 4 && !placed; probe++) {
            for (int j = 0; j < 4 && !placed; j++) {  // in slot order, so a bucket's slots fill in order
                if (bk[b].keys[j] != BB_EMPTY_KEY) continue;
                const unsigned long long old =
                    atomicCAS((unsigned long long *)&bk[b].keys[j], (unsigned long long)BB_EMPTY_KEY, (unsigned long long)key);
                if (old == BB_EMPTY_KEY) {
                    bk[b].vals[j] = val;
                    placed = true;
                }
            }
            b = (b + 1) & bmask;
        }
        if (!placed) *overflow = 1;
    }
}
__global__ void sl_unmark_kernel(const int32_t *__restrict__ in, int32_t *out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i] & 0x7FFFFFFF;
}

// ---- matching ----------------------------------------------------------------------------------------------------
// A unit's distinct ids in first-seen order with their counts (the reference's idList + countArray): entries
// [0, SL_LCAP) in shared memory behind a per-warp open-addressed hash id -> entry (SL_HASH slots, at most half full),
// the rest in the warp's global scratch, searched linearly. n / last_* / overflow are warp-uniform.
struct SlList {
    int32_t *s_id, *s_cnt;  // shared
    int32_t *g_id, *g_cnt;  // global spill (read with ld.cg, counted with atomics: never through a stale L1 line)
    int32_t *h_id, *h_j;    // shared hash over the entries below SL_LCAP (the list being filled owns it)
    int32_t *gh_id, *gh_j;  // global hash over the spilled entries (same owner)
    int n, last_id, last_j, overflow, g_dirty;

    __device__ __forceinline__ void clear_global_hash(int lane) {
        for (int s = lane; s < SL_GHASH; s += 32) gh_id[s] = 0;
        g_dirty = 0;
    }
    __device__ __forceinline__ void reset(int lane) {
        n = 0;
        last_id = 0;
        last_j = -1;
        for (int s = lane; s < SL_HASH; s += 32) h_id[s] = 0;
        if (g_dirty) clear_global_hash(lane);
        __syncwarp();
    }
    __device__ __forceinline__ int32_t id_at(int j) const { return j < SL_LCAP ? s_id[j] : __ldcg(g_id + (j - SL_LCAP)); }
    __device__ __forceinline__ int32_t cnt_at(int j) const { return j < SL_LCAP ? s_cnt[j] : __ldcg(g_cnt + (j - SL_LCAP)); }
    __device__ __forceinline__ void bump(int j, int c) {
        if (j < SL_LCAP) s_cnt[j] += c;
        else atomicAdd(g_cnt + (j - SL_LCAP), c);
    }
    __device__ __forceinline__ void bump_atomic(int j, int c) {
        if (j < SL_LCAP) atomicAdd(s_cnt + j, c);
        else atomicAdd(g_cnt + (j - SL_LCAP), c);
    }
    __device__ __forceinline__ void put(int j, int32_t id) {  // ids are >= 1, 0 = empty; lanes of one batch may race for a slot
        if (j < SL_LCAP) {
            s_id[j] = id;
            s_cnt[j] = 0;
            unsigned slot = ((unsigned)id * 0x9E3779B1u) >> (32 - SL_HASH_BITS);
            for (;;) {
                const int old = atomicCAS(h_id + slot, 0, id);
                if (old == 0) {
                    h_j[slot] = j;
                    break;
                }
                slot = (slot + 1) & (SL_HASH - 1);
            }
        } else {
            g_id[j - SL_LCAP] = id;
            g_cnt[j - SL_LCAP] = 0;
            unsigned slot = ((unsigned)id * 0x9E3779B1u) >> (32 - SL_GHASH_BITS);
            for (;;) {
                const int old = atomicCAS(gh_id + slot, 0, id);
                if (old == 0) {
                    gh_j[slot] = j;
                    break;
                }
                slot = (slot + 1) & (SL_GHASH - 1);
            }
        }
    }
    // entry of `id` among the hashed entries in shared memory, -1 if absent (per lane, any id)
    __device__ __forceinline__ int find_hashed(int32_t id) const {
        unsigned slot = ((unsigned)id * 0x9E3779B1u) >> (32 - SL_HASH_BITS);
        for (;;) {
            const int32_t x = h_id[slot];
            if (x == id) return h_j[slot];
            if (x == 0) return -1;
            slot = (slot + 1) & (SL_HASH - 1);
        }
    }
    // entry of `id` anywhere in the list, -1 if absent (per lane, any id)
    __device__ __forceinline__ int find_any(int32_t id) const {
        const int j = find_hashed(id);
        if (j >= 0 || n <= SL_LCAP) return j;
        unsigned slot = ((unsigned)id * 0x9E3779B1u) >> (32 - SL_GHASH_BITS);
        for (;;) {
            const int32_t x = __ldcg(gh_id + slot);
            if (x == id) return __ldcg(gh_j + slot);
            if (x == 0) return -1;
            slot = (slot + 1) & (SL_GHASH - 1);
        }
    }
    // hits[id] += c for one warp-uniform id; first sight appends (jgi/Seal.java:2895-2898)
    __device__ __forceinline__ void add(int32_t id, int c, int lane) {
        if (id != last_id) {
            int found = find_any(id);
            if (found < 0) {
                if (n >= SL_LCAP + SL_SPILL) {
                    overflow = 1;
                    return;
                }
                found = n++;
                if (found >= SL_LCAP) g_dirty = 1;
                if (lane == 0) put(found, id);
            }
            last_id = id;
            last_j = found;
        }
        if (lane == 0) bump(last_j, c);
        __syncwarp();
    }
    // hits[id] += c for every id of one key's value list (ascending ids, the last one carries bit 31): 32 ids per step,
    // one per lane; new ids are appended in list order
    __device__ __forceinline__ void add_list(const int32_t *__restrict__ ent, int64_t q, int c, int lane) {
        last_id = 0;
        last_j = -1;
        for (;;) {
            const int32_t e = __ldg(ent + q + lane);  // the entry array is padded by 32 words
            const unsigned endm = __ballot_sync(0xffffffffu, e < 0);
            const int nval = endm ? __ffs(endm) : 32;
            const bool valid = lane < nval;
            const int32_t id = e & 0x7FFFFFFF;
            int j = -1;
            if (valid) {
                j = find_any(id);
            }
            const unsigned newm = __ballot_sync(0xffffffffu, valid && j < 0);
            if (newm) {
                const int total = __popc(newm);
                if (n + total > SL_LCAP + SL_SPILL) {
                    overflow = 1;
                    return;
                }
                if (valid && j < 0) {
                    j = n + __popc(newm & ((1u << lane) - 1u));
                    put(j, id);
                }
                n += total;
                if (n > SL_LCAP) g_dirty = 1;
            }
            if (valid) bump(j, c);
            __syncwarp();
            if (endm) break;
            q += 32;
        }
    }
    __device__ __forceinline__ int max_count(int lane) const {
        int m = 0;
        for (int j = lane; j < n; j += 32) m = max(m, cnt_at(j));
        return __reduce_max_sync(0xffffffffu, m);
    }
};

struct SlWarpSmem {
    uint8_t bytes[SL_CH + 64];
    uint64_t hit[SL_CH];  // pass A: the position's key (0 = nothing to probe); pass B: the value found, sign-extended
    int32_t id[2][SL_LCAP];
    int32_t cnt[2][SL_LCAP];
    int32_t h_id[SL_HASH];
    int32_t h_j[SL_HASH];
};

// findBestMatch (jgi/Seal.java:2864-2907) of one read into `list`; returns numValidKmers (stream/Read.java:1673-1683)
// when want_valid, else 0.
__device__ __forceinline__ int sl_scan_read(const SealParams &p, const SealTable &tb, const uint8_t *__restrict__ bases, int L, SlList &list,
                            SlWarpSmem &sm, int lane, bool want_valid, bool table_empty) {
    const int k = p.k;
    if (L < k) return 0;  // no window, no valid k-mer
    const int start = p.restrictRight < 1 ? 0 : max(0, L - p.restrictRight);
    const int stop = p.restrictLeft < 1 ? L : min(L, p.restrictLeft);
    const int lo = want_valid ? 0 : start, hi = want_valid ? L : stop;
    int nvalid = 0;
    bool done = table_empty;  // storedKmers<1 -> return 0 (:2865)
    for (int cs = lo; cs < hi; cs += SL_CH) {
        const int ce = min(hi, cs + SL_CH), n = ce - cs;
        const int sb = max(lo, cs - k);
        __syncwarp();
        for (int j = lane; j < ce - sb; j += 32) {  // classify every base once: code | defined<<2 | isN<<3 | complement<<4
            const uint32_t b = bases[sb + j];
            const uint32_t def = bb_defined(b) ? 1u : 0u, x = def ? bb_code_raw(b) : 0u;
            sm.bytes[j] = (uint8_t)(x | (def << 2) | ((b == 'N' ? 1u : 0u) << 3) | ((def ? 3u - x : 0u) << 4));
        }
        __syncwarp();
        const int S = (n + 31) >> 5;
        const int p0 = cs + lane * S, p1 = min(ce, p0 + S);
        if (p0 < p1) {
            uint64_t kmer = 0, rkmer = 0;
            int len = 0, dlen = 0;
            auto roll = [&](int i) {
                if (i == start) {  // the reference's registers start here
                    kmer = 0;
                    rkmer = 0;
                    len = 0;
                }
                const uint32_t b = sm.bytes[i - sb];
                const uint64_t x = b & 3u, x2 = b >> 4;
                kmer = ((kmer << 2) | x) & p.mask;
                rkmer = (rkmer >> 2) | (x2 << p.shift2);  // never exceeds 2k bits
                if ((b & 8u) && p.forbidNs) {
                    len = 0;
                    rkmer = 0;
                } else len++;
                if (want_valid) dlen = (b & 4u) ? dlen + 1 : 0;
            };
            // warm-up over the k positions in front of the lane's run: the same trip count on every lane, so the probing
            // iterations below stay converged
            for (int t = 0; t < k; t++) {
                const int i = p0 - k + t;
                if (i >= sb) roll(i);
            }
            // pass A: keys of the lane's positions; both sectors a probe will touch (the bucket's keys and its values) are
            // prefetched, so the probes of pass B overlap instead of paying two dependent DRAM trips each
            for (int i = p0; i < p1; i++) {
                roll(i);
                if (want_valid) nvalid += (dlen >= k) ? 1 : 0;
                uint64_t key = 0;
                if (!done && i >= start && i < stop && len >= p.minlen2 && i >= k - 1 && !(p.qskip > 1 && (i % p.qskip != 0))) {
                    key = sl_to_value(p, kmer, rkmer);
                    if (sl_passes_speed(p.speed, key)) {
                        const uint64_t bkt = bb_bucket(bb_fhash64(key), tb.bucket_shift);
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(tb.bk[bkt].keys));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(tb.bk[bkt].vals));
                    } else key = 0;
                }
                sm.hit[i - cs] = key;
            }
            for (int i = p0; i < p1; i++) {  // pass B
                const uint64_t key = sm.hit[i - cs];
                int32_t v = 0;
                if (key) {
                    v = sl_table_get(tb, key);
                    if (v < 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(tb.ent_ids + (-(int64_t)v - 2)));  // the fold reads it next
                }
                sm.hit[i - cs] = (uint64_t)(int64_t)v;
            }
        }
        __syncwarp();
        if (!done) {
            for (int c = 0; c < n; c += 32) {
                const int32_t v = (c + lane < n) ? (int32_t)sm.hit[c + lane] : 0;
                unsigned hm = __ballot_sync(0xffffffffu, v != 0);
                if (p.match != SEAL_MATCH_ALL) {  // break after the first (single-id) hit (:2902)
                    const unsigned sm_ = p.match == SEAL_MATCH_FIRST ? hm : __ballot_sync(0xffffffffu, v > 0);
                    if (sm_) {
                        const int cut = __ffs(sm_) - 1;
                        hm &= (cut == 31) ? 0xffffffffu : ((2u << cut) - 1u);
                        done = true;
                    }
                }
                if (hm) {
                    // Positions whose ids are all listed already only add to counts, which commutes: they are applied
                    // in parallel (equal single ids batched, value lists walked one per lane with atomics). Positions
                    // that bring a new id follow in position order, so the list keeps the reference's first-seen order.
                    const bool mine = (hm >> lane) & 1u;
                    bool known = true;
                    int j0 = 0, j1 = 0, j2 = 0, j3 = 0, nl = 0;  // entries of the first four ids of my value list
                    if (mine) {
                        if (v > 0) known = list.find_any(v) >= 0;
                        else {
                            int64_t q = -(int64_t)v - 2;
                            for (;;) {
                                const int32_t e = __ldg(tb.ent_ids + q);
                                const int j = list.find_any(e & 0x7FFFFFFF);
                                if (j < 0) {
                                    known = false;
                                    break;
                                }
                                if (nl == 0) j0 = j;
                                else if (nl == 1) j1 = j;
                                else if (nl == 2) j2 = j;
                                else if (nl == 3) j3 = j;
                                nl++;
                                if (e < 0) break;
                                q++;
                            }
                        }
                    }
                    unsigned slow = __ballot_sync(0xffffffffu, mine && !known);
                    unsigned fs = __ballot_sync(0xffffffffu, mine && known && v > 0);
                    while (fs) {
                        const int32_t hv = __shfl_sync(0xffffffffu, v, __ffs(fs) - 1);
                        const unsigned same = __ballot_sync(0xffffffffu, v == hv) & fs;
                        list.add(hv, __popc(same), lane);
                        fs &= ~same;
                    }
                    if (mine && known && v < 0) {
                        if (nl <= 4) {
                            list.bump_atomic(j0, 1);
                            if (nl > 1) list.bump_atomic(j1, 1);
                            if (nl > 2) list.bump_atomic(j2, 1);
                            if (nl > 3) list.bump_atomic(j3, 1);
                        } else {
                            int64_t q = -(int64_t)v - 2;
                            for (;;) {
                                const int32_t e = __ldg(tb.ent_ids + q);
                                list.bump_atomic(list.find_any(e & 0x7FFFFFFF), 1);
                                if (e < 0) break;
                                q++;
                            }
                        }
                    }
                    __syncwarp();
                    while (slow) {
                        const int32_t hv = __shfl_sync(0xffffffffu, v, __ffs(slow) - 1);
                        const unsigned same = __ballot_sync(0xffffffffu, v == hv) & slow;  // same id, or the same key's list
                        if (hv > 0) list.add(hv, __popc(same), lane);
                        else list.add_list(tb.ent_ids, -(int64_t)hv - 2, __popc(same), lane);
                        slow &= ~same;
                    }
                }
                if (done) break;
            }
        }
    }
    if (!want_valid) return 0;
    return __reduce_add_sync(0xffffffffu, nvalid);
}

struct SlAcc {
    unsigned long long reads_in, bases_in, reads_m, bases_m, reads_u, bases_u;
};

// filterTopScaffolds_withClearzone + the start/stop choice + the counters of assignTogether / assignIndependently
// (jgi/Seal.java:2697-2708, :2393-2408, :2414-2449). count_below: the pair is "unmatched" when max < minhits (kpt only).
__device__ __forceinline__ void sl_assign(const SealParams &p, const SlList &list, int mx, int cz, int minhits, int64_t unit, long long numericID,
                          int readSum, int lenSum, bool frag, bool count_below, const seal_out &out, unsigned long long *sc_reads,
                          unsigned long long *sc_bases, unsigned long long *sc_frags, unsigned long long *sc_ambig, SlAcc &acc,
                          int lane) {
    const int thresh = max(1, mx - cz);
    int sites = 0;
    int32_t min_id = 0x7FFFFFFF;
    for (int base = 0; base < list.n; base += 32) {
        const int j = base + lane;
        const bool in = j < list.n && list.cnt_at(j) >= thresh;
        sites += __popc(__ballot_sync(0xffffffffu, in));
        if (in) min_id = min(min_id, list.id_at(j));
    }
    min_id = __reduce_min_sync(0xffffffffu, min_id);
    int start = 0, stop = 0;
    const bool ok = mx >= minhits;
    if (ok) {
        if (sites < 2 || p.ambig == SEAL_AMBIG_ALL) stop = sites;
        else if (p.ambig == SEAL_AMBIG_TOSS) stop = 0;
        else if (p.ambig == SEAL_AMBIG_FIRST) stop = 1;  // after finalList.sort(): the smallest id
        else {
            start = (int)(numericID % sites);
            stop = start + 1;
        }
    }
    const int stride = out.ids ? p.ids_stride : 0;
    if (stride > 0) {
        for (int j = lane; j < stride; j += 32) out.ids[unit * stride + j] = 0;
        __syncwarp();
    }
    int32_t first = 0;
    if (stop > start) {
        if (p.ambig == SEAL_AMBIG_FIRST && sites >= 2) {
            first = min_id;
            if (lane == 0) {
                if (stride > 0) out.ids[unit * stride] = min_id;
                atomicAdd(sc_reads + min_id, (unsigned long long)readSum);
                atomicAdd(sc_bases + min_id, (unsigned long long)lenSum);
                if (frag) atomicAdd(sc_frags + min_id, 1ull);
                atomicAdd(sc_ambig + min_id, (unsigned long long)readSum);
            }
        } else {
            int rank0 = 0;  // finals before this chunk
            for (int base = 0; base < list.n; base += 32) {
                const int j = base + lane;
                const bool in = j < list.n && list.cnt_at(j) >= thresh;
                const unsigned m = __ballot_sync(0xffffffffu, in);
                const int r = rank0 + __popc(m & ((1u << lane) - 1u));
                const bool sel = in && r >= start && r < stop;
                if (sel) {
                    const int32_t id = list.id_at(j);
                    if (r - start < stride) out.ids[unit * stride + (r - start)] = id;
                    atomicAdd(sc_reads + id, (unsigned long long)readSum);
                    atomicAdd(sc_bases + id, (unsigned long long)lenSum);
                    if (frag) atomicAdd(sc_frags + id, 1ull);
                    if (sites > 1) atomicAdd(sc_ambig + id, (unsigned long long)readSum);
                }
                const unsigned fm = __ballot_sync(0xffffffffu, sel && r == start);
                if (fm) first = __shfl_sync(0xffffffffu, sel ? list.id_at(j) : 0, __ffs(fm) - 1);
                rank0 += __popc(m);
                if (rank0 >= stop) break;
            }
        }
    }
    if (lane == 0) {
        if (out.n_assigned) out.n_assigned[unit] = stop - start;
        if (out.first_id) out.first_id[unit] = first;
        if (out.n_sites) out.n_sites[unit] = sites;
        if (out.max_hits) out.max_hits[unit] = mx;
    }
    if (ok || count_below) {
        if (stop > start) {
            acc.reads_m += readSum;
            acc.bases_m += lenSum;
        } else {
            acc.reads_u += readSum;
            acc.bases_u += lenSum;
        }
    }
}

// the second storage is in use after the scans iff a second read was scanned apart from the first
__device__ __forceinline__ bool m_last_is_second(int nm, int kpt) { return nm == 2 && !kpt; }

__device__ __forceinline__ int sl_cz(const SealParams &p, int nvalid) {
    if (!(p.czf > 0)) return p.clearzone;
    return max(p.clearzone, (int)ceilf(__fmul_rn(p.czf, (float)nvalid)));  // :2213-2214
}
__device__ __forceinline__ int sl_minhits(const SealParams &p, int nk) {
    return max(p.minKmerHits, (int)__fmul_rn(p.mkf, (float)nk));  // :2223
}

__global__ void __launch_bounds__(SL_WARPS * 32, SL_BLOCKS_PER_SM)
seal_match_kernel(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ offsets, int64_t n_frag, int paired,
                  long long first_numeric_id, SealParams p, SealTable tb, int table_empty, seal_out out, int32_t *spill,
                  unsigned long long *sc_reads, unsigned long long *sc_bases, unsigned long long *sc_frags,
                  unsigned long long *sc_ambig, unsigned long long *stats, int *err) {
    __shared__ SlWarpSmem smem[SL_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * SL_WARPS + wib, n_warps = (int64_t)gridDim.x * SL_WARPS;
    SlWarpSmem &sm = smem[wib];
    // one list object, two storages: the second one holds read 2's list when pairs are not kept together (read 1's
    // list must survive until both maxima are known, jgi/Seal.java:2507, :2576). One call site each for the scan and
    // the assignment keeps the kernel's code small.
    SlList cur;
    int32_t *sp = spill + warp * (int64_t)SL_SCRATCH;
    auto use = [&](int which) {
        cur.s_id = sm.id[which];
        cur.s_cnt = sm.cnt[which];
        cur.g_id = sp + which * 2 * SL_SPILL;
        cur.g_cnt = cur.g_id + SL_SPILL;
    };
    cur.h_id = sm.h_id;
    cur.h_j = sm.h_j;
    cur.gh_id = sp + 4 * SL_SPILL;
    cur.gh_j = sp + 4 * SL_SPILL + SL_GHASH;
    cur.clear_global_hash(lane);
    cur.overflow = 0;
    SlAcc acc = {0, 0, 0, 0, 0, 0};
    const bool want_valid = p.czf > 0;
    const int k = p.k;
    const int nm = paired ? 2 : 1;
    for (int64_t f = warp; f < n_frag; f += n_warps) {
        const int64_t i1 = paired ? 2 * f : f;
        const uint32_t o0 = offsets[i1], o1 = offsets[i1 + 1];
        const uint32_t o2 = paired ? offsets[i1 + 2] : o1;
        const int L1 = (int)(o1 - o0), L2 = (int)(o2 - o1);
        const long long nid = first_numeric_id + f;
        acc.reads_in += nm;
        acc.bases_in += (unsigned long long)(L1 + L2);
        int nv1 = 0, nv2 = 0, max1 = 0, max2 = 0, n1 = 0, n2 = 0;
        use(0);
        cur.reset(lane);
        for (int m = 0; m < nm; m++) {
            if (m == 1 && !p.kpt) {
                n1 = cur.n;
                max1 = cur.max_count(lane);
                use(1);
                cur.reset(lane);  // the hashes now serve read 2's list; read 1's entries are only read by index from here on
            }
            const int nv = sl_scan_read(p, tb, bases + (m ? o1 : o0), m ? L2 : L1, cur, sm, lane, want_valid, table_empty);
            if (m) nv2 = nv;
            else nv1 = nv;
        }
        if (m_last_is_second(nm, p.kpt)) {
            n2 = cur.n;
            max2 = cur.max_count(lane);
        } else {
            n1 = cur.n;
            max1 = cur.max_count(lane);
        }
        const int nk1 = max(L1 - k + 1, 0), nk2 = paired ? max(L2 - k + 1, 0) : 0;
        const int na = p.kpt ? 1 : nm;
        for (int m = 0; m < na; m++) {
            int mx, nv, nk, rs, ls;
            int64_t unit;
            bool frag;
            if (p.kpt) {
                mx = max1;
                nv = nv1 + nv2;
                nk = nk1 + nk2;
                rs = nm;
                ls = L1 + L2;
                unit = f;
                frag = true;
            } else {
                use(m);
                cur.n = m ? n2 : n1;
                mx = m ? max2 : max1;
                nv = m ? nv2 : nv1;
                nk = m ? nk2 : nk1;
                rs = 1;
                ls = m ? L2 : L1;
                unit = i1 + m;
                frag = m ? (max2 > max1) : (max1 >= max2);
            }
            sl_assign(p, cur, mx, sl_cz(p, nv), sl_minhits(p, nk), unit, nid, rs, ls, frag, p.kpt != 0, out, sc_reads, sc_bases, sc_frags,
                      sc_ambig, acc, lane);
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (acc.reads_in) {
            atomicAdd(stats + 0, acc.reads_in);
            atomicAdd(stats + 1, acc.bases_in);
        }
        if (acc.reads_m) {
            atomicAdd(stats + 2, acc.reads_m);
            atomicAdd(stats + 3, acc.bases_m);
        }
        if (acc.reads_u) {
            atomicAdd(stats + 4, acc.reads_u);
            atomicAdd(stats + 5, acc.bases_u);
        }
        if (cur.overflow) *err = 1;
    }
}

thread_local std::string g_serr;

}  // namespace

struct seal_handle {
    seal_cfg c;
    SealParams p;
    int device = 0, sm_count = 148, hdist = 0, rskip = 0;
    std::vector<uint8_t> ref;
    std::vector<int64_t> off{0};
    bool finalized = false;
    // table
    SlBucket *d_bk = nullptr;
    int64_t n_slots = 0;
    uint64_t *d_ekeys = nullptr;  // sorted distinct (key, id) entries
    int32_t *d_eids = nullptr;
    int64_t n_entries = 0, stored = 0, ref_kmers = 0;
    int32_t n_seqs = 0;
    // matching
    int32_t *d_spill = nullptr;
    uint8_t *b_bases = nullptr;  // staging of seal_b200_process
    uint32_t *b_off = nullptr;
    int32_t *b_res = nullptr;
    int64_t cap_bases = 0, cap_reads = 0;
    unsigned long long *d_sc = nullptr;  // 4 * (n_seqs+1)
    unsigned long long *d_stats = nullptr;
    int *d_err = nullptr;
    std::atomic<int64_t> launches{0};
    std::mutex mu;
    std::string err;
};

namespace {

int serr(seal_handle *h, const std::string &m) {
    if (h) h->err = m;
    g_serr = m;
    return 1;
}

#define SCK(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            char b_[512];                                                                                  \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return serr(h, b_);                                                                            \
        }                                                                                                  \
    } while (0)

int64_t pow2ceil64(int64_t x) {
    int64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

uint32_t shift_of(uint64_t slot_mask) {
    int m = 0;
    while ((slot_mask >> m) & 1ull) m++;
    return (uint32_t)(34 - m);
}

SealTable view(const seal_handle *h) {
    SealTable t;
    memset(&t, 0, sizeof t);
    t.bk = h->d_bk;
    t.bmask = (uint64_t)(h->n_slots >> 2) - 1;
    t.bucket_shift = shift_of((uint64_t)h->n_slots - 1);
    t.ent_ids = h->d_eids;
    return t;
}

void release_table(seal_handle *h) {
    cudaFree(h->d_bk);
    cudaFree(h->d_ekeys);
    cudaFree(h->d_eids);
    cudaFree(h->d_spill);
    cudaFree(h->b_bases);
    cudaFree(h->b_off);
    cudaFree(h->b_res);
    h->b_bases = nullptr;
    h->b_off = nullptr;
    h->b_res = nullptr;
    h->cap_bases = h->cap_reads = 0;
    cudaFree(h->d_sc);
    cudaFree(h->d_stats);
    cudaFree(h->d_err);
    h->d_bk = nullptr;
    h->d_ekeys = nullptr;
    h->d_eids = nullptr;
    h->d_spill = nullptr;
    h->d_sc = nullptr;
    h->d_stats = nullptr;
    h->d_err = nullptr;
}

int grid_for(const seal_handle *h) { return h->sm_count * SL_BLOCKS_PER_SM; }

int launch_match(seal_handle *h, const uint8_t *d_bases, const uint32_t *d_off, int64_t n_reads, int paired, int64_t first_id,
                 const seal_out &d_out, unsigned long long *d_stats, cudaStream_t st) {
    const int64_t n_frag = paired ? n_reads / 2 : n_reads;
    if (n_frag <= 0) return 0;
    const unsigned long long *sc = h->d_sc;
    const size_t a = (size_t)h->n_seqs + 1;
    seal_match_kernel<<<grid_for(h), SL_WARPS * 32, 0, st>>>(
        d_bases, d_off, n_frag, paired ? 1 : 0, (long long)first_id, h->p, view(h), h->stored < 1 ? 1 : 0, d_out, h->d_spill,
        const_cast<unsigned long long *>(sc), const_cast<unsigned long long *>(sc) + a, const_cast<unsigned long long *>(sc) + 2 * a,
        const_cast<unsigned long long *>(sc) + 3 * a, d_stats, h->d_err);
    h->launches += 1;
    SCK(cudaGetLastError());
    return 0;
}

int check_overflow(seal_handle *h) {
    int e = 0;
    SCK(cudaMemcpy(&e, h->d_err, sizeof e, cudaMemcpyDeviceToHost));
    if (e) {
        cudaMemset(h->d_err, 0, sizeof(int));
        char b[160];
        snprintf(b, sizeof b, "a pair hit more than %d distinct reference ids (the per-pair list is bounded)", SL_LCAP + SL_SPILL);
        return serr(h, b);
    }
    return 0;
}

}  // namespace

extern "C" {

const char *seal_b200_last_error(seal_handle *h) {
    if (h) g_serr = h->err;
    return g_serr.c_str();
}

void seal_b200_cfg_default(seal_cfg *c) {
    if (!c) return;
    memset(c, 0, sizeof *c);
    c->struct_size = (int32_t)sizeof(seal_cfg);
    c->k = 31;                 // jgi/Seal.java:109
    c->rcomp = 1;              // :104
    c->mask_middle = 1;        // :3088
    c->ambig_mode = SEAL_AMBIG_RANDOM;  // :124
    c->match_mode = SEAL_MATCH_ALL;     // :125
    c->keep_pairs_together = 1;         // :126
    c->min_kmer_hits = 1;      // :111
    c->table_load_pct = 50;
    c->ids_stride = 4;
}

int seal_b200_create(const seal_cfg *cfg, seal_handle **out) {
    if (!out) return serr(nullptr, "out is NULL");
    *out = nullptr;
    if (!cfg || cfg->struct_size != (int32_t)sizeof(seal_cfg)) return serr(nullptr, "seal_cfg.struct_size does not match this library");
    if (cfg->k < 1 || cfg->k > 31) return serr(nullptr, "K must range from 1 to 31.");  // jgi/Seal.java:559
    if (cfg->hdist < 0 || cfg->hdist > 2) return serr(nullptr, "hdist must be 0, 1 or 2 on the device path");
    if (cfg->min_kmer_hits < 1) return serr(nullptr, "minKmerHits must be at least 1");  // :555
    if (cfg->min_kmer_fraction > 1) return serr(nullptr, "minKmerFraction must range from 0 to 1");  // :556
    if (cfg->ambig_mode < SEAL_AMBIG_ALL || cfg->ambig_mode > SEAL_AMBIG_RANDOM) return serr(nullptr, "unknown ambiguous mode");
    if (cfg->match_mode < SEAL_MATCH_ALL || cfg->match_mode > SEAL_MATCH_UNIQUE) return serr(nullptr, "unknown match mode");
    if (cfg->ids_stride < 0) return serr(nullptr, "ids_stride must not be negative");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1)
        return serr(nullptr, "no CUDA device: seal_b200 has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return serr(nullptr, "device ordinal out of range");
    seal_handle *h = new seal_handle();
    h->c = *cfg;
    h->device = cfg->device;
    cudaSetDevice(h->device);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    SealParams &p = h->p;
    memset(&p, 0, sizeof p);
    const int k = cfg->k;
    p.k = k;
    p.rcomp = cfg->rcomp ? 1 : 0;
    h->hdist = cfg->hdist;
    h->rskip = std::max(0, cfg->rskip);
    p.forbidNs = (cfg->forbid_ns || h->hdist < 1) ? 1 : 0;  // :492
    const bool mm = cfg->mask_middle || cfg->mid_mask_len > 0;  // :255-260
    const int mml = mm ? (cfg->mid_mask_len > 0 ? cfg->mid_mask_len : 2 - (k & 1)) : 0;  // :548-552
    if (mm) {
        if (!(k > mml + 1)) {
            delete h;
            return serr(nullptr, "Middle-masking requires k>midMaskLen+1. Increase k, shorten the mask, or disable middle-masking with mm=f.");
        }
        const int bits = mml * 2, shift = ((k - mml) / 2) * 2;  // :565-568
        p.middleMask = ~((~((~0ull) << bits)) << shift);
    } else p.middleMask = ~0ull;
    p.minlen2 = mm ? (k - mml) / 2 : k;  // :2868
    p.shift2 = 2 * k - 2;
    p.mask = ~((~0ull) << (2 * k));
    p.kmask = 1ull << (2 * k);  // lengthMasks[k], :3338
    p.speed = cfg->speed;
    p.qskip = cfg->qskip;
    p.restrictLeft = std::max(cfg->restrict_left, 0);
    p.restrictRight = std::max(cfg->restrict_right, 0);
    p.ambig = cfg->ambig_mode;
    p.match = cfg->match_mode;
    p.kpt = cfg->keep_pairs_together ? 1 : 0;
    p.clearzone = cfg->clearzone;
    p.czf = cfg->clearzone_fraction;
    p.minKmerHits = cfg->min_kmer_hits;
    p.mkf = std::max(cfg->min_kmer_fraction, 0.0f);  // :554
    p.ids_stride = cfg->ids_stride;
    *out = h;
    return 0;
}

int seal_b200_add_ref(seal_handle *h, const uint8_t *bases, const int64_t *offsets, int32_t n_seqs) {
    if (!h) return serr(nullptr, "handle is NULL");
    std::lock_guard<std::mutex> g(h->mu);
    if (h->finalized) return serr(h, "add_ref after finalize");
    if (n_seqs < 0 || (n_seqs > 0 && (!bases || !offsets))) return serr(h, "bad arguments");
    for (int32_t s = 0; s < n_seqs; s++) {
        if (offsets[s + 1] < offsets[s]) return serr(h, "offsets must not decrease");
        h->ref.insert(h->ref.end(), bases + offsets[s], bases + offsets[s + 1]);
        h->off.push_back((int64_t)h->ref.size());
    }
    return 0;
}

int seal_b200_finalize(seal_handle *h, int64_t *v) {
    if (!h) return serr(nullptr, "handle is NULL");
    std::lock_guard<std::mutex> g(h->mu);
    if (h->finalized) return serr(h, "finalize called twice");
    SCK(cudaSetDevice(h->device));
    const int k = h->p.k;
    const int64_t total = (int64_t)h->ref.size();
    h->n_seqs = (int32_t)(h->off.size() - 1);
    const int blocks_g = h->sm_count * 8;
    uint64_t *d_seed_kmer = nullptr, *d_k0 = nullptr, *d_k1 = nullptr;
    int32_t *d_seed_id = nullptr;
    uint32_t *d_i0 = nullptr, *d_i1 = nullptr, *d_keep = nullptr, *d_pos = nullptr;
    uint8_t *d_ref = nullptr;
    int64_t *d_off = nullptr;
    unsigned long long *d_ctr = nullptr;
    void *d_tmp = nullptr;
    int *d_ovf = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_seed_kmer);
        cudaFree(d_k0);
        cudaFree(d_k1);
        cudaFree(d_seed_id);
        cudaFree(d_i0);
        cudaFree(d_i1);
        cudaFree(d_keep);
        cudaFree(d_pos);
        cudaFree(d_ref);
        cudaFree(d_off);
        cudaFree(d_ctr);
        cudaFree(d_tmp);
        cudaFree(d_ovf);
    };
    struct Guard {
        decltype(cleanup) &f;
        ~Guard() { f(); }
    } guard{cleanup};
    unsigned long long ctr[4] = {0, 0, 0, 0};
    int64_t n_seeds = 0, n_pairs = 0;
    SCK(cudaMalloc(&d_ctr, sizeof ctr));
    SCK(cudaMemset(d_ctr, 0, sizeof ctr));
    if (total > 0) {
        SCK(cudaMalloc(&d_ref, (size_t)total));
        SCK(cudaMalloc(&d_off, h->off.size() * sizeof(int64_t)));
        SCK(cudaMemcpy(d_ref, h->ref.data(), (size_t)total, cudaMemcpyHostToDevice));
        SCK(cudaMemcpy(d_off, h->off.data(), h->off.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        SCK(cudaMalloc(&d_seed_kmer, (size_t)total * 8));
        SCK(cudaMalloc(&d_seed_id, (size_t)total * 4));
        sl_seed_kernel<<<(unsigned)((total + 255) / 256), 256>>>(d_ref, d_off, h->n_seqs, 1, total, k, h->rskip, d_seed_kmer, d_seed_id,
                                                                 d_ctr);
        h->launches += 1;
        SCK(cudaGetLastError());
        SCK(cudaMemcpy(ctr, d_ctr, sizeof ctr, cudaMemcpyDeviceToHost));
        n_seeds = (int64_t)ctr[0];
        h->ref_kmers = (int64_t)ctr[1];
    }
    if (n_seeds > 0) {
        int64_t ball = 1;
        if (h->hdist >= 1) ball += 3 * k;
        if (h->hdist >= 2) ball += 9ll * k * (k - 1) / 2;
        const int64_t cap = n_seeds * ball;
        if (cap >= (1ll << 31)) return serr(h, "reference too large for this hdist: more than 2^31 (k-mer, id) entries before deduplication");
        SCK(cudaMalloc(&d_k0, (size_t)cap * 8));
        SCK(cudaMalloc(&d_k1, (size_t)cap * 8));
        SCK(cudaMalloc(&d_i0, (size_t)cap * 4));
        SCK(cudaMalloc(&d_i1, (size_t)cap * 4));
        const int n1 = h->hdist > 0 ? 1 + 3 * k : 1;
        const int64_t threads = n_seeds * n1;
        sl_expand_kernel<<<(unsigned)((threads + 255) / 256), 256>>>(d_seed_kmer, d_seed_id, n_seeds, h->hdist, h->p, d_k0, d_i0, d_ctr + 2);
        h->launches += 1;
        SCK(cudaGetLastError());
        SCK(cudaMemcpy(ctr, d_ctr, sizeof ctr, cudaMemcpyDeviceToHost));
        n_pairs = (int64_t)ctr[2];
    }
    int64_t n_entries = 0, stored = 0;
    if (n_pairs > 0) {
        // stable sorts: by id, then by key -> (key, id) order
        int id_bits = 1;
        while ((1ll << id_bits) <= (int64_t)h->n_seqs) id_bits++;
        size_t b1 = 0, b2 = 0, b3 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, b1, d_i0, d_i1, d_k0, d_k1, (int)n_pairs, 0, id_bits);
        cub::DeviceRadixSort::SortPairs(nullptr, b2, d_k1, d_k0, d_i1, d_i0, (int)n_pairs, 0, 2 * k + 1);
        SCK(cudaMalloc(&d_keep, (size_t)n_pairs * 4));
        SCK(cudaMalloc(&d_pos, (size_t)n_pairs * 4));
        cub::DeviceScan::ExclusiveSum(nullptr, b3, d_keep, d_pos, (int)n_pairs);
        const size_t tmp_bytes = std::max(b1, std::max(b2, b3)) + 256;
        SCK(cudaMalloc(&d_tmp, tmp_bytes));
        size_t tb = tmp_bytes;
        SCK(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_i0, d_i1, d_k0, d_k1, (int)n_pairs, 0, id_bits));
        tb = tmp_bytes;
        SCK(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_k1, d_k0, d_i1, d_i0, (int)n_pairs, 0, 2 * k + 1));
        sl_keep_kernel<<<blocks_g, 256>>>(d_k0, d_i0, n_pairs, d_keep);
        tb = tmp_bytes;
        SCK(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_keep, d_pos, (int)n_pairs));
        uint32_t last_pos = 0, last_keep = 0;
        SCK(cudaMemcpy(&last_pos, d_pos + n_pairs - 1, 4, cudaMemcpyDeviceToHost));
        SCK(cudaMemcpy(&last_keep, d_keep + n_pairs - 1, 4, cudaMemcpyDeviceToHost));
        n_entries = (int64_t)last_pos + last_keep;
        SCK(cudaMalloc(&h->d_ekeys, (size_t)n_entries * 8));
        SCK(cudaMalloc(&h->d_eids, (size_t)(n_entries + 32) * 4));
        SCK(cudaMemset(h->d_eids + n_entries, 0, 32 * 4));
        sl_scatter_kernel<<<blocks_g, 256>>>(d_k0, d_i0, d_keep, d_pos, n_pairs, h->d_ekeys, h->d_eids);
        sl_count_heads_kernel<<<blocks_g, 256>>>(h->d_ekeys, n_entries, d_ctr + 3);
        h->launches += 6;
        SCK(cudaGetLastError());
        SCK(cudaMemcpy(ctr, d_ctr, sizeof ctr, cudaMemcpyDeviceToHost));
        stored = (int64_t)ctr[3];
    }
    h->n_entries = n_entries;
    h->stored = stored;
    int load = h->c.table_load_pct;
    if (load < 10 || load > 90) load = 50;
    h->n_slots = std::max<int64_t>(1024, pow2ceil64(stored * 100 / load + 4));
    SCK(cudaMalloc(&h->d_bk, (size_t)(h->n_slots >> 2) * sizeof(SlBucket)));
    sl_fill_kernel<<<blocks_g, 256>>>(h->d_bk, h->n_slots >> 2);
    h->launches += 1;
    if (n_entries > 0) {
        SCK(cudaMalloc(&d_ovf, sizeof(int)));
        SCK(cudaMemset(d_ovf, 0, sizeof(int)));
        const uint64_t slot_mask = (uint64_t)h->n_slots - 1;
        sl_insert_kernel<<<blocks_g, 256>>>(h->d_ekeys, h->d_eids, n_entries, h->d_bk, (slot_mask >> 2), shift_of(slot_mask), d_ovf);
        h->launches += 1;
        SCK(cudaGetLastError());
        int ovf = 0;
        SCK(cudaMemcpy(&ovf, d_ovf, sizeof ovf, cudaMemcpyDeviceToHost));
        if (ovf) return serr(h, "hash array overflow while inserting the reference k-mers");
    } else {
        // lookups never dereference ent_ids without a multi-id key, but keep the pointers valid
        SCK(cudaMalloc(&h->d_ekeys, 8));
        SCK(cudaMalloc(&h->d_eids, 33 * 4));
    }
    const size_t a = (size_t)h->n_seqs + 1;
    SCK(cudaMalloc(&h->d_sc, 4 * a * sizeof(unsigned long long)));
    SCK(cudaMemset(h->d_sc, 0, 4 * a * sizeof(unsigned long long)));
    SCK(cudaMalloc(&h->d_stats, 8 * sizeof(unsigned long long)));
    SCK(cudaMalloc(&h->d_err, sizeof(int)));
    SCK(cudaMemset(h->d_err, 0, sizeof(int)));
    SCK(cudaMalloc(&h->d_spill, (size_t)grid_for(h) * SL_WARPS * SL_SCRATCH * sizeof(int32_t)));
    SCK(cudaDeviceSynchronize());
    h->finalized = true;
    std::vector<uint8_t>().swap(h->ref);
    if (v) {
        v[0] = stored;
        v[1] = n_entries;
        v[2] = h->ref_kmers;
    }
    return 0;
}

int64_t seal_b200_n_units(const seal_handle *h, int64_t n_reads, int32_t paired) {
    if (!h) return -1;
    return (paired && h->p.kpt) ? n_reads / 2 : n_reads;
}

int seal_b200_process_device(seal_handle *h, const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_reads, int32_t paired,
                             int64_t first_numeric_id, const seal_out *d_out, unsigned long long *d_stats, void *stream) {
    if (!h) return serr(nullptr, "handle is NULL");
    if (!h->finalized) return serr(h, "process before finalize");
    if (!d_out || !d_stats) return serr(h, "d_out / d_stats is NULL");
    if (n_reads < 0 || (paired && (n_reads & 1))) return serr(h, "paired batches need an even number of reads");
    if (first_numeric_id < 0) return serr(h, "first_numeric_id must not be negative");
    std::lock_guard<std::mutex> g(h->mu);  // one launch at a time per handle: the warps' spill scratch belongs to the handle
    SCK(cudaSetDevice(h->device));
    return launch_match(h, d_bases, d_offsets, n_reads, paired, first_numeric_id, *d_out, d_stats, (cudaStream_t)stream);
}

int seal_b200_process(seal_handle *h, const uint8_t *bases, const int64_t *offsets, int64_t n_reads, int32_t paired,
                      int64_t first_numeric_id, const seal_out *out, seal_stats *stats) {
    if (!h) return serr(nullptr, "handle is NULL");
    if (!h->finalized) return serr(h, "process before finalize");
    if (!out) return serr(h, "out is NULL");
    if (n_reads < 0 || (paired && (n_reads & 1))) return serr(h, "paired batches need an even number of reads");
    if (n_reads > 0 && (!bases || !offsets)) return serr(h, "bases / offsets is NULL");
    if (first_numeric_id < 0) return serr(h, "first_numeric_id must not be negative");
    std::lock_guard<std::mutex> g(h->mu);
    SCK(cudaSetDevice(h->device));
    const int step = paired ? 2 : 1;
    const int stride = out->ids ? h->p.ids_stride : 0;
    const bool per_pair = paired && h->p.kpt;
    SCK(cudaMemset(h->d_stats, 0, 8 * sizeof(unsigned long long)));
    // chunks of whole fragments, < 1 GiB of bases and <= 8 Mi reads each
    const int64_t max_bases = 1ll << 30, max_reads = 1ll << 23;
    // device staging of the host entry point: kept with the handle and only ever grown (cudaMalloc / cudaFree per call
    // cost more than the kernel on small batches)
    uint8_t *&d_bases = h->b_bases;
    uint32_t *&d_off = h->b_off;
    int32_t *&d_res = h->b_res;
    int64_t &cap_bases = h->cap_bases, &cap_reads = h->cap_reads;
    std::vector<uint32_t> off32;
    int rc = 0;
    int64_t r0 = 0;
    while (r0 < n_reads && rc == 0) {
        int64_t r1 = r0;
        while (r1 < n_reads && r1 - r0 + step <= max_reads && offsets[r1 + step] - offsets[r0] <= max_bases) r1 += step;
        if (r1 == r0) {
            rc = serr(h, "a single read / pair exceeds 1 GiB");
            break;
        }
        const int64_t nr = r1 - r0, nb = offsets[r1] - offsets[r0];
        const int64_t nu = per_pair ? nr / 2 : nr;
        if (nb + 16 > cap_bases || nr > cap_reads) {
            cudaFree(d_bases);
            cudaFree(d_off);
            cudaFree(d_res);
            d_bases = nullptr;
            d_off = nullptr;
            d_res = nullptr;
            cap_bases = std::max(cap_bases, nb + 16);
            cap_reads = std::max(cap_reads, nr);
            if (cudaMalloc(&d_bases, (size_t)cap_bases) != cudaSuccess || cudaMalloc(&d_off, (size_t)(cap_reads + 1) * 4) != cudaSuccess ||
                cudaMalloc(&d_res, (size_t)cap_reads * (4 + (size_t)h->p.ids_stride) * 4 + 16) != cudaSuccess) {
                cudaFree(d_bases);
                cudaFree(d_off);
                cudaFree(d_res);
                d_bases = nullptr;
                d_off = nullptr;
                d_res = nullptr;
                cap_bases = cap_reads = 0;
                rc = serr(h, "cudaMalloc failed for the batch buffers");
                break;
            }
        }
        off32.resize((size_t)nr + 1);
        for (int64_t i = 0; i <= nr; i++) {
            if (offsets[r0 + i] < offsets[r0] || (i > 0 && offsets[r0 + i] < offsets[r0 + i - 1])) {
                rc = serr(h, "offsets must not decrease");
                break;
            }
            off32[(size_t)i] = (uint32_t)(offsets[r0 + i] - offsets[r0]);
        }
        if (rc) break;
        cudaError_t e = cudaSuccess;
        if (nb > 0) e = cudaMemcpy(d_bases, bases + offsets[r0], (size_t)nb, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_off, off32.data(), (size_t)(nr + 1) * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            rc = serr(h, std::string("host to device copy failed: ") + cudaGetErrorString(e));
            break;
        }
        seal_out d;
        d.n_assigned = d_res;
        d.first_id = d_res + nu;
        d.n_sites = d_res + 2 * nu;
        d.max_hits = d_res + 3 * nu;
        d.ids = stride > 0 ? d_res + 4 * nu : nullptr;
        const int64_t fid = first_numeric_id + (paired ? r0 / 2 : r0);
        rc = launch_match(h, d_bases, d_off, nr, paired, fid, d, h->d_stats, 0);
        if (rc) break;
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            rc = serr(h, std::string("seal_match_kernel failed: ") + cudaGetErrorString(e));
            break;
        }
        const int64_t u0 = per_pair ? r0 / 2 : r0;
        if (out->n_assigned) cudaMemcpy(out->n_assigned + u0, d.n_assigned, (size_t)nu * 4, cudaMemcpyDeviceToHost);
        if (out->first_id) cudaMemcpy(out->first_id + u0, d.first_id, (size_t)nu * 4, cudaMemcpyDeviceToHost);
        if (out->n_sites) cudaMemcpy(out->n_sites + u0, d.n_sites, (size_t)nu * 4, cudaMemcpyDeviceToHost);
        if (out->max_hits) cudaMemcpy(out->max_hits + u0, d.max_hits, (size_t)nu * 4, cudaMemcpyDeviceToHost);
        if (stride > 0) cudaMemcpy(out->ids + u0 * stride, d.ids, (size_t)nu * stride * 4, cudaMemcpyDeviceToHost);
        e = cudaGetLastError();
        if (e != cudaSuccess) {
            rc = serr(h, std::string("device to host copy failed: ") + cudaGetErrorString(e));
            break;
        }
        r0 = r1;
    }
    if (rc) return rc;
    if (check_overflow(h)) return 1;
    if (stats) {
        unsigned long long s[8];
        SCK(cudaMemcpy(s, h->d_stats, sizeof s, cudaMemcpyDeviceToHost));
        memset(stats, 0, sizeof *stats);
        stats->reads_in = (int64_t)s[0];
        stats->bases_in = (int64_t)s[1];
        stats->reads_matched = (int64_t)s[2];
        stats->bases_matched = (int64_t)s[3];
        stats->reads_unmatched = (int64_t)s[4];
        stats->bases_unmatched = (int64_t)s[5];
    }
    return 0;
}

int seal_b200_scaffold_counts(seal_handle *h, int64_t *reads, int64_t *bases, int64_t *frags, int64_t *ambig, int32_t n) {
    if (!h) return serr(nullptr, "handle is NULL");
    if (!h->finalized) return serr(h, "scaffold_counts before finalize");
    std::lock_guard<std::mutex> g(h->mu);
    SCK(cudaSetDevice(h->device));
    SCK(cudaDeviceSynchronize());
    if (check_overflow(h)) return 1;
    const size_t a = (size_t)h->n_seqs + 1;
    const size_t m = std::min<size_t>(a, (size_t)std::max(n, 0));
    int64_t *dst[4] = {reads, bases, frags, ambig};
    for (int j = 0; j < 4; j++)
        if (dst[j] && m) SCK(cudaMemcpy(dst[j], h->d_sc + j * a, m * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int seal_b200_table_export(seal_handle *h, uint64_t *keys, int32_t *ids, int64_t cap, int64_t *n_out) {
    if (!h) return serr(nullptr, "handle is NULL");
    if (!h->finalized) return serr(h, "table_export before finalize");
    std::lock_guard<std::mutex> g(h->mu);
    SCK(cudaSetDevice(h->device));
    if (n_out) *n_out = h->n_entries;
    const int64_t m = std::min(cap, h->n_entries);
    if (m <= 0) return 0;
    if (keys) SCK(cudaMemcpy(keys, h->d_ekeys, (size_t)m * 8, cudaMemcpyDeviceToHost));
    if (ids) {
        int32_t *d_tmp = nullptr;
        SCK(cudaMalloc(&d_tmp, (size_t)m * 4));
        sl_unmark_kernel<<<h->sm_count * 4, 256>>>(h->d_eids, d_tmp, m);
        h->launches += 1;
        cudaError_t e = cudaMemcpy(ids, d_tmp, (size_t)m * 4, cudaMemcpyDeviceToHost);
        cudaFree(d_tmp);
        SCK(e);
    }
    return 0;
}

int64_t seal_b200_launch_count(const seal_handle *h) { return h ? h->launches.load() : 0; }

void seal_b200_destroy(seal_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    release_table(h);
    delete h;
}

}  // extern "C"
