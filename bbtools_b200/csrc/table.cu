// table.cu -- one-time reference-table build on the device.
//
// Replaces the reference's LoadThread scan + mutate recursion (jgi/BBDuk.java:2210-2452) and the
// kmer.HashArray1D inserts behind it. The reference walks the (3k)^d mutation tree once per way and
// throws 6/7 of it away; here every (reference k-mer, mutation prefix) is one thread, mutants are
// inserted with atomicCAS on the key and atomicMin on the id, which realises "first writer wins with
// scaffolds in file order" == smallest scaffold id, and duplicate mutants
// collapse in the insert itself.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <vector>

#include "bbduk_dev.cuh"
#include "table.h"

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            snprintf(err, errlen, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 1;                                                                                 \
        }                                                                                             \
    } while (0)

__host__ __device__ __forceinline__ uint32_t bucket_shift_of(uint64_t slot_mask) {
    // n_slots = slot_mask+1 = 2^m  ->  34 - m
    int m = 0;
    while ((slot_mask >> m) & 1ull) m++;
    return (uint32_t)(34 - m);
}

struct Seed {
    uint64_t kmer;
    int32_t id;
    int8_t extra;  // next reference base (0..3) or -1; "extraBase" of jgi/BBDuk.java:2277
    int8_t kind;   // short seeds: 1 = prefix of a scaffold's first k-mer, 2 = suffix of its last k-mer
    int8_t pad[2];
};

// ---- pass 1: reference scan (jgi/BBDuk.java:2210-2288 addToMap(Read,skip)) --------------------------
// One thread per reference base i. run[i] = number of consecutive defined bases ending at i inside the
// scaffold (the reference's `len`), precomputed for skip>1, else derived from a k-base look-back.
// Emits the full-length seed at i when len>=k (&& len%skip==0), and -- when useShortKmers -- the
// prefix seeds at i==k-1 (addToMapRightShift, :2327-2346) and suffix seeds at i==L-1
// (addToMapLeftShift, :2299-2317).
__global__ void ref_seed_kernel(const uint8_t *__restrict__ bases, const int64_t *__restrict__ offsets,
                                const int32_t *__restrict__ scaf_of_block, const int32_t *__restrict__ skips,
                                int32_t first_id, int64_t total, BBParams p, Seed *full_seeds,
                                unsigned long long *n_full, Seed *short_seeds /* [32][cap_short] */,
                                unsigned long long *n_short /* [32] */, int64_t cap_short,
                                unsigned long long *ref_kmers) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    // locate the scaffold of this base: blocks are annotated with the scaffold of their first base
    int32_t s = scaf_of_block[blockIdx.x];
    while (offsets[s + 1] <= g) s++;
    const int64_t s0 = offsets[s], s1 = offsets[s + 1];
    const int64_t i = g - s0, L = s1 - s0;
    const int k = p.k;
    if (L < k || i < k - 1) return;
    // rolling state at i, closed form: all of the last k bases must be defined for len>=k
    uint64_t kmer = 0, rkmer = 0;
    for (int j = 0; j < k; j++) {
        const uint32_t c = bases[g - (k - 1) + j];
        if (!bb_defined(c)) return;
        const uint32_t x = bb_code_raw(c);
        kmer = (kmer << 2) | x;
        rkmer = (rkmer >> 2) | ((uint64_t)(3u - x) << p.shift2);
    }
    namespace cg = cooperative_groups;
    int skip = skips[s];
    {
        cg::coalesced_group grp = cg::coalesced_threads();
        if (grp.thread_rank() == 0) atomicAdd(ref_kmers, (unsigned long long)grp.size());
    }
    if (skip > 1) {
        // exact run length needed: walk back to the last undefined base (rare path, skip>1 only)
        int64_t len = k;
        int64_t q = g - k;
        while (q >= s0 && bb_defined(bases[q])) {
            len++;
            q--;
        }
        if (len % skip != 0) return;
    }
    const int id = first_id + s;
    const int extra = (i >= L - 1) ? -1 : bb_code_m1(bases[g + 1]);
    {
        cg::coalesced_group grp = cg::coalesced_threads();
        unsigned long long w = 0;
        if (grp.thread_rank() == 0) w = atomicAdd(n_full, (unsigned long long)grp.size());
        w = grp.shfl(w, 0) + grp.thread_rank();
        Seed sd;
        sd.kmer = kmer;
        sd.id = id;
        sd.extra = (int8_t)extra;
        sd.kind = 0;
        full_seeds[w] = sd;
    }
    if (p.useShortKmers) {
        if (i == k - 1) {  // prefixes of the first k-mer, lengths k-1..mink
            uint64_t km = kmer;
            for (int n = k - 1; n >= p.mink; n--) {
                const int eb = (int)(km & 3);
                km >>= 2;
                const unsigned long long w = atomicAdd(n_short + n, 1ull);
                Seed sd;
                sd.kmer = km;
                sd.id = id;
                sd.extra = (int8_t)eb;
                sd.kind = 1;
                short_seeds[(int64_t)n * cap_short + w] = sd;
            }
        }
        if (i == L - 1) {  // suffixes of the last k-mer
            for (int n = k - 1; n >= p.mink; n--) {
                const uint64_t km = kmer & ((1ull << (2 * n)) - 1);
                const unsigned long long w = atomicAdd(n_short + n, 1ull);
                Seed sd;
                sd.kmer = km;
                sd.id = id;
                sd.extra = (int8_t)extra;
                sd.kind = 2;
                short_seeds[(int64_t)n * cap_short + w] = sd;
            }
        }
    }
    (void)rkmer;
}

// ---- pass 2: neighbourhood expansion (jgi/BBDuk.java:2359-2452 addToMap + mutate) ------------------
// One edit operation of mutate()'s loops, addressed by a flat index:
//   [0, 4*len)                 Sub   j=op/len, i=op%len                       (:2413-2421)
//   [.., +len-1)               Del   i=1..len-1, needs extraBase in 0..3      (:2425-2433)
//   [.., +4*(len-1))           Ins   i=1..len-1, j=0..3                       (:2436-2446)
// The indel ranges exist only when the global editDistance>0 (:2423).
__device__ __forceinline__ int n_ops_for(int len, bool edits) { return 4 * len + (edits ? 5 * (len - 1) : 0); }

__device__ __forceinline__ bool apply_op(uint64_t kmer, int len, int extra, int op, uint64_t &out, int &extra_out) {
    if (op < 4 * len) {
        const int j = op / len, i = op - j * len;
        out = (kmer & ~(3ull << (2 * i))) | ((uint64_t)j << (2 * i));
        extra_out = extra;
        return out != kmer;
    }
    op -= 4 * len;
    if (op < len - 1) {
        const int i = op + 1;
        if (extra < 0 || extra > 3) return false;
        const uint64_t left = ~0ull << (2 * i), right = ~left;
        out = (kmer & left) | ((kmer << 2) & right) | (uint64_t)extra;
        extra_out = -1;
        return out != kmer;
    }
    op -= (len - 1);
    const int i = (op >> 2) + 1, j = op & 3;
    const uint64_t left = ~0ull << (2 * i), right = ~left;
    out = ((kmer & left) | ((kmer & right) >> 2)) | ((uint64_t)j << (2 * (i - 1)));
    extra_out = (int)(kmer & 3);
    return out != kmer;
}

__device__ __forceinline__ int put_kmer(uint64_t *keys, int32_t *vals, uint64_t slot_mask, const BBParams &p,
                                        uint64_t kmer, int len, int id, int *overflow) {
    const uint64_t key = bb_to_value(p, kmer, bb_rcomp(kmer, len), 1ull << (2 * len));
    return bb_table_put(keys, vals, slot_mask, bucket_shift_of(slot_mask), key, id, overflow);
}

// dist = number of mutate() recursion levels; prefix_levels = dist-1 levels are decoded from the
// thread index, the last level is a loop. Every intermediate string is inserted too (mutate inserts
// its own key before recursing, :2395-2407).
__global__ void expand_kernel(const Seed *__restrict__ seeds, int64_t n_seeds, int len, int dist, int use_extra,
                              BBParams p, uint64_t *keys, int32_t *vals, uint64_t slot_mask,
                              unsigned long long *created, int *overflow) {
    const bool edits = p.editDistance > 0;
    const int nops = n_ops_for(len, edits);
    int64_t n_prefix = 1;
    for (int d = 1; d < dist; d++) n_prefix *= nops;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_seeds * n_prefix) return;
    const int64_t si = t / n_prefix;
    int64_t pre = t - si * n_prefix;
    const Seed sd = seeds[si];
    uint64_t km = sd.kmer;
    int extra = use_extra ? (int)sd.extra : -1;
    int made = 0;
    if (dist == 0) {
        // addToMap hdist==0 branch: speed filter applies only here (:2366-2372)
        const uint64_t key = bb_to_value(p, km, bb_rcomp(km, len), 1ull << (2 * len));
        if (bb_passes_speed(p, key)) made += bb_table_put(keys, vals, slot_mask, bucket_shift_of(slot_mask), key, sd.id, overflow);
        if (made) atomicAdd(created, (unsigned long long)made);
        return;
    }
    made += put_kmer(keys, vals, slot_mask, p, km, len, sd.id, overflow);
    for (int d = 1; d < dist; d++) {
        const int op = (int)(pre % nops);
        pre /= nops;
        uint64_t nk;
        int ne;
        if (!apply_op(km, len, extra, op, nk, ne)) {  // temp==kmer or impossible deletion: no subtree
            if (made) atomicAdd(created, (unsigned long long)made);
            return;
        }
        km = nk;
        extra = ne;
        made += put_kmer(keys, vals, slot_mask, p, km, len, sd.id, overflow);
    }
    for (int op = 0; op < nops; op++) {
        uint64_t nk;
        int ne;
        if (apply_op(km, len, extra, op, nk, ne)) made += put_kmer(keys, vals, slot_mask, p, nk, len, sd.id, overflow);
    }
    if (made) atomicAdd(created, (unsigned long long)made);
}

__global__ void fill_kernel(uint64_t *keys, int32_t *vals, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        keys[i] = BB_EMPTY_KEY;
        vals[i] = 0x7FFFFFFF;
    }
}

__global__ void rehash_kernel(const uint64_t *__restrict__ okeys, const int32_t *__restrict__ ovals, int64_t n_old,
                              uint64_t *keys, int32_t *vals, uint64_t slot_mask, int *overflow) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_old; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = okeys[i];
        if (k != BB_EMPTY_KEY) bb_table_put(keys, vals, slot_mask, bucket_shift_of(slot_mask), k, ovals[i], overflow);
    }
}

__global__ void filter_build_kernel(const uint64_t *__restrict__ keys, int64_t n, uint32_t *filter, uint32_t n_words) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        if (k != BB_EMPTY_KEY) {
            const uint32_t t = bb_fhash((uint32_t)k, (uint32_t)(k >> 32));
            atomicOr(filter + bb_filter_word(t, n_words), bb_filter_bits(t));
        }
    }
}

// L2-resident filter: ONE bit per key, word = high-half range reduction of the 32-bit hash, bit = its low 5 bits
__global__ void big_filter_build_kernel(const uint64_t *__restrict__ keys, int64_t n, uint32_t *filter, uint32_t n_words) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        if (k != BB_EMPTY_KEY) {
            const uint32_t t = bb_fhash64(k);
            atomicOr(filter + bb_big_word(t, n_words), 1u << (t & 31u));
        }
    }
}

// short-key bloom: keys whose length marker sits below bit 2k (the mink..k-1 tails)
__global__ void short_filter_build_kernel(const uint64_t *__restrict__ keys, int64_t n, uint64_t kmask, uint32_t *filter,
                                          uint32_t n_words) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        if (k != BB_EMPTY_KEY && k < kmask) {
            const uint32_t t = bb_fhash64(k);
            atomicOr(filter + bb_filter_word(t, n_words), bb_filter_bits(t));
        }
    }
}

// part filter: the part values of every full-length reference k-mer, both strands (if rcomp)
__global__ void part_filter_build_kernel(const Seed *__restrict__ seeds, int64_t n, int k, int rcomp, int n_parts, int w,
                                         int lag0, int lag1, int lag2, int lag3, uint32_t *filter, uint32_t n_words,
                                         uint8_t *samp) {
    const int lags[4] = {lag0, lag1, lag2, lag3};
    const uint32_t vm = (w >= 16) ? 0xFFFFFFFFu : ((1u << (2 * w)) - 1u);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t f = seeds[i].kmer;
        const uint64_t r = bb_rcomp(f, k);
        for (int o = 0; o < (rcomp ? 2 : 1); o++) {
            const uint64_t x = o ? r : f;
            for (int j = 0; j < n_parts; j++) {
                const uint32_t v = (uint32_t)(x >> (2 * lags[j])) & vm;
                for (int d = 0; d <= w - BB_PART_WD; d++) {  // every 9-mer of the part (bbduk_dev.cuh)
                    const uint32_t y = v >> (2 * d);
                    atomicOr(filter + bb_part_word(y), bb_part_bit(y));
                    if (samp) {  // both 8-mers of the 9-mer: every 8-mer inside a part is covered (probe_fast2.cu)
                        samp[(y >> 2) & 0xFFFFu] = 1;
                        samp[y & 0xFFFFu] = 1;
                    }
                }
            }
        }
    }
}

// Tail bitmaps (probe_fast2.cu): the short k-mers are the prefixes S[0:n] of every scaffold's first k-mer and the suffixes
// of its last k-mer, n = mink..k-1, each with its hdist2 substitution neighbourhood, stored canonically. A read's last n
// bases can only match one of them if (type I) read[L-n : L-n+q] lies within hdist2 substitutions of the FIRST q bases of
// a prefix string or of the reverse complement of a suffix string, or (type II) the read's last q bases lie within hdist2
// of the LAST q bases of a suffix string or of the reverse complement of a prefix string (q = min(mink, 12); ktrim=l uses
// the mirror image). B1 = bitmap of the type-I q-mers, B2 = of the type-II q-mers, each 4^q bits, direct-addressed.
__device__ __forceinline__ void tail_set(uint32_t *bm, uint32_t v) { atomicOr(bm + (v >> 5), 1u << (v & 31u)); }
__device__ void tail_ball(uint32_t *bm, uint32_t v, int q, int dist) {  // every q-mer within `dist` (<= 3) substitutions of v
    tail_set(bm, v);
    if (dist < 1) return;
    for (int i = 0; i < q; i++)
        for (uint32_t a = 1; a < 4; a++) {
            const uint32_t v1 = v ^ (a << (2 * i));
            tail_set(bm, v1);
            if (dist < 2) continue;
            for (int j = i + 1; j < q; j++)
                for (uint32_t b = 1; b < 4; b++) {
                    const uint32_t v2 = v1 ^ (b << (2 * j));
                    tail_set(bm, v2);
                    if (dist < 3) continue;
                    for (int l = j + 1; l < q; l++)
                        for (uint32_t c = 1; c < 4; c++) tail_set(bm, v2 ^ (c << (2 * l)));
                }
        }
}
__global__ void tail_filter_build_kernel(const Seed *__restrict__ seeds, int64_t n, int len, int q, int dist, int rcomp,
                                         uint32_t *b1, uint32_t *b2) {
    const uint32_t qm = (q >= 16) ? 0xFFFFFFFFu : ((1u << (2 * q)) - 1u);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Seed sd = seeds[i];
        const uint32_t first = (uint32_t)(sd.kmer >> (2 * (len - q))) & qm, last = (uint32_t)sd.kmer & qm;
        if (sd.kind == 1) {
            tail_ball(b1, first, q, dist);
            if (rcomp) tail_ball(b2, (uint32_t)bb_rcomp(first, q), q, dist);
        } else {
            tail_ball(b2, last, q, dist);
            if (rcomp) tail_ball(b1, (uint32_t)bb_rcomp(last, q), q, dist);
        }
    }
}

// Level-0 images of the two tail bitmaps for the shared memory of probe_fast2.cu: bit (v >> 2(q-8)) of an 8-mer bitmap is set
// iff some q-mer v with that 8-base prefix is set in the full bitmap (q >= 9). 2048 words each, behind B1 | B2.
__global__ void tail_prefix_kernel(const uint32_t *__restrict__ b, int64_t words, int q, uint32_t *l0) {
    const int sh = 2 * (q - 8);
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (int64_t)gridDim.x * blockDim.x) {
        uint32_t m = b[w];
        while (m) {
            const uint32_t v = (uint32_t)(32 * w) + (uint32_t)(__ffs(m) - 1);
            m &= m - 1;
            const uint32_t u = v >> sh;
            atomicOr(l0 + (u >> 5), 1u << (u & 31u));
        }
    }
}

static int64_t pow2ceil(int64_t x) {
    int64_t p = 1024;
    while (p < x) p <<= 1;
    return p;
}

// number of strings visited by mutate() (upper bound on keys created by one seed)
static double ball_bound(int len, int dist, bool edits) {
    const double nops = 4.0 * len + (edits ? 5.0 * (len - 1) : 0.0);
    double tot = 1, lvl = 1;
    for (int d = 1; d <= dist; d++) {
        lvl *= nops;
        tot += lvl;
    }
    if (!edits) {  // exact Hamming-ball size
        double b = 0, c = 1;
        for (int d = 0; d <= dist; d++) {
            b += c;
            c = c * (len - d) / (d + 1) * 3.0;
        }
        return b;
    }
    return tot;
}

void DeviceTable::release() {
    if (owns) {
        cudaFree(d_keys);
        cudaFree(d_vals);
        cudaFree(d_filter);
    }
    d_keys = nullptr;
    d_vals = nullptr;
    d_filter = nullptr;
    n_slots = 0;
}

int DeviceTable::alloc(int64_t slots, uint32_t total_words, char *err, int errlen) {
    release();
    n_slots = slots;
    owns = true;
    CK(cudaMalloc(&d_keys, sizeof(uint64_t) * (size_t)slots));
    CK(cudaMalloc(&d_vals, sizeof(int32_t) * (size_t)slots));
    CK(cudaMalloc(&d_filter, sizeof(uint32_t) * (size_t)std::max<uint32_t>(total_words, 1)));
    return 0;
}

BBTable DeviceTable::view() const {
    BBTable t;
    t.keys = d_keys;
    t.vals = d_vals;
    t.slot_mask = (uint64_t)n_slots - 1;
    t.bucket_shift = bucket_shift_of(t.slot_mask);
    t.filter = d_filter;
    t.n_filter_words = n_filter_words;
    t.part_words = part_words;
    t.short_words = short_words;
    t.big_words = big_words;
    t.samp_words = samp_words;
    t.tail_words = tail_words;
    t.tail_q = tail_q;
    t.n_parts = n_parts;
    t.part_w = part_w;
    for (int j = 0; j < 4; j++) t.part_lag[j] = part_lag[j];
    t.n_scaffolds = n_scaffolds;
    t.stored = stored;
    return t;
}

int DeviceTable::build(const BBParams &p, const std::vector<uint8_t> &ref, const std::vector<int64_t> &offsets,
                       int load_pct, uint32_t filter_words, cudaStream_t st, int64_t *launches, char *err, int errlen) {
    const int32_t n_scaf = (int32_t)offsets.size() - 1;
    n_scaffolds = n_scaf;
    stored = 0;
    ref_kmers = 0;
    const int64_t total = n_scaf > 0 ? offsets.back() : 0;
    if (load_pct <= 0 || load_pct > 90) load_pct = 50;

    uint8_t *d_ref = nullptr;
    int64_t *d_off = nullptr;
    int32_t *d_blk = nullptr, *d_skip = nullptr;
    Seed *d_full = nullptr, *d_short = nullptr;
    unsigned long long *d_cnt = nullptr;  // [0]=n_full [1..32]=n_short[len] [34]=created [35]=ref_kmers
    const int TPB = 256;
    const int64_t n_blocks = (total + TPB - 1) / TPB;
    const int64_t cap_short = 2 * (int64_t)std::max(n_scaf, 1);
    unsigned long long h_cnt[40] = {0};

    auto cleanup = [&]() {
        cudaFree(d_ref);
        cudaFree(d_off);
        cudaFree(d_blk);
        cudaFree(d_skip);
        cudaFree(d_full);
        cudaFree(d_short);
        cudaFree(d_cnt);
    };
#define CKC(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            snprintf(err, errlen, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            cleanup();                                                                                \
            return 1;                                                                                 \
        }                                                                                             \
    } while (0)

    CKC(cudaMalloc(&d_cnt, sizeof(h_cnt)));
    CKC(cudaMemsetAsync(d_cnt, 0, sizeof(h_cnt), st));
    if (total > 0) {
        // per-scaffold skip (jgi/BBDuk.java:2190, :2211) and per-block scaffold hints
        std::vector<int32_t> skips(n_scaf), blk(n_blocks);
        for (int32_t s = 0; s < n_scaf; s++) {
            const int64_t L = offsets[s + 1] - offsets[s];
            const int sk = L > 20000000 ? p.k : L > 5000000 ? 11 : L > 500000 ? 2 : 0;
            skips[s] = std::max(p.minSkip, std::min(p.maxSkip, sk));
        }
        {
            int32_t s = 0;
            for (int64_t b = 0; b < n_blocks; b++) {
                const int64_t g = b * TPB;
                while (s + 1 < n_scaf && offsets[s + 1] <= g) s++;
                blk[b] = s;
            }
        }
        CKC(cudaMalloc(&d_ref, (size_t)total + 16));
        CKC(cudaMalloc(&d_off, sizeof(int64_t) * (n_scaf + 1)));
        CKC(cudaMalloc(&d_blk, sizeof(int32_t) * n_blocks));
        CKC(cudaMalloc(&d_skip, sizeof(int32_t) * n_scaf));
        CKC(cudaMalloc(&d_full, sizeof(Seed) * (size_t)total));
        CKC(cudaMalloc(&d_short, sizeof(Seed) * (size_t)(32 * cap_short)));
        CKC(cudaMemcpyAsync(d_ref, ref.data(), (size_t)total, cudaMemcpyHostToDevice, st));
        CKC(cudaMemcpyAsync(d_off, offsets.data(), sizeof(int64_t) * (n_scaf + 1), cudaMemcpyHostToDevice, st));
        CKC(cudaMemcpyAsync(d_blk, blk.data(), sizeof(int32_t) * n_blocks, cudaMemcpyHostToDevice, st));
        CKC(cudaMemcpyAsync(d_skip, skips.data(), sizeof(int32_t) * n_scaf, cudaMemcpyHostToDevice, st));
        ref_seed_kernel<<<(unsigned)n_blocks, TPB, 0, st>>>(d_ref, d_off, d_blk, d_skip, 1, total, p, d_full, d_cnt,
                                                            d_short, d_cnt + 1, cap_short, d_cnt + 35);
        (*launches)++;
        CKC(cudaGetLastError());
        CKC(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
        CKC(cudaStreamSynchronize(st));
    }
    ref_kmers = (int64_t)h_cnt[35];
    const int64_t n_full = (int64_t)h_cnt[0];

    // distances per class, as addToMap picks them (jgi/BBDuk.java:2366-2378)
    auto class_dist = [&](int hd, int ed, int *use_extra) {
        *use_extra = 0;
        if (hd == 0) return 0;
        if (ed > 0) {
            *use_extra = 1;
            return ed;
        }
        return hd;
    };
    const bool edits = p.editDistance > 0;
    int ux_full = 0, ux_short = 0;
    const int dist_full = class_dist(p.hammingDistance, p.editDistance, &ux_full);
    const int dist_short = class_dist(p.hammingDistance2, p.editDistance2, &ux_short);

    double bound = (double)n_full * ball_bound(p.k, dist_full, edits && dist_full > 0);
    for (int n = 1; n < 32; n++) bound += (double)h_cnt[1 + n] * ball_bound(n, dist_short, edits && dist_short > 0);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    int64_t slots = pow2ceil((int64_t)(bound * 100.0 / load_pct) + 4);
    if (slots > (1ll << 34)) slots = 1ll << 34;
    while ((double)slots * 12.0 > 0.8 * (double)free_b && slots > 1024) slots >>= 1;
    // pigeonhole part geometry (substitution neighbourhoods only): hdist+1 disjoint parts of w bases that
    // avoid the masked middle; part j ends lag[j] bases before the window end
    n_filter_words = filter_words;
    part_words = short_words = samp_words = tail_words = 0;
    n_parts = part_w = tail_q = 0;
    if (!edits && dist_full <= 3 && n_full > 0) {
        const int P = dist_full + 1, k = p.k, mml = p.midMaskLen;
        int offs[4] = {0, 0, 0, 0}, w = 0;
        if (mml == 0) {
            const int stride = k / P;
            w = std::min(16, stride);
            for (int j = 0; j < P; j++) offs[j] = (j == P - 1) ? k - w : j * stride;
        } else {
            const int a = k - 1 - ((k - mml) / 2 + mml - 1);  // first masked window position
            const int lenL = a, lenR = k - a - mml;
            const int PL = (P + 1) / 2, PR = P / 2;
            const int strideL = lenL / PL, strideR = PR ? lenR / PR : 99;
            w = std::min(16, std::min(strideL, strideR));
            for (int j = 0; j < PL; j++) offs[j] = j * strideL;
            for (int j = 0; j < PR; j++) offs[PL + j] = (j == PR - 1) ? k - w : a + mml + j * strideR;
        }
        // 9-mers the bitmap receives (an upper bound: adapter-like references repeat themselves); past half
        // of the 4^9 values the AND over a part's 9-mers stops rejecting anything
        const double entries = (double)n_full * (p.rcomp ? 2 : 1) * P * (w - BB_PART_WD + 1);
        const uint32_t pw = BB_PART_WORDS;  // 32 KB
        if (w >= BB_PART_WD && entries <= (double)pw * 16.0) {
            n_parts = P;
            part_w = w;
            for (int j = 0; j < P; j++) part_lag[j] = k - offs[j] - w;
            part_words = pw;
            short_words = p.useShortKmers ? 16384 : 0;  // 64 KB: at 32 KB every third tail iteration of a warp went to the table for a false positive
            // probe_fast2.cu: the sampled scan needs a part to hold a 4-aligned 8-mer whatever its phase (w >= 11)
            samp_words = (w >= 11) ? 16384 : 0;  // 4^8 bytes
            tail_q = 0;
            tail_words = 0;
            if (samp_words && p.useShortKmers && dist_short <= 3 && p.mink >= 1) {
                tail_q = std::min(p.mink, 12);
                tail_words = 2 * std::max<uint32_t>(1u, (1u << (2 * tail_q)) >> 5);
                if (tail_q >= 9) tail_words += 2 * BB_TAIL0_WORDS;  // 8-mer level-0 images of both bitmaps (tail_prefix_kernel)
            }
        }
    }
    if (alloc(slots, total_filter_words(), err, errlen)) {
        cleanup();
        return 1;
    }
    fill_kernel<<<1184, 256, 0, st>>>(d_keys, d_vals, slots);
    (*launches)++;
    CKC(cudaMemsetAsync(d_filter, 0, sizeof(uint32_t) * total_filter_words(), st));
    if (part_words) {
        part_filter_build_kernel<<<296, 256, 0, st>>>(d_full, n_full, p.k, p.rcomp, n_parts, part_w, part_lag[0], part_lag[1],
                                                      part_lag[2], part_lag[3], d_filter + n_filter_words, part_words,
                                                      samp_words ? reinterpret_cast<uint8_t *>(d_filter + n_filter_words + part_words + short_words)
                                                                 : nullptr);
        (*launches)++;
    }
    if (tail_words) {
        uint32_t *b1 = d_filter + n_filter_words + part_words + short_words + samp_words;
        const uint32_t main_words = std::max<uint32_t>(1u, (1u << (2 * tail_q)) >> 5);
        tail_filter_build_kernel<<<64, 64, 0, st>>>(d_short + (int64_t)(p.k - 1) * cap_short, (int64_t)h_cnt[1 + p.k - 1], p.k - 1,
                                                    tail_q, dist_short, p.rcomp, b1, b1 + main_words);
        (*launches)++;
        if (tail_words > 2 * main_words) {
            tail_prefix_kernel<<<296, 256, 0, st>>>(b1, main_words, tail_q, b1 + 2 * main_words);
            tail_prefix_kernel<<<296, 256, 0, st>>>(b1 + main_words, main_words, tail_q, b1 + 2 * main_words + BB_TAIL0_WORDS);
            (*launches) += 2;
        }
    }

    auto launch_expand = [&](const Seed *seeds, int64_t n, int len, int dist, int use_extra) -> int {
        if (n <= 0) return 0;
        const double nops = 4.0 * len + (edits ? 5.0 * (len - 1) : 0.0);
        double n_prefix = 1;
        for (int d = 1; d < dist; d++) n_prefix *= nops;
        const double threads = (double)n * n_prefix;
        // keep each launch below 2^31 blocks' worth of threads by slicing the seed list
        const int64_t max_seeds = std::max<int64_t>(1, (int64_t)(4.0e9 / n_prefix));
        for (int64_t s0 = 0; s0 < n; s0 += max_seeds) {
            const int64_t ns = std::min(max_seeds, n - s0);
            const int64_t nt = (int64_t)((double)ns * n_prefix);
            const int64_t nb = (nt + 127) / 128;
            expand_kernel<<<(unsigned)nb, 128, 0, st>>>(seeds + s0, ns, len, dist, use_extra, p, d_keys, d_vals,
                                                        (uint64_t)slots - 1, d_cnt + 34, (int *)(d_cnt + 36));
            (*launches)++;
        }
        (void)threads;
        return 0;
    };
    launch_expand(d_full, n_full, p.k, dist_full, ux_full);
    if (p.useShortKmers)
        for (int n = p.k - 1; n >= p.mink && n >= 1; n--)
            launch_expand(d_short + (int64_t)n * cap_short, (int64_t)h_cnt[1 + n], n, dist_short, ux_short);
    CKC(cudaGetLastError());
    CKC(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    CKC(cudaStreamSynchronize(st));
    stored = (int64_t)h_cnt[34];
    if (h_cnt[36] != 0 || (double)stored > 0.9 * (double)slots) {
        snprintf(err, errlen, "device hash array overfull (%lld keys in %lld slots)", (long long)stored, (long long)slots);
        cleanup();
        return 1;
    }

    // duplicates collapse heavily for adapter-like references: shrink to the requested load factor
    const int64_t want = pow2ceil((int64_t)((double)stored * 100.0 / load_pct) + 4);
    if (want < slots) {
        uint64_t *nk = nullptr;
        int32_t *nv = nullptr;
        CKC(cudaMalloc(&nk, sizeof(uint64_t) * (size_t)want));
        CKC(cudaMalloc(&nv, sizeof(int32_t) * (size_t)want));
        fill_kernel<<<1184, 256, 0, st>>>(nk, nv, want);
        rehash_kernel<<<1184, 256, 0, st>>>(d_keys, d_vals, slots, nk, nv, (uint64_t)want - 1, (int *)(d_cnt + 36));
        (*launches) += 2;
        CKC(cudaStreamSynchronize(st));
        cudaFree(d_keys);
        cudaFree(d_vals);
        d_keys = nk;
        d_vals = nv;
        n_slots = want;
    }
    filter_build_kernel<<<1184, 256, 0, st>>>(d_keys, n_slots, d_filter, n_filter_words);
    (*launches)++;
    if (short_words) {
        short_filter_build_kernel<<<1184, 256, 0, st>>>(d_keys, n_slots, p.kmask, d_filter + n_filter_words + part_words,
                                                         short_words);
        (*launches)++;
    }
    // HBM-resident arrays (too many keys for the on-chip images): an L2-resident one-bit-per-key filter, up to
    // a cap well inside the 126 MB L2, appended to the filter buffer so that it replicates with it
    big_words = 0;
    if (stored > (int64_t)n_filter_words * 24) {
        // ~2.75 bits per key: measured best on cfg 3 (1e8 keys -> 32 MB); beyond 64 MB the filter starts to fall
        // out of the L2 next to the streaming reads (cfg 4, 2.9e8 keys: 64 MB best, 80 MB slower)
        const uint64_t want_words = std::max<uint64_t>(1u << 16, (uint64_t)((double)stored * 2.75 / 32.0));
        uint64_t cap_mb = 64;
        if (const char *e = getenv("BBDUK_B200_BIGFILTER_MB")) cap_mb = (uint64_t)std::max(1, atoi(e));
        const uint32_t bw = (uint32_t)std::min<uint64_t>(want_words, (cap_mb << 20) / 4);
        uint32_t *nf = nullptr;
        const size_t old_words = total_filter_words();
        CKC(cudaMalloc(&nf, sizeof(uint32_t) * (old_words + bw)));
        CKC(cudaMemcpyAsync(nf, d_filter, sizeof(uint32_t) * old_words, cudaMemcpyDeviceToDevice, st));
        CKC(cudaMemsetAsync(nf + old_words, 0, sizeof(uint32_t) * (size_t)bw, st));
        big_filter_build_kernel<<<1184, 256, 0, st>>>(d_keys, n_slots, nf + old_words, bw);
        (*launches)++;
        CKC(cudaStreamSynchronize(st));
        cudaFree(d_filter);
        d_filter = nf;
        big_words = bw;
    }
    CKC(cudaGetLastError());
    CKC(cudaStreamSynchronize(st));
    cleanup();
    return 0;
#undef CKC
}
