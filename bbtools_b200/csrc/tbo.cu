// tbo.cu -- BBDuk's trim-by-overlap step (tbo=t) on the device: bbduk_b200_tbo / bbduk_b200_tbo_device.
//
// Replaces jgi/BBDuk.java:2878-2926 (guard, reverse complement of r2, BBMergeOverlapper.mateByOverlapRatio, minInsert
// cut, trimToPosition of both mates) with jgi/BBMergeOverlapper.java:411-621 (mateByOverlapRatioJava) and :785-836
// (findBestRatio) underneath. One lane per pair, the reference's two insert loops unchanged in structure and in
// single-precision evaluation order (__fmul_rn / __fadd_rn / __fdiv_rn: nothing is contracted). What changes is the
// inner base-by-base loop: the reference adds 0.95f per matching / mismatching base and leaves the loop once
// bad > badlimit. Because those partial sums only ever grow, "the loop ran to its end" is the same as
// T[mismatches] <= badlimit with T[c] = 0.95f added c times, so the kernel COUNTS matches and mismatches of an
// alignment 16 bases at a time on 2-bit packed copies of the trimmed mates (xor, fold, popc; N and "exactly one is N"
// from a second bit stream) and looks the float sums up in T. Mates holding any byte other than A C G T N (lower case,
// IUPAC) take an exact byte-wise path, since the reference compares raw bytes.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/bbduk_b200.h"

namespace {

constexpr int TBO_THREADS = 128;
constexpr int TBO_MAX_LEN = 1008;
constexpr int EXTRA_BADLIMIT = 20;  // jgi/BBMergeOverlapper.java:1464

struct TboDev {
    int minOverlap0, minOverlap, minInsert0, minInsert;  // BBDuk's values (jgi/BBDuk.java:5368-5371)
    float maxRatio, minSecondRatio, margin, offset, meeFilter;
    int qualOffset;
    int W;  // words per packed stream per lane
};

struct PairCtx {
    // packed fast path: little-endian 2-bit streams in shared memory, word w of this lane at base[w * TBO_THREADS]
    const uint32_t *a2, *aN, *b2, *bN;
    // exact path
    const uint8_t *a_bytes;      // r1 trimmed, forward
    const uint8_t *b_rev_bytes;  // r2 trimmed, LAST base (b[j] = comp[b_rev_bytes[-j]])
    const uint8_t *comp;
    bool exact;
    bool has_n;  // some base of either mate is 'N' (else the N streams are all zero and are skipped)
};

// mismatches / non-N matches of a[istart..istart+ov) against b[jstart..jstart+ov). The reference leaves its loop once
// bad > badlimit; T only grows, so the caller needs exact counts only while T[nbad] <= badlimit: counting stops as
// soon as nbad exceeds the largest count whose partial sum still fits (then ngood is not used).
template <bool GENERAL>
__device__ __forceinline__ void count_alignment(const PairCtx &c, int istart, int jstart, int ov, float badlimit,
                                                const float *T, int n_T, int &nbad, int &ngood) {
    nbad = 0;
    ngood = 0;
    (void)n_T;
    if (!GENERAL || !c.exact) {
        constexpr int S = TBO_THREADS;
        int wa = (istart >> 4) * S, wb = (jstart >> 4) * S;
        const int sa = (istart & 15) * 2, sb = (jstart & 15) * 2;
        uint32_t a_lo = c.a2[wa], b_lo = c.b2[wb];
        if (!GENERAL) {
            int t = 0;
            for (; t + 16 <= ov; t += 16) {
                wa += S;
                wb += S;
                const uint32_t a_hi = c.a2[wa], b_hi = c.b2[wb];
                const uint32_t x = __funnelshift_r(a_lo, a_hi, sa) ^ __funnelshift_r(b_lo, b_hi, sb);
                nbad += __popc((x | (x >> 1)) & 0x55555555u);
                a_lo = a_hi;
                b_lo = b_hi;
                if (T[nbad] > badlimit) return;
            }
            if (t < ov) {
                const uint32_t x = __funnelshift_r(a_lo, c.a2[wa + S], sa) ^ __funnelshift_r(b_lo, c.b2[wb + S], sb);
                nbad += __popc((x | (x >> 1)) & 0x55555555u & ((1u << (2 * (ov - t))) - 1u));
            }
            ngood = ov - nbad;
        } else {
            uint32_t an_lo = c.aN[wa], bn_lo = c.bN[wb];
            for (int t = 0; t < ov; t += 16) {
                wa += S;
                wb += S;
                const uint32_t a_hi = c.a2[wa], b_hi = c.b2[wb], an_hi = c.aN[wa], bn_hi = c.bN[wb];
                const uint32_t x = __funnelshift_r(a_lo, a_hi, sa) ^ __funnelshift_r(b_lo, b_hi, sb);
                const uint32_t na = __funnelshift_r(an_lo, an_hi, sa), nb = __funnelshift_r(bn_lo, bn_hi, sb);
                const int m = min(16, ov - t);
                const uint32_t mask = (m >= 16) ? 0x55555555u : (((1u << (2 * m)) - 1u) & 0x55555555u);
                const uint32_t d = (x | (x >> 1)) & 0x55555555u;
                const uint32_t nn = ~(na | nb) & 0x55555555u;
                nbad += __popc(((d & nn) | (na ^ nb)) & mask);
                ngood += __popc(~d & nn & mask);
                a_lo = a_hi;
                b_lo = b_hi;
                an_lo = an_hi;
                bn_lo = bn_hi;
                if (T[nbad] > badlimit) return;
            }
        }
    } else {
        for (int t = 0; t < ov; t++) {
            const uint8_t ca = c.a_bytes[istart + t];
            const uint8_t cb = c.comp[c.b_rev_bytes[-(jstart + t)] & 127];
            if (ca == cb) {
                if (ca != 'N') ngood++;
            } else {
                nbad++;
                if (T[nbad] > badlimit) return;
            }
        }
    }
}

// jgi/BBMergeOverlapper.java:785-836
template <bool GENERAL>
__device__ float find_best_ratio(const PairCtx &c, int alen, int blen, int minOverlap0, int minOverlap, int minInsert,
                                 float maxRatio, float offset, const float *T, int n_T) {
    float bestRatio = __fadd_rn(maxRatio, 0.0001f);
    const float halfmax = __fmul_rn(maxRatio, 0.5f);
    for (int insert = alen + blen - minOverlap; insert >= minInsert; insert--) {
        const int istart = (insert <= blen ? 0 : insert - blen);
        const int jstart = (insert >= blen ? 0 : blen - insert);
        const int ov = min(alen - istart, min(blen - jstart, insert));
        const float badlimit = __fadd_rn(__fmul_rn(bestRatio, (float)ov), (float)EXTRA_BADLIMIT);
        int nbad, ngood;
        count_alignment<GENERAL>(c, istart, jstart, ov, badlimit, T, n_T, nbad, ngood);
        const float bad = T[nbad];
        if (bad <= badlimit) {
            const float good = T[ngood];
            if (bad == 0.0f && good > (float)minOverlap0 && good < (float)minOverlap) return 100.0f;
            const float ratio = __fdiv_rn(__fadd_rn(bad, offset), (float)ov);
            if (ratio < bestRatio) {
                bestRatio = ratio;
                if (good >= (float)minOverlap && ratio < halfmax) return bestRatio;
            }
        }
    }
    return bestRatio;
}

// jgi/BBMergeOverlapper.java:411-621 (TAG_CUSTOM = MAKE_VECTOR = false); returns bestInsert, sets ambig
// STAGE 0: both loops. STAGE 1: findBestRatio only; returns -3 and *x_io if the second loop has to run.
// STAGE 2: the second loop, with findBestRatio's result handed in through *x_io.
template <bool GENERAL, int STAGE>
__device__ int mate_by_overlap_ratio(const PairCtx &c, int alen, int blen, const TboDev &p, const float *T, int n_T,
                                     bool &ambig_out, float *x_io) {
    const int minOverlap = max(4, max(p.minOverlap0, p.minOverlap));
    int minOverlap0;
    {  // Tools.mid(4, minOverlap0, minOverlap): the median
        const int x = 4, y = p.minOverlap0, z = minOverlap;
        minOverlap0 = x < y ? (y < z ? y : max(x, z)) : (x < z ? x : max(y, z));
    }
    const int minLength = min(alen, blen);
    float maxRatio = p.maxRatio;
    ambig_out = false;
    {
        float x;
        if (STAGE == 2) x = *x_io;
        else x = find_best_ratio<GENERAL>(c, alen, blen, minOverlap0, minOverlap, p.minInsert, maxRatio, p.offset, T, n_T);
        if (x > maxRatio) return -1;  // rvector[4] = 0
        if (STAGE == 1) {
            *x_io = x;
            return -3;
        }
        maxRatio = fminf(maxRatio, x);
    }
    const float margin = p.margin, offset = p.offset;
    const float margin2 = __fdiv_rn(__fadd_rn(margin, offset), (float)minLength);
    int bestInsert = -1;
    float bestRatio = 1.0f, secondBestRatio = 1.0f;
    bool ambig = false;
    for (int insert = alen + blen - minOverlap0; insert >= p.minInsert0; insert--) {
        const int istart = (insert <= blen ? 0 : insert - blen);
        const int jstart = (insert >= blen ? 0 : blen - insert);
        const int ov = min(alen - istart, min(blen - jstart, insert));
        const float badlimit =
            __fadd_rn(__fadd_rn(__fmul_rn(1.2f, __fmul_rn(__fmul_rn(fminf(bestRatio, maxRatio), margin), (float)ov)), 1.0f),
                      (float)EXTRA_BADLIMIT);
        int nbad, ngood;
        count_alignment<GENERAL>(c, istart, jstart, ov, badlimit, T, n_T, nbad, ngood);
        const float bad = T[nbad];
        if (bad <= badlimit) {
            const float good = T[ngood];
            if (bad == 0.0f && good > (float)minOverlap0 && good < (float)minOverlap) {
                ambig_out = true;
                return -1;
            }
            const float ratio = __fdiv_rn(__fadd_rn(bad, offset), (float)ov);
            if (ratio < __fmul_rn(bestRatio, margin)) {
                ambig = (__fmul_rn(ratio, margin) >= bestRatio || good < (float)minOverlap);
                if (ratio < bestRatio) {
                    secondBestRatio = bestRatio;
                    bestInsert = insert;
                    bestRatio = ratio;
                } else if (ratio < secondBestRatio) {
                    secondBestRatio = ratio;
                }
                if ((ambig && bestRatio < margin2) || secondBestRatio < p.minSecondRatio) {
                    ambig_out = true;
                    return -1;
                }
            }
        }
    }
    if (!ambig && bestRatio > maxRatio) bestInsert = -1;
    ambig_out = ambig;
    return bestInsert;
}

__device__ __forceinline__ bool fully_defined(uint8_t b) {
    const uint8_t y = b | 0x20;
    return b < 128 && (y == 'a' || y == 'c' || y == 'g' || y == 't' || y == 'u');
}

// Three launches keep the lanes of a warp on one code path with similar trip counts:
//   MODE 0  every pair whose mates are made of A C G T only: pack, findBestRatio. ~70 % of the pairs end here (no overlap
//           worth a second look); the others go to list M with their ratio; pairs with an N or any other byte go to list G.
//   MODE 1  list M, compacted: the second insert loop of mateByOverlapRatioJava.
//   MODE 2  list G, compacted: both loops with the N streams / the exact byte path.
template <int MODE>
__global__ void __launch_bounds__(TBO_THREADS)
tbo_kernel(const uint8_t *__restrict__ bases, const uint8_t *__restrict__ quals, const uint32_t *__restrict__ offsets,
           int64_t n_pairs, const int32_t *__restrict__ lo, int32_t *hi, uint8_t *flags, int32_t *insert_out, TboDev p,
           const float *__restrict__ T_g, int n_T, const float *__restrict__ prob_error_g, const uint8_t *__restrict__ comp_g,
           unsigned long long *stats, int32_t *list_g, unsigned int *list_g_n, int32_t *list_m, float *list_m_x,
           unsigned int *list_m_n) {
    constexpr bool GENERAL = MODE == 2;
    extern __shared__ __align__(16) uint32_t smem[];
    float *T = reinterpret_cast<float *>(smem);
    float *prob_error = T + n_T;
    uint8_t *comp = reinterpret_cast<uint8_t *>(prob_error + 128);
    uint32_t *streams = reinterpret_cast<uint32_t *>(comp + 128);
    for (int i = threadIdx.x; i < n_T; i += TBO_THREADS) T[i] = T_g[i];
    for (int i = threadIdx.x; i < 128; i += TBO_THREADS) {
        prob_error[i] = prob_error_g[i];
        comp[i] = comp_g[i];
    }
    __syncthreads();
    const int W = p.W;
    uint32_t *a2 = streams + threadIdx.x, *aN = a2 + (size_t)W * TBO_THREADS, *b2 = aN + (size_t)W * TBO_THREADS,
             *bN = b2 + (size_t)W * TBO_THREADS;
    unsigned long long n_trim = 0, b_trim = 0;
    const int64_t n_items = MODE == 2 ? (int64_t)*list_g_n : MODE == 1 ? (int64_t)*list_m_n : n_pairs;
    for (int64_t item = (int64_t)blockIdx.x * TBO_THREADS + threadIdx.x; item < n_items; item += (int64_t)gridDim.x * TBO_THREADS) {
        const int64_t pair = MODE == 2 ? (int64_t)list_g[item] : MODE == 1 ? (int64_t)list_m[item] : item;
        const int64_t i1 = 2 * pair, i2 = i1 + 1;
        int best = -1;
        bool ambig = false;
        const bool removed = (flags[i1] & BBDUK_F_REMOVED) != 0;
        const int lo1 = lo[i1], lo2 = lo[i2];
        const int alen = hi[i1] - lo1, blen = hi[i2] - lo2;
        const uint8_t *a = bases + offsets[i1] + lo1;
        const uint8_t *b0 = bases + offsets[i2] + lo2;
        bool run = !removed;
        if (MODE != 1 && run && quals) {  // expectedErrors(r1, r2) < meeFilter (jgi/BBDuk.java:2878, stream/Read.java:2985-3003)
            float ea = 0.0f, eb = 0.0f;
            const uint8_t *qa = quals + offsets[i1] + lo1, *qb = quals + offsets[i2] + lo2;
            for (int i = 0; i < alen; i++)
                if (fully_defined(a[i])) ea = __fadd_rn(ea, prob_error[(qa[i] - p.qualOffset) & 127]);
            for (int i = 0; i < blen; i++)
                if (fully_defined(b0[i])) eb = __fadd_rn(eb, prob_error[(qb[i] - p.qualOffset) & 127]);
            run = fmaxf(ea, eb) < p.meeFilter;
        }
        if (run && (alen > 16 * (W - 1) || blen > 16 * (W - 1))) run = false;  // guarded on the host: cannot happen
        if (run) {
            // pack r1 forward and r2 reverse-complemented; any byte outside A C G T N switches to the exact path
            bool exact = false, has_n = false;
            auto pack = [&](const uint8_t *src, int len, bool rc, uint32_t *s2, uint32_t *sN) {
                uint32_t w2 = 0, wN = 0;
                int wi = 0;
                for (int i = 0; i < len; i++) {
                    const uint8_t ch = rc ? src[len - 1 - i] : src[i];
                    uint32_t code = 0, isn = 0;
                    if (ch == 'A') code = 0;
                    else if (ch == 'C') code = 1;
                    else if (ch == 'G') code = 2;
                    else if (ch == 'T') code = 3;
                    else if (ch == 'N') {
                        isn = 1;
                        has_n = true;
                    }
                    else exact = true;
                    if (rc) code = 3u - code;
                    const int sh = 2 * (i & 15);
                    w2 |= code << sh;
                    wN |= isn << sh;
                    if ((i & 15) == 15) {
                        s2[wi * TBO_THREADS] = w2;
                        sN[wi * TBO_THREADS] = wN;
                        wi++;
                        w2 = wN = 0;
                    }
                }
                s2[wi * TBO_THREADS] = w2;
                sN[wi * TBO_THREADS] = wN;
                if (wi + 1 < W) {
                    s2[(wi + 1) * TBO_THREADS] = 0;
                    sN[(wi + 1) * TBO_THREADS] = 0;
                }
            };
            pack(a, alen, false, a2, aN);
            pack(b0, blen, true, b2, bN);
            PairCtx c;
            c.a2 = a2;
            c.aN = aN;
            c.b2 = b2;
            c.bN = bN;
            c.a_bytes = a;
            c.b_rev_bytes = b0 + blen - 1;
            c.comp = comp;
            c.exact = exact;
            c.has_n = has_n;
            if (MODE == 0 && (has_n || exact)) {  // left to the general launch
                list_g[atomicAdd(list_g_n, 1u)] = (int32_t)pair;
                continue;
            }
            float x = MODE == 1 ? list_m_x[item] : 0.0f;
            best = mate_by_overlap_ratio<GENERAL, MODE == 0 ? 1 : MODE == 1 ? 2 : 0>(c, alen, blen, p, T, n_T, ambig, &x);
            if (MODE == 0 && best == -3) {  // the second loop runs in the compacted launch
                const unsigned int w = atomicAdd(list_m_n, 1u);
                list_m[w] = (int32_t)pair;
                list_m_x[w] = x;
                continue;
            }
            if (best < p.minInsert) best = -1;
            if (best > 0 && !ambig) {  // TrimRead.trimToPosition(r, 0, bestInsert-1, 1)
                if (best < alen) {
                    hi[i1] = lo1 + best;
                    flags[i1] |= BBDUK_F_TBO;
                    n_trim++;
                    b_trim += alen - best;
                }
                if (best < blen) {
                    hi[i2] = lo2 + best;
                    flags[i2] |= BBDUK_F_TBO;
                    n_trim++;
                    b_trim += blen - best;
                }
            }
        }
        if (insert_out) insert_out[pair] = (best > 0 && !ambig) ? best : (ambig ? -2 : -1);
    }
    for (int o = 16; o > 0; o >>= 1) {
        n_trim += __shfl_xor_sync(0xFFFFFFFFu, n_trim, o);
        b_trim += __shfl_xor_sync(0xFFFFFFFFu, b_trim, o);
    }
    if (stats && (threadIdx.x & 31) == 0 && n_trim) {
        atomicAdd(stats, n_trim);
        atomicAdd(stats + 1, b_trim);
    }
}

struct TboTables {
    float *d_T = nullptr, *d_pe = nullptr;
    uint8_t *d_comp = nullptr;
    int32_t *d_list = nullptr, *d_list_m = nullptr;  // pairs left to the general launch / to the second-loop launch
    float *d_list_x = nullptr;
    unsigned int *d_list_n = nullptr;                // [0] general, [1] second loop
    int64_t list_cap = 0;
    int device = -1;
    float incr = 0;
};
std::mutex g_tab_mu;
std::vector<TboTables> g_tabs;

int get_tables(int device, int64_t n_pairs, TboTables *out) {
    std::lock_guard<std::mutex> g(g_tab_mu);
    for (auto &t : g_tabs)
        if (t.device == device) {
            if (n_pairs > t.list_cap) {
                cudaDeviceSynchronize();
                cudaFree(t.d_list);
                cudaFree(t.d_list_m);
                cudaFree(t.d_list_x);
                t.list_cap = n_pairs + n_pairs / 8 + 1024;
                if (cudaMalloc(&t.d_list, sizeof(int32_t) * t.list_cap) != cudaSuccess ||
                    cudaMalloc(&t.d_list_m, sizeof(int32_t) * t.list_cap) != cudaSuccess ||
                    cudaMalloc(&t.d_list_x, sizeof(float) * t.list_cap) != cudaSuccess)
                    return 1;
            }
            *out = t;
            return 0;
        }
    TboTables t;
    t.device = device;
    std::vector<float> T(TBO_MAX_LEN + 2), pe(128);
    T[0] = 0.0f;
    for (size_t c = 1; c < T.size(); c++) {
        volatile float s = T[c - 1] + 0.95f;  // bad+=bIncr, single precision, one rounding per add
        T[c] = s;
    }
    for (int i = 0; i < 128; i++) pe[i] = (float)pow(10.0, 0 - .1 * i);  // align2/QualityTools.java:688-698
    pe[0] = .75f;
    pe[1] = .7f;
    uint8_t comp[128];
    {  // dna/AminoAcid.java:1315-1332 with the tables of :206-231
        const char *nb = " ACMGRSVTWYHKDBNX       ", *nc = " TGKCYSBAWRDMHVNX       ";
        for (int i = 0; i < 128; i++) comp[i] = (uint8_t)i;
        for (int i = 0; i < 24; i++) {
            const unsigned char x = (unsigned char)nb[i], x2 = (unsigned char)nc[i];
            comp[x] = x2;
            comp[(x >= 'A' && x <= 'Z') ? x + 32 : x] = (uint8_t)((x2 >= 'A' && x2 <= 'Z') ? x2 + 32 : x2);
        }
        comp['U'] = 'A';
        comp['u'] = 'a';
    }
    if (cudaMalloc(&t.d_T, sizeof(float) * T.size()) != cudaSuccess || cudaMalloc(&t.d_pe, sizeof(float) * 128) != cudaSuccess ||
        cudaMalloc(&t.d_comp, 128) != cudaSuccess)
        return 1;
    cudaMemcpy(t.d_T, T.data(), sizeof(float) * T.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(t.d_pe, pe.data(), sizeof(float) * 128, cudaMemcpyHostToDevice);
    cudaMemcpy(t.d_comp, comp, 128, cudaMemcpyHostToDevice);
    t.list_cap = n_pairs + n_pairs / 8 + 1024;
    if (cudaMalloc(&t.d_list, sizeof(int32_t) * t.list_cap) != cudaSuccess ||
        cudaMalloc(&t.d_list_m, sizeof(int32_t) * t.list_cap) != cudaSuccess ||
        cudaMalloc(&t.d_list_x, sizeof(float) * t.list_cap) != cudaSuccess || cudaMalloc(&t.d_list_n, 2 * sizeof(unsigned int)) != cudaSuccess)
        return 1;
    g_tabs.push_back(t);
    *out = t;
    return 0;
}

}  // namespace

// launcher used by abi.cu; returns 0 on success, 1 on CUDA failure, 2 if a read is too long for the device path
int launch_tbo(int device, int sm_count, const bbduk_tbo_cfg *cfg, const uint8_t *d_bases, const uint8_t *d_quals,
               const uint32_t *d_offsets, int64_t n_reads, int max_len, const int32_t *d_lo, int32_t *d_hi, uint8_t *d_flags,
               int32_t *d_insert, unsigned long long *d_stats, cudaStream_t st) {
    if (n_reads < 2) return 0;
    if (max_len > TBO_MAX_LEN) return 2;
    TboTables tab;
    if (get_tables(device, n_reads / 2, &tab)) return 1;
    TboDev p;
    p.minOverlap0 = cfg->min_overlap0 >= 0 ? cfg->min_overlap0 : 7;
    p.minOverlap = cfg->min_overlap >= 0 ? std::max(cfg->min_overlap, 1) : 14;
    p.minOverlap0 = std::min(p.minOverlap0, p.minOverlap);  // jgi/BBDuk.java:657-660
    p.minInsert0 = cfg->min_insert0 >= 0 ? cfg->min_insert0 : 16;
    p.minInsert = cfg->min_insert >= 0 ? std::max(cfg->min_insert, 1) : 40;
    p.minInsert0 = std::min(p.minInsert0, p.minInsert);  // :662-665
    p.minSecondRatio = 0.12f;
    if (cfg->strict_overlap) {  // :712-719
        p.maxRatio = 0.05f;
        p.margin = 9.0f;
        p.offset = 0.5f;
        p.meeFilter = 15.0f;
    } else {  // :720-727
        p.maxRatio = 0.10f;
        p.margin = 5.0f;
        p.offset = 0.4f;
        p.meeFilter = 999999999.0f;
    }
    if (cfg->mee_filter > 0.0f) p.meeFilter = cfg->mee_filter;
    p.qualOffset = cfg->qual_offset > 0 ? cfg->qual_offset : 33;
    p.W = (std::max(max_len, 16) + 15) / 16 + 2;
    const int n_T = TBO_MAX_LEN + 2;
    const size_t smem = sizeof(float) * (n_T + 128) + 128 + sizeof(uint32_t) * 4 * (size_t)p.W * TBO_THREADS;
    if (cudaFuncSetAttribute(tbo_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(tbo_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(tbo_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return 1;
    const int64_t n_pairs = n_reads / 2;
    const int blocks = (int)std::min<int64_t>((n_pairs + TBO_THREADS - 1) / TBO_THREADS, (int64_t)sm_count * 8);
    if (cudaMemsetAsync(tab.d_list_n, 0, 2 * sizeof(unsigned int), st) != cudaSuccess) return 1;
#define TBO_ARGS d_bases, d_quals, d_offsets, n_pairs, d_lo, d_hi, d_flags, d_insert, p, tab.d_T, n_T, tab.d_pe, tab.d_comp, d_stats, \
                 tab.d_list, tab.d_list_n, tab.d_list_m, tab.d_list_x, tab.d_list_n + 1
    tbo_kernel<0><<<blocks, TBO_THREADS, smem, st>>>(TBO_ARGS);
    tbo_kernel<1><<<std::max(1, blocks / 2), TBO_THREADS, smem, st>>>(TBO_ARGS);
    tbo_kernel<2><<<std::max(1, blocks / 4), TBO_THREADS, smem, st>>>(TBO_ARGS);
#undef TBO_ARGS
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
