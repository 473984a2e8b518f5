// tbo.cu -- BBDuk's trim-by-overlap step (tbo=t) on the device: bbduk_b200_tbo / bbduk_b200_tbo_device.
//
// Replaces jgi/BBDuk.java:2878-2926 (guard, reverse complement of r2, BBMergeOverlapper.mateByOverlapRatio, minInsert
// cut, trimToPosition of both mates) with jgi/BBMergeOverlapper.java:411-621 (mateByOverlapRatioJava) and :785-836
// (findBestRatio) underneath. One lane per pair; the per-pair arithmetic lives in tbo_core.cuh (bit planes of 32 bases
// per word, mismatch COUNTS + a table of the float partial sums instead of the reference's base-by-base float adds,
// a 64-base register screen in front of every alignment). Mates holding any byte other than A C G T N (lower case,
// IUPAC) take an exact byte-wise path, since the reference compares raw bytes.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/bbduk_b200.h"
#include "tbo_core.cuh"

namespace {

constexpr int TBO_THREADS = 128;
constexpr int TBO_MAX_LEN = tbo::MAX_LEN;
using TboDev = tbo::Params;

__device__ __forceinline__ bool fully_defined(uint8_t b) {
    const uint8_t y = b | 0x20;
    return b < 128 && (y == 'a' || y == 'c' || y == 'g' || y == 't' || y == 'u');
}

// Four launches keep the lanes of a warp on one code path with similar trip counts:
//   MODE 0  every pair whose mates are made of A C G T only: pack, findBestRatio. ~70 % of the pairs end here (no overlap
//           worth a second look); the others go to list M with their ratio; pairs with an N or any other byte go to list G.
//   MODE 1  list M, compacted: the second insert loop of mateByOverlapRatioJava.
//   MODE 2  list G, compacted: findBestRatio with the N planes / the exact byte path; survivors go to list G2.
//   MODE 3  list G2, compacted: the second insert loop of those (in one launch with MODE 2's work the warps ran at 9 of 32
//           lanes: most pairs are done after findBestRatio).
template <int MODE>
__global__ void __launch_bounds__(TBO_THREADS)
tbo_kernel(const uint8_t *__restrict__ bases, const uint8_t *__restrict__ quals, const uint32_t *__restrict__ offsets,
           int64_t n_pairs, const int32_t *__restrict__ lo, int32_t *hi, uint8_t *flags, int32_t *insert_out, TboDev p,
           const float *__restrict__ T_g, int n_T, const float *__restrict__ prob_error_g, const uint8_t *__restrict__ comp_g,
           unsigned long long *stats, int32_t *list_g, unsigned int *list_g_n, const int32_t *in_list, const tbo::Handoff *in_x,
           const unsigned int *in_n, int32_t *out_list, tbo::Handoff *out_x, unsigned int *out_n) {
    // list_g: where MODE 0 leaves the pairs it cannot pack; in_*: the launch's compacted input (MODE 1, 2, 3); out_*: where a
    // findBestRatio launch (MODE 0, 2) leaves the pairs whose second loop has to run
    constexpr bool GENERAL = MODE >= 2;
    constexpr int STAGE = (MODE == 0 || MODE == 2) ? 1 : 2;
    constexpr int S = TBO_THREADS;
    extern __shared__ __align__(16) uint32_t smem[];
    float *T = reinterpret_cast<float *>(smem);
    float *prob_error = T + n_T;
    uint8_t *comp = reinterpret_cast<uint8_t *>(prob_error + 128);
    uint32_t *planes = reinterpret_cast<uint32_t *>(comp + 128) + threadIdx.x;
    for (int i = threadIdx.x; i < n_T; i += TBO_THREADS) T[i] = T_g[i];
    for (int i = threadIdx.x; i < 128; i += TBO_THREADS) {
        prob_error[i] = prob_error_g[i];
        comp[i] = comp_g[i];
    }
    __syncthreads();
    const int W = p.W;
    unsigned long long n_trim = 0, b_trim = 0;
    const int64_t n_items = MODE == 0 ? n_pairs : (int64_t)*in_n;
    for (int64_t item = (int64_t)blockIdx.x * TBO_THREADS + threadIdx.x; item < n_items; item += (int64_t)gridDim.x * TBO_THREADS) {
        const int64_t pair = MODE == 0 ? item : (int64_t)in_list[item];
        if (MODE == 0) {  // the block's next 128 pairs are one contiguous stretch of the batch: pull it into L2 meanwhile
            const int64_t nxt = item - threadIdx.x + (int64_t)gridDim.x * TBO_THREADS;
            if (nxt < n_items) {
                const int64_t last = nxt + TBO_THREADS < n_items ? nxt + TBO_THREADS : n_items;
                const uint32_t b0p = offsets[2 * nxt] & ~127u, b1p = offsets[2 * last];
                for (uint32_t a_ = b0p + 128u * threadIdx.x; a_ < b1p; a_ += 128u * TBO_THREADS)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(bases + a_));
            }
        }
        const int64_t i1 = 2 * pair, i2 = i1 + 1;
        int best = -1;
        bool ambig = false;
        const bool removed = (flags[i1] & BBDUK_F_REMOVED) != 0;
        const int lo1 = lo[i1], lo2 = lo[i2];
        const int alen = hi[i1] - lo1, blen = hi[i2] - lo2;
        const uint8_t *a = bases + offsets[i1] + lo1;
        const uint8_t *b0 = bases + offsets[i2] + lo2;
        bool run = !removed;
        // expectedErrors(r1, r2) < meeFilter (jgi/BBDuk.java:2878, stream/Read.java:2985-3003); a sum of <= 1008 terms
        // of at most 0.75 cannot reach a filter above 756, so the default of strictoverlap=f skips the loop
        if (STAGE == 1 && run && quals && p.meeFilter <= 0.75f * TBO_MAX_LEN) {
            float ea = 0.0f, eb = 0.0f;
            const uint8_t *qa = quals + offsets[i1] + lo1, *qb = quals + offsets[i2] + lo2;
            for (int i = 0; i < alen; i++)
                if (fully_defined(a[i])) ea = __fadd_rn(ea, prob_error[(qa[i] - p.qualOffset) & 127]);
            for (int i = 0; i < blen; i++)
                if (fully_defined(b0[i])) eb = __fadd_rn(eb, prob_error[(qb[i] - p.qualOffset) & 127]);
            run = fmaxf(ea, eb) < p.meeFilter;
        }
        if (run && (tbo::plane_words(alen) > W || tbo::plane_words(blen) > W)) run = false;  // guarded on the host: cannot happen
        if (run) {
            tbo::Ctx<S> c;
            tbo::Cands<S> q;
            c.comp = comp;
            const uint32_t what = tbo::pack_pair<GENERAL, S>(a, alen, b0, blen, planes, W, c, q);
            if (MODE == 0 && (what & 1u)) {  // an N or another byte: left to the general launch
                list_g[atomicAdd(list_g_n, 1u)] = (int32_t)pair;
                continue;
            }
            tbo::Handoff x = STAGE == 2 ? in_x[item] : tbo::Handoff{0.0f, -1};
            best = tbo::mate_by_overlap_ratio<GENERAL, STAGE, S>(c, q, alen, blen, p, T, n_T, ambig, &x);
            if (STAGE == 1 && best == -3) {  // the second loop runs in the compacted launch
                const unsigned int w = atomicAdd(out_n, 1u);
                out_list[w] = (int32_t)pair;
                out_x[w] = x;
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
    tbo::Handoff *d_list_x = nullptr;
    unsigned int *d_list_n = nullptr;                // [0] list G, [1] list M, [2] list G2
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
                    cudaMalloc(&t.d_list_x, sizeof(tbo::Handoff) * t.list_cap) != cudaSuccess)
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
        cudaMalloc(&t.d_list_x, sizeof(tbo::Handoff) * t.list_cap) != cudaSuccess || cudaMalloc(&t.d_list_n, 4 * sizeof(unsigned int)) != cudaSuccess)
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
    p.W = tbo::plane_words(std::max(max_len, 16));
    const int n_T = TBO_MAX_LEN + 2;
    // per lane: 8 arrays (AH AL BH BL, 2 raw planes, 2 candidate bitmaps) in the A C G T launches, 11 (+ AN BN + 1 raw) in the general one
    const size_t smem_fixed = sizeof(float) * (n_T + 128) + 128;
    const size_t smem6 = smem_fixed + sizeof(uint32_t) * tbo::N_PLANES_ACGT * (size_t)p.W * TBO_THREADS;
    const size_t smem9 = smem_fixed + sizeof(uint32_t) * tbo::N_PLANES_GENERAL * (size_t)p.W * TBO_THREADS;
    if (cudaFuncSetAttribute(tbo_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6) != cudaSuccess ||
        cudaFuncSetAttribute(tbo_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6) != cudaSuccess ||
        cudaFuncSetAttribute(tbo_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem9) != cudaSuccess ||
        cudaFuncSetAttribute(tbo_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem9) != cudaSuccess)
        return 1;
    const int64_t n_pairs = n_reads / 2;
    const int blocks = (int)std::min<int64_t>((n_pairs + TBO_THREADS - 1) / TBO_THREADS, (int64_t)sm_count * 8);
    // counters: [0] list G (pairs with N / other bytes), [1] list M (second loop of the A C G T pairs), [2] list G2 (second loop
    // of the G pairs; it reuses M's arrays, which launch 1 has consumed by then)
    if (cudaMemsetAsync(tab.d_list_n, 0, 4 * sizeof(unsigned int), st) != cudaSuccess) return 1;
    unsigned int *n_g = tab.d_list_n, *n_m = tab.d_list_n + 1, *n_g2 = tab.d_list_n + 2;
#define TBO_ARGS d_bases, d_quals, d_offsets, n_pairs, d_lo, d_hi, d_flags, d_insert, p, tab.d_T, n_T, tab.d_pe, tab.d_comp, d_stats, tab.d_list, n_g
    tbo_kernel<0><<<blocks, TBO_THREADS, smem6, st>>>(TBO_ARGS, nullptr, nullptr, nullptr, tab.d_list_m, tab.d_list_x, n_m);
    tbo_kernel<1><<<blocks, TBO_THREADS, smem6, st>>>(TBO_ARGS, tab.d_list_m, tab.d_list_x, n_m, nullptr, nullptr, nullptr);
    tbo_kernel<2><<<blocks, TBO_THREADS, smem9, st>>>(TBO_ARGS, tab.d_list, nullptr, n_g, tab.d_list_m, tab.d_list_x, n_g2);
    tbo_kernel<3><<<blocks, TBO_THREADS, smem9, st>>>(TBO_ARGS, tab.d_list_m, tab.d_list_x, n_g2, nullptr, nullptr, nullptr);
#undef TBO_ARGS
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
