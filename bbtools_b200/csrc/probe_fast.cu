// probe_fast.cu -- the tuned per-read kernel for the common BBDuk configurations on sm_100a.
//
// Covers ktrim=r, ktrim=l and kfilter (countSetKmers, maxbadkmers=0) with qhdist=0, speed=0, qskip=1,
// no restrictleft/right, k<=31, minkmerfraction=0; everything else, and any tile of reads that does
// not fit the staging, is handed to probe_generic.cu. Results are identical by construction: this
// kernel only decides WHICH read positions can possibly hit (an on-chip blocked-bloom image of all
// table keys, no false negatives) and then evaluates those positions with the exact key formula and
// the exact hash array.
//
// Per warp, per tile of 32 consecutive reads (one lane per read; mates are neighbouring lanes):
//   A. the tile's contiguous ASCII bytes are read once with coalesced 16-byte loads and converted
//      to a 2-bit big-endian stream F (+ 1 bit/base "defined" stream D) in shared memory;
//   B. each lane walks its read 16 positions at a time; forward k-mer and reverse-complement k-mer
//      come from funnel shifts of the packed stream with compile-time shift amounts, then
//      canonical max -> middle mask -> 32-bit hash -> one shared-memory filter word -> bit test;
//      survivors (plus every window that touches an undefined base) become candidates;
//   C. candidates of ALL lanes are pooled in a warp queue (shuffle prefix sums) and evaluated
//      32 at a time with every lane busy: exact key, one 32-byte bucket load from the hash array
//      in L2/HBM; hits fold into per-read (first position,id) / last position with shared atomics;
//      the short-k-mer tails run for reads without a full-length hit;
//   D. trim arithmetic (TrimRead rules), minlen, pair logic (rieb, tpe) via lane shuffles, coalesced
//      result stores, warp-aggregated counters.
// Rolling-state semantics follow jgi/BBDuk.java:3882-3900 in the closed form of SURVEY.md A.2.
#include <algorithm>
#include <cstdlib>

#include "bbduk_dev.cuh"
#include "fast_common.cuh"
#include "probe.h"

namespace {

using namespace bbfast;
// Holding back a read's candidates beyond CAND_CAP per drain (and releasing them only if all earlier ones miss) was
// meant for the long streaks of true hits inside an adapter. Measured on cfg 2 it buys nothing at any cap from 2 to 20
// (gpurun_out/candcap.txt: 2.84-2.91 G reads/s) -- the streak that matters is the FALSE one in front of the first true
// hit, which has to be evaluated in full -- and its bookkeeping costs 3 % of the instructions, so it is compiled out.
constexpr bool DEFER = false;
constexpr int CAND_CAP = 20;         // candidates a read without undefined bases may have unresolved at a time
constexpr int QCAP = 32 + 32 * 16;   // queue entries: the drain threshold plus one full step of all lanes

struct FastGeom {
    int warps;        // warps per block
    int nch;          // staged 16-base chunks per warp (capacity, without padding)
    int warp_bytes;   // shared memory per warp
    int nbadw;        // words of the per-tile "chunk has an undefined base" bit mask
    int csteps;       // 16-position steps the candidate buffer holds per lane
    uint32_t nfw;     // words of the main on-chip filter (canonical bloom, or the part filter)
    uint32_t nsw;     // words of the short-key bloom that follows it (part-filter kernels only)
    uint32_t src_off; // word offset of the main filter inside BBTable::filter
    int cand_cap;     // candidates a read without undefined bases may release between two full drains
};

// Debug builds only (nvcc ... -DBB_FAST_COUNT or -DBB_FAST_CLOCK, see scratch/dbgcount.py): where the candidates of the fast kernel come from.
// 0 tiles, 1 steps, 2 candidates queued, 3 evaluation rounds, 4 windows forced by undefined bases, 5 of them left by the
// T-reading pre-pass, 6 iterations of the serial enqueue loop, 7 tiles with an undefined base; warp cycles (clock64, lane 0) per
// tile phase: 8 offsets + stage A, 9 undefined-base pre-pass, 10 scan loop incl. 11 its evaluation rounds, 12 tails, 13 stage D.
// Compiles to nothing otherwise.
#if defined(BB_FAST_COUNT) || defined(BB_FAST_CLOCK)
__device__ unsigned long long bb_fast_dbg[16];
#endif
#ifdef BB_FAST_COUNT  // per-lane atomics on a handful of addresses: the counts are exact, the run is an order of magnitude slower
#define DBG_ADD(i, v)                                                       \
    do {                                                                    \
        const unsigned long long v_ = (unsigned long long)(v);              \
        if (v_) atomicAdd(&bb_fast_dbg[i], v_); /* per lane: any context */ \
    } while (0)
#else
#define DBG_ADD(i, v)
#endif
#ifdef BB_FAST_CLOCK  // a separate build: seven atomics per warp and tile, so that the cycles mean something
#define DBG_CLK(name) const long long name = clock64()
#define DBG_PHASE(i, t0, t1)                                                                   \
    do {                                                                                       \
        if (lane == 0) atomicAdd(&bb_fast_dbg[i], (unsigned long long)((t1) - (t0)));          \
        if (lane == 0 && (i) == 8) atomicAdd(&bb_fast_dbg[0], 1ull);                           \
    } while (0)
#else
#define DBG_CLK(name)
#define DBG_PHASE(i, t0, t1)
#endif

// PARTS selects the scan: false = canonical k-mer against the bloom image of all keys;
// true = forward part_w-mer against the pigeonhole part filter (see bbduk_dev.cuh), one lookup per
// position, the hdist+1 parts of a window being the same lookup at different lags.
// PRE = the first/last full-length hits of every read were already found by probe_direct.cu (pre_first /
// pre_last): stages A-C compile away and only the per-read epilogue (stage D) runs.
// PACKED = the host already converted the batch (hostpack.cpp): stage A copies the F / D words of the tile
// instead of classifying ASCII (pk_F[i], pk_D[i] = bases 16i..16i+15 of the batch; `bases` is unused).
template <int FMODE, bool RCOMP, bool K16, bool PARTS, bool PRE = false, bool PACKED = false>
__global__ void __launch_bounds__(1024, 1)
bbduk_fast_kernel(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ offsets, int64_t n_reads, int paired,
                  BBParams p, BBTable t, bbduk_out out, bbduk_stats *stats, unsigned long long *scaf_reads,
                  unsigned long long *scaf_bases, int32_t *handoff, unsigned int *handoff_n, FastGeom geo,
                  const unsigned long long *__restrict__ pre_first, const int *__restrict__ pre_last,
                  const uint32_t *__restrict__ pk_F, const uint16_t *__restrict__ pk_D) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *filt = smem;
    const uint8_t *filt_bytes = reinterpret_cast<const uint8_t *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t *filt_short = PARTS ? smem + geo.nfw : smem;  // bloom consulted by the short-k-mer tails
    const uint32_t n_short = PARTS ? geo.nsw : geo.nfw;
    uint8_t *wbase = reinterpret_cast<uint8_t *>(smem + geo.nfw + geo.nsw) + (size_t)warp * geo.warp_bytes;
    unsigned long long *first64 = reinterpret_cast<unsigned long long *>(wbase);  // [32] (pos<<32 | id) of the first hit
    int *lastpos = reinterpret_cast<int *>(first64 + 32);                         // [32] last hit position
    uint32_t *badw = reinterpret_cast<uint32_t *>(lastpos + 32);                  // [nbadw] chunks with a non-ACGTU base
    uint32_t *Fs = badw + geo.nbadw;
    uint16_t *Ds = reinterpret_cast<uint16_t *>(Fs + geo.nch + PAD + TAIL);
    uint16_t *queue = Ds + ((geo.nch + PAD + TAIL + 1) & ~1);                     // [QCAP] (owner lane << 11) | position
    uint16_t *cand = queue + QCAP;                                                // [csteps][32] candidate bits per step

    for (uint32_t i = threadIdx.x; i < geo.nfw + geo.nsw; i += blockDim.x) filt[i] = __ldg(t.filter + geo.src_off + i);
    for (int i = lane; i < PAD; i += 32) {
        Fs[i] = 0;
        Ds[i] = 0;
    }
    __syncthreads();

    const Stream st{Fs, Ds};
    const int k = p.k;
    const uint32_t mask_hi = (uint32_t)(p.mask >> 32), mask_lo = (uint32_t)p.mask;
    const uint32_t mm_hi = (uint32_t)(p.middleMask >> 32), mm_lo = (uint32_t)p.middleMask;
    const uint32_t km_hi = (uint32_t)(p.kmask >> 32), km_lo = (uint32_t)p.kmask;
    const int psh0 = 32 - t.part_lag[0], psh1 = 32 - t.part_lag[1], psh2 = 32 - t.part_lag[2],
              psh3 = 32 - t.part_lag[t.n_parts > 3 ? 3 : 2];
    const uintptr_t base_addr = PACKED ? (uintptr_t)0 : reinterpret_cast<uintptr_t>(bases);
    const int64_t n_tiles = (n_reads + 31) >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    long long s_rk = 0, s_bk = 0, s_rf = 0, s_bf = 0, s_ro = 0, s_bo = 0, s_ri = 0, s_bi = 0;

    for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; tile < n_tiles; tile += warps_total) {
        DBG_CLK(t_a0);
        const int64_t r = tile * 32 + lane;
        const bool live = r < n_reads;
        const uint32_t o0 = live ? offsets[r] : 0, o1 = live ? offsets[r + 1] : 0;
        const uint32_t tile_lo = __shfl_sync(0xFFFFFFFFu, o0, 0);
        const int last_lane = (int)min((long long)31, (long long)(n_reads - 1 - tile * 32));
        const uint32_t tile_hi = __shfl_sync(0xFFFFFFFFu, o1, last_lane);
        const uintptr_t a0 = (base_addr + tile_lo) & ~(uintptr_t)15;
        const int nchunks = (int)((base_addr + tile_hi - a0 + 15) >> 4);
        const int L = (int)(o1 - o0);
        const int maxL = __reduce_max_sync(0xFFFFFFFFu, L);
        if (!PRE && (nchunks > geo.nch || maxL > MAX_FAST_LEN || ((maxL + 15) >> 4) > geo.csteps)) {
            // tile does not fit the staging: hand its units to the generic kernel
            if (live && (!paired || !(lane & 1))) {
                const unsigned int w = atomicAdd(handoff_n, 1u);
                handoff[w] = (int32_t)(paired ? (r >> 1) : r);
            }
            continue;
        }
        // ---- A. stage + convert ---------------------------------------------------------------
#pragma unroll 1
        for (int i = lane; i < geo.nbadw; i += 32) badw[i] = 0;
        __syncwarp();
        // the ASCII loads run two trips ahead of their use: a warp waits for DRAM once per tile instead of once per trip
        // (a 4-fold unrolled body did the same but cost 7 KB of instruction cache and 20 % of the step)
        const uint4 *src16 = reinterpret_cast<const uint4 *>(a0);
        uint4 n1 = make_uint4(0, 0, 0, 0), n2 = make_uint4(0, 0, 0, 0);
        if (!PACKED && !PRE) {
            if (lane < nchunks) n1 = __ldg(src16 + lane);
            if (lane + 32 < nchunks) n2 = __ldg(src16 + lane + 32);
        }
#pragma unroll 1
        for (int c = lane; c < (PRE ? 0 : nchunks + TAIL); c += 32) {
            const uint4 v = n1;
            if (!PACKED) {
                n1 = n2;
                if (c + 64 < nchunks) n2 = __ldg(src16 + c + 64);
            }
            uint32_t f = 0, dbits = 0;
            if (PACKED) {
                if (c < nchunks) {
                    f = __ldg(pk_F + (a0 >> 4) + c);
                    dbits = __ldg(pk_D + (a0 >> 4) + c);
                    if (dbits != 0xFFFFu) {
                        f &= spread16(dbits);  // undefined bases read as code 0 (jgi/BBDuk.java:3882)
                        atomicOr(badw + (c >> 5), 1u << (c & 31));
                    }
                }
            } else if (c < nchunks) {
                uint32_t cw[4], bw[4];
                classify4(v.x, cw[0], bw[0]);
                classify4(v.y, cw[1], bw[1]);
                classify4(v.z, cw[2], bw[2]);
                classify4(v.w, cw[3], bw[3]);
                f = (pack4(cw[0]) << 24) | (pack4(cw[1]) << 16) | (pack4(cw[2]) << 8) | pack4(cw[3]);
                dbits = 0xFFFFu;
                if ((bw[0] | bw[1] | bw[2] | bw[3]) != 0) {  // rare: some base of the chunk is not ACGTU
                    dbits = (valid4(bw[0]) << 12) | (valid4(bw[1]) << 8) | (valid4(bw[2]) << 4) | valid4(bw[3]);
                    f &= spread16(dbits);  // undefined bases read as code 0 (jgi/BBDuk.java:3882)
                    atomicOr(badw + (c >> 5), 1u << (c & 31));
                }
            }
            Fs[c + PAD] = f;
            Ds[c + PAD] = (uint16_t)dbits;
        }
        first64[lane] = ~0ull;
        lastpos[lane] = -1;
        __syncwarp();

        // ---- B. per-lane scan, C. pooled exact evaluation ----------------------------------------
        const int s = (int)(base_addr + o0 - a0);  // stream base of read position 0
        const int nsteps = (L + 15) >> 4;
        const int max_steps = PRE ? 0 : ((maxL + 15) >> 4);
        const int pairnum = (paired && (lane & 1)) ? 1 : 0;
        const bool skip = (p.skipR1 && pairnum == 0) || (p.skipR2 && pairnum == 1);
        const bool scan = live && L >= k && t.stored > 0 && !skip;
        bool has_undef = false;  // some chunk overlapping this read holds an undefined base
        if (!PRE && L > 0) {
            const int c0 = s >> 4, c1 = (s + L - 1) >> 4;
#pragma unroll 1
            for (int w = c0 >> 5; w <= (c1 >> 5); w++) {
                uint32_t m = badw[w];
                if (w == (c0 >> 5)) m &= 0xFFFFFFFFu << (c0 & 31);
                if (w == (c1 >> 5)) m &= 0xFFFFFFFFu >> (31 - (c1 & 31));
                has_undef |= (m != 0);
            }
        }
        const uint32_t undef_mask = __ballot_sync(0xFFFFFFFFu, has_undef);
        const bool any_undef = __any_sync(0xFFFFFFFFu, has_undef && scan);
        // Windows with an undefined base (forbidNs off): the forward k-mer reads it as A, the reverse one as
        // the complement of T (x = x2 = 0, jgi/BBDuk.java:3882-3888), so key = max(kmer, rkmer) can only be in
        // the table if the window passes the part test with its undefined bases read as A (what the scan over
        // F computes anyway) or with all of them read as T. Only the part_w-mers that CONTAIN an undefined base
        // differ between the two readings: they are looked up here, pooled over the tile (one lane per
        // w-mer end, one pass per chunk with an undefined base), and the rare hits are left in the lane's
        // column of cand[] -- bit b of cand[j][lane] = "the T reading of the w-mer ending at position 16j+b
        // passes" -- which step j picks up before it stores its own candidate bits there.
        DBG_CLK(t_a1);
        DBG_PHASE(8, t_a0, t_a1);
        const bool tvar = PARTS && !PRE && any_undef && !p.forbidNs;
        DBG_ADD(0, lane == 0);
        DBG_ADD(7, lane == 0 && any_undef);
        if (tvar) {
#pragma unroll 1
            for (int i = lane; i < max_steps * 16; i += 32) reinterpret_cast<uint32_t *>(cand)[i] = 0u;
            __syncwarp();
            const int w = t.part_w;
            const uint32_t wbits = (1u << w) - 1u;  // w <= 16
#pragma unroll 1
            for (int bwi = 0; bwi < geo.nbadw; bwi++) {
                uint32_t m = badw[bwi];
                while (m) {
                    const int c = bwi * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    const int e = 16 * c + lane;  // this lane's w-mer ends at stream base e
                    bool hit = false;
                    if (lane < 16 + w - 1) {
                        const uint32_t u = ~st.d16(e - 15) & wbits;  // bit t = base e-t is undefined
                        if (u) {
                            hit = bb_part_test(filt, st.f16(e - 15) | spread16(u), w);
                        }
                    }
                    uint32_t hb = __ballot_sync(0xFFFFFFFFu, hit);
                    while (hb) {  // rare: every lane checks whether the base lies in its own read
                        const int ee = 16 * c + __ffs(hb) - 1;
                        hb &= hb - 1;
                        if (live && ee >= s && ee < s + L) cand[((ee - s) >> 4) * 32 + lane] |= (uint16_t)(1u << ((ee - s) & 15));
                    }
                }
            }
            __syncwarp();
        }
        DBG_CLK(t_b0);
        DBG_PHASE(9, t_a1, t_b0);
        uint64_t thist = 0;  // tvar: the T-reading bits, laid out like mhist
        int qn = 0;  // warp-uniform queue fill
        // Every step's candidate bits are kept in cand[step][lane]. A read releases them to the warp queue
        // as it goes, but a read without undefined bases releases at most CAND_CAP per step: a window run
        // entering an adapter is a long streak of true hits of which only the first matters (ktrim=r,
        // kfilter). The surplus -- and everything after it, to keep position order -- is deferred; it is
        // only ever evaluated if all released candidates of the read turn out to be misses.
        bool done = !scan;      // the read's first hit is confirmed (or it is not scanned at all)
        bool deferred = false;  // candidates from (dj, dbits) onwards are held back
        int rel = 0;            // candidates released since the last full drain
        int dj = 0;
        uint32_t dbits = 0;
        const uint32_t lt_mask = (1u << lane) - 1u;

        auto drain = [&](int n_take) {
            // the last n_take (<= 32) queue entries, one per lane
            DBG_ADD(3, lane == 0);
            const uint32_t ent = (lane < n_take) ? queue[qn - n_take + lane] : 0u;
            const int owner = (int)(ent >> 11), pos = (int)(ent & 0x7FFu);
            const int s_owner = __shfl_sync(0xFFFFFFFFu, s, owner);
            if (lane < n_take) {
                const int id = exact_full(st, s_owner + pos, !((undef_mask >> owner) & 1u), p, t);
                if (id > 0) {
                    atomicMin(first64 + owner, ((unsigned long long)pos << 32) | (unsigned int)id);
                    if (FMODE == FM_KTRIM_L) atomicMax(lastpos + owner, pos);
                }
            }
            qn -= n_take;
            __syncwarp();
            if (first64[lane] != ~0ull) done = true;
        };

        uint32_t f_m2 = 0, f_m1 = 0, f_0 = 0, r_0 = 0, r_1 = 0, r_2 = 0;
        int last_und = -(1 << 20);  // position of the most recent undefined base before the current step
        uint64_t mhist = 0;     // PARTS: part-filter results, bit 32+b = position 16j+b, lower bits = older positions
        const int part_nd = t.part_w - BB_PART_WD;
        uint32_t raw_prev = 0;  // PARTS: the previous step's 9-mer bits (bit b = position 16(j-1)+b)
        if (!PARTS && scan && RCOMP) {
            r_0 = pair_reverse_complement(st.f16(s - (k - 1)));
            r_1 = pair_reverse_complement(st.f16(s + 16 - (k - 1)));
        }
        // One loop drives both the scan steps (j < max_steps) and, afterwards, the rare release of deferred
        // candidates, so that the evaluation code below exists once in the instruction stream.
        for (int j = 0;; j++) {
            if (j < max_steps) {
                uint32_t cbits = 0;
                bool stepping = scan && j < nsteps;
                if (FMODE != FM_KTRIM_L && done) stepping = false;  // first hit known: nothing further matters
                if (stepping) {
                    f_0 = st.f16(s + 16 * j);
                    if (PARTS) {
                        uint32_t mb = 0;  // bit 15-b = the 9-mer ending at position 16j+b is in the bitmap
    #pragma unroll
                        for (int b = 0; b < 16; b++) {
                            // the 32-bit window ending at 16j+b: bits 5..17 pick the word, bits 0..4 the bit (bbduk_dev.cuh)
                            const uint32_t v = __funnelshift_r(f_0, f_m1, 2 * (15 - b));
                            const uint32_t fw = *reinterpret_cast<const uint32_t *>(filt_bytes + ((v >> 3) & (4u * (BB_PART_WORDS - 1u))));
                            mb = __funnelshift_l(__funnelshift_r(fw, fw, v), mb, 1);
                        }
                        // a part of part_w bases passes iff all of its part_w-8 overlapping 9-mers do: AND of the
                        // bit with its part_nd (<= 7) predecessors, on (this step : previous step) in one register
                        const uint32_t raw = (__brev(mb) & 0xFFFF0000u) | raw_prev;  // bit 16+b = position 16j+b
                        uint32_t mw = raw;
#pragma unroll 1
                        for (int d = 1; d <= part_nd; d++) mw &= raw << d;
                        raw_prev = raw >> 16;
                        mhist |= (uint64_t)(mw >> 16) << 32;
                        cbits = (uint32_t)(mhist >> psh0);
                        if (t.n_parts > 1) cbits |= (uint32_t)(mhist >> psh1);
                        if (t.n_parts > 2) cbits |= (uint32_t)(mhist >> psh2) | (uint32_t)(mhist >> psh3);
                        cbits &= 0xFFFFu;
                        mhist >>= 16;
                        f_m1 = f_0;
                    } else {
                        if (RCOMP) r_2 = pair_reverse_complement(st.f16(s + 16 * (j + 2) - (k - 1)));
    #pragma unroll
                        for (int b = 0; b < 16; b++) {
                            const int sh = 2 * (15 - b);
                            uint32_t klo = __funnelshift_r(f_0, f_m1, sh);
                            uint32_t khi = __funnelshift_r(f_m1, f_m2, sh) & mask_hi;
                            if (!K16) klo &= mask_lo;
                            if (RCOMP) {
                                uint32_t rlo = __funnelshift_r(r_0, r_1, 2 * b);
                                const uint32_t rhi = __funnelshift_r(r_1, r_2, 2 * b) & mask_hi;
                                if (!K16) rlo &= mask_lo;
                                const bool gt = (((uint64_t)rhi << 32) | rlo) > (((uint64_t)khi << 32) | klo);
                                klo = gt ? rlo : klo;
                                khi = gt ? rhi : khi;
                            }
                            klo = (klo & mm_lo) | km_lo;
                            khi = (khi & mm_hi) | km_hi;
                            const uint32_t tt = bb_fhash(klo, khi);
                            const uint32_t pat = bb_filter_bits(tt);
                            const bool pass = (filt[bb_filter_word(tt, geo.nfw)] & pat) == pat;
                            cbits |= pass ? (1u << b) : 0u;
                        }
                        f_m2 = f_m1;
                        f_m1 = f_0;
                        r_0 = r_1;
                        r_1 = r_2;
                    }
                    if (any_undef) {
                        // every window that contains an undefined base is decided by the exact evaluator
                        uint32_t dd = st.d16(s + 16 * j);
                        const int rem = L - 16 * j;
                        if (rem < 16) dd |= (0xFFFFu >> rem);
                        const uint32_t und = (~dd) & 0xFFFFu;  // bit 15-b = position 16j+b undefined
                        // undefined bases inside this step cover the k positions from themselves on (towards lower
                        // bits); for k >= 16 that is everything at or below the highest undefined bit
                        uint32_t forced = (k >= 16) ? (und ? ((2u << (31 - __clz(und))) - 1u) : 0u) : smear_right(und, k);
                        const int carry = k - (16 * j - last_und);     // positions of this step still covered by an older one
                        if (carry > 0) forced |= (carry >= 16) ? 0xFFFFu : ((0xFFFFu << (16 - carry)) & 0xFFFFu);
                        uint32_t fb = __brev(forced) >> 16;  // -> bit b
                        DBG_ADD(4, __popc(fb));
                        if (tvar) {
                            // of the windows with an undefined base only those whose A or T reading passes the part test
                            thist |= (uint64_t)cand[j * 32 + lane] << 32;
                            uint32_t ext = (uint32_t)(thist >> psh0);
                            if (t.n_parts > 1) ext |= (uint32_t)(thist >> psh1);
                            if (t.n_parts > 2) ext |= (uint32_t)(thist >> psh2) | (uint32_t)(thist >> psh3);
                            thist >>= 16;
                            fb &= ext;
                        }
                        DBG_ADD(5, __popc(fb));
                        cbits |= fb;
                        if (und) last_und = 16 * j + 15 - (__ffs(und) - 1);
                    }
                    // keep positions k-1 <= i < L
                    const int i0 = 16 * j;
                    uint32_t vm = 0xFFFFu;
                    if (i0 < k - 1) vm &= (k - 1 - i0 >= 16) ? 0u : (0xFFFFu << (k - 1 - i0));
                    if (L - i0 < 16) vm &= (1u << (L - i0)) - 1u;
                    cbits &= vm;
                }
                if (scan && j < nsteps) cand[j * 32 + lane] = (uint16_t)cbits;
                uint32_t pb = (done || deferred) ? 0u : cbits;
                if (DEFER && !has_undef && rel + __popc(pb) > geo.cand_cap) {
                    // keep the lowest (CAND_CAP - rel) bits, hold everything else back
                    uint32_t rest = pb;
    #pragma unroll 1
                for (int c = rel; c < geo.cand_cap; c++) rest &= rest - 1;
                    pb ^= rest;
                    deferred = true;
                    dj = j;
                    dbits = rest;
                }
                if (DEFER) rel += __popc(pb);

                // pool this step's released candidates: exclusive prefix sum of the per-lane counts
                const int cnt = __popc(pb);
                DBG_ADD(1, lane == 0);
                DBG_ADD(2, cnt);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += v;
                }
                const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                if (total) {
                    int w = qn + incl - cnt;
                    const uint32_t tag = ((uint32_t)lane << 11) + 16u * (uint32_t)j;
                    while (pb) {
                        DBG_ADD(6, (__activemask() & ((1u << lane) - 1u)) == 0);
                        const int b = __ffs(pb) - 1;
                        pb &= pb - 1;
                        queue[w++] = (uint16_t)(tag + b);
                    }
                    qn += total;
                }
            } else {
                // rare: a read whose released candidates all missed still holds deferred ones; release them
                // in position order, 8 at a time, until a hit is confirmed or they are exhausted
                if (!__any_sync(0xFFFFFFFFu, deferred && !done)) break;
#pragma unroll 1
                for (int c = 0; c < 8; c++) {
                    bool has = false;
                    if (deferred && !done) {
                        while (dbits == 0 && dj < nsteps - 1) {
                            dj++;
                            dbits = cand[dj * 32 + lane];
                        }
                        has = dbits != 0;
                        if (!has) deferred = false;
                    }
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, has);
                    if (!bal) break;
                    if (has) {
                        const int b = __ffs(dbits) - 1;
                        dbits &= dbits - 1;
                        queue[qn + __popc(bal & lt_mask)] = (uint16_t)((lane << 11) + 16 * dj + b);
                    }
                    qn += __popc(bal);
                }
            }
            __syncwarp();
            if (qn >= 32 || (j >= max_steps - 1 && qn > 0)) {
                // evaluate everything released so far; afterwards nothing of any lane is unresolved
                DBG_CLK(t_r0);
                while (qn > 0) drain(min(32, qn));
                DBG_CLK(t_r1);
                DBG_PHASE(11, t_r0, t_r1);
                rel = 0;
            }
        }

        DBG_CLK(t_c);
        DBG_PHASE(10, t_b0, t_c);
        if (PRE) {
            first64[lane] = live ? pre_first[r] : ~0ull;
            lastpos[lane] = (live && pre_last) ? pre_last[r] : -1;
        }
        int found = 0, id0 = -1, minLoc = 999999999, maxLoc = -1, count = 0;
        int lo = 0, hi = L;
        bool discarded = false, ktrimmed = false;
        {
            const unsigned long long f64 = first64[lane];
            if (scan && f64 != ~0ull) {
                const int pos = (int)(f64 >> 32);
                id0 = (int)(unsigned int)f64;
                found = 1;
                minLoc = pos - k + 1;
                maxLoc = pos;
            }
        }
        if (FMODE == FM_KTRIM_L && PRE && found) maxLoc = max(maxLoc, lastpos[lane]);
        if (FMODE == FM_KTRIM_L && !PRE) {
            // ktrim=l also needs the LAST hit: walk the buffered candidates downwards from the read end
            // until one beyond the first hit is confirmed (same in-flight cap, highest position first)
            // hits at or below the forward phase's last confirmed hit are already accounted for
            const int firstpos = found ? max(maxLoc, lastpos[lane]) : -1;
            bool bdone = !found;
            int bj = nsteps;
            uint32_t bbits = 0;
            while (true) {
                for (int c = 0; c < 4; c++) {
                    bool has = false;
                    int pos = 0;
                    if (!bdone) {
                        while (bbits == 0 && bj > 0) {
                            bj--;
                            bbits = cand[bj * 32 + lane];
                        }
                        if (bbits != 0) {
                            const int b = 31 - __clz(bbits);
                            pos = 16 * bj + b;
                            bbits &= ~(1u << b);
                            if (pos > firstpos)
                                has = true;
                            else
                                bdone = true;  // reached the first hit: it is also the last one
                        } else {
                            bdone = true;
                        }
                    }
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, has);
                    if (!bal) break;
                    if (has) queue[qn + __popc(bal & lt_mask)] = (uint16_t)((lane << 11) + pos);
                    qn += __popc(bal);
                }
                __syncwarp();
                if (qn == 0) break;
                while (qn > 0) drain(min(32, qn));
                if (lastpos[lane] > firstpos) bdone = true;
            }
            if (found) maxLoc = max(firstpos, lastpos[lane]);
        }
        if (FMODE == FM_KFILTER) {  // countSetKmers with maxBadKmers==0 (jgi/BBDuk.java:3395-3457)
            count = found;
            discarded = found > 0;
        } else {
            // ktrim guard (jgi/BBDuk.java:3868): reads shorter than k still get the short-k-mer tails
            const bool tscan = !PRE && live && t.stored > 0 && p.useShortKmers && L >= max(1, min(k, p.mink)) && !found && !skip;
            int minLocX = 999999999, maxLocX = -1;
            if (found) {
                minLocX = minLoc + k;
                maxLocX = maxLoc - k;
            }
            if (tscan) {
                // All tail k-mers of a read are sub-windows of ONE 32-base window: the suffix tails
                // (ktrim=r, :3945-3975) of the window ending at the last base, the prefix tails
                // (ktrim=l, :3910-3942) of the window ending at base min(k,L)-1. Its reverse complement is
                // taken once; per length n the key costs a mask, a shift and a compare. The tails read
                // undefined bases as code 0 / complement 0 ("no N handling in tails").
                const int nmax = (FMODE == FM_KTRIM_R) ? min(k - 1, L) : min(k, L);
                const int e = (FMODE == FM_KTRIM_R) ? (s + L - 1) : (s + nmax - 1);
                uint64_t W = st.win(e);  // slot t = base e-t
                uint64_t RC;
                if (has_undef) {
                    const uint64_t E = spread2(st.dwin(e));
                    W &= E;
                    RC = bb_rcomp(W, 32) & rev2(E, 32);
                } else {
                    RC = bb_rcomp(W, 32);
                }
#pragma unroll 1
                for (int n = max(p.mink, 1); n <= nmax; n++) {
                    const uint64_t nm = (1ull << (2 * n)) - 1ull;
                    uint64_t kmer, rkmer;
                    int i;
                    if (FMODE == FM_KTRIM_R) {  // last n bases; reference loop index i = L-n
                        kmer = W & nm;
                        rkmer = RC >> (2 * (32 - n));
                        i = L - n;
                    } else {  // first n bases; reference loop index i = n-1
                        kmer = (W >> (2 * (nmax - n))) & nm;
                        rkmer = (RC >> (2 * (32 - nmax))) & nm;
                        i = n - 1;
                    }
                    const uint64_t key = bb_to_value(p, kmer, rkmer, 1ull << (2 * n));
                    // the short-key bloom only knows keys shorter than k; a prefix of length k goes straight to the table
                    if (n == k || filter_pass(filt_short, n_short, key)) {
                        const int id = bb_table_get(t, key);
                        if (id > 0) {
                            if (id0 < 0) id0 = id;
                            if (FMODE == FM_KTRIM_R) {
                                minLoc = i;
                                minLocX = min(minLocX, L);
                                maxLoc = L - 1;
                                maxLocX = max(maxLocX, i - 1);
                            } else {
                                minLoc = 0;
                                minLocX = min(minLocX, i + 1);
                                maxLoc = max(maxLoc, i);
                                maxLocX = max(maxLocX, 0);
                            }
                            found++;
                        }
                    }
                }
            }
            if (found) {  // :3981-4012
                if (p.trimPad != 0) {
                    maxLoc = mid3(0, maxLoc + p.trimPad, L);
                    minLoc = mid3(0, minLoc - p.trimPad, L);
                    maxLocX = mid3(0, maxLocX + p.trimPad, L);
                    minLocX = mid3(0, minLocX - p.trimPad, L);
                }
                if (FMODE == FM_KTRIM_L) {
                    const int leftLoc = p.ktrimExclusive ? maxLocX + 1 : maxLoc + 1;
                    count = trim_amounts(lo, hi, leftLoc, 0, 1);  // trimToPosition(r, leftLoc, L-1, 1)
                } else {
                    const int rightLoc = p.ktrimExclusive ? minLocX - 1 : minLoc - 1;
                    count = trim_amounts(lo, hi, 0, L - rightLoc - 1, 1);
                }
                ktrimmed = count > 0;
            }
        }
        DBG_CLK(t_d);
        DBG_PHASE(12, t_c, t_d);
        if (live && id0 > 0 && scaf_reads) {
            atomicAdd(scaf_reads + id0, 1ull);
            atomicAdd(scaf_bases + id0, (unsigned long long)L);
        }

        // ---- D. per-read minlen, pair logic (jgi/BBDuk.java:2750-2813, :2844-2871), outputs ------
        const bool active = live && t.stored > 0;  // doKmerTrimming / doKmerFiltering need stored k-mers
        const int minlenR = (int)fmaxf(__fmul_rn((float)L, p.minLenFraction), (float)p.minReadLength);
        const int len_pre = hi - lo;  // rlen1 / rlen2: captured before setDiscarded
        if (active && (FMODE == FM_KFILTER ? discarded : (len_pre < minlenR))) {
            // setDiscarded (jgi/BBDuk.java:3260-3266)
            if (p.trimFailuresTo1bp) {
                discarded = false;
                if (hi - lo > 1) trim_amounts(lo, hi, 0, hi - lo - 1, 1);
            } else {
                discarded = true;
            }
        }
        int len_cur = hi - lo;
        const bool disc_eff = discarded || (p.trimFailuresTo1bp && len_cur == 1);
        const bool disc_mate = __shfl_xor_sync(0xFFFFFFFFu, (int)disc_eff, 1) != 0;
        const int len_pre_mate = __shfl_xor_sync(0xFFFFFFFFu, len_pre, 1);
        const int len_cur_mate = __shfl_xor_sync(0xFFFFFFFFu, len_cur, 1);
        const int cnt_mate = __shfl_xor_sync(0xFFFFFFFFu, count, 1);
        const int L_mate = __shfl_xor_sync(0xFFFFFFFFu, L, 1);
        const bool remove = paired ? (p.removePairsIfEitherBad ? (disc_eff || disc_mate) : (disc_eff && disc_mate)) : disc_eff;
        bool tpe = false;
        int x_tpe = 0;
        if (FMODE == FM_KTRIM_R && paired && active && !remove && p.trimPairsEvenly && (count + cnt_mate) > 0 &&
            len_cur > len_cur_mate) {
            // the longer mate is cut to the shorter one's length: trimToPosition(longer, 0, shorterLen-1, 1)
            x_tpe = trim_amounts(lo, hi, 0, len_cur - len_cur_mate, 1);
            tpe = true;
            len_cur = hi - lo;
        }
        const int x_tpe_mate = __shfl_xor_sync(0xFFFFFFFFu, x_tpe, 1);
        const bool tpe_mate = __shfl_xor_sync(0xFFFFFFFFu, (int)tpe, 1) != 0;  // never inside a short-circuit
        const bool tpe_pair = paired && (tpe || tpe_mate);
        if (active && (!paired || !(lane & 1))) {  // one lane per unit accounts
            if (FMODE != FM_KFILTER) {
                int xsum = count + (paired ? cnt_mate : 0);
                int rkt = (count > 0) + ((paired && cnt_mate > 0) ? 1 : 0);
                if (remove) {
                    xsum += len_pre + (paired ? len_pre_mate : 0);
                    rkt = paired ? 2 : 1;
                } else if (tpe_pair) {
                    if (rkt < 2) rkt++;
                    xsum += x_tpe + x_tpe_mate;
                }
                s_bk += xsum;
                s_rk += rkt;
            } else if (remove) {
                s_rf += paired ? 2 : 1;
                s_bf += L + (paired ? L_mate : 0);
            }
        }
        if (live) {
            s_ri += 1;
            s_bi += L;
            if (!remove) {
                s_ro += 1;
                s_bo += len_cur;
            }
            if (out.id0) out.id0[r] = id0;
            if (out.id0b) out.id0b[r] = -1;
            if (out.lo) out.lo[r] = lo;
            if (out.hi) out.hi[r] = hi;
            if (out.count) out.count[r] = count;
            if (out.flags)
                out.flags[r] = (uint8_t)((discarded ? BBDUK_F_DISCARDED : 0) | (remove ? BBDUK_F_REMOVED : 0) |
                                         (ktrimmed ? BBDUK_F_KTRIMMED : 0) | (tpe ? BBDUK_F_TPE : 0));
        }
        DBG_CLK(t_e);
        DBG_PHASE(13, t_d, t_e);
    }
    if (stats) {
        // per-warp sums with one REDUX per counter (a launch holds < 2^32 bases, so 32 bits suffice per warp)
        const unsigned int v[8] = {(unsigned int)s_ri, (unsigned int)s_bi, (unsigned int)s_rk, (unsigned int)s_bk,
                                   (unsigned int)s_rf, (unsigned int)s_bf, (unsigned int)s_ro, (unsigned int)s_bo};
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const unsigned int x = __reduce_add_sync(0xFFFFFFFFu, v[q]);
            if (lane == 0 && x) atomicAdd((unsigned long long *)stats + q, (unsigned long long)x);
        }
    }
}

bool parts_ok(const BBParams &p, const BBTable &t) {
    return t.part_words > 0 && t.n_parts > 0 && p.editDistance == 0 && (!p.useShortKmers || t.short_words > 0);
}

bool canon_ok(const BBTable &t) {
    // beyond ~24 keys per filter word the image is saturated and every position would be a candidate
    return t.filter != nullptr && t.n_filter_words > 0 && t.stored <= (int64_t)t.n_filter_words * 24;
}

FastGeom make_geom(const BBParams &p, const BBTable &t, int max_read_len, bool pre = false) {
    FastGeom g;
    const int lmax = std::max(max_read_len, 16);
    g.nch = pre ? 0 : (32 * lmax + 15 + 15) / 16 + 1;
    g.nbadw = (g.nch + 31) / 32 + 1;
    g.csteps = pre ? 0 : (lmax + 15) / 16 + 1;
    int wb = 32 * 8 + 32 * 4 + g.nbadw * 4 + (g.nch + PAD + TAIL) * 4 + ((g.nch + PAD + TAIL + 1) & ~1) * 2 + QCAP * 2 +
             g.csteps * 32 * 2;
    wb = (wb + 15) & ~15;
    g.warp_bytes = wb;
    if (pre) {
        g.nfw = g.nsw = g.src_off = 0;
    } else if (parts_ok(p, t)) {
        g.nfw = t.part_words;
        g.nsw = t.short_words;
        g.src_off = t.n_filter_words;
    } else {
        g.nfw = t.n_filter_words;
        g.nsw = 0;
        g.src_off = 0;
    }
    const int avail = FAST_SMEM_LIMIT - (int)(g.nfw + g.nsw) * 4 - 64;
    g.warps = std::min(32, avail / wb);
    g.cand_cap = CAND_CAP;
    if (const char *e = getenv("BBDUK_B200_CAND_CAP")) g.cand_cap = std::max(1, atoi(e));
    return g;
}

}  // namespace

#if defined(BB_FAST_COUNT) || defined(BB_FAST_CLOCK)
extern "C" __attribute__((visibility("default"))) int bbduk_b200_debug_fast_counters(unsigned long long *out16, int reset) {
    if (cudaMemcpyFromSymbol(out16, bb_fast_dbg, sizeof(unsigned long long) * 16) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long z[16] = {0};
        if (cudaMemcpyToSymbol(bb_fast_dbg, z, sizeof z) != cudaSuccess) return 1;
    }
    return 0;
}
#endif

// the packed stage A exists for the part-filter kernels and the canonical rcomp/k>=16 kernels
bool packed_ok(const BBParams &p, const BBTable &t) { return parts_ok(p, t) || (p.rcomp && p.k >= 16); }

FastPlan plan_fast(const BBParams &p, const BBTable &t, int max_read_len) {
    FastPlan pl{false, max_read_len, 0, 0};
    const bool mode_ok = (p.mode == MODE_KTRIM) || (p.mode == MODE_KFILTER && p.maxBadKmers0 == 0);
    if (!mode_ok) return pl;
    if (p.qHammingDistance != 0 || (p.useShortKmers && p.qHammingDistance2 != 0)) return pl;
    if (p.speed != 0 || p.qSkip != 1 || p.restrictLeft != 0 || p.restrictRight != 0) return pl;
    if (p.kbig > p.k || p.minKmerFraction != 0.0f) return pl;
    if (p.k < 2) return pl;
    if (max_read_len > MAX_FAST_LEN) max_read_len = MAX_FAST_LEN;  // longer reads are handed off per tile
    if (!parts_ok(p, t) && !canon_ok(t)) return pl;  // HBM-resident tables: handled by the generic kernel for now
    const FastGeom g = make_geom(p, t, max_read_len);
    if (g.warps < 8) return pl;
    pl.usable = true;
    pl.max_read_len = max_read_len;
    pl.smem_bytes = (int)(g.nfw + g.nsw) * 4 + g.warps * g.warp_bytes + 64;
    pl.filter_words = (int)(g.nfw + g.nsw);
    return pl;
}

int launch_fast(const FastPlan &plan, const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_reads, int paired,
                const BBParams &p, const BBTable &t, const bbduk_out &out, bbduk_stats *d_stats,
                unsigned long long *scaf_reads, unsigned long long *scaf_bases, int32_t *d_handoff,
                unsigned int *d_handoff_n, int sm_count, cudaStream_t st, const uint32_t *pk_F, const uint16_t *pk_D) {
    const FastGeom g = make_geom(p, t, plan.max_read_len);
    const int threads = g.warps * 32;
    const int64_t n_tiles = (n_reads + 31) / 32;
    const int blocks = (int)std::min<int64_t>(sm_count, (n_tiles + g.warps - 1) / g.warps);
    auto go = [&](auto kern) -> int {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes) != cudaSuccess) return -1;
        kern<<<blocks, threads, plan.smem_bytes, st>>>(d_bases, d_offsets, n_reads, paired, p, t, out, d_stats, scaf_reads,
                                                       scaf_bases, d_handoff, d_handoff_n, g, nullptr, nullptr, pk_F, pk_D);
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    };
    const bool k16 = p.k >= 16;
    const bool parts = parts_ok(p, t);
#define BB_DISPATCH(FM)                                                                                               \
    (parts ? go(bbduk_fast_kernel<FM, true, true, true>)                                                              \
           : p.rcomp ? (k16 ? go(bbduk_fast_kernel<FM, true, true, false>) : go(bbduk_fast_kernel<FM, true, false, false>))    \
                     : (k16 ? go(bbduk_fast_kernel<FM, false, true, false>) : go(bbduk_fast_kernel<FM, false, false, false>)))
    if (pk_F) {  // host-packed batch (only offered when packed_ok())
        if (!pk_D || !packed_ok(p, t)) return -1;
#define BB_PK(FM) (parts ? go(bbduk_fast_kernel<FM, true, true, true, false, true>) : go(bbduk_fast_kernel<FM, true, true, false, false, true>))
        if (p.mode == MODE_KFILTER) return BB_PK(FM_KFILTER);
        if (p.ktrimLeft) return BB_PK(FM_KTRIM_L);
        return BB_PK(FM_KTRIM_R);
#undef BB_PK
    }
    if (p.mode == MODE_KFILTER) return BB_DISPATCH(FM_KFILTER);
    if (p.ktrimLeft) return BB_DISPATCH(FM_KTRIM_L);
    return BB_DISPATCH(FM_KTRIM_R);
#undef BB_DISPATCH
}

// stage D only: ktrim arithmetic, minlen, pair logic, counters and outputs from precomputed first/last hits
int launch_epilogue(const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_reads, int paired, const BBParams &p,
                    const BBTable &t, const bbduk_out &out, bbduk_stats *d_stats, unsigned long long *scaf_reads,
                    unsigned long long *scaf_bases, const unsigned long long *d_first64, const int *d_lastpos, int sm_count,
                    cudaStream_t st) {
    const FastGeom g = make_geom(p, t, 16, true);
    const int threads = 1024;
    const int smem = g.warps * g.warp_bytes + 64;
    const int64_t n_tiles = (n_reads + 31) / 32;
    const int blocks = (int)std::min<int64_t>(sm_count, (n_tiles + g.warps - 1) / g.warps);
    auto go = [&](auto kern) -> int {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
        kern<<<blocks, threads, smem, st>>>(d_bases, d_offsets, n_reads, paired, p, t, out, d_stats, scaf_reads, scaf_bases,
                                            nullptr, nullptr, g, d_first64, d_lastpos, nullptr, nullptr);
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    };
    if (p.mode == MODE_KFILTER) return go(bbduk_fast_kernel<FM_KFILTER, true, true, true, true>);
    if (p.ktrimLeft) return go(bbduk_fast_kernel<FM_KTRIM_L, true, true, true, true>);
    return go(bbduk_fast_kernel<FM_KTRIM_R, true, true, true, true>);
}
