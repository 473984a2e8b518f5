// probe_fast2.cu -- the sampled-pigeonhole per-read kernel (round 2) for ktrim=r, ktrim=l and kfilter with a
// substitution-only neighbourhood whose pigeonhole parts are at least 11 bases wide (k=23 hdist<=1, k=31 hdist<=1, ...).
//
// Same contract as probe_fast.cu (which it replaces where it applies): it only decides WHICH read positions can hit and
// then evaluates those with the exact key formula against the exact hash array, so results are identical by construction.
// What changed is how few instructions the deciding takes (probe_fast.cu tests a 9-mer at EVERY read position, 6
// instructions each, and its scan loop carries the queue bookkeeping):
//
//   A.  stage the tile's ASCII as a 2-bit stream F (+ defined bits D) in shared memory (as probe_fast.cu);
//   B1. sampled scan: a window can only hit if one of its hdist+1 parts (w >= 11 bases) equals a reference part; such a
//       part contains a whole 8-mer that starts on a 4-base boundary of the STREAM, whatever the read's phase. So the
//       lane looks up only the byte-aligned 16-bit fields of F -- one PRMT, one byte load from a 64 KB map of all 8-mers
//       that occur inside a reference part, one multiply-add into the hit mask; 4 lookups per 16 bases, no queue work;
//   B2. the sample hits of all 32 reads (on random sequence 1.5 per read for adapters.fa) are pooled and confirmed 32 at
//       a time against the 9-mer bitmap: the 2w-16 9-mers around a hit give the parts that pass entirely, whose end
//       positions are OR-ed into the owner's seed bits S (shared memory, one column per lane);
//   C.  candidate windows = seed bits shifted by the part lags (+ every window with an undefined base); they are released
//       in position order, a few per read and round so that a round never exceeds 32, evaluated exactly by the pooled
//       evaluator, and a read stops at its first confirmed hit (ktrim=r, kfilter; ktrim=l then walks down from the end);
//   T.  short-k-mer tails: two direct bitmaps over q-mers (table.cu: tail_filter_build_kernel) say which tail lengths can
//       hit at all -- on random reads 2 % of them -- and only those are looked up;
//   D.  TrimRead arithmetic, minlen, pair logic, outputs (as probe_fast.cu).
// Rolling-state semantics: jgi/BBDuk.java:3882-3900 in the closed form of SURVEY.md A.2; ktrim :3866-4013; kfilter :3395-3457.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "bbduk_dev.cuh"
#include "fast_common.cuh"
#include "probe.h"

namespace {

using namespace bbfast;

constexpr int QCAP2 = 256;     // pooled queue entries (16 bit each): sample hits / tail lengths of one pass (512 B per warp: 32 warps fit)
constexpr int ITEM_CAP = 8;    // items one read contributes per pass (32 x 8 = QCAP2)
constexpr int L1_WORDS = 12;   // stream words (16 bases each) one unrolled block of the sampled scan covers
constexpr int SPELL_WORDS = 256;  // the "what do these four codes spell" table of stage A

struct Fast2Geom {
    int warps;        // warps per block
    int nch;          // staged 16-base chunks per warp (capacity, without padding)
    int warp_bytes;   // shared memory per warp
    int nbadw;        // words of the per-tile "chunk has an undefined base" bit mask
    int sw;           // 32-position words of the per-lane seed bit columns
    uint32_t part_off, samp_off, tail_off;  // word offsets of the part bitmap / 8-mer byte map / tail bitmaps in BBTable::filter
    uint32_t tail0_off;  // word offset of this mode's 8-mer level-0 tail bitmap in BBTable::filter; 0 = none
    int mw;              // kmask: 32-position words of the per-lane mask bit columns (= sw), else 0
};

// OR of x << d for d in [0, n), n <= 32
__device__ __forceinline__ uint64_t smear_left64(uint64_t x, int n) {
    int have = 1;
    while (have < n) {
        const int s = min(have, n - have);
        x |= x << s;
        have += s;
    }
    return x;
}

// sets bits a..b (inclusive, a <= b, at most two words apart) of a lane's bit column (word w of lane l at col[w * 32])
__device__ __forceinline__ void or_range(uint32_t *col, int a, int b) {
    const int wa = a >> 5, wb = b >> 5;
    const uint32_t ma = 0xFFFFFFFFu << (a & 31), mb = 0xFFFFFFFFu >> (31 - (b & 31));
    if (wa == wb) {
        atomicOr(col + wa * 32, ma & mb);
    } else {
        atomicOr(col + wa * 32, ma);
        for (int w = wa + 1; w < wb; w++) atomicOr(col + w * 32, 0xFFFFFFFFu);
        atomicOr(col + wb * 32, mb);
    }
}

// Out-of-line helpers: the per-tile loop has to fit the 32 KB instruction cache with 8 warps per scheduler in different
// phases of it, so everything that runs a few times per tile at most is one shared copy behind a call.
__device__ __noinline__ int warp_scan_incl(int v, int lane) {  // inclusive prefix sum over the warp
#pragma unroll 1
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}
#if defined(BB_FAST_COUNT)
__device__ unsigned long long bb_fast2_dbg[16];
#define DBG2(i, v)                                                  \
    do {                                                            \
        const unsigned long long v_ = (unsigned long long)(v);      \
        if (v_) atomicAdd(&bb_fast2_dbg[i], v_);                    \
    } while (0)
#else
#define DBG2(i, v)
#endif

// PW11: the parts are 11 bases wide (k=23 hdist=1, k=23 mm=t hdist=0): the confirmation loops are unrolled for it.
// FN = forbidNs: every window with an undefined base is forced to the exact evaluator; otherwise the all-T pre-pass runs instead --
// each build carries only its own path (the kernel is bound by instruction fetch)
template <int FMODE, bool PACKED, bool PW11, bool FN>
__global__ void __launch_bounds__(1024, 1)
bbduk_fast2_kernel(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ offsets, int64_t n_reads, int paired, BBParams p,
                   BBTable t, bbduk_out out, bbduk_stats *stats, unsigned long long *scaf_reads, unsigned long long *scaf_bases,
                   int32_t *handoff, unsigned int *handoff_n, Fast2Geom geo, const uint32_t *__restrict__ pk_F,
                   const uint16_t *__restrict__ pk_D) {
    extern __shared__ __align__(16) uint32_t smem[];
    // [256] the four upper-case bases a packed byte of F stands for (base 0 = bits 7:6 = lowest address): stage A proves a
    // word of ASCII valid by comparing it with what its own codes spell
    uint32_t *spell = smem;
    const uint8_t *samp = reinterpret_cast<const uint8_t *>(smem + SPELL_WORDS);  // [65536] 8-mer byte map
    const uint32_t *filt = smem + SPELL_WORDS + 16384;                              // [BB_PART_WORDS] 9-mer bitmap
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // [BB_TAIL0_WORDS] 8-mer level-0 image of the tail bitmap that is asked once per tail length (none: the tails go to L2 directly)
    const uint32_t *tail0 = smem + SPELL_WORDS + 16384 + BB_PART_WORDS;
    const uint32_t tail0_words = geo.tail0_off ? BB_TAIL0_WORDS : 0u;
    uint8_t *wbase = reinterpret_cast<uint8_t *>(smem + SPELL_WORDS + 16384 + BB_PART_WORDS + tail0_words) + (size_t)warp * geo.warp_bytes;
    uint32_t *first32 = reinterpret_cast<uint32_t *>(wbase);                      // [32] (pos << 22 | id) of the first hit, ~0 = none
    uint32_t *owners = first32 + 32;                                              // [32] evaluation rounds: lane of the rank-th releasing read
    int *lastpos = reinterpret_cast<int *>(owners + 32);                          // [32] last hit position
    uint32_t *misc = reinterpret_cast<uint32_t *>(lastpos + 32);                  // [4] misc[0] != 0: the tile has a seed bit
    uint32_t *badw = misc + 4;                                                    // [nbadw] chunks with a non-ACGTU base
    uint32_t *Fs = badw + geo.nbadw;
    uint16_t *Ds = reinterpret_cast<uint16_t *>(Fs + geo.nch + PAD + TAIL);
    uint16_t *queue = Ds + ((geo.nch + PAD + TAIL + 1) & ~1);                     // [QCAP2] (owner lane << 11) | position
    uint32_t *S = reinterpret_cast<uint32_t *>(queue + QCAP2);                    // [sw][32] seed bits, one column per lane
    uint32_t *M = S + geo.sw * 32;                                                // [mw][32] kmask: covered bases, one column per lane

    {
        const uint32_t *src = t.filter + geo.samp_off;
        for (uint32_t i = threadIdx.x; i < 16384u; i += blockDim.x) smem[SPELL_WORDS + i] = __ldg(src + i);
        src = t.filter + geo.part_off;
        for (uint32_t i = threadIdx.x; i < BB_PART_WORDS; i += blockDim.x) smem[SPELL_WORDS + 16384 + i] = __ldg(src + i);
        src = t.filter + geo.tail0_off;
        for (uint32_t i = threadIdx.x; i < tail0_words; i += blockDim.x) smem[SPELL_WORDS + 16384 + BB_PART_WORDS + i] = __ldg(src + i);
        for (uint32_t i = threadIdx.x; i < (uint32_t)SPELL_WORDS; i += blockDim.x) {
            uint32_t e = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) e |= ((0x54474341u >> (8 * ((i >> (6 - 2 * j)) & 3u))) & 0xFFu) << (8 * j);  // "ACGT"
            spell[i] = e;
        }
    }
    for (int i = lane; i < PAD; i += 32) {
        Fs[i] = 0;
        Ds[i] = 0;
    }
    __syncthreads();

    const Stream st{Fs, Ds};
    const int k = p.k;
    const int pw = PW11 ? 11 : t.part_w;          // >= 11
    const int n9 = 2 * pw - 16;                   // 9-mers around a sample hit that decide the parts containing it
    const uint32_t amask = (1u << (pw - 7)) - 1u; // part ends a sample hit can discover
    const int lag0 = t.part_lag[0], lag1 = t.part_lag[1], lag2 = t.part_lag[2], lag3 = t.part_lag[t.n_parts > 3 ? 3 : 2];
    const uint32_t *tailb1 = t.filter + geo.tail_off;
    const uint32_t *tailb2 = tailb1 + max(1u, (1u << (2 * t.tail_q)) >> 5);
    const uintptr_t base_addr = PACKED ? (uintptr_t)0 : reinterpret_cast<uintptr_t>(bases);
    const int64_t n_tiles = (n_reads + 31) >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    const uint32_t lt_mask = (1u << lane) - 1u;
    long long s_rk = 0, s_bk = 0, s_rf = 0, s_bf = 0, s_ro = 0, s_bo = 0, s_ri = 0, s_bi = 0;

    for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; tile < n_tiles; tile += warps_total) {
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
        if (nchunks > geo.nch || maxL > MAX_FAST_LEN || ((maxL + 16) >> 5) + 2 > geo.sw) {
            // tile does not fit the staging: hand its units to the generic kernel
            if (live && (!paired || !(lane & 1))) {
                const unsigned int w = atomicAdd(handoff_n, 1u);
                handoff[w] = (int32_t)(paired ? (r >> 1) : r);
            }
            continue;
        }
        // this warp's NEXT tile on its way from DRAM to L2 while this one is worked on (its ASCII is as long as this tile's to a
        // first approximation: a prefetch is only a hint, a wrong guess costs nothing but the line)
        if (!PACKED) {
            const int64_t rn0 = (tile + warps_total) * 32;
            if (rn0 < n_reads) {
                const uintptr_t an = (base_addr + __ldg(offsets + rn0)) & ~(uintptr_t)127;
                const int nl = (int)((tile_hi - tile_lo + 255) >> 7);
                for (int i = lane; i < nl; i += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(an + (uintptr_t)i * 128));
            }
        }
        // ---- A. stage + convert ---------------------------------------------------------------
#pragma unroll 1
        for (int i = lane; i < geo.nbadw; i += 32) badw[i] = 0;
#pragma unroll 1
        for (int i = lane; i < (geo.sw + geo.mw) * 32; i += 32) S[i] = 0;  // S and, behind it, M
        __syncwarp();
        const uint4 *src16 = reinterpret_cast<const uint4 *>(a0);
        uint4 n1 = make_uint4(0, 0, 0, 0), n2 = make_uint4(0, 0, 0, 0);
        if (!PACKED) {  // the ASCII loads run two trips ahead of their use
            if (lane < nchunks) n1 = __ldg(src16 + lane);
            if (lane + 32 < nchunks) n2 = __ldg(src16 + lane + 32);
        }
#pragma unroll 1
        for (int c = lane; c < nchunks + TAIL; c += 32) {
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
                // codes = ((w >> 1) ^ (w >> 2)) & 3 per byte; one multiply packs four of them into the product's top byte
                // (bits 23:22 of the product are always zero, so product >> 22 is that byte times 4: the table offset).
                // A word is A C G T only (either case) iff it equals what its codes spell; anything else (N, IUPAC, U) takes
                // the exact classification below.
                const uint32_t aw[4] = {v.x, v.y, v.z, v.w};
                uint32_t pr[4], odd = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    pr[j] = (((aw[j] >> 1) ^ (aw[j] >> 2)) & 0x03030303u) * 0x40100401u;
                    const uint32_t e = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(spell) + (pr[j] >> 22));
                    odd |= (aw[j] & 0xDFDFDFDFu) ^ e;
                }
                f = __byte_perm(__byte_perm(pr[2], pr[3], 0x0037), __byte_perm(pr[0], pr[1], 0x0037), 0x5410);
                dbits = 0xFFFFu;
                if (odd != 0) {  // rare: some base of the chunk is not ACGT
                    uint32_t cw, bw[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) classify4(aw[j], cw, bw[j]);
                    if ((bw[0] | bw[1] | bw[2] | bw[3]) != 0) {  // and not U either
                        dbits = (valid4(bw[0]) << 12) | (valid4(bw[1]) << 8) | (valid4(bw[2]) << 4) | valid4(bw[3]);
                        f &= spread16(dbits);  // undefined bases read as code 0 (jgi/BBDuk.java:3882)
                        atomicOr(badw + (c >> 5), 1u << (c & 31));
                    }
                }
            }
            Fs[c + PAD] = f;
            Ds[c + PAD] = (uint16_t)dbits;
        }
        first32[lane] = ~0u;
        lastpos[lane] = -1;
        if (lane == 0) misc[0] = 0u;
        __syncwarp();

        const int s = live ? (int)(base_addr + o0 - a0) : 0;  // stream base of read position 0
        const int pairnum = (paired && (lane & 1)) ? 1 : 0;
        const bool skip = (p.skipR1 && pairnum == 0) || (p.skipR2 && pairnum == 1);
        const bool scan = live && L >= k && t.stored > 0 && !skip;
        bool has_undef = false;  // some chunk overlapping this read holds an undefined base
        if (L > 0) {
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
        // Windows with an undefined base. forbidNs: all of them go to the exact evaluator (force_und). Otherwise the forward
        // k-mer reads the base as A and the reverse k-mer as the complement of T (x = x2 = 0, jgi/BBDuk.java:3882-3888), so
        // key = max(kmer, rkmer) can only be in the table if the window passes the part test with its undefined bases read
        // as A -- what the scan over F finds anyway -- or with all of them read as T. Only parts that CONTAIN an undefined
        // base differ between the readings: for every undefined base (up to 4 per read, else force_und) one pooled item
        // looks up the 2pw-9 9-mers around it in the T reading and adds the passing part ends to the seed bits.
        const bool force_und = FN && has_undef && scan;
        if (!FN && __any_sync(0xFFFFFFFFu, has_undef && scan)) {
            uint16_t *ntmp = queue + 128;  // [4][32] this lane's undefined positions of this pass (the pass queues at most 128 items in front of it)
            int skip_n = 0;                // positions earlier passes have dealt with (a pass takes four per read)
            bool more;
          do {
            int n_und = 0, seen = 0;
            if (has_undef && scan) {
                const int c0 = s >> 4, c1 = (s + L - 1) >> 4;
#pragma unroll 1
                for (int w = c0 >> 5; w <= (c1 >> 5); w++) {
                    uint32_t m = badw[w];
                    if (w == (c0 >> 5)) m &= 0xFFFFFFFFu << (c0 & 31);
                    if (w == (c1 >> 5)) m &= 0xFFFFFFFFu >> (31 - (c1 & 31));
                    while (m) {
                        const int c = 32 * w + __ffs(m) - 1;
                        m &= m - 1;
                        uint32_t ub = __brev(~(uint32_t)Ds[c + PAD]) >> 16;  // bit b = stream base 16c+b is undefined
                        const int first = s - 16 * c, end = s + L - 16 * c;  // the read covers [first, end) of this chunk
                        if (first > 0) ub &= 0xFFFFu << first;
                        if (end < 16) ub &= (1u << end) - 1u;
                        while (ub) {
                            const int b = __ffs(ub) - 1;
                            ub &= ub - 1;
                            if (seen >= skip_n && n_und < 4) ntmp[n_und++ * 32 + lane] = (uint16_t)(16 * c + b - s);
                            seen++;
                        }
                    }
                }
            }
            more = seen > skip_n + 4;
            skip_n += 4;
            const int incl = warp_scan_incl(n_und, lane);
            const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);  // <= 128
            DBG2(7, lane == 0 ? total : 0);
            DBG2(8, lane == 0);
            __syncwarp();
            for (int c = 0; c < n_und; c++) queue[incl - n_und + c] = (uint16_t)(((uint32_t)lane << 11) + ntmp[c * 32 + lane]);
            __syncwarp();
            const int nT = 2 * pw - 9;
#pragma unroll 1
            for (int base = 0; base < total; base += 32) {
                const bool on = base + lane < total;
                const uint32_t ent = on ? queue[base + lane] : 0u;
                const int owner = (int)(ent >> 11), u = (int)(ent & 0x7FFu);
                const int s_owner = __shfl_sync(0xFFFFFFFFu, s, owner);
                if (on) {
                    const int e_hi = s_owner + u + pw - 1;
                    const uint64_t W = st.win(e_hi) | ~spread2(st.dwin(e_hi));  // undefined bases read as T
                    uint32_t B = 0;  // bit i = the 9-mer ending at read position u - (pw-9) + i passes in the T reading
#pragma unroll 1
                    for (int d = 0; d < nT; d++) {
                        const uint32_t x = (uint32_t)(W >> (2 * d));
                        const uint32_t fw = filt[bb_part_word(x)];
                        B = __funnelshift_l(__funnelshift_r(fw, fw, x), B, 1);
                    }
                    uint32_t A = B;  // bit v = the part ending at read position u + v passes
#pragma unroll 1
                    for (int c = 1; c < pw - 8; c++) A &= B >> c;
                    A &= (1u << pw) - 1u;
                    if (A) {
                        misc[0] = 1u;
                        const int sh = u & 31;
                        uint32_t *col = S + (u >> 5) * 32 + owner;
                        atomicOr(col, A << sh);
                        if (sh && (A >> (32 - sh)) && (u >> 5) + 1 < geo.sw) atomicOr(col + 32, A >> (32 - sh));
                    }
                }
            }
            __syncwarp();
          } while (__any_sync(0xFFFFFFFFu, more));
        }
        const bool any_force = FN && __any_sync(0xFFFFFFFFu, force_und);

        // ---- B1. sampled scan --------------------------------------------------------------------
        // sample g of the read = the 8-mer that starts at stream byte 4*w0 - 1 + g (w0 = first stream word of the read);
        // it counts if it lies inside the read: g in [glo, ghi]
        const int w0 = s >> 4;
        const int nw = scan ? ((s + L - 1) >> 4) - w0 + 1 : 0;  // stream words the read touches
        const int nwmax = __reduce_max_sync(0xFFFFFFFFu, nw);
        const int glo = ((s + 3) >> 2) - 4 * w0 + 1, ghi = ((s + L - 8) >> 2) - 4 * w0 + 1;  // L >= k >= 11 here
#pragma unroll 1
        for (int blk = 0; blk * L1_WORDS < nwmax; blk++) {
            uint32_t acc_lo = 0, acc_hi = 0;  // sample bits 48*blk + [0, 32) and + [32, 48)
            const int wb = w0 + blk * L1_WORDS;
            const int left = nw - blk * L1_WORDS;  // this lane's words in the block
            uint32_t fprev = Fs[PAD + wb - 1];
#pragma unroll 1
            for (int i0 = 0; i0 < L1_WORDS; i0 += 2) {  // partly unrolled: the fully unrolled loop is 6 KB of a 32 KB instruction cache
                uint32_t acc = 0;
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int i = i0 + j;
                    if (blk * L1_WORDS + i < nwmax) {  // warp-uniform
                        const bool mine = i < left;
                        const uint32_t f = mine ? Fs[PAD + wb + i] : 0u;
                        const uint32_t x0 = __byte_perm(__funnelshift_l(f, fprev, 8), 0u, 0x4410);  // (last byte of the previous word, first of this)
                        const uint32_t x1 = __byte_perm(f, 0u, 0x4432);
                        const uint32_t x2 = __byte_perm(f, 0u, 0x4421);
                        const uint32_t x3 = __byte_perm(f, 0u, 0x4410);
                        uint32_t h = (uint32_t)samp[x0] + 2u * (uint32_t)samp[x1] + 4u * (uint32_t)samp[x2] + 8u * (uint32_t)samp[x3];
                        if (!mine) h = 0;
                        acc += h << (4 * j);
                        fprev = f;
                    }
                }
                if (i0 < 8)
                    acc_lo += acc << (4 * i0);
                else
                    acc_hi += acc << (4 * (i0 - 8));
            }
            // keep the samples inside the read
            {
                const int g0 = 48 * blk;
                const int lo_ = max(glo - g0, 0), hi_ = min(ghi - g0, 47);
                uint64_t vm = 0;
                if (hi_ >= lo_) vm = ((hi_ >= 63 ? ~0ull : ((1ull << (hi_ + 1)) - 1ull)) >> lo_) << lo_;
                acc_lo &= (uint32_t)vm;
                acc_hi &= (uint32_t)(vm >> 32);
            }
            DBG2(1, __popc(acc_lo) + __popc(acc_hi));
            // ---- B2. pooled confirmation of the sample hits against the 9-mer bitmap ----------------
            while (__any_sync(0xFFFFFFFFu, (acc_lo | acc_hi) != 0u)) {
                const int have = __popc(acc_lo) + __popc(acc_hi);
                const int cnt = min(have, ITEM_CAP);
                const int incl = warp_scan_incl(cnt, lane);
                const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                {
                    int w = incl - cnt;
                    // rel = stream position of the sample start minus s = 4 * (4*w0 - 1 + g) - s
                    const int rel0 = 4 * (4 * w0 - 1 + 48 * blk) - s;
#pragma unroll 1
                    for (int c = 0; c < cnt; c++) {
                        int g;
                        if (acc_lo) {
                            g = __ffs(acc_lo) - 1;
                            acc_lo &= acc_lo - 1;
                        } else {
                            g = 32 + __ffs(acc_hi) - 1;
                            acc_hi &= acc_hi - 1;
                        }
                        queue[w++] = (uint16_t)(((uint32_t)lane << 11) + (uint32_t)(rel0 + 4 * g));
                    }
                }
                __syncwarp();
#pragma unroll 1
                for (int base = 0; base < total; base += 32) {
                    DBG2(2, lane == 0);
                    const bool on = base + lane < total;
                    const uint32_t ent = on ? queue[base + lane] : 0u;
                    const int owner = (int)(ent >> 11), rel = (int)(ent & 0x7FFu);
                    const int s_owner = __shfl_sync(0xFFFFFFFFu, s, owner);
                    if (on) {
                        // 9-mers ending at stream positions e_hi - d, d = 0 .. n9-1, e_hi = sample start + pw - 1; bit i of B =
                        // the 9-mer ending at (sample start + 16 - pw + i) is in the bitmap (rotating the filter word right by
                        // the 9-mer leaves its bit in bit 31, one funnel shift pushes it into B: bbduk_dev.cuh)
                        uint32_t B = 0;
                        if (PW11) {
                            const uint32_t w32 = st.f16(s_owner + rel + pw - 1 - 15);  // n9 = 6: 2*5 + 18 bits of the newest 16 bases
#pragma unroll
                            for (int d = 0; d < 6; d++) {
                                const uint32_t x = w32 >> (2 * d);
                                const uint32_t fw = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(filt) + ((x >> 3) & (4u * (BB_PART_WORDS - 1u))));
                                B = __funnelshift_l(__funnelshift_r(fw, fw, x), B, 1);
                            }
                        } else {
                            const uint64_t W = st.win(s_owner + rel + pw - 1);
#pragma unroll 1
                            for (int d = 0; d < n9; d++) {
                                const uint32_t x = (uint32_t)(W >> (2 * d));
                                const uint32_t fw = filt[bb_part_word(x)];
                                B = __funnelshift_l(__funnelshift_r(fw, fw, x), B, 1);
                            }
                        }
                        uint32_t A = B;  // bit u = the part ending at (sample start + 7 + u) passes: its pw-8 9-mers are all set
                        if (PW11) {
                            A = B & (B >> 1) & (B >> 2);
                        } else {
#pragma unroll 1
                            for (int c = 1; c < pw - 8; c++) A &= B >> c;
                        }
                        A &= amask;
                        if (A) {
                            misc[0] = 1u;
                            const int e0 = rel + 7;  // read-relative end of the u = 0 part
                            const int sh = e0 & 31;
                            uint32_t *col = S + (e0 >> 5) * 32 + owner;
                            atomicOr(col, A << sh);
                            if (sh && (A >> (32 - sh)) && (e0 >> 5) + 1 < geo.sw) atomicOr(col + 32, A >> (32 - sh));
                        }
                    }
                }
                __syncwarp();
            }
        }

        // ---- C. candidate windows, released in position order, evaluated exactly in pooled rounds --------
        // first the candidate bits of every 32-position word, in place of the seed bits (descending, word c needs seed word c-1)
        const int ncw = scan ? ((L + 31) >> 5) : 0;  // 32-position words of this read (<= 32)
        uint32_t nzw = 0;                             // bit c = candidate word c is not empty
        {
            const int ncwmax = __reduce_max_sync(0xFFFFFFFFu, ncw);
            uint32_t und_c = 0;  // force_und: undefined bits of word c
            auto und_word = [&](int c) -> uint32_t {  // bit b = read position 32c+b is undefined
                if (c < 0 || c >= ncw) return 0u;
                const uint32_t d = (st.d16(s + 32 * c) << 16) | st.d16(s + 32 * c + 16);  // bit 31-b = position 32c+b defined
                uint32_t u = __brev(~d);
                const int rem = L - 32 * c;
                if (rem < 32) u &= (1u << rem) - 1u;
                return u;
            };
            if (FN && any_force && force_und) und_c = und_word(ncwmax - 1);
            // no seed bit in the whole tile and nothing forced (most tiles of reads without adapters): S is all zero already
            const int c_top = (any_force || misc[0] != 0u) ? ncwmax - 1 : -1;
#pragma unroll 1
            for (int c = c_top; c >= 0; c--) {
                const uint32_t sc = S[c * 32 + lane], sp = c > 0 ? S[(c - 1) * 32 + lane] : 0u;
                uint32_t cb = __funnelshift_l(sp, sc, lag0);
                if (t.n_parts > 1) cb |= __funnelshift_l(sp, sc, lag1);
                if (t.n_parts > 2) cb |= __funnelshift_l(sp, sc, lag2) | __funnelshift_l(sp, sc, lag3);
                if (FN && any_force) {  // every window that contains an undefined base is decided by the exact evaluator
                    if (force_und) {
                        const uint32_t und_p = und_word(c - 1);
                        cb |= (uint32_t)(smear_left64(((uint64_t)und_c << 32) | und_p, k) >> 32);
                        und_c = und_p;
                    }
                }
                // keep positions k-1 <= i < L
                const int lowcut = min(max(k - 1 - 32 * c, 0), 32), hicut = min(max(L - 32 * c, 0), 32);
                const uint32_t mlow = lowcut >= 32 ? 0u : (0xFFFFFFFFu << lowcut), mhigh = hicut >= 32 ? 0xFFFFFFFFu : ((1u << hicut) - 1u);
                cb &= mlow & mhigh;
                if (c >= ncw) cb = 0;
                S[c * 32 + lane] = cb;
                nzw |= (cb != 0u ? 1u : 0u) << c;
            }
        }
        // A round: the reads that still look for their first hit share the 32 lanes: with na such reads each gets R = 8, 4, 2
        // or 1 lanes, lane j evaluates the (j mod R)-th lowest unreleased candidate of the (j div R)-th read. Nothing is
        // queued: the evaluating lane fetches the owner's current candidate word by shuffle and skips to its own bit.
        // down: highest candidates first (ktrim=l's search for the last hit).
        auto round = [&](bool has, int cw, uint32_t &creg, bool down) -> bool {
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, has);
            if (!bal) return false;
            const int na = __popc(bal);
            // (measured: giving each read floor(32/na) lanes instead of a power of two evaluates more candidates behind
            // the first hit than it saves rounds: 2.09 vs 1.81 ms per 8.4 M cfg-2 reads)
            // (a cap of 4 lanes changes nothing, 2 lanes cost 3 %, 1 lane 20 %)
            const int lgR = (na > 16) ? 0 : (na > 8) ? 1 : (na > 4) ? 2 : 3;
            DBG2(3, lane == 0);
            DBG2(6, lane == 0 ? na : 0);
            if (has) owners[__popc(bal & lt_mask)] = (uint32_t)lane;
            __syncwarp();
            const bool on = lane < (na << lgR);
            const int owner = on ? (int)owners[lane >> lgR] : 0;
            uint32_t x = __shfl_sync(0xFFFFFFFFu, creg, owner);
            const int cw_o = __shfl_sync(0xFFFFFFFFu, cw, owner);
            const int s_owner = __shfl_sync(0xFFFFFFFFu, s, owner);
            const int idx = lane & ((1 << lgR) - 1);
            if (down) {
#pragma unroll 1
                for (int c = 0; c < idx && x; c++) x &= ~(0x80000000u >> __clz(x));
            } else {
#pragma unroll 1
                for (int c = 0; c < idx; c++) x &= x - 1;
            }
            if (on && x) {
                DBG2(4, 1);
                const int pos = 32 * cw_o + (down ? 31 - __clz(x) : __ffs(x) - 1);
                const int id = exact_full(st, s_owner + pos, !((undef_mask >> owner) & 1u), p, t);
                if (id > 0) {
                    atomicMin(first32 + owner, ((uint32_t)pos << 22) | (uint32_t)id);
                    if (FMODE == FM_KTRIM_L) atomicMax(lastpos + owner, pos);
                    if (FMODE == FM_KMASK) or_range(M + owner, max(0, pos - k + 1), pos);  // bs.set(max(0, i-(k-1)), i+1), trimpad 0
                }
            }
            if (has) {  // the candidates this round took
                if (down) {
#pragma unroll 1
                    for (int c = 0; c < (1 << lgR) && creg; c++) creg &= ~(0x80000000u >> __clz(creg));
                } else {
#pragma unroll 1
                    for (int c = 0; c < (1 << lgR) && creg; c++) creg &= creg - 1;
                }
            }
            __syncwarp();
            return true;
        };

        {
            bool done = !scan;
            int cw = 0;
            uint32_t creg = 0;
            while (true) {
                if (!done && creg == 0 && nzw) {
                    cw = __ffs(nzw) - 1;
                    nzw &= nzw - 1;
                    creg = S[cw * 32 + lane];
                }
                if (!round(!done && creg != 0, cw, creg, false)) break;
                if (FMODE != FM_KMASK && first32[lane] != ~0u) done = true;  // everything still unreleased lies behind the confirmed hit
            }
        }

        int found = 0, id0 = -1, minLoc = 999999999, maxLoc = -1, count = 0;
        int lo = 0, hi = L;
        bool discarded = false, ktrimmed = false;
        {
            const uint32_t f32 = first32[lane];
            if (scan && f32 != ~0u) {
                const int pos = (int)(f32 >> 22);
                id0 = (int)(f32 & 0x3FFFFFu);
                found = 1;
                minLoc = pos - k + 1;
                maxLoc = pos;
            }
        }
        if (FMODE == FM_KTRIM_L) {
            // ktrim=l also needs the LAST hit: release the candidates from the read end downwards until one beyond the
            // forward phase's last confirmed hit is confirmed
            const int firstpos = found ? max(maxLoc, lastpos[lane]) : -1;
            bool bdone = !found;
            // candidate words from the top; nothing at or below the known hit matters
            uint32_t nzd = 0;
            if (found) {
#pragma unroll 1
                for (int c = firstpos >> 5; c < ncw; c++) nzd |= (S[c * 32 + lane] != 0u ? 1u : 0u) << c;
            }
            int cw = 0;
            uint32_t creg = 0;
            while (true) {
                if (!bdone && creg == 0) {
                    if (nzd) {
                        cw = 31 - __clz(nzd);
                        nzd &= ~(1u << cw);
                        creg = S[cw * 32 + lane];
                        if (32 * cw <= firstpos) {
                            const int keep_from = firstpos + 1 - 32 * cw;  // positions >= firstpos+1
                            creg = keep_from >= 32 ? 0u : (creg & (0xFFFFFFFFu << keep_from));
                        }
                    }
                    if (creg == 0 && nzd == 0) bdone = true;
                }
                if (!round(!bdone && creg != 0, cw, creg, true)) {
                    if (!__any_sync(0xFFFFFFFFu, !bdone)) break;
                    continue;
                }
                if (lastpos[lane] > firstpos) bdone = true;  // the highest hit of a round is the last hit of the read
            }
            if (found) maxLoc = max(firstpos, lastpos[lane]);
        }

        if (FMODE == FM_KFILTER) {  // countSetKmers with maxBadKmers==0 (jgi/BBDuk.java:3395-3457)
            count = found;
            discarded = found > 0;
        } else {
            // ---- T. short-k-mer tails (jgi/BBDuk.java:3910-3975; kmask :4080-4150) -------------------------
            // One side at a time: RIGHT = the suffixes of the read (lengths mink..k-1), else its prefixes (mink..k). Leaves the
            // lengths that hit in hm and the id of the shortest one in idt. Called by all lanes.
            auto tails = [&](auto right_tag, const bool tscan, uint32_t &hm, int &idt) {
                constexpr bool RIGHT = decltype(right_tag)::value;
                hm = 0u;
                idt = -1;
                if (!__any_sync(0xFFFFFFFFu, tscan)) return;
                // All tail k-mers of a read are sub-windows of ONE 32-base window: the suffix tails (ktrim=r) of the
                // window ending at the last base, the prefix tails (ktrim=l) of the window ending at base min(k,L)-1.
                // The tails read undefined bases as code 0 / complement 0 ("no N handling in tails").
                const int nmax = (RIGHT) ? min(k - 1, L) : min(k, L);
                const int nlo = max(p.mink, 1);
                const int e = (RIGHT) ? (s + L - 1) : (s + nmax - 1);
                uint64_t W = 0, RC = 0;
                uint32_t todo = 0;  // bit n = the tail of n bases has to be looked up
                if (tscan) {
                    W = st.win(e);  // slot t = base e-t
                    bool undef_here = false;
                    if (has_undef) {
                        const uint32_t dw = st.dwin(e);
                        const uint32_t need = nmax >= 32 ? 0xFFFFFFFFu : ((1u << nmax) - 1u);
                        undef_here = (dw & need) != need;
                        const uint64_t E = spread2(dw);
                        W &= E;
                        RC = bb_rcomp(W, 32) & rev2(E, 32);
                    } else {
                        RC = bb_rcomp(W, 32);
                    }
                    const uint32_t all = (nmax >= nlo) ? ((nmax >= 31 ? 0xFFFFFFFFu : ((1u << (nmax + 1)) - 1u)) & ~((1u << nlo) - 1u)) : 0u;
                    if (t.tail_words == 0 || undef_here) {
                        todo = all;
                    } else {
                        const int q = t.tail_q;
                        const uint32_t qm = (1u << (2 * q)) - 1u;  // q <= 12
                        const int ntop = min(nmax, k - 1);
                        // one bitmap (b_all) is tested once and decides all lengths, the other (b_len) once per length;
                        // ktrim=r: b_all = type II on the read's last q bases, b_len = type I on read[L-n : L-n+q];
                        // ktrim=l (mirror image): b_all = type I on read[0:q], b_len = type II on read[n-q : n]
                        const uint32_t *b_all = (RIGHT) ? tailb2 : tailb1, *b_len = (RIGHT) ? tailb1 : tailb2;
                        const uint32_t va = (RIGHT) ? ((uint32_t)W & qm) : ((uint32_t)(W >> (2 * max(nmax - q, 0))) & qm);
                        const uint32_t wa = __ldg(b_all + (va >> 5));
                        uint32_t ask = (ntop >= nlo) ? ((ntop >= 31 ? 0xFFFFFFFFu : ((1u << (ntop + 1)) - 1u)) & ~((1u << nlo) - 1u)) : 0u;
                        if (tail0_words && RIGHT == (p.ktrimLeft == 0)) {
                            // level 0 in shared memory: only lengths whose q-mer starts with an 8-mer that some listed q-mer
                            // starts with go on to the full bitmap in L2 (on random reads one length in ten)
                            // the 8 leading bases of the q-mer of length n are 16 bits of W: from bit 2(n-8) (ktrim=r: the q-mer is
                            // read[L-n : L-n+q]) or from bit 2(nmax-n+q-8) (ktrim=l: read[n-q : n]); four independent lookups per trip
                            const uint32_t wlo = (uint32_t)W, whi = (uint32_t)(W >> 32);
                            uint32_t pass = 0;
#pragma unroll 1
                            for (int n = nlo; n <= ntop; n += 2) {
                                uint32_t u[2], w0[2];
#pragma unroll
                                for (int j = 0; j < 2; j++) {
                                    const int sh = (RIGHT) ? 2 * (n + j - 8) : 2 * (nmax - n - j + q - 8);
                                    const uint32_t lo_ = (sh & 32) ? whi : wlo, hi_ = (sh & 32) ? 0u : whi;
                                    u[j] = __funnelshift_r(lo_, hi_, sh & 31) & 0xFFFFu;
                                    w0[j] = tail0[u[j] >> 5];
                                }
#pragma unroll
                                for (int j = 0; j < 2; j++) pass |= ((w0[j] >> (u[j] & 31u)) & 1u) << ((n + j) & 31);
                            }
                            ask &= pass;
                        }
#pragma unroll 1
                        while (ask) {  // up to four lookups in flight
                            uint32_t v[2], wv[2];
                            int nn[2];
#pragma unroll
                            for (int j = 0; j < 2; j++) {
                                nn[j] = ask ? __ffs(ask) - 1 : -1;
                                ask &= ask - 1;
                                const int sh = (RIGHT) ? 2 * (nn[j] - q) : 2 * (nmax - nn[j]);
                                v[j] = (uint32_t)(W >> (sh & 63)) & qm;
                                wv[j] = nn[j] >= 0 ? __ldg(b_len + (v[j] >> 5)) : 0u;
                            }
#pragma unroll
                            for (int j = 0; j < 2; j++) todo |= ((wv[j] >> (v[j] & 31u)) & 1u) << (nn[j] & 31);
                        }
                        if (nmax < q || ((wa >> (va & 31u)) & 1u)) todo = all;
                        if (!RIGHT && nmax == k) todo |= (k >= 31 ? 0x80000000u : (1u << k));  // a prefix of k bases is a full-length key
                        todo &= all;
                    }
                    DBG2(5, __popc(todo));
                }
                // The lookups of all 32 reads share the lanes: (read, length) pairs are queued and evaluated 32 at a time, so a read
                // whose tail holds an undefined base (all its lengths are asked) costs the warp one round, not a dozen dependent
                // trips to L2. A hit leaves its length in the owner's hit mask and (smallest length << 22 | id) in first32; the
                // reference's loop (ascending lengths, :3929-3975) is then replayed from the mask in closed form.
                if (__any_sync(0xFFFFFFFFu, todo != 0u)) {
                    uint32_t *hitm = reinterpret_cast<uint32_t *>(lastpos);
                    hitm[lane] = 0u;
                    first32[lane] = ~0u;
                    __syncwarp();
#pragma unroll 1
                    while (__any_sync(0xFFFFFFFFu, todo != 0u)) {
                        const int cnt = min(__popc(todo), ITEM_CAP);
                        const int incl = warp_scan_incl(cnt, lane);
                        const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                        {
                            int w = incl - cnt;
#pragma unroll 1
                            for (int c = 0; c < cnt; c++) {
                                queue[w++] = (uint16_t)(((uint32_t)lane << 11) + (uint32_t)(__ffs(todo) - 1));
                                todo &= todo - 1;
                            }
                        }
                        __syncwarp();
#pragma unroll 1
                        for (int base = 0; base < total; base += 32) {
                            const bool on = base + lane < total;
                            const uint32_t ent = on ? queue[base + lane] : 0u;
                            const int owner = (int)(ent >> 11), n = (int)(ent & 0x7FFu);
                            const uint64_t Wo = ((uint64_t)__shfl_sync(0xFFFFFFFFu, (uint32_t)(W >> 32), owner) << 32) |
                                                __shfl_sync(0xFFFFFFFFu, (uint32_t)W, owner);
                            const uint64_t RCo = ((uint64_t)__shfl_sync(0xFFFFFFFFu, (uint32_t)(RC >> 32), owner) << 32) |
                                                 __shfl_sync(0xFFFFFFFFu, (uint32_t)RC, owner);
                            const int nmax_o = (RIGHT) ? 1 : __shfl_sync(0xFFFFFFFFu, min(k, L), owner);
                            if (on) {
                                const uint64_t nm = (1ull << (2 * n)) - 1ull;
                                uint64_t kmer, rkmer;
                                if (RIGHT) {  // last n bases
                                    kmer = Wo & nm;
                                    rkmer = RCo >> (2 * (32 - n));
                                } else {  // first n bases
                                    kmer = (Wo >> (2 * (nmax_o - n))) & nm;
                                    rkmer = (RCo >> (2 * (32 - nmax_o))) & nm;
                                }
                                const int id = bb_table_get(t, bb_to_value(p, kmer, rkmer, 1ull << (2 * n)));
                                if (id > 0) {
                                    atomicOr(hitm + owner, 1u << n);
                                    atomicMin(first32 + owner, ((uint32_t)n << 22) | (uint32_t)id);
                                }
                            }
                        }
                        __syncwarp();
                    }
                    hm = hitm[lane];
                    if (hm) idt = (int)(first32[lane] & 0x3FFFFFu);
                }
            };
            int minLocX = 999999999, maxLocX = -1;
            if (found) {
                minLocX = minLoc + k;
                maxLocX = maxLoc - k;
            }
            uint32_t hm = 0, hm2 = 0;
            int idt = -1, idt2 = -1;
            if (FMODE == FM_KMASK) {
                // kmask scans both sides whatever the full-length scan found (:4080); reads shorter than k are left alone (:4027)
                const bool tscan = live && t.stored > 0 && p.useShortKmers && L >= k && !skip;
                tails(std::false_type{}, tscan, hm, idt);   // prefixes first, as the reference
                tails(std::true_type{}, tscan, hm2, idt2);
                if (hm | hm2) {
                    if (id0 < 0) id0 = hm ? idt : idt2;
                    found += __popc(hm) + __popc(hm2);
                    if (hm) or_range(M + lane, 0, (31 - __clz(hm)) - 1);          // bs.set(0, min(L, i+1)), i = n-1
                    if (hm2) or_range(M + lane, L - (31 - __clz(hm2)), L - 1);    // bs.set(i, L), i = L-n
                }
            } else {
                // ktrim guard (:3868): reads shorter than k still get the tails
                const bool tscan = live && t.stored > 0 && p.useShortKmers && L >= max(1, min(k, p.mink)) && !found && !skip;
                if (FMODE == FM_KTRIM_R) tails(std::true_type{}, tscan, hm, idt);
                else tails(std::false_type{}, tscan, hm, idt);
                if (hm) {
                    // lengths n_lo..n_hi hit; reference loop index i = L-n (ktrim=r, descending) or n-1 (ktrim=l, ascending)
                    const int n_lo = __ffs(hm) - 1, n_hi = 31 - __clz(hm);
                    if (id0 < 0) id0 = idt;
                    found += __popc(hm);
                    if (FMODE == FM_KTRIM_R) {
                        minLoc = L - n_hi;
                        minLocX = min(minLocX, L);
                        maxLoc = L - 1;
                        maxLocX = max(maxLocX, L - n_lo - 1);
                    } else {
                        minLoc = 0;
                        minLocX = min(minLocX, n_lo);
                        maxLoc = max(maxLoc, n_hi - 1);
                        maxLocX = max(maxLocX, 0);
                    }
                }
            }
            if (found && FMODE != FM_KMASK) {  // :3981-4012
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
        if (FMODE == FM_KMASK) {  // cardinality of the BitSet; the mask words go out as they are (zero when nothing was found, :4170)
            const int nw = (L + 31) >> 5;
            uint32_t *dst = (live && out.maskbits && out.mask_off) ? out.maskbits + out.mask_off[r] : nullptr;
#pragma unroll 1
            for (int w = 0; w < nw; w++) {
                const uint32_t m = found ? M[w * 32 + lane] : 0u;
                count += __popc(m);
                if (dst) dst[w] = m;
            }
            ktrimmed = count > 0;
            if (!found) id0 = -1;
        }
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
                    if (FMODE != FM_KMASK) {
                        xsum += len_pre + (paired ? len_pre_mate : 0);
                        rkt = paired ? 2 : 1;
                    }
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
    }
    if (stats) {
        const unsigned int v[8] = {(unsigned int)s_ri, (unsigned int)s_bi, (unsigned int)s_rk, (unsigned int)s_bk,
                                   (unsigned int)s_rf, (unsigned int)s_bf, (unsigned int)s_ro, (unsigned int)s_bo};
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const unsigned int x = __reduce_add_sync(0xFFFFFFFFu, v[q]);
            if (lane == 0 && x) atomicAdd((unsigned long long *)stats + q, (unsigned long long)x);
        }
    }
}

Fast2Geom make_geom2(const BBParams &p, const BBTable &t, int max_read_len) {
    Fast2Geom g;
    const int lmax = std::max(max_read_len, 16);
    g.nch = (32 * lmax + 15 + 15) / 16 + 1;
    g.nbadw = (g.nch + 31) / 32 + 1;
    g.sw = ((lmax + 16) >> 5) + 2;
    g.mw = p.mode == MODE_KMASK ? g.sw : 0;
    int wb = 32 * 8 + 32 * 4 + 16 + g.nbadw * 4 + (g.nch + PAD + TAIL) * 4 + ((g.nch + PAD + TAIL + 1) & ~1) * 2 + QCAP2 * 2 + g.sw * 32 * 4 + g.mw * 32 * 4;
    wb = (wb + 15) & ~15;
    g.warp_bytes = wb;
    g.part_off = t.n_filter_words;
    g.samp_off = t.n_filter_words + t.part_words + t.short_words;
    g.tail_off = g.samp_off + t.samp_words;
    // the 8-mer level-0 image of the per-length tail bitmap: type I (first) for ktrim=r, type II (second) for ktrim=l
    const uint32_t tail_main = std::max<uint32_t>(1u, (1u << (2 * t.tail_q)) >> 5);
    static const bool use_tail0 = [] {
        const char *e = getenv("BBDUK_B200_TAIL0");
        return !(e && atoi(e) == 0);
    }();
    g.tail0_off = 0;
    if (use_tail0 && p.mode == MODE_KTRIM && p.useShortKmers && t.tail_words >= 2 * tail_main + 2 * BB_TAIL0_WORDS)
        g.tail0_off = g.tail_off + 2 * tail_main + (p.ktrimLeft ? BB_TAIL0_WORDS : 0u);
    const int avail = FAST_SMEM_LIMIT - (SPELL_WORDS + 16384 + (int)BB_PART_WORDS + (g.tail0_off ? (int)BB_TAIL0_WORDS : 0)) * 4 - 64;
    g.warps = std::min(32, avail / wb);
    static const int warp_cap = [] {  // A/B knob: fewer warps = fewer phases of the loop in flight per scheduler
        const char *e = getenv("BBDUK_B200_FAST2_WARPS");
        return e ? atoi(e) : 0;
    }();
    if (warp_cap >= 8) g.warps = std::min(g.warps, warp_cap);
    return g;
}

}  // namespace

#if defined(BB_FAST_COUNT)
extern "C" __attribute__((visibility("default"))) int bbduk_b200_debug_fast2_counters(unsigned long long *out16, int reset) {
    if (cudaMemcpyFromSymbol(out16, bb_fast2_dbg, sizeof(unsigned long long) * 16) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long z[16] = {0};
        if (cudaMemcpyToSymbol(bb_fast2_dbg, z, sizeof z) != cudaSuccess) return 1;
    }
    return 0;
}
#endif

FastPlan plan_fast2(const BBParams &p, const BBTable &t, int max_read_len) {
    FastPlan pl{false, max_read_len, 0, 0};
    static const bool enabled = [] {
        const char *e = getenv("BBDUK_B200_FAST2");
        return !(e && atoi(e) == 0);
    }();
    if (!enabled) return pl;
    // kmask: the plain BitSet mode without padding (kmaskfullycovered and trimpad stay with the generic kernel)
    const bool mode_ok = (p.mode == MODE_KTRIM) || (p.mode == MODE_KFILTER && p.maxBadKmers0 == 0) ||
                         (p.mode == MODE_KMASK && !p.kmaskFullyCovered && p.trimPad == 0);
    if (!mode_ok) return pl;
    if (p.qHammingDistance != 0 || (p.useShortKmers && p.qHammingDistance2 != 0)) return pl;
    if (p.speed != 0 || p.qSkip != 1 || p.restrictLeft != 0 || p.restrictRight != 0) return pl;
    if (p.kbig > p.k || p.minKmerFraction != 0.0f) return pl;
    if (p.editDistance != 0 || t.part_words == 0 || t.n_parts < 1 || t.samp_words == 0 || t.part_w < 11) return pl;
    if (t.n_scaffolds >= (1 << 22)) return pl;  // the first hit of a read is kept as (position << 22 | id) in one 32-bit word
    if (max_read_len > MAX_FAST_LEN) max_read_len = MAX_FAST_LEN;
    const Fast2Geom g = make_geom2(p, t, max_read_len);
    if (g.warps < 8) return pl;
    pl.usable = true;
    pl.max_read_len = max_read_len;
    pl.smem_bytes = (SPELL_WORDS + 16384 + (int)BB_PART_WORDS + (g.tail0_off ? (int)BB_TAIL0_WORDS : 0)) * 4 + g.warps * g.warp_bytes + 64;
    pl.filter_words = 16384 + (int)BB_PART_WORDS;
    return pl;
}

int launch_fast2(const FastPlan &plan, const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_reads, int paired,
                 const BBParams &p, const BBTable &t, const bbduk_out &out, bbduk_stats *d_stats, unsigned long long *scaf_reads,
                 unsigned long long *scaf_bases, int32_t *d_handoff, unsigned int *d_handoff_n, int sm_count, cudaStream_t st,
                 const uint32_t *pk_F, const uint16_t *pk_D) {
    const Fast2Geom g = make_geom2(p, t, plan.max_read_len);
    const int threads = g.warps * 32;
    const int64_t n_tiles = (n_reads + 31) / 32;
    const int blocks = (int)std::min<int64_t>(sm_count, (n_tiles + g.warps - 1) / g.warps);
    auto go = [&](auto kern) -> int {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes) != cudaSuccess) return -1;
        kern<<<blocks, threads, plan.smem_bytes, st>>>(d_bases, d_offsets, n_reads, paired, p, t, out, d_stats, scaf_reads,
                                                       scaf_bases, d_handoff, d_handoff_n, g, pk_F, pk_D);
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    };
    const bool pw11 = t.part_w == 11;
    const bool fn = p.forbidNs != 0;
#define BB_GO2(FM, PK)                                                                                                  \
    (pw11 ? (fn ? go(bbduk_fast2_kernel<FM, PK, true, true>) : go(bbduk_fast2_kernel<FM, PK, true, false>))             \
          : (fn ? go(bbduk_fast2_kernel<FM, PK, false, true>) : go(bbduk_fast2_kernel<FM, PK, false, false>)))
    if (p.mode == MODE_KMASK) return pk_F ? -1 : BB_GO2(FM_KMASK, false);  // the host entry never packs for kmask (case matters to its caller)
    if (pk_F) {
        if (!pk_D) return -1;
        if (p.mode == MODE_KFILTER) return BB_GO2(FM_KFILTER, true);
        if (p.ktrimLeft) return BB_GO2(FM_KTRIM_L, true);
        return BB_GO2(FM_KTRIM_R, true);
    }
    if (p.mode == MODE_KFILTER) return BB_GO2(FM_KFILTER, false);
    if (p.ktrimLeft) return BB_GO2(FM_KTRIM_L, false);
    return BB_GO2(FM_KTRIM_R, false);
#undef BB_GO2
}
