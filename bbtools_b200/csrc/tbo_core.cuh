// tbo_core.cuh -- the per-pair arithmetic of BBDuk's trim-by-overlap step, written once for the device (tbo.cu, one lane
// per pair, per-lane arrays interleaved in shared memory with stride S) and for a host build (S = 1) that the CPU tests
// check (tests/test_tbo_core_cpu.py), so that the bit-plane packing and the two insert loops can be verified without a GPU.
//
// Follows jgi/BBMergeOverlapper.java:411-621 (mateByOverlapRatioJava) and :785-836 (findBestRatio); single-precision
// evaluation order kept (no contraction). The reference adds 0.95f per matching / mismatching base and leaves its
// base loop once bad > badlimit. Those partial sums only grow, so "the loop ran to its end" <=> T[mismatches] <= badlimit
// with T[c] = 0.95f added c times: the code COUNTS mismatches and looks the float sums up in T.
//
// Layout: every mate is held as bit planes, 32 bases per word, base i = bit 31-(i&31) of word i>>5 ("big-endian"):
// H / L = the two bits of the base code, N = "the byte is 'N'". A mismatch word of 32 aligned bases is
// (Ha^Hb)|(La^Lb): 2 logic ops + 1 popc per 32 bases. One of the two mates always starts at base 0 of an alignment
// (insert >= blen: r2' from 0, r1 from insert-blen; insert < blen: r1 from 0, r2' from blen-insert), so only the other
// one is funnel-shifted. Every alignment is first screened on its first 64 bases with the sliding mate held in
// registers (scan_side): more mismatches there than the largest count any badlimit of the loop admits => skipped.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TBO_HD __host__ __device__ __forceinline__
#else
#define TBO_HD static inline
#endif

namespace tbo {

constexpr int MAX_LEN = 1008;
constexpr int EXTRA_BADLIMIT = 20;  // jgi/BBMergeOverlapper.java:1464

struct Params {
    int minOverlap0, minOverlap, minInsert0, minInsert;  // BBDuk's values (jgi/BBDuk.java:5368-5371)
    float maxRatio, minSecondRatio, margin, offset, meeFilter;
    int qualOffset;
    int W;  // words per bit plane per lane
};

// words per plane for reads of up to max_len bases: the screen's sliding window touches word (len-1)/32 + 4
TBO_HD int plane_words(int max_len) { return (max_len + 31) / 32 + 4; }

// (hi << s) | (lo >> (32 - s)), s taken mod 32
TBO_HD uint32_t fsl(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    s &= 31u;
    return s ? ((hi << s) | (lo >> (32u - s))) : hi;
#endif
}
TBO_HD int popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
TBO_HD uint32_t brev(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(x);
#endif
}
TBO_HD uint32_t bperm(uint32_t a, uint32_t b, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, s);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
#endif
}
TBO_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
TBO_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
TBO_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b;
    return r;
#endif
}
TBO_HD int imin(int a, int b) { return a < b ? a : b; }
TBO_HD int imax(int a, int b) { return a > b ? a : b; }
// 0xFF in the n lowest byte lanes (n clamped to [0,4])
TBO_HD uint32_t low_bytes(int n) { return n >= 4 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ((1u << (8 * n)) - 1u)); }
// top `n` bits set (n clamped to [0,32]): the first n bases of a plane word
TBO_HD uint32_t head_mask(int n) { return n >= 32 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ~(0xFFFFFFFFu >> n)); }

// ---- packing ----------------------------------------------------------------------------------------------------------
// 16 aligned bytes at q as four little-endian words. The device reads them with one vector load (an aligned 16-byte
// block that holds a valid byte lies inside the allocation); the host build only touches [v_lo, v_hi).
TBO_HD void load_chunk(const uint8_t *q, uint32_t x[4], const uint8_t *v_lo, const uint8_t *v_hi) {
#if defined(__CUDA_ARCH__)
    (void)v_lo;
    (void)v_hi;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(q));
    x[0] = v.x;
    x[1] = v.y;
    x[2] = v.z;
    x[3] = v.w;
#else
    for (int k = 0; k < 4; k++) {
        uint32_t w = 0;
        for (int j = 0; j < 4; j++) {
            const uint8_t *b = q + 4 * k + j;
            if (b >= v_lo && b < v_hi) w |= (uint32_t)*b << (8 * j);
        }
        x[k] = w;
    }
#endif
}

// Raw planes of the aligned 16-byte chunks that cover p[0..len): stream position of base i = u0 + i, u0 = p & 15.
// Bytes of the first / last chunk outside the read are replaced by 'A' (code 0, valid). Writes nw = ceil(chunks/2) words
// per plane and zeroes the rest of the W words; returns nw. bad != 0 afterwards <=> some byte of the read is not one
// of A C G T (GENERAL: ... and not N; n_any != 0 <=> some byte is N). tn is only written if GENERAL.
//   code bits: bit0 = b1^b2, bit1 = b2^b3 of the ASCII byte (dna/AminoAcid.java:1289-1320 for A C G T);
//   four byte lanes -> four plane bits with one multiply (all partial products land on distinct bits);
//   validity: (b4,b2,b1) indexes an 8-entry table of the only byte that may sit there; PRMT does four lookups at once.
template <bool GENERAL, int S>
TBO_HD int pack_raw(const uint8_t *p, int len, uint32_t *th, uint32_t *tl, uint32_t *tn, int W, uint32_t &u0_out, uint32_t &bad,
                    uint32_t &n_any) {
    const uint32_t u0 = len > 0 ? (uint32_t)((uintptr_t)p & 15u) : 0u;
    const uint8_t *q = p - u0;
    const int nchunks = len > 0 ? (int)((u0 + (uint32_t)len + 15u) >> 4) : 0;
    uint32_t ah = 0, al = 0, an = 0;
    int wi = 0;
    auto chunk = [&](int c, bool edge) {
        uint32_t x[4];
        load_chunk(q + 16 * c, x, p, p + len);
        if (edge) {
            // bytes [v0, v1) of this chunk belong to the read: bit j of m = byte j is kept
            const int v0 = imax((int)u0 - 16 * c, 0), v1 = imin((int)u0 + len - 16 * c, 16);
            const uint32_t m = ((1u << v1) - 1u) & ~((1u << v0) - 1u);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < 4; k++) {
                // nibble k -> 0xFF per kept byte lane (the partial products of the multiply land on distinct bits)
                const uint32_t keep = ((((m >> (4 * k)) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
                x[k] = (x[k] & keep) | (0x41414141u & ~keep);
            }
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; k++) {
            uint32_t xx = x[k];
            if (GENERAL) {
                const uint32_t y = xx ^ 0x4E4E4E4Eu;                                     // zero byte <=> 'N'
                const uint32_t z = ((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y;                // bit 7 of a lane <=> byte != 0
                const uint32_t nb = (~z >> 7) & 0x01010101u;
                an = fsl(nb * 0x80402010u, an, 4);
                n_any |= nb;
                xx ^= nb * 0x0Fu;  // 'N' -> 'A'
            }
            const uint32_t B1 = xx >> 1, B2 = xx >> 2, t = B1 ^ B2;
            al = fsl((t & 0x01010101u) * 0x80402010u, al, 4);
            ah = fsl((t & 0x02020202u) * 0x40201008u, ah, 4);
            const uint32_t idx = (B1 & 0x03030303u) | (B2 & 0x04040404u);
            const uint32_t sel = (idx | (idx >> 12)) & 0xFFFFu;  // nibbles: lane 0, lane 2, lane 1, lane 3
            bad |= bperm(0x47004341u, 0x00540000u, sel) ^ bperm(xx, xx, 0x3120u);
        }
        if (c & 1) {
            th[wi * S] = ah;
            tl[wi * S] = al;
            if (GENERAL) tn[wi * S] = an;
            wi++;
        }
    };
    // the first and the last chunk are peeled: every lane of a warp reaches them at the same point of the code
    if (nchunks > 0) chunk(0, true);
    for (int c = 1; c < nchunks - 1; c++) chunk(c, false);
    if (nchunks > 1) chunk(nchunks - 1, true);
    if (nchunks & 1) {
        th[wi * S] = ah << 16;
        tl[wi * S] = al << 16;
        if (GENERAL) tn[wi * S] = an << 16;
        wi++;
    }
    const int nw = wi;
    for (; wi < W; wi++) {
        th[wi * S] = 0;
        tl[wi * S] = 0;
        if (GENERAL) tn[wi * S] = 0;
    }
    u0_out = u0;
    return nw;
}

// plane of the read itself: base i -> bit i of the output (stream position u0 + i of the raw plane)
template <int S>
TBO_HD void finish_forward(const uint32_t *raw, uint32_t *out, int len, uint32_t u0, int W) {
    uint32_t cur = raw[0];
    for (int w = 0; w < W; w++) {
        const uint32_t nxt = (w + 1 < W) ? raw[(w + 1) * S] : 0u;
        out[w * S] = fsl(nxt, cur, u0) & head_mask(len - 32 * w);
        cur = nxt;
    }
}

// plane of the reverse(-complemented) read: out base j = raw base len-1-j, inverted if `complement` (3 - code = ~code)
template <int S>
TBO_HD void finish_reverse(const uint32_t *raw, uint32_t *out, int len, uint32_t u0, int nw, int W, bool complement) {
    const int start = 32 * nw - (int)u0 - len;  // position of out base 0 in the reversed stream
    const int w0 = start >> 5;
    const uint32_t r = (uint32_t)start & 31u;
    auto R = [&](int i) -> uint32_t { return (i < nw) ? brev(raw[(nw - 1 - i) * S]) : 0u; };
    uint32_t cur = R(w0);
    for (int w = 0; w < W; w++) {
        const uint32_t nxt = R(w0 + w + 1);
        uint32_t v = fsl(nxt, cur, r);
        if (complement) v = ~v;
        out[w * S] = v & head_mask(len - 32 * w);
        cur = nxt;
    }
}

// ---- counting ---------------------------------------------------------------------------------------------------------
template <int S>
struct Ctx {
    const uint32_t *ah, *al, *an, *bh, *bl, *bn;  // planes of r1 (forward) and r2' (reverse complement); word w at [w * S]
    // exact path: a mate holds a byte other than A C G T N, the reference compares raw bytes
    const uint8_t *a_bytes;      // r1 trimmed, forward
    const uint8_t *b_rev_bytes;  // r2 trimmed, LAST base (b[j] = comp[b_rev_bytes[-j]])
    const uint8_t *comp;
    bool exact;
};

// mismatches / non-N matches of a[istart..istart+ov) against b[jstart..jstart+ov); one of istart, jstart is 0.
// Exact while T[nbad] <= badlimit; counting stops once it is exceeded (then ngood is not used).
template <bool GENERAL, int S>
TBO_HD void count_exact(const Ctx<S> &c, int istart, int jstart, int ov, float badlimit, const float *T, int &nbad, int &ngood) {
    nbad = 0;
    ngood = 0;
    if (GENERAL && c.exact) {
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
        return;
    }
    const bool slide_a = istart > 0;
    const uint32_t *fh = slide_a ? c.bh : c.ah, *fl = slide_a ? c.bl : c.al, *fn = slide_a ? c.bn : c.an;
    const uint32_t *sh = slide_a ? c.ah : c.bh, *sl = slide_a ? c.al : c.bl, *sn = slide_a ? c.an : c.bn;
    const int s = slide_a ? istart : jstart;
    int w = (s >> 5) * S;
    const uint32_t r = (uint32_t)s & 31u;
    uint32_t h_cur = sh[w], l_cur = sl[w], n_cur = GENERAL ? sn[w] : 0u;
    for (int t = 0, tw = 0; t < ov; t += 32, tw += S) {
        w += S;
        const uint32_t h_nxt = sh[w], l_nxt = sl[w];
        const uint32_t d = (fsl(h_nxt, h_cur, r) ^ fh[tw]) | (fsl(l_nxt, l_cur, r) ^ fl[tw]);
        const uint32_t m = head_mask(ov - t);
        if (GENERAL) {
            const uint32_t n_nxt = sn[w];
            const uint32_t na = fsl(n_nxt, n_cur, r), nb = fn[tw];
            nbad += popc(((d & ~(na | nb)) | (na ^ nb)) & m);
            ngood += popc(~d & ~(na | nb) & m);
            n_cur = n_nxt;
        } else {
            nbad += popc(d & m);
        }
        h_cur = h_nxt;
        l_cur = l_nxt;
        if (T[nbad] > badlimit) return;
    }
    if (!GENERAL) ngood = ov - nbad;
}

// ---- the screen ---------------------------------------------------------------------------------------------------------
// Alignments are visited by falling insert size. While insert > blen r1 slides (from base s = insert - blen, s falling)
// against r2' from base 0 ("side A"); afterwards r2' slides (from s = blen - insert, s rising) against r1 from base 0
// ("side B"). The screen runs BEFORE the insert loops and out of their order: for every s of a side it counts the
// mismatches (GENERAL: by the reference's N rules) among the first min(ov, 32*NW) bases and sets bit s of the lane's
// candidate bitmap if they do not exceed `cap`, the largest count with which an alignment can still change the state
// of the coming loop. All lanes of a
// warp walk the same (word, shift) sequence, so the warp stays on one code path whatever the lengths of its pairs:
// per word index the fixed mate's NW words and the sliding mate's NW+1 words sit in registers and one alignment costs
// 2*NW funnel shifts, 2*NW logic ops and NW popc. The insert loops then visit the set bits only.
TBO_HD bool any_lane(bool p) {
#if defined(__CUDA_ARCH__)
    return __any_sync(__activemask(), p) != 0;  // a hint only: either answer is correct for every lane that votes true
#else
    return p;
#endif
}
TBO_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
TBO_HD int ffs32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x);
#else
    return __builtin_ffs((int)x);
#endif
}
// bits r of word w whose position 32w + r lies in [lo, hi]
TBO_HD uint32_t range_bits(int w, int lo, int hi) {
    const int r_lo = imax(lo - 32 * w, 0), r_hi = imin(hi - 32 * w, 31);
    if (r_lo > r_hi) return 0u;
    return ((2u << r_hi) - 1u) & ~((1u << r_lo) - 1u);
}

// An upper bound, cheap enough for the inner loop, of the largest mismatch count a badlimit of the form a * ov + b admits:
// T[c] >= 0.9499 * c, so c <= limit * 1.0527; fixed point with constants rounded up, + 2 for the float roundings of the
// reference's own limit (checked against cap_of for every ov by the host test).
struct CapLine {
    int a_fx, b_int;  // ((a_fx * ov) >> 16) + b_int
#if defined(__CUDACC__)
    __host__ __device__ __forceinline__
#endif
    int of(int ov) const { return ((a_fx * ov) >> 16) + b_int; }
};
// the same for a limit on the ratio: (T[c] + offset) / ov < ratio  =>  c <= of(ov)
TBO_HD CapLine cap_line(float a, float b);
TBO_HD CapLine ratio_cap_line(float ratio, float offset) { return cap_line(ratio, offset < 0.0f ? -offset : 0.0f); }
TBO_HD CapLine cap_line(float a, float b) {
    CapLine c;
    if (a > 2.0f) a = 2.0f;  // a count never exceeds ov, so a slope above 1 already admits everything (and a_fx * ov stays in range)
    c.a_fx = (int)(a * 1.0527f * 65536.0f) + 2;
    c.b_int = (int)(b * 1.0527f) + 3;
    return c;
}

// One word of candidate bits: shifts r = 0..31 of the sliding mate's words v* against the fixed mate's first NW words f*.
// ncap = ~cap of the word (the sign bit of ncap + count says "at most cap mismatches"). MASKED: positions beyond either
// mate's end do not count -- vv are the sliding mate's validity words (they shift with it), fm the fixed mate's.
template <bool GENERAL, int NW, bool MASKED>
TBO_HD uint32_t scan_word(const uint32_t (&fh)[NW], const uint32_t (&fl)[NW], const uint32_t (&fn)[NW], const uint32_t (&fm)[NW],
                          const uint32_t (&vh)[NW + 1], const uint32_t (&vl)[NW + 1], const uint32_t (&vn)[NW + 1],
                          const uint32_t (&vv)[NW + 1], int ncap) {
    uint32_t bits = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll(NW == 1 ? 8 : 4)
#endif
    for (uint32_t r = 0; r < 32; r++) {
        int n = ncap;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < NW; k++) {
            uint32_t d = (fsl(vh[k + 1], vh[k], r) ^ fh[k]) | (fsl(vl[k + 1], vl[k], r) ^ fl[k]);
            if (GENERAL) {
                const uint32_t na = fsl(vn[k + 1], vn[k], r);
                d = (d & ~(na | fn[k])) | (na ^ fn[k]);
            }
            if (MASKED) d &= fsl(vv[k + 1], vv[k], r) & fm[k];
            n += popc(d);
        }
        bits = fsl((uint32_t)n, bits, 1);  // one funnel shift takes n's sign bit in: after 32 steps bit 31-r holds the verdict of shift r
    }
    return brev(bits);
}

// candidate bits of one side for s in [s_lo, s_hi] (words s_lo>>5 .. s_hi>>5 of cand are written). X = length of the
// sliding mate, Y = of the fixed one: ov(s) = min(X - s, Y). Every word gets the cap of its longest overlap (s = 32w):
// cl.of grows with ov, so that cap admits whatever the cap of a later s of the word admits.
// Words on the far side of `cut` (tight_below: words below it, else words from it on) belong to alignments the loop
// visits after the one that set maxRatio; they get the tighter cap (cap2, cl2) -- see mate_by_overlap_ratio.
struct Cap2 {
    int cut_word;      // a word index
    bool tight_below;  // which side of cut_word is tight
    int cap;
    CapLine cl;
};
template <bool GENERAL, int S, int NW>
TBO_HD void scan_side(const uint32_t *sh, const uint32_t *sl, const uint32_t *sn, const uint32_t *fhp, const uint32_t *flp,
                      const uint32_t *fnp, int X, int Y, int s_lo, int s_hi, int cap, CapLine cl, const Cap2 &c2, uint32_t *cand) {
    if (s_hi < s_lo) return;
    uint32_t fh[NW], fl[NW], fn[NW], fm[NW];
    for (int k = 0; k < NW; k++) {
        fh[k] = fhp[k * S];
        fl[k] = flp[k * S];
        fn[k] = GENERAL ? fnp[k * S] : 0u;
        fm[k] = head_mask(Y - 32 * k);
    }
    for (int w = s_lo >> 5; w <= (s_hi >> 5); w++) {
        uint32_t vh[NW + 1], vl[NW + 1], vn[NW + 1], vv[NW + 1];
        for (int k = 0; k <= NW; k++) {
            vh[k] = sh[(w + k) * S];
            vl[k] = sl[(w + k) * S];
            vn[k] = GENERAL ? sn[(w + k) * S] : 0u;
            vv[k] = head_mask(X - 32 * (w + k));
        }
        const int ov0 = imax(imin(X - 32 * w, Y), 0);
        int cap_w = imin(cap, cl.of(ov0));
        if (c2.tight_below ? (w < c2.cut_word) : (w >= c2.cut_word)) cap_w = imin(cap_w, imin(c2.cap, c2.cl.of(ov0)));
        const int ncap = ~cap_w;
        // the shortest overlap of the word is at its last s; unmasked only if no lane of the warp needs the masks
        const bool masked = any_lane(imin(X - (32 * w + 31), Y) < 32 * NW);
        const uint32_t bits = masked ? scan_word<GENERAL, NW, true>(fh, fl, fn, fm, vh, vl, vn, vv, ncap)
                                     : scan_word<GENERAL, NW, false>(fh, fl, fn, fm, vh, vl, vn, vv, ncap);
        cand[w * S] = bits & range_bits(w, s_lo, s_hi);
    }
}

// the alignments of one insert loop (inserts i_top down to i_bot) as two runs of s, and their candidate bitmaps
template <int S>
struct Cands {
    uint32_t *A, *B;   // bit s of side A / side B; word w at [w * S]
    int a_lo, a_hi;    // side A: s = insert - blen in [a_lo, a_hi], visited from a_hi down
    int b_lo, b_hi;    // side B: s = blen - insert in [b_lo, b_hi], visited from b_lo up
    int phase, w;      // iterator: 0 = side A, 1 = side B, 2 = done
    uint32_t bits;
};

#ifndef TBO_CAP_NW1
#define TBO_CAP_NW1 20
#endif
template <bool GENERAL, int S>
TBO_HD void build_cands(const Ctx<S> &c, int alen, int blen, int i_top, int i_bot, int cap, CapLine cl, Cands<S> &q,
                        int i_tight = -(1 << 30), int cap_tight = 0, CapLine cl_tight = CapLine{0, 0}) {
    q.a_hi = i_top - blen;
    q.a_lo = imax(1, i_bot - blen);
    q.b_lo = imax(0, blen - i_top);
    q.b_hi = blen - i_bot;
    if (i_top < i_bot) {
        q.a_hi = q.a_lo - 1;
        q.b_hi = q.b_lo - 1;
    }
    const int longest = imin(alen, blen);
    // inserts below i_tight take the tight cap. Side A: s = insert - blen < i_tight - blen, rounded down to a word so
    // that the word holding the boundary stays loose; side B: s = blen - insert > blen - i_tight, rounded up likewise.
    Cap2 ca, cb;
    ca.tight_below = true;
    ca.cut_word = imax(i_tight - blen, 0) >> 5;
    cb.tight_below = false;
    cb.cut_word = blen - i_tight < 0 ? 0 : ((blen - i_tight) >> 5) + 1;
    ca.cap = cb.cap = cap_tight;
    ca.cl = cb.cl = cl_tight;
    if ((GENERAL && c.exact) || cap >= imin(longest, 90)) {  // no screen: the byte path, or a cap the screen cannot beat
        for (int w = imax(q.a_lo, 0) >> 5; w <= (q.a_hi >> 5) && q.a_hi >= q.a_lo; w++) q.A[w * S] = range_bits(w, q.a_lo, q.a_hi);
        for (int w = q.b_lo >> 5; w <= (q.b_hi >> 5) && q.b_hi >= q.b_lo; w++) q.B[w * S] = range_bits(w, q.b_lo, q.b_hi);
    } else if (!any_lane(cap > TBO_CAP_NW1)) {  // one window width per warp (a wider window than a lane needs is still a valid screen)
        // The screen never needs the N planes: an N reads as A in both mates (pack_pair), and with that reading a position
        // counts as a mismatch at most as often as by the reference's N rules (N against a base: always bad there, bad here
        // unless the base is A; N against N: bad in neither) -- a lower bound is all the screen promises.
        // 32 bases already reject a chance alignment (24 expected mismatches) with probability > 0.9 at this cap; the few
        // that slip through cost one exact count each, less than a second window on every alignment
        scan_side<false, S, 1>(c.ah, c.al, c.an, c.bh, c.bl, c.bn, alen, blen, q.a_lo, q.a_hi, cap, cl, ca, q.A);
        scan_side<false, S, 1>(c.bh, c.bl, c.bn, c.ah, c.al, c.an, blen, alen, q.b_lo, q.b_hi, cap, cl, cb, q.B);
    } else if (!any_lane(cap > 43)) {
        scan_side<false, S, 2>(c.ah, c.al, c.an, c.bh, c.bl, c.bn, alen, blen, q.a_lo, q.a_hi, cap, cl, ca, q.A);
        scan_side<false, S, 2>(c.bh, c.bl, c.bn, c.ah, c.al, c.an, blen, alen, q.b_lo, q.b_hi, cap, cl, cb, q.B);
    } else {
        scan_side<false, S, 4>(c.ah, c.al, c.an, c.bh, c.bl, c.bn, alen, blen, q.a_lo, q.a_hi, cap, cl, ca, q.A);
        scan_side<false, S, 4>(c.bh, c.bl, c.bn, c.ah, c.al, c.an, blen, alen, q.b_lo, q.b_hi, cap, cl, cb, q.B);
    }
    if (q.a_hi >= q.a_lo) {
        q.phase = 0;
        q.w = q.a_hi >> 5;
        q.bits = q.A[q.w * S];
    } else if (q.b_hi >= q.b_lo) {
        q.phase = 1;
        q.w = q.b_lo >> 5;
        q.bits = q.B[q.w * S];
    } else {
        q.phase = 2;
        q.w = 0;
        q.bits = 0;
    }
}

// next candidate in visiting order: returns false when there is none; else side (0 = A) and s
template <int S>
TBO_HD bool next_cand(Cands<S> &q, int &side, int &s) {
    for (;;) {
        if (q.phase == 2) return false;
        if (q.bits) {
            const int r = q.phase == 0 ? 31 - clz32(q.bits) : ffs32(q.bits) - 1;
            q.bits &= ~(1u << r);
            side = q.phase;
            s = 32 * q.w + r;
            return true;
        }
        if (q.phase == 0) {
            if (q.w > (q.a_lo >> 5)) {
                q.w--;
                q.bits = q.A[q.w * S];
            } else if (q.b_hi >= q.b_lo) {
                q.phase = 1;
                q.w = q.b_lo >> 5;
                q.bits = q.B[q.w * S];
            } else {
                q.phase = 2;
            }
        } else {
            if (q.w < (q.b_hi >> 5)) {
                q.w++;
                q.bits = q.B[q.w * S];
            } else {
                q.phase = 2;
            }
        }
    }
}

// largest count c with T[c] <= limit (T grows strictly; T[0] = 0 <= limit always holds for the limits used here)
TBO_HD int cap_of(float limit, const float *T, int n_T) {
    int c = (int)fdiv(limit, 0.95f) + 2;
    if (c > n_T - 1) c = n_T - 1;
    if (c < 0) c = 0;
    while (c > 0 && T[c] > limit) c--;
    return c;
}

// jgi/BBMergeOverlapper.java:785-836
template <bool GENERAL, int S>
TBO_HD float find_best_ratio(const Ctx<S> &c, Cands<S> &q, int alen, int blen, int minOverlap0, int minOverlap, int minInsert,
                             float maxRatio, float offset, const float *T, int n_T, int &best_ins) {
    best_ins = -1;  // the insert whose ratio is returned (if any alignment lowered the initial value)
    float bestRatio = fadd(maxRatio, 0.0001f);
    const float halfmax = fmul(maxRatio, 0.5f);
    // An alignment changes this loop's state only if it has no mismatch at all or its ratio (bad + offset) / ov is below
    // the running bestRatio, which never exceeds its initial value; the far looser badlimit (+20) only bounds the
    // reference's counting. So the screen admits the counts c with c = 0 or T[c] + offset < bestRatio0 * ov, i.e.
    // (T[c] >= 0.9499 c) c <= ratio_cap_line(ov); checked against the float evaluation for every ov by the host test.
    const CapLine cl = ratio_cap_line(bestRatio, offset);
    const int cap_max = cl.of(imin(alen, blen));
    build_cands<GENERAL, S>(c, alen, blen, alen + blen - minOverlap, minInsert, cap_max, cl, q);
    int side, s;
    while (next_cand<S>(q, side, s)) {  // for (insert = alen + blen - minOverlap; insert >= minInsert; insert--), candidates only
        const int insert = side == 0 ? s + blen : blen - s;
        const int istart = side == 0 ? s : 0, jstart = side == 0 ? 0 : s;
        const int ov = imin(alen - istart, imin(blen - jstart, insert));
        int nbad, ngood;
        const float badlimit = fadd(fmul(bestRatio, (float)ov), (float)EXTRA_BADLIMIT);
        count_exact<GENERAL, S>(c, istart, jstart, ov, badlimit, T, nbad, ngood);
        const float bad = T[nbad];
        if (bad <= badlimit) {
            const float good = T[ngood];
            if (bad == 0.0f && good > (float)minOverlap0 && good < (float)minOverlap) return 100.0f;
            const float ratio = fdiv(fadd(bad, offset), (float)ov);
            if (ratio < bestRatio) {
                bestRatio = ratio;
                best_ins = insert;
                if (good >= (float)minOverlap && ratio < halfmax) return bestRatio;
            }
        }
    }
    return bestRatio;
}

// jgi/BBMergeOverlapper.java:411-621 (TAG_CUSTOM = MAKE_VECTOR = false); returns bestInsert, sets ambig
// STAGE 0: both loops. STAGE 1: findBestRatio only; returns -3 and *h if the second loop has to run.
// STAGE 2: the second loop, with findBestRatio's result handed in through *h.
struct Handoff {
    float x;   // findBestRatio's ratio
    int ins;   // the insert that has it
};
template <bool GENERAL, int STAGE, int S>
TBO_HD int mate_by_overlap_ratio(const Ctx<S> &c, Cands<S> &q, int alen, int blen, const Params &p, const float *T, int n_T,
                                 bool &ambig_out, Handoff *h) {
    const int minOverlap = imax(4, imax(p.minOverlap0, p.minOverlap));
    int minOverlap0;
    {  // Tools.mid(4, minOverlap0, minOverlap): the median
        const int x = 4, y = p.minOverlap0, z = minOverlap;
        minOverlap0 = x < y ? (y < z ? y : imax(x, z)) : (x < z ? x : imax(y, z));
    }
    const int minLength = imin(alen, blen);
    float maxRatio = p.maxRatio;
    int x_ins;
    ambig_out = false;
    {
        float x;
        if (STAGE == 2) {
            x = h->x;
            x_ins = h->ins;
        } else {
            x = find_best_ratio<GENERAL, S>(c, q, alen, blen, minOverlap0, minOverlap, p.minInsert, maxRatio, p.offset, T, n_T, x_ins);
        }
        if (x > maxRatio) return -1;  // rvector[4] = 0
        if (STAGE == 1) {
            h->x = x;
            h->ins = x_ins;
            return -3;
        }
        maxRatio = x < maxRatio ? x : maxRatio;
    }
    const float margin = p.margin, offset = p.offset;
    const float margin2 = fdiv(fadd(margin, offset), (float)minLength);
    int bestInsert = -1;
    float bestRatio = 1.0f, secondBestRatio = 1.0f;
    bool ambig = false;
    // min(bestRatio, maxRatio) <= maxRatio and ov <= minLength: the screen's cap for this loop.
    // maxRatio is now the ratio of the alignment at insert x_ins (x <= p.maxRatio got us here, and an x below the initial
    // value is always some alignment's ratio). That alignment is visited by this loop too, passes its badlimit (bad <
    // x * ov) and leaves bestRatio <= maxRatio. Every alignment visited after it -- the inserts below x_ins -- then changes
    // the state only if it has no mismatch or ratio < bestRatio * margin <= maxRatio * margin: those inserts get the cap
    // of that ratio test. (Before x_ins bestRatio may still be anything up to 1 and only the badlimit bounds the count.)
    const int cap_max =
        cap_of(fadd(fadd(fmul(1.2f, fmul(fmul(maxRatio, margin), (float)minLength)), 1.0f), (float)EXTRA_BADLIMIT), T, n_T);
    const CapLine cl_t = ratio_cap_line(fmul(maxRatio, margin), offset);
    build_cands<GENERAL, S>(c, alen, blen, alen + blen - minOverlap0, p.minInsert0, cap_max,
                            cap_line(fmul(1.2f, fmul(maxRatio, margin)), 1.0f + (float)EXTRA_BADLIMIT), q, x_ins, cl_t.of(minLength),
                            cl_t);
    int side, s;
    while (next_cand<S>(q, side, s)) {  // for (insert = alen + blen - minOverlap0; insert >= minInsert0; insert--), candidates only
        const int insert = side == 0 ? s + blen : blen - s;
        const int istart = side == 0 ? s : 0, jstart = side == 0 ? 0 : s;
        const int ov = imin(alen - istart, imin(blen - jstart, insert));
        const float rmin = bestRatio < maxRatio ? bestRatio : maxRatio;
        const float badlimit = fadd(fadd(fmul(1.2f, fmul(fmul(rmin, margin), (float)ov)), 1.0f), (float)EXTRA_BADLIMIT);
        int nbad, ngood;
        count_exact<GENERAL, S>(c, istart, jstart, ov, badlimit, T, nbad, ngood);
        const float bad = T[nbad];
        if (bad <= badlimit) {
            const float good = T[ngood];
            if (bad == 0.0f && good > (float)minOverlap0 && good < (float)minOverlap) {
                ambig_out = true;
                return -1;
            }
            const float ratio = fdiv(fadd(bad, offset), (float)ov);
            if (ratio < fmul(bestRatio, margin)) {
                ambig = (fmul(ratio, margin) >= bestRatio || good < (float)minOverlap);
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

// pack both mates of a pair into the lane's planes; arrays: planes[k * W * S], k = 0..7 (AH AL BH BL, two raw planes, the
// two candidate bitmaps) and, if GENERAL, 8..10 (AN BN + a raw plane). Returns bit 0 = a byte outside A C G T (GENERAL:
// outside A C G T N), bit 1 = an 'N' seen.
constexpr int N_PLANES_ACGT = 8, N_PLANES_GENERAL = 11;
template <bool GENERAL, int S>
TBO_HD uint32_t pack_pair(const uint8_t *a, int alen, const uint8_t *b0, int blen, uint32_t *planes, int W, Ctx<S> &c, Cands<S> &q) {
    const int P = W * S;
    uint32_t *ah = planes, *al = planes + P, *bh = planes + 2 * P, *bl = planes + 3 * P, *th = planes + 4 * P,
             *tl = planes + 5 * P;
    q.A = planes + 6 * P;
    q.B = planes + 7 * P;
    uint32_t *an = GENERAL ? planes + 8 * P : nullptr, *bn = GENERAL ? planes + 9 * P : nullptr,
             *tn = GENERAL ? planes + 10 * P : nullptr;
    uint32_t bad = 0, n_any = 0, u0;
    int nw = pack_raw<GENERAL, S>(a, alen, th, tl, tn, W, u0, bad, n_any);
    finish_forward<S>(th, ah, alen, u0, W);
    finish_forward<S>(tl, al, alen, u0, W);
    if (GENERAL) finish_forward<S>(tn, an, alen, u0, W);
    nw = pack_raw<GENERAL, S>(b0, blen, th, tl, tn, W, u0, bad, n_any);
    finish_reverse<S>(th, bh, blen, u0, nw, W, true);
    finish_reverse<S>(tl, bl, blen, u0, nw, W, true);
    if (GENERAL) {
        finish_reverse<S>(tn, bn, blen, u0, nw, W, false);
        // An N of r2' must read as A like an N of r1 (the packer's N -> A was complemented to T above): the screen counts
        // code differences without the N planes, and N against N is no mismatch for the reference. (The exact counts mask
        // both code planes with the N planes, so nothing else sees these bits.)
        for (int w = 0; w < W; w++) {
            const uint32_t keep = ~bn[w * S];
            bh[w * S] &= keep;
            bl[w * S] &= keep;
        }
    }
    c.ah = ah;
    c.al = al;
    c.an = an;
    c.bh = bh;
    c.bl = bl;
    c.bn = bn;
    c.a_bytes = a;
    c.b_rev_bytes = b0 + blen - 1;
    c.exact = GENERAL && bad != 0;
    return (bad != 0 ? 1u : 0u) | (n_any != 0 ? 2u : 0u);
}

}  // namespace tbo
