// Host build of bbtools_b200/csrc/tbo_core.cuh (stride 1): the same packing, screening and insert loops the device lanes
// run, driven pair by pair the way tbo.cu's four launches do. TEST INFRASTRUCTURE: compiled by tests/test_tbo_core_cpu.py
// with g++ and compared with the oracle; nothing in bbtools_b200 loads it.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../bbtools_b200/csrc/tbo_core.cuh"

extern "C" {

// planes of one read as the device lanes build them: out[0..W) = H, [W..2W) = L, [2W..3W) = N; returns the flags of pack_pair
int tbo_host_pack(const uint8_t *buf, int start, int len, int reverse, int W, uint32_t *out) {
    std::vector<uint32_t> raw(3 * (size_t)W);
    uint32_t bad = 0, n_any = 0, u0 = 0;
    const int nw = tbo::pack_raw<true, 1>(buf + start, len, raw.data(), raw.data() + W, raw.data() + 2 * W, W, u0, bad, n_any);
    for (int k = 0; k < 3; k++) {
        if (reverse) tbo::finish_reverse<1>(raw.data() + k * W, out + k * W, len, u0, nw, W, k < 2);
        else tbo::finish_forward<1>(raw.data() + k * W, out + k * W, len, u0, W);
    }
    return (bad ? 1 : 0) | (n_any ? 2 : 0);
}

// CapLine must never undercut cap_of: returns the number of overlap lengths (0..MAX_LEN) where it does, for both limit
// formulas (loop 1: ratio * ov + 20; loop 2: 1.2 * (ratio * margin * ov) + 1 + 20) in the reference's float evaluation order
int tbo_host_check_cap_line(float ratio, float margin) {
    std::vector<float> T(tbo::MAX_LEN + 2);
    T[0] = 0.0f;
    for (size_t c = 1; c < T.size(); c++) {
        volatile float s = T[c - 1] + 0.95f;
        T[c] = s;
    }
    int bad = 0;
    const tbo::CapLine l1 = tbo::cap_line(ratio, 20.0f), l2 = tbo::cap_line(tbo::fmul(1.2f, tbo::fmul(ratio, margin)), 21.0f);
    for (int ov = 0; ov <= tbo::MAX_LEN; ov++) {
        const float lim1 = tbo::fadd(tbo::fmul(ratio, (float)ov), 20.0f);
        const float lim2 = tbo::fadd(tbo::fadd(tbo::fmul(1.2f, tbo::fmul(tbo::fmul(ratio, margin), (float)ov)), 1.0f), 20.0f);
        if (l1.of(ov) < tbo::cap_of(lim1, T.data(), (int)T.size())) bad++;
        if (l2.of(ov) < tbo::cap_of(lim2, T.data(), (int)T.size())) bad++;
    }
    return bad;
}

// ratio_cap_line must admit every count that can lower findBestRatio's running ratio: returns the number of overlap
// lengths where some c >= 1 with (T[c] + offset) / ov < ratio (the reference's float evaluation) exceeds of(ov)
int tbo_host_check_ratio_cap_line(float ratio, float offset) {
    std::vector<float> T(tbo::MAX_LEN + 2);
    T[0] = 0.0f;
    for (size_t c = 1; c < T.size(); c++) {
        volatile float s = T[c - 1] + 0.95f;
        T[c] = s;
    }
    int bad = 0;
    const tbo::CapLine l = tbo::ratio_cap_line(ratio, offset);
    for (int ov = 1; ov <= tbo::MAX_LEN; ov++) {
        int cmax = 0;
        for (int c = 1; c <= ov; c++)
            if (tbo::fdiv(tbo::fadd(T[c], offset), (float)ov) < ratio) cmax = c;
        if (l.of(ov) < cmax) bad++;
    }
    return bad;
}

// same contract as oracle/tbo_oracle.c:tbo_ora_process (quals are ignored: the expectedErrors guard is not part of the core)
void tbo_host_process(const uint8_t *bases, const int64_t *offsets, int64_t n_reads, const int32_t *lo, int32_t *hi,
                      const uint8_t *flags, int min_overlap0, int min_overlap, int min_insert0, int min_insert, float max_ratio,
                      float min_second_ratio, float margin, float offset, const uint8_t *comp, int32_t *insert_out,
                      uint8_t *ambig_out, int64_t *stats, int64_t *path_counts) {
    std::vector<float> T(tbo::MAX_LEN + 2);
    T[0] = 0.0f;
    for (size_t c = 1; c < T.size(); c++) {
        volatile float s = T[c - 1] + 0.95f;
        T[c] = s;
    }
    const int n_T = (int)T.size();
    tbo::Params p;
    p.minOverlap0 = std::min(min_overlap0, min_overlap);
    p.minOverlap = min_overlap;
    p.minInsert0 = std::min(min_insert0, min_insert);
    p.minInsert = min_insert;
    p.maxRatio = max_ratio;
    p.minSecondRatio = min_second_ratio;
    p.margin = margin;
    p.offset = offset;
    p.meeFilter = 0;
    p.qualOffset = 33;
    for (int64_t u = 0; u + 1 < n_reads; u += 2) {
        const int64_t i1 = u, i2 = u + 1;
        insert_out[u / 2] = -1;
        ambig_out[u / 2] = 0;
        if (flags[i1] & 0x02) continue;
        const uint8_t *a = bases + offsets[i1] + lo[i1], *b0 = bases + offsets[i2] + lo[i2];
        const int alen = hi[i1] - lo[i1], blen = hi[i2] - lo[i2];
        const int W = tbo::plane_words(std::max(std::max(alen, blen), 16));
        p.W = W;
        std::vector<uint32_t> planes(tbo::N_PLANES_GENERAL * (size_t)W);
        tbo::Ctx<1> c;
        tbo::Cands<1> q;
        c.comp = comp;
        int best;
        bool ambig = false;
        const uint32_t what = tbo::pack_pair<false, 1>(a, alen, b0, blen, planes.data(), W, c, q);
        if (what & 1u) {  // MODE 2, then MODE 3 after a re-pack as on the device
            path_counts[2]++;
            tbo::pack_pair<true, 1>(a, alen, b0, blen, planes.data(), W, c, q);
            if (c.exact) path_counts[3]++;
            tbo::Handoff x{0.0f, -1};
            best = tbo::mate_by_overlap_ratio<true, 1, 1>(c, q, alen, blen, p, T.data(), n_T, ambig, &x);
            if (best == -3) {
                std::fill(planes.begin(), planes.end(), 0xDEADBEEFu);
                tbo::pack_pair<true, 1>(a, alen, b0, blen, planes.data(), W, c, q);
                best = tbo::mate_by_overlap_ratio<true, 2, 1>(c, q, alen, blen, p, T.data(), n_T, ambig, &x);
            }
        } else {  // MODE 0, then MODE 1 after a re-pack as on the device
            path_counts[0]++;
            tbo::Handoff x{0.0f, -1};
            best = tbo::mate_by_overlap_ratio<false, 1, 1>(c, q, alen, blen, p, T.data(), n_T, ambig, &x);
            if (best == -3) {
                path_counts[1]++;
                std::fill(planes.begin(), planes.end(), 0xDEADBEEFu);
                tbo::pack_pair<false, 1>(a, alen, b0, blen, planes.data(), W, c, q);
                best = tbo::mate_by_overlap_ratio<false, 2, 1>(c, q, alen, blen, p, T.data(), n_T, ambig, &x);
            }
        }
        if (best < p.minInsert) best = -1;
        insert_out[u / 2] = best;
        ambig_out[u / 2] = ambig ? 1 : 0;
        if (best > 0 && !ambig) {
            if (best < alen) {
                hi[i1] = lo[i1] + best;
                stats[0] += 1;
                stats[1] += alen - best;
            }
            if (best < blen) {
                hi[i2] = lo[i2] + best;
                stats[0] += 1;
                stats[1] += blen - best;
            }
        }
    }
}
}
