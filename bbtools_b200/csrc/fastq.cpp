// fastq.cpp -- include/fastq_b200.h: native FASTQ record index / gather / format (host only).
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/bbduk_b200.h"
#include "../../include/fastq_b200.h"

namespace {

template <class F>
void parallel_for(int threads, int64_t n, F fn, int64_t min_chunk = 4096) {
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n / min_chunk + 1));
    if (threads > n && n > 0) threads = (int)n;
    if (threads == 1) {
        fn(0, n, 0);
        return;
    }
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; t++) ts.emplace_back([=] { fn(n * t / threads, n * (t + 1) / threads, t); });
    for (auto &t : ts) t.join();
}

inline int64_t line_end(const uint8_t *text, int64_t pos, int64_t n) {  // index of '\n' or n
    const void *p = memchr(text + pos, '\n', (size_t)(n - pos));
    return p ? (const uint8_t *)p - text : n;
}

// one record starting at pos; returns the position after it, or -1 if incomplete, -2 if malformed
int64_t one_record(const uint8_t *text, int64_t pos, int64_t n, bool final, int64_t *r4) {
    if (text[pos] != '@') return -2;
    int64_t e[4], p = pos;
    for (int l = 0; l < 4; l++) {
        if (p > n) return -1;
        if (p == n) {
            if (!(final && l == 3)) return -1;  // only an empty last quality line may be missing entirely
            e[l] = n;
            break;
        }
        e[l] = line_end(text, p, n);
        if (e[l] == n && !(final && l == 3)) return -1;
        p = e[l] + 1;
    }
    const int64_t s0 = e[0] + 1, q0 = e[2] + 1;
    if (text[e[1] + 1] != '+') return -2;
    int64_t slen = e[1] - s0, qlen = e[3] - q0;
    if (slen > 0 && text[s0 + slen - 1] == '\r') slen--;
    if (qlen > 0 && text[q0 + qlen - 1] == '\r') qlen--;
    if (qlen != slen) return -2;
    r4[0] = pos;
    r4[1] = s0;
    r4[2] = slen;
    r4[3] = q0;
    return std::min(n, e[3] + 1);
}

}  // namespace

extern "C" {

int fastq_b200_index(const uint8_t *text, int64_t n, int32_t final, int64_t max_records, int64_t stride, int64_t first,
                     int64_t *rec, int64_t *n_records, int64_t *consumed, int32_t threads) {
    if (!text || !n_records || !consumed || n < 0 || stride < 1) return 1;
    *n_records = 0;
    *consumed = 0;
    if (n == 0 || max_records <= 0) return 0;
    // pass 1 (parallel): newlines per block -> the line number every block starts at
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::max(threads, 1), n / (1 << 20) + 1));
    std::vector<int64_t> nl(T + 1, 0);
    parallel_for(T, T, [&](int64_t a, int64_t b, int) {
        for (int64_t t = a; t < b; t++) {
            const int64_t p0 = n * t / T, p1 = n * (t + 1) / T;
            int64_t c = 0;
            const uint8_t *p = text + p0, *e = text + p1;
            while (p < e && (p = (const uint8_t *)memchr(p, '\n', (size_t)(e - p)))) {
                c++;
                p++;
            }
            nl[t + 1] = c;
        }
    }, 1);
    for (int t = 0; t < T; t++) nl[t + 1] += nl[t];
    const int64_t total_lines = nl[T] + ((final && text[n - 1] != '\n') ? 1 : 0);
    const int64_t n_rec = std::min(max_records, total_lines / 4);
    if (n_rec == 0) return 0;
    if (!rec) {  // count only
        *n_records = n_rec;
        return 0;
    }
    // pass 2 (parallel): block t parses the records that START in it; it first skips to the next line whose number is
    // a multiple of 4
    std::vector<int> err(T, 0);
    std::vector<int64_t> last_end(T, 0);
    parallel_for(T, T, [&](int64_t a, int64_t b, int) {
        for (int64_t t = a; t < b; t++) {
            const int64_t p0 = n * t / T, p1 = n * (t + 1) / T;
            int64_t line = nl[t], pos = p0;
            if (p0 > 0 && text[p0 - 1] != '\n') {  // inside a line: it belongs to the previous block
                pos = line_end(text, p0, n) + 1;
                line++;
            }
            while (pos < p1 && (line & 3)) {
                pos = line_end(text, pos, n) + 1;
                line++;
            }
            while (pos < p1 && pos < n) {
                const int64_t r = line >> 2;
                if (r >= n_rec) break;
                int64_t r4[4];
                const int64_t nx = one_record(text, pos, n, final != 0, r4);
                if (nx == -1) break;
                if (nx < 0) {
                    err[t] = 1;
                    break;
                }
                memcpy(rec + 4 * (first + r * stride), r4, sizeof r4);
                last_end[t] = nx;
                pos = nx;
                line += 4;
            }
        }
    }, 1);
    for (int t = 0; t < T; t++)
        if (err[t]) return 2;
    int64_t end = 0;
    for (int t = 0; t < T; t++) end = std::max(end, last_end[t]);
    *n_records = n_rec;
    *consumed = end;
    return 0;
}

int fastq_b200_gather(const uint8_t *text1, const uint8_t *text2, const int64_t *rec, int64_t n_reads, uint8_t *bases,
                      int64_t *offsets, int32_t threads) {
    if (!rec || !offsets || n_reads < 0 || (n_reads > 0 && !text1)) return 1;
    offsets[0] = 0;
    for (int64_t i = 0; i < n_reads; i++) offsets[i + 1] = offsets[i] + rec[4 * i + 2];
    if (!bases) return 0;
    parallel_for(std::max(threads, 1), n_reads, [&](int64_t a, int64_t b, int) {
        for (int64_t i = a; i < b; i++) {
            const uint8_t *src = (text2 && (i & 1)) ? text2 : text1;
            memcpy(bases + offsets[i], src + rec[4 * i + 1], (size_t)rec[4 * i + 2]);
        }
    });
    return 0;
}

int fastq_b200_format(const uint8_t *text1, const uint8_t *text2, const int64_t *rec, int64_t n_reads, int32_t per,
                      const int32_t *lo, const int32_t *hi, const uint8_t *flags, int32_t want_removed, int32_t mate_sel,
                      int32_t trim_removed, uint8_t *out, int64_t out_cap, int64_t *out_len, int32_t threads) {
    if (!rec || !out_len || !flags || !lo || !hi || (per != 1 && per != 2) || n_reads < 0 || (n_reads % per)) return 1;
    const int64_t n_units = n_reads / per;
    auto text_of = [&](int64_t i) { return (text2 && (i & 1)) ? text2 : text1; };
    auto hdr_len = [&](int64_t i) {  // header line without the line terminator
        const uint8_t *t = text_of(i);
        int64_t len = rec[4 * i + 1] - 1 - rec[4 * i];
        if (len > 0 && t[rec[4 * i] + len - 1] == '\r') len--;
        return len;
    };
    auto span = [&](int64_t i, bool removed, int64_t &a, int64_t &b) {
        a = 0;
        b = rec[4 * i + 2];
        if (!removed || trim_removed) {
            a = lo[i];
            b = hi[i];
        }
        if (a < 0) a = 0;
        if (b > rec[4 * i + 2]) b = rec[4 * i + 2];
        if (b < a) b = a;
    };
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::max(threads, 1), n_units / 4096 + 1));
    std::vector<int64_t> part(T + 1, 0);
    auto selected = [&](int64_t u, int q) {
        const bool removed = (flags[u * per] & BBDUK_F_REMOVED) != 0;
        if (removed != (want_removed != 0)) return false;
        return mate_sel == 0 || (per == 2 && mate_sel == q + 1) || (per == 1 && mate_sel == 1);
    };
    auto size_of = [&](int64_t i, bool removed) {
        int64_t a, b;
        span(i, removed, a, b);
        return hdr_len(i) + 1 + (b - a) + 3 + (b - a) + 1;
    };
    parallel_for(T, T, [&](int64_t ta, int64_t tb, int) {
        for (int64_t t = ta; t < tb; t++) {
            int64_t sz = 0;
            for (int64_t u = n_units * t / T; u < n_units * (t + 1) / T; u++) {
                const bool removed = (flags[u * per] & BBDUK_F_REMOVED) != 0;
                for (int q = 0; q < per; q++)
                    if (selected(u, q)) sz += size_of(u * per + q, removed);
            }
            part[t + 1] = sz;
        }
    }, 1);
    for (int t = 0; t < T; t++) part[t + 1] += part[t];
    *out_len = part[T];
    if (!out) return 0;
    if (out_cap < part[T]) return 3;
    parallel_for(T, T, [&](int64_t ta, int64_t tb, int) {
        for (int64_t t = ta; t < tb; t++) {
            uint8_t *w = out + part[t];
            for (int64_t u = n_units * t / T; u < n_units * (t + 1) / T; u++) {
                const bool removed = (flags[u * per] & BBDUK_F_REMOVED) != 0;
                for (int q = 0; q < per; q++) {
                    if (!selected(u, q)) continue;
                    const int64_t i = u * per + q;
                    const uint8_t *tx = text_of(i);
                    int64_t a, b;
                    span(i, removed, a, b);
                    const int64_t hl = hdr_len(i);
                    memcpy(w, tx + rec[4 * i], (size_t)hl);
                    w += hl;
                    *w++ = '\n';
                    memcpy(w, tx + rec[4 * i + 1] + a, (size_t)(b - a));
                    w += b - a;
                    *w++ = '\n';
                    *w++ = '+';
                    *w++ = '\n';
                    memcpy(w, tx + rec[4 * i + 3] + a, (size_t)(b - a));
                    w += b - a;
                    *w++ = '\n';
                }
            }
        }
    }, 1);
    return 0;
}

}  // extern "C"
