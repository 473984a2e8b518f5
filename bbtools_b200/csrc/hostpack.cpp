// hostpack.cpp -- see hostpack.h. AVX2 path selected at run time on x86-64; portable scalar path otherwise.
#include "hostpack.h"

#include <cstdlib>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

struct Tables {
    uint8_t code[256], valid[256];
    Tables() {
        memset(code, 0, sizeof code);
        memset(valid, 0, sizeof valid);
        const char *s = "ACGT";
        for (int i = 0; i < 4; i++) {
            code[(uint8_t)s[i]] = code[(uint8_t)(s[i] | 0x20)] = (uint8_t)i;
            valid[(uint8_t)s[i]] = valid[(uint8_t)(s[i] | 0x20)] = 1;
        }
        code['U'] = code['u'] = 3;
        valid['U'] = valid['u'] = 1;
    }
};
const Tables g_tab;

void pack_scalar(const uint8_t *b, int64_t n, int64_t g0, int64_t g1, uint32_t *F, uint16_t *D) {
    for (int64_t g = g0; g < g1; g++) {
        uint32_t f = 0, d = 0;
        const int64_t base = g * 16;
        const int m = (int)((n - base) < 16 ? (n - base) : 16);
        for (int j = 0; j < m; j++) {
            const uint8_t c = b[base + j];
            f |= (uint32_t)g_tab.code[c] << (30 - 2 * j);
            d |= (uint32_t)g_tab.valid[c] << (15 - j);
        }
        // undefined bases keep code 0 (code[] is 0 for them already)
        F[g] = f;
        D[g] = (uint16_t)d;
    }
}

#if defined(__x86_64__)
// 32 bases -> two F words and two D halfwords (in one uint32)
__attribute__((target("avx2"))) static inline void pack32(const uint8_t *src, uint32_t &f0, uint32_t &f1, uint32_t &dd) {
    const __m256i three = _mm256_set1_epi8(3);
    const __m256i lower = _mm256_set1_epi8(0x20);
    const char X = (char)0xFF;
    // expected lower-case byte by low nibble: a(1) c(3) t(4) u(5) g(7)
    const __m256i lut = _mm256_setr_epi8(X, 0x61, X, 0x63, 0x74, 0x75, X, 0x67, X, X, X, X, X, X, X, X,
                                         X, 0x61, X, 0x63, 0x74, 0x75, X, 0x67, X, X, X, X, X, X, X, X);
    const __m256i m41 = _mm256_set1_epi16(0x0104);      // bytes (4,1): 4*c0 + c1
    const __m256i m161 = _mm256_set1_epi32(0x00010010);  // words (16,1): 16*x0 + x1
    const __m256i gather = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i rev = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0,
                                         15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src));
    // codes = ((c>>1) ^ (c>>2)) & 3 per byte (bits shifted in from the neighbour byte are masked away)
    __m256i c = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(v, 1), _mm256_srli_epi16(v, 2)), three);
    // defined <=> (c|0x20) is one of a c g t u and c < 128 (pshufb zeroes lanes whose index has bit 7 set)
    const __m256i y = _mm256_or_si256(v, lower);
    const __m256i ok = _mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, y), y);
    c = _mm256_and_si256(c, ok);  // undefined -> code 0
    const __m256i p4 = _mm256_madd_epi16(_mm256_maddubs_epi16(c, m41), m161);  // one byte per 4 bases in each dword
    const __m256i w = _mm256_shuffle_epi8(p4, gather);
    f0 = (uint32_t)_mm256_extract_epi32(w, 0);
    f1 = (uint32_t)_mm256_extract_epi32(w, 4);
    dd = (uint32_t)_mm256_movemask_epi8(_mm256_shuffle_epi8(ok, rev));  // bit 15-b = base b, per 16-bit half
}

__attribute__((target("avx2"))) void pack_avx2(const uint8_t *b, int64_t n, int64_t g0, int64_t g1, uint32_t *F, uint16_t *D) {
    int64_t g = g0;
    const int64_t full = n / 16;  // groups whose 16 bytes are all inside the batch
    // peel to a 64-byte boundary of F (16 groups) so that whole cache lines can be written without reading them
    while (g + 1 < g1 && g + 1 < full && ((reinterpret_cast<uintptr_t>(F + g) & 63) != 0 || (reinterpret_cast<uintptr_t>(D + g) & 63) != 0)) {
        uint32_t f0, f1, dd;
        pack32(b + g * 16, f0, f1, dd);
        F[g] = f0;
        F[g + 1] = f1;
        D[g] = (uint16_t)dd;
        D[g + 1] = (uint16_t)(dd >> 16);
        g += 2;
    }
    // 32 groups (512 bases) per iteration: two 64-byte lines of F and one of D, streamed past the cache
    if ((reinterpret_cast<uintptr_t>(F + g) & 63) == 0 && (reinterpret_cast<uintptr_t>(D + g) & 63) == 0) {
        for (; g + 32 <= g1 && g + 32 <= full; g += 32) {
            alignas(64) uint32_t fb[32];
            alignas(64) uint32_t db[16];
#pragma GCC unroll 4
            for (int q = 0; q < 16; q++) pack32(b + (g + 2 * q) * 16, fb[2 * q], fb[2 * q + 1], db[q]);
            for (int q = 0; q < 4; q++)
                _mm256_stream_si256(reinterpret_cast<__m256i *>(F + g) + q, _mm256_load_si256(reinterpret_cast<const __m256i *>(fb) + q));
            for (int q = 0; q < 2; q++)
                _mm256_stream_si256(reinterpret_cast<__m256i *>(D + g) + q, _mm256_load_si256(reinterpret_cast<const __m256i *>(db) + q));
        }
        _mm_sfence();
    }
    for (; g + 1 < g1 && g + 1 < full; g += 2) {
        uint32_t f0, f1, dd;
        pack32(b + g * 16, f0, f1, dd);
        F[g] = f0;
        F[g + 1] = f1;
        D[g] = (uint16_t)dd;
        D[g + 1] = (uint16_t)(dd >> 16);
    }
    if (g < g1) pack_scalar(b, n, g, g1, F, D);
}
// AVX-512 VBMI: 64 bases per step. One 128-entry byte table lookup (vpermi2b) turns ASCII into (code | 0x80 if undefined),
// two multiply-adds pack 4 codes per byte, one byte permute gathers the 16 packed bytes (4 F words, big-endian inside a word),
// and the defined bits come straight out of a byte-test mask of the per-16-byte reversed vector (bit 15-b = base b).
struct Lut512 {
    alignas(64) uint8_t lo[64], hi[64];
    Lut512() {
        for (int i = 0; i < 128; i++) {
            const uint8_t v = g_tab.valid[i] ? g_tab.code[i] : (uint8_t)0x80;
            (i < 64 ? lo[i] : hi[i - 64]) = v;
        }
    }
};
const Lut512 g_lut512;

__attribute__((target("avx512f,avx512bw,avx512vbmi,avx512vl"))) static inline void pack64(const uint8_t *src, const __m512i lutlo,
                                                                                          const __m512i luthi, __m128i &f4, uint64_t &d4) {
    const __m512i v = _mm512_loadu_si512(src);
    const __mmask64 high = _mm512_movepi8_mask(v);                          // bytes >= 128: undefined
    __m512i t = _mm512_permutex2var_epi8(lutlo, v, luthi);                  // index = low 7 bits
    t = _mm512_mask_mov_epi8(t, high, _mm512_set1_epi8((char)0x80));
    const __m512i c = _mm512_and_si512(t, _mm512_set1_epi8(3));            // undefined -> code 0 (table value 0x80)
    const __m512i p4 = _mm512_madd_epi16(_mm512_maddubs_epi16(c, _mm512_set1_epi16(0x0104)), _mm512_set1_epi32(0x00010010));
    // byte 4j of p4 = bases 4j..4j+3; F word w (16 bases) = bytes (16w+0, 16w+4, 16w+8, 16w+12) most significant first
    const __m512i gather = _mm512_castsi128_si512(_mm_setr_epi8(12, 8, 4, 0, 28, 24, 20, 16, 44, 40, 36, 32, 60, 56, 52, 48));
    f4 = _mm512_castsi512_si128(_mm512_permutexvar_epi8(gather, p4));
    const __m512i rev = _mm512_broadcast_i32x4(_mm_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0));
    d4 = ~(uint64_t)_mm512_movepi8_mask(_mm512_shuffle_epi8(t, rev));       // bit 7 set = undefined
}

// groups with an undefined base, collected while their defined bits are still in a register
struct ExcOut {
    uint64_t *out;
    int64_t cap, n;
    bool overflow;
    inline void add4(int64_t g, uint64_t d4) {  // d4 = the D values of groups g..g+3, 16 bits each
        for (int j = 0; j < 4; j++) {
            const uint64_t dj = (d4 >> (16 * j)) & 0xFFFFu;
            if (dj == 0xFFFFu) continue;
            if (n >= cap) {
                overflow = true;
                return;
            }
            out[n++] = ((uint64_t)(g + j) << 16) | dj;
        }
    }
};

__attribute__((target("avx512f,avx512bw,avx512vbmi,avx512vl"))) void pack_avx512(const uint8_t *b, int64_t n, int64_t g0, int64_t g1,
                                                                                  uint32_t *F, uint16_t *D, ExcOut *ex) {
    const __m512i lutlo = _mm512_load_si512(g_lut512.lo), luthi = _mm512_load_si512(g_lut512.hi);
    int64_t g = g0;
    const int64_t full = n / 16;
    const int64_t stop = g1 < full ? g1 : full;
    // peel to a 64-byte boundary of F and D (32 groups make one line of D and two of F)
    while (g + 4 <= stop && (((reinterpret_cast<uintptr_t>(F + g)) & 63) != 0 || ((reinterpret_cast<uintptr_t>(D + g)) & 63) != 0)) {
        __m128i f4;
        uint64_t d4;
        pack64(b + g * 16, lutlo, luthi, f4, d4);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(F + g), f4);
        memcpy(D + g, &d4, 8);
        if (ex && d4 != ~0ull) ex->add4(g, d4);
        g += 4;
    }
    if ((reinterpret_cast<uintptr_t>(F + g) & 63) == 0 && (reinterpret_cast<uintptr_t>(D + g) & 63) == 0) {
        for (; g + 32 <= stop; g += 32) {  // 512 bases: two lines of F, one of D, streamed past the cache
            __m128i f[8];
            alignas(64) uint64_t d[8];
#pragma GCC unroll 8
            for (int q = 0; q < 8; q++) pack64(b + (g + 4 * q) * 16, lutlo, luthi, f[q], d[q]);
            if (ex && (d[0] & d[1] & d[2] & d[3] & d[4] & d[5] & d[6] & d[7]) != ~0ull)
                for (int q = 0; q < 8; q++)
                    if (d[q] != ~0ull) ex->add4(g + 4 * q, d[q]);
            const __m512i fa = _mm512_inserti64x4(_mm512_castsi256_si512(_mm256_set_m128i(f[1], f[0])), _mm256_set_m128i(f[3], f[2]), 1);
            const __m512i fb = _mm512_inserti64x4(_mm512_castsi256_si512(_mm256_set_m128i(f[5], f[4])), _mm256_set_m128i(f[7], f[6]), 1);
            _mm512_stream_si512(reinterpret_cast<__m512i *>(F + g), fa);
            _mm512_stream_si512(reinterpret_cast<__m512i *>(F + g + 16), fb);
            _mm512_stream_si512(reinterpret_cast<__m512i *>(D + g), _mm512_load_si512(d));
        }
        _mm_sfence();
    }
    for (; g + 4 <= stop; g += 4) {
        __m128i f4;
        uint64_t d4;
        pack64(b + g * 16, lutlo, luthi, f4, d4);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(F + g), f4);
        memcpy(D + g, &d4, 8);
        if (ex && d4 != ~0ull) ex->add4(g, d4);
    }
    if (g < g1) {
        pack_scalar(b, n, g, g1, F, D);
        if (ex)
            for (; g < g1; g++)
                if (D[g] != 0xFFFFu) ex->add4(g, 0xFFFFFFFFFFFF0000ull | D[g]);
    }
}
#endif

}  // namespace

static void pack_range_impl(const uint8_t *bases, int64_t n, int64_t g0, int64_t g1, uint32_t *F, uint16_t *D, ExcOut *ex) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    static const bool have_vbmi = __builtin_cpu_supports("avx512vbmi") && __builtin_cpu_supports("avx512bw") && !getenv("BBDUK_B200_NO_AVX512");
    if (have_vbmi) {
        pack_avx512(bases, n, g0, g1, F, D, ex);
        return;
    }
#endif
    if (ex) {  // the other packers do not collect: list from the array they wrote
        pack_range_impl(bases, n, g0, g1, F, D, nullptr);
        const int64_t c = list_undefined_groups(D, g0, g1, ex->out + ex->n, ex->cap - ex->n);
        if (c < 0) ex->overflow = true;
        else ex->n += c;
        return;
    }
#if defined(__x86_64__)
    if (have_avx2) {
        pack_avx2(bases, n, g0, g1, F, D);
        return;
    }
#endif
    pack_scalar(bases, n, g0, g1, F, D);
}

void pack_bases_range(const uint8_t *bases, int64_t n, int64_t g0, int64_t g1, uint32_t *F, uint16_t *D) {
    pack_range_impl(bases, n, g0, g1, F, D, nullptr);
}

int64_t pack_bases_range_listing(const uint8_t *bases, int64_t n, int64_t g0, int64_t g1, uint32_t *F, uint16_t *D, uint64_t *exc,
                                 int64_t cap) {
    ExcOut ex{exc, cap, 0, false};
    pack_range_impl(bases, n, g0, g1, F, D, &ex);
    return ex.overflow ? -1 : ex.n;
}

void pack_bases(const uint8_t *bases, int64_t n, uint32_t *F, uint16_t *D) { pack_bases_range(bases, n, 0, (n + 15) / 16, F, D); }

namespace {
// every byte of b[i0, i1) is one of A C G T N (upper case): then F + D lose nothing (codes spell A C G T, undefined = N)
bool plain_scalar(const uint8_t *b, int64_t i0, int64_t i1) {
    for (int64_t i = i0; i < i1; i++) {
        const uint8_t c = b[i];
        if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N')) return false;
    }
    return true;
}
#if defined(__x86_64__)
__attribute__((target("avx512f,avx512bw"))) bool plain_avx512(const uint8_t *b, int64_t i0, int64_t i1) {
    __mmask64 bad = 0;
    int64_t i = i0;
    const __m512i A = _mm512_set1_epi8('A'), Cc = _mm512_set1_epi8('C'), G = _mm512_set1_epi8('G'), T = _mm512_set1_epi8('T'),
                  N = _mm512_set1_epi8('N');
    for (; i + 64 <= i1; i += 64) {
        const __m512i v = _mm512_loadu_si512(b + i);
        const __mmask64 ok = _mm512_cmpeq_epi8_mask(v, A) | _mm512_cmpeq_epi8_mask(v, Cc) | _mm512_cmpeq_epi8_mask(v, G) |
                             _mm512_cmpeq_epi8_mask(v, T) | _mm512_cmpeq_epi8_mask(v, N);
        bad |= ~ok;
    }
    return bad == 0 && plain_scalar(b, i, i1);
}
__attribute__((target("avx2"))) bool plain_avx2(const uint8_t *b, int64_t i0, int64_t i1) {
    __m256i all = _mm256_set1_epi8((char)0xFF);
    int64_t i = i0;
    const __m256i A = _mm256_set1_epi8('A'), Cc = _mm256_set1_epi8('C'), G = _mm256_set1_epi8('G'), T = _mm256_set1_epi8('T'),
                  N = _mm256_set1_epi8('N');
    for (; i + 32 <= i1; i += 32) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(b + i));
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, A), _mm256_cmpeq_epi8(v, Cc)),
                                           _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, G), _mm256_cmpeq_epi8(v, T)),
                                                           _mm256_cmpeq_epi8(v, N)));
        all = _mm256_and_si256(all, ok);
    }
    return _mm256_movemask_epi8(all) == -1 && plain_scalar(b, i, i1);
}
#endif
bool plain_range(const uint8_t *b, int64_t i0, int64_t i1) {
#if defined(__x86_64__)
    static const bool have_512 = __builtin_cpu_supports("avx512bw") && !getenv("BBDUK_B200_NO_AVX512");
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_512) return plain_avx512(b, i0, i1);
    if (have_avx2) return plain_avx2(b, i0, i1);
#endif
    return plain_scalar(b, i0, i1);
}
}  // namespace

bool pack_bases_range_plain(const uint8_t *bases, int64_t n, int64_t g0, int64_t g1, uint32_t *F, uint16_t *D) {
    // 64 KB of bases at a time: the check re-reads what the packer has just pulled into the cache
    bool plain = true;
    for (int64_t g = g0; g < g1; g += 4096) {
        const int64_t ge = g + 4096 < g1 ? g + 4096 : g1;
        pack_bases_range(bases, n, g, ge, F, D);
        if (plain) plain = plain_range(bases, 16 * g, 16 * ge < n ? 16 * ge : n);
    }
    return plain;
}

int64_t list_undefined_groups(const uint16_t *D, int64_t g0, int64_t g1, uint64_t *out, int64_t cap) {
    int64_t n = 0, g = g0;
    auto one = [&](int64_t i) -> bool {
        if (D[i] == 0xFFFFu) return true;
        if (n >= cap) return false;
        out[n++] = ((uint64_t)i << 16) | D[i];
        return true;
    };
    for (; g < g1 && (reinterpret_cast<uintptr_t>(D + g) & 7); g++)
        if (!one(g)) return -1;
    for (; g + 4 <= g1; g += 4) {
        uint64_t w;
        memcpy(&w, D + g, 8);
        if (w == ~0ull) continue;
        for (int j = 0; j < 4; j++)
            if (!one(g + j)) return -1;
    }
    for (; g < g1; g++)
        if (!one(g)) return -1;
    return n;
}

HostPool::HostPool(int n_threads) {
    for (int i = 1; i < n_threads; i++) workers.emplace_back([this, i] { loop(i); });
}

HostPool::~HostPool() {
    {
        std::lock_guard<std::mutex> g(mu);
        stop = true;
        epoch++;
    }
    cv_go.notify_all();
    for (auto &t : workers) t.join();
}

void HostPool::loop(int idx) {
    uint64_t seen = 0;
    while (true) {
        const std::function<void(int, int)> *fn = nullptr;
        {
            std::unique_lock<std::mutex> g(mu);
            cv_go.wait(g, [&] { return epoch != seen; });
            seen = epoch;
            if (stop) return;
            fn = job;
        }
        (*fn)(idx, size());
        {
            std::lock_guard<std::mutex> g(mu);
            if (--pending == 0) cv_done.notify_all();
        }
    }
}

void HostPool::run(const std::function<void(int, int)> &fn) {
    {
        std::lock_guard<std::mutex> g(mu);
        job = &fn;
        pending = (int)workers.size();
        epoch++;
    }
    cv_go.notify_all();
    fn(0, size());
    std::unique_lock<std::mutex> g(mu);
    cv_done.wait(g, [&] { return pending == 0; });
}
