/*
 * kcount_oracle.c -- CPU restatement of KmerCountExact's k-mer counting loop (BASELINE.json config 5).
 *
 * TEST INFRASTRUCTURE ONLY: linked by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; the
 * product (bbtools_b200/) never calls it. PARITY UNPINNED: the reference ships no golden vectors for this
 * path and there is no JVM here, so this port is pinned only by an independent numpy restatement in
 * tests/test_kcount_oracle.py (np.unique over explicitly enumerated canonical k-mers).
 *
 * Follows, statement by statement (paths relative to /root/reference/current):
 *   kmer/KmerTableSet.java:652-716   addKmersToTable: rolling kmer/rkmer, x<0 resets len AND both k-mers,
 *                                    len>=k -> key=toValue(kmer,rkmer) -> incrementAndReturnNumCreated
 *   kmer/KmerTableSet.java:1887-1895 toValue: rcomp ? max(kmer,rkmer) : kmer   (MASK_CORE / MASK_MIDDLE are
 *                                    false by default, kmer/AbstractKmerTableSet.java:806-808)
 *   kmer/HashArray1D.java:68-89      increment with saturation at Integer.MAX_VALUE
 *   kmer/HashArray.java:577-588      fillHistogram: ca[min(count,max)]++
 *   kmer/KmerTableSet.java:530,:538  reads shorter than k contribute nothing
 *   dna/AminoAcid.java:269-285       baseToNumber / baseToComplementNumber (-1 for anything but ACGTUacgtu)
 * The hash-table layout is not part of the semantics (key -> count map).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KC_EMPTY 0xFFFFFFFFFFFFFFFFull

typedef struct kc_oracle {
    int k, rcomp;
    uint64_t *keys;
    int32_t *vals;
    uint64_t n_slots, size;
    int64_t reads_in, bases_in, kmers_in;
} kc_oracle;

static int8_t g_num[256], g_cnum[256];
static int g_init = 0;
static void kc_init_tables(void) {
    /* dna/AminoAcid.java:1289-1320 */
    if (g_init) return;
    memset(g_num, -1, sizeof g_num);
    memset(g_cnum, -1, sizeof g_cnum);
    const char *s = "ACGT";
    for (int i = 0; i < 4; i++) {
        g_num[(uint8_t)s[i]] = (int8_t)i;
        g_num[(uint8_t)(s[i] | 0x20)] = (int8_t)i;
        g_cnum[(uint8_t)s[i]] = (int8_t)(3 - i);
        g_cnum[(uint8_t)(s[i] | 0x20)] = (int8_t)(3 - i);
    }
    g_num['U'] = g_num['u'] = 3;
    g_cnum['U'] = g_cnum['u'] = 0;
    g_init = 1;
}

static uint64_t kc_mix(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

static void kc_put(kc_oracle *o, uint64_t key, int64_t incr);

static void kc_grow(kc_oracle *o) {
    uint64_t *ok = o->keys;
    int32_t *ov = o->vals;
    const uint64_t on = o->n_slots;
    o->n_slots = on * 2;
    o->keys = (uint64_t *)malloc(sizeof(uint64_t) * o->n_slots);
    o->vals = (int32_t *)calloc(o->n_slots, sizeof(int32_t));
    memset(o->keys, 0xFF, sizeof(uint64_t) * o->n_slots);
    o->size = 0;
    for (uint64_t i = 0; i < on; i++)
        if (ok[i] != KC_EMPTY) kc_put(o, ok[i], ov[i]);
    free(ok);
    free(ov);
}

/* kmer/HashArray1D.java:68-89: values[cell]+=incr; if(values[cell]<0) values[cell]=Integer.MAX_VALUE */
static void kc_put(kc_oracle *o, uint64_t key, int64_t incr) {
    if ((o->size + 1) * 10 > o->n_slots * 6) kc_grow(o);
    uint64_t s = kc_mix(key) & (o->n_slots - 1);
    while (o->keys[s] != KC_EMPTY && o->keys[s] != key) s = (s + 1) & (o->n_slots - 1);
    if (o->keys[s] == KC_EMPTY) {
        o->keys[s] = key;
        o->size++;
    }
    int64_t v = (int64_t)o->vals[s] + incr;
    if (v > 0x7FFFFFFFll) v = 0x7FFFFFFFll;
    o->vals[s] = (int32_t)v;
}

kc_oracle *kc_ora_create(int k, int rcomp) {
    if (k < 1 || k > 31) return NULL;
    kc_init_tables();
    kc_oracle *o = (kc_oracle *)calloc(1, sizeof *o);
    o->k = k;
    o->rcomp = rcomp;
    o->n_slots = 1024;
    o->keys = (uint64_t *)malloc(sizeof(uint64_t) * o->n_slots);
    o->vals = (int32_t *)calloc(o->n_slots, sizeof(int32_t));
    memset(o->keys, 0xFF, sizeof(uint64_t) * o->n_slots);
    return o;
}

void kc_ora_destroy(kc_oracle *o) {
    if (!o) return;
    free(o->keys);
    free(o->vals);
    free(o);
}

/* kmer/KmerTableSet.java:652-716 for every read of the batch */
void kc_ora_add_reads(kc_oracle *o, const uint8_t *bases, const int64_t *offsets, int64_t n_reads) {
    const int k = o->k;
    const int shift = 2 * k, shift2 = shift - 2;
    const uint64_t mask = (shift > 63) ? ~0ull : ~((~0ull) << shift);
    for (int64_t r = 0; r < n_reads; r++) {
        const uint8_t *b = bases + offsets[r];
        const int64_t L = offsets[r + 1] - offsets[r];
        o->reads_in++;
        o->bases_in += L;
        if (L < k) continue; /* :530/:538 and the bases.length<k guard of :667 */
        uint64_t kmer = 0, rkmer = 0;
        int len = 0;
        for (int64_t i = 0; i < L; i++) {
            const int64_t x = g_num[b[i]], x2 = g_cnum[b[i]];
            kmer = ((kmer << 2) | (uint64_t)x) & mask;
            rkmer = ((rkmer >> 2) | ((uint64_t)x2 << shift2)) & mask;
            if (x < 0) {
                len = 0;
                kmer = rkmer = 0;
            } else {
                len++;
            }
            if (len >= k) {
                o->kmers_in++;
                const uint64_t key = o->rcomp ? (kmer > rkmer ? kmer : rkmer) : kmer;
                kc_put(o, key, 1);
            }
        }
    }
}

/* add pre-counted entries (the multi-GPU exchange's merge step; counts saturate like increment()) */
void kc_ora_merge(kc_oracle *o, const uint64_t *keys, const int32_t *counts, int64_t n) {
    for (int64_t i = 0; i < n; i++) kc_put(o, keys[i], counts[i]);
}

/* v = {reads_in, bases_in, kmers_in, unique_kmers} */
void kc_ora_stats(const kc_oracle *o, int64_t *v) {
    v[0] = o->reads_in;
    v[1] = o->bases_in;
    v[2] = o->kmers_in;
    v[3] = (int64_t)o->size;
}

/* kmer/HashArray.java:577-588: hist[min(count,histmax)]++ , hist has histmax+1 entries */
void kc_ora_khist(const kc_oracle *o, int32_t histmax, int64_t *hist) {
    memset(hist, 0, sizeof(int64_t) * ((size_t)histmax + 1));
    for (uint64_t i = 0; i < o->n_slots; i++)
        if (o->keys[i] != KC_EMPTY) hist[o->vals[i] < histmax ? o->vals[i] : histmax]++;
}

/* all (key,count) with mincount<=count<=maxcount, unordered; returns how many (writes at most cap) */
int64_t kc_ora_dump(const kc_oracle *o, int32_t mincount, int32_t maxcount, uint64_t *keys, int32_t *counts, int64_t cap) {
    int64_t n = 0;
    for (uint64_t i = 0; i < o->n_slots; i++) {
        if (o->keys[i] == KC_EMPTY || o->vals[i] < mincount || o->vals[i] > maxcount) continue;
        if (n < cap) {
            keys[n] = o->keys[i];
            counts[n] = o->vals[i];
        }
        n++;
    }
    return n;
}
