/*
 * seal_oracle.c -- CPU restatement of Seal's k-mer loader and matching block.
 *
 * TEST INFRASTRUCTURE ONLY: linked by tests/ (and tools/); the product (bbtools_b200/) never calls it.
 * PARITY UNPINNED: the reference ships no golden vectors for this path and there is no JVM here, so this port
 * is pinned only by an independent Python restatement (oracle/seal.py: dictionaries of explicitly enumerated
 * k-mers, tests/test_seal_oracle.py) and hand-computed cases.
 *
 * Follows, statement by statement (paths relative to /root/reference/current):
 *   jgi/Seal.java:486-571     constructor derivations: hammingDistance, forbidNs, midMaskLen, middleMask
 *   jgi/Seal.java:1760-1829   LoadThread.addToMap(Read, skip): rolling kmer/rkmer, x<0 resets len and rkmer,
 *                             len>=k (and len%skip==0) -> addToMap(kmer, rkmer, ...)
 *   jgi/Seal.java:1839-1871   addToMap(kmer...): hdist==0 -> speed filter + map.set; else mutate
 *   jgi/Seal.java:1890-1946   mutate: set own key, then all substitutions recursively (edist=0 only here)
 *   kmer/HashArray.java:187-218 + IntList3 ASCENDING: map.set appends an id unless already there; ids arrive
 *                             in increasing order, so a key's value list is its ascending set of ids, and the
 *                             return value is 1 only for a new key (storedKmers = distinct keys)
 *   jgi/Seal.java:2864-2907   findBestMatch(Read, sets, int[] hits, IntList idList) (default count array)
 *   jgi/Seal.java:2790-2804   getValuesInner: qskip, toValue, passesSpeed, lookup
 *   jgi/Seal.java:2654-2667   condenseLoose(int[], ...): counts in first-seen order, max
 *   jgi/Seal.java:2697-2708   filterTopScaffolds_withClearzone
 *   jgi/Seal.java:2186-2276   the matching block of ProcessThread.run (kpt and not kpt)
 *   jgi/Seal.java:2386-2453   assignTogether; :2462-2606 assignIndependently
 *   stream/Read.java:1665-1683 numValidKmers / numValidPairKmers; jgi/Seal.java:1198-1207 numKmers
 *   dna/AminoAcid.java:269-285 baseToNumber(0) / baseToComplementNumber(0)
 * The hash-table layout is not part of the semantics: the table here is a sorted array of (key, id) entries behind an
 * open-addressed index of the distinct keys.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/seal_b200.h"

typedef struct {
    uint64_t key;
    int32_t id;
} sl_pair;

typedef struct sl_oracle {
    seal_cfg c;
    int k, hammingDistance, forbidNs, maskMiddle, midMaskLen, rcomp;
    uint64_t middleMask;
    /* reference */
    uint8_t *ref;
    int64_t ref_len, ref_cap;
    int64_t *seq_off;
    int32_t n_seqs, seq_cap;
    /* table */
    sl_pair *pairs;
    int64_t n_pairs, cap_pairs;
    int64_t *index;      /* open-addressed hash over the distinct keys: first entry of the key in pairs[], -1 = empty */
    uint64_t index_mask;
    int64_t stored, ref_kmers;
    /* per-scaffold counters, index = id */
    int64_t *s_reads, *s_bases, *s_frags, *s_ambig;
    /* work */
    int32_t *countArray;
    char err[256];
} sl_oracle;

static int8_t s_num[256], s_cnum[256], s_num0[256], s_cnum0[256];
static int s_init = 0;
static void sl_init_tables(void) {
    if (s_init) return;
    memset(s_num, -1, sizeof s_num);
    memset(s_cnum, -1, sizeof s_cnum);
    const char *s = "ACGT";
    for (int i = 0; i < 4; i++) {
        s_num[(uint8_t)s[i]] = s_num[(uint8_t)(s[i] | 0x20)] = (int8_t)i;
        s_cnum[(uint8_t)s[i]] = s_cnum[(uint8_t)(s[i] | 0x20)] = (int8_t)(3 - i);
    }
    s_num['U'] = s_num['u'] = 3;
    s_cnum['U'] = s_cnum['u'] = 0;
    for (int i = 0; i < 256; i++) {
        s_num0[i] = s_num[i] < 0 ? 0 : s_num[i];
        s_cnum0[i] = s_cnum[i] < 0 ? 0 : s_cnum[i];
    }
    s_init = 1;
}

/* dna/AminoAcid.java:585-603 */
static uint64_t sl_rcomp(uint64_t kmer, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; i++) {
        r = (r << 2) | (3 - (kmer & 3));
        kmer >>= 2;
    }
    return r;
}

static uint64_t sl_mix(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

static uint64_t sl_to_value(const sl_oracle *o, uint64_t kmer, uint64_t rkmer, uint64_t lengthMask) {
    uint64_t v = o->rcomp ? (kmer > rkmer ? kmer : rkmer) : kmer;
    return (v & o->middleMask) | lengthMask;
}
static int sl_passes_speed(const sl_oracle *o, uint64_t key) {
    return o->c.speed < 1 || (int)((key & 0x7FFFFFFFFFFFFFFFull) % 17) >= o->c.speed;
}

void *sl_ora_create(const seal_cfg *cfg) {
    sl_init_tables();
    if (!cfg || cfg->struct_size != (int32_t)sizeof(seal_cfg)) return NULL;
    if (cfg->k < 1 || cfg->k > 31 || cfg->hdist < 0 || cfg->hdist > 2 || cfg->min_kmer_hits < 1) return NULL;
    sl_oracle *o = (sl_oracle *)calloc(1, sizeof *o);
    o->c = *cfg;
    o->k = cfg->k;
    o->rcomp = cfg->rcomp != 0;
    o->hammingDistance = cfg->hdist;
    o->forbidNs = (cfg->forbid_ns || o->hammingDistance < 1);
    o->maskMiddle = cfg->mask_middle || cfg->mid_mask_len > 0; /* :255-260 */
    if (o->maskMiddle) o->midMaskLen = cfg->mid_mask_len > 0 ? cfg->mid_mask_len : 2 - (o->k & 1);
    else o->midMaskLen = 0;
    if (o->maskMiddle) {
        if (!(o->k > o->midMaskLen + 1)) {
            free(o);
            return NULL;
        }
        int bits = o->midMaskLen * 2;
        int shift = ((o->k - o->midMaskLen) / 2) * 2;
        o->middleMask = ~((~((~0ull) << bits)) << shift);
    } else o->middleMask = ~0ull;
    o->seq_cap = 16;
    o->seq_off = (int64_t *)calloc(o->seq_cap + 1, sizeof(int64_t));
    return o;
}

void sl_ora_destroy(void *h) {
    sl_oracle *o = (sl_oracle *)h;
    if (!o) return;
    free(o->ref);
    free(o->seq_off);
    free(o->pairs);
    free(o->index);
    free(o->s_reads);
    free(o->s_bases);
    free(o->s_frags);
    free(o->s_ambig);
    free(o->countArray);
    free(o);
}

int sl_ora_add_ref(void *h, const uint8_t *bases, const int64_t *offsets, int32_t n) {
    sl_oracle *o = (sl_oracle *)h;
    for (int32_t s = 0; s < n; s++) {
        const int64_t L = offsets[s + 1] - offsets[s];
        if (o->ref_len + L > o->ref_cap) {
            o->ref_cap = (o->ref_len + L) * 2 + 64;
            o->ref = (uint8_t *)realloc(o->ref, (size_t)o->ref_cap);
        }
        memcpy(o->ref + o->ref_len, bases + offsets[s], (size_t)L);
        o->ref_len += L;
        if (o->n_seqs + 1 > o->seq_cap) {
            o->seq_cap *= 2;
            o->seq_off = (int64_t *)realloc(o->seq_off, (size_t)(o->seq_cap + 1) * sizeof(int64_t));
        }
        o->n_seqs++;
        o->seq_off[o->n_seqs] = o->ref_len;
    }
    return 0;
}

static void sl_set(sl_oracle *o, uint64_t key, int32_t id) {
    if (o->n_pairs == o->cap_pairs) {
        o->cap_pairs = o->cap_pairs ? o->cap_pairs * 2 : 4096;
        o->pairs = (sl_pair *)realloc(o->pairs, (size_t)o->cap_pairs * sizeof(sl_pair));
    }
    o->pairs[o->n_pairs].key = key;
    o->pairs[o->n_pairs].id = id;
    o->n_pairs++;
}

/* jgi/Seal.java:1890-1946, editDistance == 0 */
static void sl_mutate(sl_oracle *o, uint64_t kmer, uint64_t rkmer, int len, int id, int dist) {
    const uint64_t key = sl_to_value(o, kmer, rkmer, 1ull << (2 * len));
    sl_set(o, key, id);
    if (dist > 0) {
        const int dist2 = dist - 1;
        for (int j = 0; j < 4; j++) {
            for (int i = 0; i < len; i++) {
                const uint64_t temp = (kmer & ~(3ull << (2 * i))) | ((uint64_t)j << (2 * i));
                if (temp != kmer) sl_mutate(o, temp, sl_rcomp(temp, len), len, id, dist2);
            }
        }
    }
}

static void sl_add_kmer(sl_oracle *o, uint64_t kmer, uint64_t rkmer, int len, int id) {
    if (o->hammingDistance == 0) {
        const uint64_t key = sl_to_value(o, kmer, rkmer, 1ull << (2 * len));
        if (!sl_passes_speed(o, key)) return;
        sl_set(o, key, id);
    } else sl_mutate(o, kmer, rkmer, len, id, o->hammingDistance);
}

static int sl_cmp(const void *a, const void *b) {
    const sl_pair *x = (const sl_pair *)a, *y = (const sl_pair *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return (x->id > y->id) - (x->id < y->id);
}

/* v[3] = {storedKmers, (key,id) entries, refKmers} */
int sl_ora_finalize(void *h, int64_t *v) {
    sl_oracle *o = (sl_oracle *)h;
    const int k = o->k, shift = 2 * k, shift2 = shift - 2;
    const uint64_t mask = shift > 63 ? ~0ull : ~((~0ull) << shift);
    const int skip = o->c.rskip > 0 ? o->c.rskip : 0; /* refSkip=max(0,refSkip), :489 */
    o->n_pairs = 0;
    o->ref_kmers = 0;
    for (int32_t s = 0; s < o->n_seqs; s++) {
        const uint8_t *bases = o->ref + o->seq_off[s];
        const int64_t L = o->seq_off[s + 1] - o->seq_off[s];
        const int id = s + 1; /* a fake first scaffold takes id 0, :129-131 */
        if (L < k) continue;
        uint64_t kmer = 0, rkmer = 0;
        int64_t len = 0;
        for (int64_t i = 0; i < L; i++) {
            const uint8_t b = bases[i];
            const int64_t x = s_num[b], x2 = s_cnum[b];
            kmer = ((kmer << 2) | (uint64_t)x) & mask;
            rkmer = ((rkmer >> 2) | ((uint64_t)x2 << shift2)) & mask;
            if (x < 0) {
                len = 0;
                rkmer = 0;
            } else len++;
            if (len >= k) {
                o->ref_kmers++;
                if (skip > 1 && len % skip != 0) continue;
                sl_add_kmer(o, kmer, rkmer, k, id);
            }
        }
    }
    qsort(o->pairs, (size_t)o->n_pairs, sizeof(sl_pair), sl_cmp);
    int64_t w = 0, stored = 0;
    for (int64_t i = 0; i < o->n_pairs; i++) {
        if (i > 0 && o->pairs[i].key == o->pairs[w - 1].key && o->pairs[i].id == o->pairs[w - 1].id) continue;
        if (w == 0 || o->pairs[i].key != o->pairs[w - 1].key) stored++;
        o->pairs[w++] = o->pairs[i];
    }
    o->n_pairs = w;
    o->stored = stored;
    /* index: the map is a hash table in the reference too (kmer/HashArray.java); layout is not part of the semantics */
    free(o->index);
    uint64_t slots = 1024;
    while (slots < (uint64_t)stored * 2 + 2) slots <<= 1;
    o->index_mask = slots - 1;
    o->index = (int64_t *)malloc(slots * sizeof(int64_t));
    for (uint64_t i = 0; i < slots; i++) o->index[i] = -1;
    for (int64_t i = 0; i < o->n_pairs; i++) {
        if (i > 0 && o->pairs[i].key == o->pairs[i - 1].key) continue;
        uint64_t h = sl_mix(o->pairs[i].key) & o->index_mask;
        while (o->index[h] >= 0) h = (h + 1) & o->index_mask;
        o->index[h] = i;
    }
    const size_t alen = (size_t)o->n_seqs + 1;
    free(o->s_reads);
    free(o->s_bases);
    free(o->s_frags);
    free(o->s_ambig);
    free(o->countArray);
    o->s_reads = (int64_t *)calloc(alen, 8);
    o->s_bases = (int64_t *)calloc(alen, 8);
    o->s_frags = (int64_t *)calloc(alen, 8);
    o->s_ambig = (int64_t *)calloc(alen, 8);
    o->countArray = (int32_t *)calloc(alen, 4);
    if (v) {
        v[0] = stored;
        v[1] = o->n_pairs;
        v[2] = o->ref_kmers;
    }
    return 0;
}

int64_t sl_ora_table(void *h, uint64_t *keys, int32_t *ids, int64_t cap) {
    sl_oracle *o = (sl_oracle *)h;
    for (int64_t i = 0; i < o->n_pairs && i < cap; i++) {
        keys[i] = o->pairs[i].key;
        ids[i] = o->pairs[i].id;
    }
    return o->n_pairs;
}

/* set.getValues(key): index of the first entry with this key and the number of entries, 0 if absent */
static int64_t sl_lookup(const sl_oracle *o, uint64_t key, int *n) {
    uint64_t h = sl_mix(key) & o->index_mask;
    *n = 0;
    for (;;) {
        const int64_t at = o->index[h];
        if (at < 0) return 0;
        if (o->pairs[at].key == key) {
            int c = 1;
            while (at + c < o->n_pairs && o->pairs[at + c].key == key) c++;
            *n = c;
            return at;
        }
        h = (h + 1) & o->index_mask;
    }
}

typedef struct {
    int32_t *a;
    int size, cap;
} sl_list;
static void sl_push(sl_list *l, int32_t x) {
    if (l->size == l->cap) {
        l->cap = l->cap ? l->cap * 2 : 64;
        l->a = (int32_t *)realloc(l->a, (size_t)l->cap * 4);
    }
    l->a[l->size++] = x;
}

/* jgi/Seal.java:2864-2907 */
static int sl_find_best_match(sl_oracle *o, const uint8_t *bases, int64_t L, int present, int32_t *hits, sl_list *idList) {
    if (!present || o->stored < 1) return 0;
    const int k = o->k;
    const int minlen = k - 1;
    const int minlen2 = (o->maskMiddle ? (k - o->midMaskLen) / 2 : k);
    const int shift = 2 * k, shift2 = shift - 2;
    const uint64_t mask = shift > 63 ? ~0ull : ~((~0ull) << shift);
    const uint64_t kmask = 1ull << (2 * k);
    uint64_t kmer = 0, rkmer = 0;
    int found = 0;
    int64_t len = 0;
    if (L < k) return -1;
    const int64_t start = (o->c.restrict_right < 1 ? 0 : (L - o->c.restrict_right > 0 ? L - o->c.restrict_right : 0));
    const int64_t stop = (o->c.restrict_left < 1 ? L : (L < o->c.restrict_left ? L : o->c.restrict_left));
    for (int64_t i = start; i < stop; i++) {
        const uint8_t b = bases[i];
        const uint64_t x = (uint64_t)s_num0[b], x2 = (uint64_t)s_cnum0[b];
        kmer = ((kmer << 2) | x) & mask;
        rkmer = ((rkmer >> 2) | (x2 << shift2)) & mask;
        if (b == 'N' && o->forbidNs) {
            len = 0;
            rkmer = 0;
        } else len++;
        if (len >= minlen2 && i >= minlen) {
            /* getValues -> getValuesInner (qHammingDistance == 0) */
            if (o->c.qskip > 1 && (i % o->c.qskip != 0)) continue;
            const uint64_t key = sl_to_value(o, kmer, rkmer, kmask);
            if (!sl_passes_speed(o, key)) continue;
            int n = 0;
            const int64_t at = sl_lookup(o, key, &n);
            if (n > 0) {
                for (int t = 0; t < n; t++) {
                    const int32_t id = o->pairs[at + t].id;
                    hits[id]++;
                    if (hits[id] == 1) sl_push(idList, id);
                }
                found++;
                if (o->c.match_mode == SEAL_MATCH_FIRST || (o->c.match_mode == SEAL_MATCH_UNIQUE && n == 1)) break;
            }
        }
    }
    return found;
}

/* jgi/Seal.java:2654-2667 */
static int sl_condense(int32_t *loose, const sl_list *packed, sl_list *counts) {
    counts->size = 0;
    if (packed->size < 1) return 0;
    int max = 0;
    for (int i = 0; i < packed->size; i++) {
        const int p = packed->a[i];
        const int c = loose[p];
        sl_push(counts, c);
        loose[p] = 0;
        if (c > max) max = c;
    }
    return max;
}

/* jgi/Seal.java:2697-2708 */
static void sl_filter_top(const sl_list *packed, const sl_list *counts, sl_list *out, int max, int cz) {
    out->size = 0;
    if (packed->size < 1) return;
    const int thresh = (max - cz > 1 ? max - cz : 1);
    for (int i = 0; i < packed->size; i++)
        if (counts->a[i] >= thresh) sl_push(out, packed->a[i]);
}

static int sl_num_valid_kmers(const uint8_t *bases, int64_t L, int k) {
    int len = 0, counted = 0;
    for (int64_t i = 0; i < L; i++) {
        if (s_num[bases[i]] < 0) len = 0;
        else len++;
        if (len >= k) counted++;
    }
    return counted;
}
static int sl_num_kmers1(int64_t L, int k) { return (int)(L - k + 1 > 0 ? L - k + 1 : 0); }

static int sl_cmp_int(const void *a, const void *b) {
    const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* the start/stop choice shared by assignTogether (:2393-2408) and assignIndependently (:2473-2488) */
static void sl_range(const sl_oracle *o, sl_list *fin, int64_t numericID, int *start, int *stop) {
    const int sites = fin->size;
    if (sites < 2 || o->c.ambig_mode == SEAL_AMBIG_ALL) {
        *start = 0;
        *stop = sites;
    } else if (o->c.ambig_mode == SEAL_AMBIG_TOSS) {
        *start = *stop = 0;
    } else if (o->c.ambig_mode == SEAL_AMBIG_FIRST) {
        qsort(fin->a, (size_t)fin->size, 4, sl_cmp_int);
        *start = 0;
        *stop = 1;
    } else {
        *start = (int)(numericID % sites);
        *stop = *start + 1;
    }
}

static void sl_emit(const sl_oracle *o, const seal_out *out, int64_t u, const sl_list *fin, int start, int stop, int sites,
                    int max) {
    if (out->n_assigned) out->n_assigned[u] = stop - start;
    if (out->first_id) out->first_id[u] = stop > start ? fin->a[start] : 0;
    if (out->n_sites) out->n_sites[u] = sites;
    if (out->max_hits) out->max_hits[u] = max;
    if (out->ids && o->c.ids_stride > 0) {
        for (int j = 0; j < o->c.ids_stride; j++)
            out->ids[u * o->c.ids_stride + j] = (start + j < stop) ? fin->a[start + j] : 0;
    }
}

/* fragments [f0, f1) of the batch: the matching block of ProcessThread.run with the thread's own count array and counters
 * (scaffoldReadCountsT ..., jgi/Seal.java:1995-2007) */
static void sl_process_range(sl_oracle *o, const uint8_t *bases, const int64_t *offsets, int32_t paired, int64_t first_numeric_id,
                             const seal_out *out, int64_t f0, int64_t f1, int32_t *countArray, int64_t *sr, int64_t *sb, int64_t *sf,
                             int64_t *sa, seal_stats *st) {
    const int k = o->k;
    sl_list idList1 = {0}, idList2 = {0}, countList1 = {0}, countList2 = {0}, finalList1 = {0}, finalList2 = {0};
    seal_stats s;
    memset(&s, 0, sizeof s);
    const int kpt = o->c.keep_pairs_together != 0;
    for (int64_t f = f0; f < f1; f++) {
        const int64_t i1 = paired ? 2 * f : f;
        const uint8_t *b1 = bases + offsets[i1];
        const int64_t L1 = offsets[i1 + 1] - offsets[i1];
        const int has2 = paired;
        const uint8_t *b2 = has2 ? bases + offsets[i1 + 1] : NULL;
        const int64_t L2 = has2 ? offsets[i1 + 2] - offsets[i1 + 1] : 0;
        const int64_t numericID = first_numeric_id + f;
        s.reads_in += 1 + has2;
        s.bases_in += L1 + L2;
        if (kpt) {
            idList1.size = 0;
            sl_find_best_match(o, b1, L1, 1, countArray, &idList1);
            sl_find_best_match(o, b2, L2, has2, countArray, &idList1);
            const int max = sl_condense(countArray, &idList1, &countList1);
            int cz = o->c.clearzone;
            if (o->c.clearzone_fraction > 0) {
                const int nv = sl_num_valid_kmers(b1, L1, k) + (has2 ? sl_num_valid_kmers(b2, L2, k) : 0);
                const int c2 = (int)ceil((double)(float)(o->c.clearzone_fraction * (float)nv));
                if (c2 > cz) cz = c2;
            }
            sl_filter_top(&idList1, &countList1, &finalList1, max, cz);
            const int sites = finalList1.size;
            const int nk = sl_num_kmers1(L1, k) + (has2 ? sl_num_kmers1(L2, k) : 0);
            const int mf = (int)(o->c.min_kmer_fraction * (float)nk);
            const int minhits = o->c.min_kmer_hits > mf ? o->c.min_kmer_hits : mf;
            int start = 0, stop = 0;
            if (max >= minhits) {
                /* assignTogether */
                const int64_t lenSum = L1 + L2;
                const int readSum = 1 + has2;
                sl_range(o, &finalList1, numericID, &start, &stop);
                for (int j = start; j < stop; j++) {
                    const int id = finalList1.a[j];
                    sr[id] += readSum;
                    sb[id] += lenSum;
                    sf[id]++;
                    if (sites > 1) sa[id] += readSum;
                }
                if (start < stop) {
                    s.reads_matched += readSum;
                    s.bases_matched += lenSum;
                } else {
                    s.reads_unmatched += readSum;
                    s.bases_unmatched += lenSum;
                }
            } else {
                s.reads_unmatched += 1 + has2;
                s.bases_unmatched += L1 + L2;
            }
            sl_emit(o, out, paired ? f : i1, &finalList1, start, stop, sites, max);
        } else {
            idList1.size = 0;
            sl_find_best_match(o, b1, L1, 1, countArray, &idList1);
            const int max1 = sl_condense(countArray, &idList1, &countList1);
            {
                int cz = o->c.clearzone;
                if (o->c.clearzone_fraction > 0) {
                    const int c2 = (int)ceil((double)(float)(o->c.clearzone_fraction * (float)sl_num_valid_kmers(b1, L1, k)));
                    if (c2 > cz) cz = c2;
                }
                sl_filter_top(&idList1, &countList1, &finalList1, max1, cz);
            }
            int max2 = 0;
            finalList2.size = 0;
            if (has2) {
                idList2.size = 0;
                sl_find_best_match(o, b2, L2, 1, countArray, &idList2);
                max2 = sl_condense(countArray, &idList2, &countList2);
                int cz = o->c.clearzone;
                if (o->c.clearzone_fraction > 0) {
                    const int c2 = (int)ceil((double)(float)(o->c.clearzone_fraction * (float)sl_num_valid_kmers(b2, L2, k)));
                    if (c2 > cz) cz = c2;
                }
                sl_filter_top(&idList2, &countList2, &finalList2, max2, cz);
            }
            /* assignIndependently */
            for (int m = 0; m < 1 + has2; m++) {
                sl_list *fin = m ? &finalList2 : &finalList1;
                const int64_t L = m ? L2 : L1;
                const int mx = m ? max2 : max1;
                const int mf = (int)(o->c.min_kmer_fraction * (float)sl_num_kmers1(L, k));
                const int minhits = o->c.min_kmer_hits > mf ? o->c.min_kmer_hits : mf;
                const int sites = fin->size;
                int start = 0, stop = 0;
                if (mx >= minhits) {
                    sl_range(o, fin, numericID, &start, &stop);
                    const int frag = m ? (max2 > max1) : (max1 >= max2);
                    for (int j = start; j < stop; j++) {
                        const int id = fin->a[j];
                        sr[id]++;
                        sb[id] += L;
                        if (frag) sf[id]++;
                        if (sites > 1) sa[id]++;
                    }
                    if (start < stop) {
                        s.reads_matched++;
                        s.bases_matched += L;
                    } else {
                        s.reads_unmatched++;
                        s.bases_unmatched += L;
                    }
                }
                sl_emit(o, out, i1 + m, fin, start, stop, sites, mx);
            }
        }
    }
    free(idList1.a);
    free(idList2.a);
    free(countList1.a);
    free(countList2.a);
    free(finalList1.a);
    free(finalList2.a);
    *st = s;
}


int sl_ora_process(void *h, const uint8_t *bases, const int64_t *offsets, int64_t n_reads, int32_t paired,
                   int64_t first_numeric_id, const seal_out *out, seal_stats *st) {
    sl_oracle *o = (sl_oracle *)h;
    if (!o->countArray) return 1;
    seal_stats s;
    sl_process_range(o, bases, offsets, paired, first_numeric_id, out, 0, paired ? n_reads / 2 : n_reads, o->countArray, o->s_reads,
                     o->s_bases, o->s_frags, o->s_ambig, &s);
    if (st) *st = s;
    return 0;
}

/* the same with `threads` ProcessThreads over contiguous slices of the batch; their counters are summed as the reference sums
 * its threads' (jgi/Seal.java:1640-1672). Results do not depend on the number of threads. */
typedef struct {
    sl_oracle *o;
    const uint8_t *bases;
    const int64_t *offsets;
    int32_t paired;
    int64_t first_numeric_id;
    const seal_out *out;
    int64_t f0, f1;
    int32_t *countArray;
    int64_t *cnt; /* 4 * alen */
    seal_stats st;
} sl_job;

static void *sl_worker(void *arg) {
    sl_job *j = (sl_job *)arg;
    const size_t alen = (size_t)j->o->n_seqs + 1;
    sl_process_range(j->o, j->bases, j->offsets, j->paired, j->first_numeric_id, j->out, j->f0, j->f1, j->countArray, j->cnt,
                     j->cnt + alen, j->cnt + 2 * alen, j->cnt + 3 * alen, &j->st);
    return NULL;
}

int sl_ora_process_mt(void *h, const uint8_t *bases, const int64_t *offsets, int64_t n_reads, int32_t paired,
                      int64_t first_numeric_id, const seal_out *out, seal_stats *st, int32_t threads) {
    sl_oracle *o = (sl_oracle *)h;
    if (!o->countArray) return 1;
    const int64_t n_frag = paired ? n_reads / 2 : n_reads;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if ((int64_t)threads > n_frag) threads = n_frag > 0 ? (int32_t)n_frag : 1;
    const size_t alen = (size_t)o->n_seqs + 1;
    sl_job *jobs = (sl_job *)calloc((size_t)threads, sizeof(sl_job));
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        sl_job *j = &jobs[t];
        j->o = o;
        j->bases = bases;
        j->offsets = offsets;
        j->paired = paired;
        j->first_numeric_id = first_numeric_id;
        j->out = out;
        j->f0 = n_frag * t / threads;
        j->f1 = n_frag * (t + 1) / threads;
        j->countArray = (int32_t *)calloc(alen, 4);
        j->cnt = (int64_t *)calloc(4 * alen, 8);
        pthread_create(&tid[t], NULL, sl_worker, j);
    }
    seal_stats s;
    memset(&s, 0, sizeof s);
    for (int t = 0; t < threads; t++) {
        pthread_join(tid[t], NULL);
        sl_job *j = &jobs[t];
        s.reads_in += j->st.reads_in;
        s.bases_in += j->st.bases_in;
        s.reads_matched += j->st.reads_matched;
        s.bases_matched += j->st.bases_matched;
        s.reads_unmatched += j->st.reads_unmatched;
        s.bases_unmatched += j->st.bases_unmatched;
        for (size_t i = 0; i < alen; i++) {
            o->s_reads[i] += j->cnt[i];
            o->s_bases[i] += j->cnt[alen + i];
            o->s_frags[i] += j->cnt[2 * alen + i];
            o->s_ambig[i] += j->cnt[3 * alen + i];
        }
        free(j->countArray);
        free(j->cnt);
    }
    free(jobs);
    free(tid);
    if (st) *st = s;
    return 0;
}

int sl_ora_scaffold_counts(void *h, int64_t *reads, int64_t *bases, int64_t *frags, int64_t *ambig, int32_t n) {
    sl_oracle *o = (sl_oracle *)h;
    if (!o->s_reads) return 1;
    for (int32_t i = 0; i < n && i <= o->n_seqs; i++) {
        if (reads) reads[i] = o->s_reads[i];
        if (bases) bases[i] = o->s_bases[i];
        if (frags) frags[i] = o->s_frags[i];
        if (ambig) ambig[i] = o->s_ambig[i];
    }
    return 0;
}
