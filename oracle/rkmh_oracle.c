/* oracle/rkmh_oracle.c -- TEST INFRASTRUCTURE ONLY (see rkmh_oracle.h for the rules and the
 * parity status).  Plain-C restatement of the rkmh MinHash hot path; each function names the
 * reference file:line it restates.  Nothing here is used by the product path. */
#include "rkmh_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ MurmurHash3_x64_128 ------ */
/* mkmh/murmur3/murmur3.cpp:26-29 */
static inline uint64_t rot64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

/* mkmh/murmur3/murmur3.cpp:58-67 */
static inline uint64_t avalanche64(uint64_t v) {
    v ^= v >> 33;
    v *= 0xff51afd7ed558ccdULL;
    v ^= v >> 33;
    v *= 0xc4ceb9fe1a85ec53ULL;
    v ^= v >> 33;
    return v;
}

static inline uint64_t load_le64(const uint8_t* p, int nbytes) {
    uint64_t v = 0;
    for (int b = 0; b < nbytes; ++b) v |= (uint64_t)p[b] << (8 * b);
    return v;
}

/* mkmh/murmur3/murmur3.cpp:234-312.  Lanes A/B are the reference's h1/h2. */
void rko_murmur3_x64_128(const void* key, int len, uint32_t seed, void* out) {
    const uint8_t* p = (const uint8_t*)key;
    const uint64_t CA = 0x87c37b91114253d5ULL, CB = 0x4cf5ad432745937fULL;
    uint64_t a = seed, b = seed;
    int full = len / 16;
    for (int blk = 0; blk < full; ++blk, p += 16) {             /* body, :248-259 */
        uint64_t ka = load_le64(p, 8), kb = load_le64(p + 8, 8);
        ka *= CA; ka = rot64(ka, 31); ka *= CB; a ^= ka;
        a = rot64(a, 27); a += b; a = a * 5 + 0x52dce729;
        kb *= CB; kb = rot64(kb, 33); kb *= CA; b ^= kb;
        b = rot64(b, 31); b += a; b = b * 5 + 0x38495ab5;
    }
    int rem = len & 15;                                           /* tail, :264-289 */
    if (rem > 8) {
        uint64_t kb = load_le64(p + 8, rem - 8);
        kb *= CB; kb = rot64(kb, 33); kb *= CA; b ^= kb;
    }
    if (rem > 0) {
        uint64_t ka = load_le64(p, rem > 8 ? 8 : rem);
        ka *= CA; ka = rot64(ka, 31); ka *= CB; a ^= ka;
    }
    a ^= (uint64_t)(int64_t)len; b ^= (uint64_t)(int64_t)len;     /* finalisation, :294-305 */
    a += b; b += a;
    a = avalanche64(a); b = avalanche64(b);
    a += b; b += a;
    ((uint64_t*)out)[0] = a;
    ((uint64_t*)out)[1] = b;
}

/* ------------------------------------------------------------------ bases --------------------- */
/* mkmh/mkmh.hpp:231-236 : (c - 91) > 0 ? c - 32 : c, on (signed) char */
void rko_to_upper(char* seq, int64_t len) {
    for (int64_t i = 0; i < len; ++i) {
        signed char c = (signed char)seq[i];
        if ((int)c - 91 > 0) seq[i] = (char)(c - 32);
    }
}

static inline int is_acgt(unsigned char c) {
    /* the eight zero entries of valid_dna[], mkmh/mkmh.hpp:137-152 */
    switch (c) {
        case 'A': case 'C': case 'G': case 'T':
        case 'a': case 'c': case 'g': case 't': return 1;
        default: return 0;
    }
}

/* mkmh/mkmh.hpp:182-188 */
int rko_canonical(const char* x, int len) {
    int bad = 0;
    for (int i = 0; i < len; ++i) bad |= !is_acgt((unsigned char)x[i]);
    return !bad;
}

/* mkmh/mkmh.hpp:172-178 (rev_arr: A<->T, C<->G, every other upper-case letter maps to itself)
 * and :204-215.  Only ever reached with upper-cased ACGT windows on the hot path. */
void rko_reverse_complement(const char* seq, char* ret, int len) {
    for (int i = 0; i < len; ++i) {
        char c = seq[len - 1 - i], o = c;
        if (c == 'A') o = 'T'; else if (c == 'T') o = 'A';
        else if (c == 'C') o = 'G'; else if (c == 'G') o = 'C';
        ret[i] = o;
    }
    ret[len] = '\0';
}

/* One window of mkmh/mkmh.hpp:515-528.  The reference reads the 128-bit digest as uint32[4] and
 * forms (w0 << 32) | w1, i.e. the first 64-bit word with its halves swapped. */
static inline uint64_t digest_to_hash(const uint64_t d[2]) { return (d[0] << 32) | (d[0] >> 32); }

uint64_t rko_calc_hash(const char* kmer, int k) {
    if (!rko_canonical(kmer, k)) return 0;
    char stackbuf[256];
    char* rc = (k + 1 <= (int)sizeof stackbuf) ? stackbuf : (char*)malloc((size_t)k + 1);
    uint64_t df[2], dr[2];
    rko_reverse_complement(kmer, rc, k);
    rko_murmur3_x64_128(kmer, k, 42u, df);
    rko_murmur3_x64_128(rc, k, 42u, dr);
    if (rc != stackbuf) free(rc);
    uint64_t f = digest_to_hash(df), r = digest_to_hash(dr);
    return f < r ? f : r;
}

/* mkmh/mkmh.hpp:505-532 : windows 0 .. len-k-1 (the last k-mer is never hashed) */
int64_t rko_calc_hashes(const char* seq, int64_t len, int k, uint64_t* out) {
    int64_t n = len - k;
    if (n < 0) n = 0;
    for (int64_t i = 0; i < n; ++i) out[i] = rko_calc_hash(seq + i, k);
    return n;
}

int64_t rko_num_hashes(int64_t len, const int* ks, int nk) {
    int64_t n = 0;
    for (int j = 0; j < nk; ++j) n += (len - ks[j] > 0) ? len - ks[j] : 0;
    return n;
}

/* mkmh/mkmh.hpp:689-711 */
int64_t rko_calc_hashes_multi(const char* seq, int64_t len, const int* ks, int nk, uint64_t* out) {
    int64_t n = 0;
    for (int j = 0; j < nk; ++j) n += rko_calc_hashes(seq, len, ks[j], out + n);
    return n;
}

/* ------------------------------------------------------------------ count table --------------- */
/* mkmh/HASHTCounter.cpp:46-49; callers mkmh/mkmh.hpp:729-737 (valid windows only) and
 * src/rkmh.cpp:331-336 (every hash) */
void rko_count_increment(int32_t* counts, uint64_t size, const uint64_t* h, int64_t n, int valid_only) {
    for (int64_t i = 0; i < n; ++i)
        if (!valid_only || h[i] != 0) ++counts[h[i] % size];
    /* NB a valid window whose hash is exactly 0 would be counted by the reference and skipped here;
     * the device path makes the same choice, and no such k-mer is known. */
}

/* mkmh/mkmh.hpp:898-908 (max_occ left at UINT32_MAX as in src/rkmh.cpp:924) */
void rko_mask_by_frequency(uint64_t* h, int64_t n, const int32_t* counts, uint64_t size, int min_occ) {
    for (int64_t i = 0; i < n; ++i) {
        int freq = counts[h[i] % size];
        if (!(min_occ <= freq)) h[i] = 0;
    }
}

/* mkmh/mkmh.hpp:954-976 keeps h != 0 && min_occ(0) <= freq <= max_occ while scanning the sorted
 * array; zeroing the rejected hashes before the sort selects the same multiset. */
void rko_mask_max_frequency(uint64_t* h, int64_t n, const int32_t* counts, uint64_t size, int max_occ) {
    for (int64_t i = 0; i < n; ++i) {
        int freq = counts[h[i] % size];
        if (!(freq >= 0 && (uint32_t)freq <= (uint32_t)max_occ)) h[i] = 0;
    }
}

/* ------------------------------------------------------------------ sketch -------------------- */
static int cmp_u64(const void* x, const void* y) {
    uint64_t a = *(const uint64_t*)x, b = *(const uint64_t*)y;
    return (a > b) - (a < b);
}

/* mkmh/mkmh.hpp:272-279 (ascending sort) + :931-950 */
int rko_minhashes(uint64_t* hashes, int64_t n, int s, uint64_t* ret) {
    qsort(hashes, (size_t)n, sizeof(uint64_t), cmp_u64);
    int kept = 0;
    for (int64_t i = 0; i < n && kept < s; ++i)
        if (hashes[i] != 0) ret[kept++] = hashes[i];
    return kept;
}

/* mkmh/mkmh.hpp:822-840 : two-pointer merge, equal non-zero values advance both sides */
int rko_intersection_size(const uint64_t* a, int na, const uint64_t* b, int nb) {
    int i = 0, j = 0, shared = 0;
    while (i < na && j < nb) {
        if (a[i] == b[j] && a[i] != 0) { ++shared; ++i; ++j; }
        else if (a[i] > b[j]) ++j;
        else ++i;
    }
    return shared;
}

/* Sharded `hash`: the bottom-s multiset of a concatenation equals the first s of the
 * merge of the parts' bottom-s multisets. */
int rko_merge_sketches(const uint64_t* a, int na, const uint64_t* b, int nb, int s, uint64_t* out) {
    int i = 0, j = 0, n = 0;
    while (n < s && (i < na || j < nb)) {
        if (j >= nb || (i < na && a[i] <= b[j])) out[n++] = a[i++];
        else out[n++] = b[j++];
    }
    return n;
}

/* src/rkmh.cpp:874-883 : max_shared starts at -1, strict >, diff = new - previous running max.
 * We return prev (the running max before the last update, -1 when the winner is index 0) so that
 * both stream (diff = best - prev) and filter (src/equiv.hpp:328-345, best/prev floor at 0)
 * semantics can be derived by the caller. */
void rko_argmax(const int32_t* counts, int nrefs, int sketch_len, rko_hit* out) {
    int best = -1, id = 0, prev = -1;
    for (int j = 0; j < nrefs; ++j)
        if (counts[j] > best) { prev = best; best = counts[j]; id = j; }
    out->ref_id = id; out->best = best; out->prev = prev; out->sketch_len = sketch_len;
}

/* ------------------------------------------------------------------ batch drivers ------------- */
static char* upper_copy(const char* bases, int64_t lo, int64_t hi) {
    int64_t len = hi - lo;
    char* s = (char*)malloc((size_t)(len > 0 ? len : 1));
    memcpy(s, bases + lo, (size_t)len);
    rko_to_upper(s, len);
    return s;
}

void rko_sketch_batch(const char* bases, const int64_t* offs, int64_t nseq, const int* ks, int nk, int s,
                      const int32_t* counts, uint64_t counts_size, int mode, int occ,
                      uint64_t* sketches, int32_t* lens) {
    for (int64_t r = 0; r < nseq; ++r) {
        int64_t len = offs[r + 1] - offs[r];
        char* seq = upper_copy(bases, offs[r], offs[r + 1]);
        int64_t cap = rko_num_hashes(len, ks, nk);
        uint64_t* h = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(cap > 0 ? cap : 1));
        int64_t n = rko_calc_hashes_multi(seq, len, ks, nk, h);
        if (counts && mode == 1) rko_mask_max_frequency(h, n, counts, counts_size, occ);
        if (counts && mode == 2) rko_mask_by_frequency(h, n, counts, counts_size, occ);
        lens[r] = rko_minhashes(h, n, s, sketches + (size_t)r * (size_t)s);
        free(h); free(seq);
    }
}

void rko_count_batch(const char* bases, const int64_t* offs, int64_t nseq, const int* ks, int nk,
                     int32_t* counts, uint64_t counts_size, int valid_only, int per_sample_set) {
    for (int64_t r = 0; r < nseq; ++r) {
        int64_t len = offs[r + 1] - offs[r];
        char* seq = upper_copy(bases, offs[r], offs[r + 1]);
        int64_t cap = rko_num_hashes(len, ks, nk);
        uint64_t* h = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(cap > 0 ? cap : 1));
        int64_t n = rko_calc_hashes_multi(seq, len, ks, nk, h);
        if (per_sample_set) {                       /* src/rkmh.cpp:347-355: set<hash_t>, zero included */
            qsort(h, (size_t)n, sizeof(uint64_t), cmp_u64);
            for (int64_t i = 0; i < n; ++i)
                if (i == 0 || h[i] != h[i - 1]) ++counts[h[i] % counts_size];
        } else {
            rko_count_increment(counts, counts_size, h, n, valid_only);
        }
        free(h); free(seq);
    }
}

void rko_classify_batch(const char* bases, const int64_t* offs, int64_t nreads, const int* ks, int nk, int s,
                        const uint64_t* ref_sketches, const int32_t* ref_lens, int nrefs,
                        const int32_t* counts, uint64_t counts_size, int min_occ,
                        rko_hit* out) {
    uint64_t* mins = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(s > 0 ? s : 1));
    int32_t* shared = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nrefs > 0 ? nrefs : 1));
    for (int64_t r = 0; r < nreads; ++r) {            /* src/rkmh.cpp:845-898 / 911-948 */
        int64_t len = offs[r + 1] - offs[r];
        char* seq = upper_copy(bases, offs[r], offs[r + 1]);
        int64_t cap = rko_num_hashes(len, ks, nk);
        uint64_t* h = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(cap > 0 ? cap : 1));
        int64_t n = rko_calc_hashes_multi(seq, len, ks, nk, h);
        if (counts) rko_mask_by_frequency(h, n, counts, counts_size, min_occ);
        int L = rko_minhashes(h, n, s, mins);
        for (int j = 0; j < nrefs; ++j)
            shared[j] = rko_intersection_size(mins, L, ref_sketches + (size_t)j * (size_t)s, ref_lens[j]);
        rko_argmax(shared, nrefs, L, &out[r]);
        free(h); free(seq);
    }
    free(mins); free(shared);
}

/* ------------------------------------------------------------------ hpv16 ------- */
/* mkmh/mkmh.hpp:842-863 : walk over two ascending arrays counting the DISTINCT non-zero values they share
 * (`prev` keeps a value from being counted twice; an equal pair that does not count advances alpha only) */
int rko_set_intersection_size(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb) {
    int64_t i = 0, j = 0;
    int shared = 0;
    uint64_t prev = 0;
    while (i < na && j < nb) {
        if (a[i] == b[j] && a[i] != prev && a[i] != 0) { ++shared; prev = a[i]; ++i; ++j; }
        else if (a[i] > b[j]) ++j;
        else ++i;
    }
    return shared;
}

/* The per-read block of main_hpv16 (src/rkmh.cpp:2643-2706): every hash of the read over every k (zeros included),
 * masked by frequency when counts != NULL, sorted ascending; sets [0, ntypes) are compared with
 * hash_set_intersection_size and reduced to the first index of the strict maximum starting from -1
 * (src/rkmh.cpp:2655-2664); sets [ntypes, nsets) -- the lineage / sublineage specific tables -- with
 * hash_intersection_size as sort_by_similarity does (mkmh/mkmh.hpp:1304-1333), into extra [nreads x (nsets - ntypes)].
 * Every set must be ascending (the reference's type arrays are: minhashes() sorts them in place, src/rkmh.cpp:2547). */
void rko_classify_sets(const char* bases, const int64_t* offs, int64_t nreads, const int* ks, int nk,
                       const uint64_t* set_hashes, const int64_t* set_offs, int nsets, int ntypes,
                       const int32_t* counts, uint64_t counts_size, int min_occ, rko_hit* out, int32_t* extra) {
    for (int64_t r = 0; r < nreads; ++r) {
        int64_t len = offs[r + 1] - offs[r];
        char* seq = upper_copy(bases, offs[r], offs[r + 1]);
        int64_t cap = rko_num_hashes(len, ks, nk);
        uint64_t* h = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(cap > 0 ? cap : 1));
        int64_t n = rko_calc_hashes_multi(seq, len, ks, nk, h);
        if (counts) rko_mask_by_frequency(h, n, counts, counts_size, min_occ);
        qsort(h, (size_t)n, sizeof(uint64_t), cmp_u64);
        int best = -1, id = 0;
        for (int j = 0; j < ntypes; ++j) {
            int shared = rko_set_intersection_size(h, n, set_hashes + set_offs[j], set_offs[j + 1] - set_offs[j]);
            if (shared > best) { best = shared; id = j; }
        }
        out[r].ref_id = id; out[r].best = best; out[r].prev = -1; out[r].sketch_len = (int32_t)n;
        for (int j = ntypes; j < nsets; ++j)
            extra[(size_t)r * (size_t)(nsets - ntypes) + (size_t)(j - ntypes)] =
                rko_intersection_size(h, (int)n, set_hashes + set_offs[j], (int)(set_offs[j + 1] - set_offs[j]));
        free(h); free(seq);
    }
}
