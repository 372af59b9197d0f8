/* oracle/rkmh_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the rkmh MinHash read-classification hot path.
 * It exists to CHECK the CUDA product; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (rkmh_b200/) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   - the reference's own Murmur3 known-answer vectors (mkmh/murmur3/test.c:42-45),
 *   - the relational checks of mkmh/mkmh_test.cpp (strand invariance, N => 0, intersection == 2),
 *   - golden vectors produced by the real reference compiled here (oracle/_ref, see oracle/Makefile)
 *     and committed under tests/golden/ by tests/golden/make_golden.py.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 */
#ifndef RKMH_ORACLE_H
#define RKMH_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t ref_id;     /* first index attaining the maximum shared count (0 when every count is 0) */
    int32_t best;       /* the maximum shared count                                                 */
    int32_t prev;       /* max of the counts at indices < ref_id, or -1 when ref_id == 0            */
    int32_t sketch_len; /* number of hashes in the read's sketch (<= s)                             */
} rko_hit;

/* mkmh/murmur3/murmur3.cpp:234-312 */
void rko_murmur3_x64_128(const void* key, int len, uint32_t seed, void* out);
/* mkmh/mkmh.hpp:231-236 */
void rko_to_upper(char* seq, int64_t len);
/* mkmh/mkmh.hpp:137-152,182-188 : 1 when every byte is one of ACGTacgt */
int rko_canonical(const char* x, int len);
/* mkmh/mkmh.hpp:172-178,204-215 */
void rko_reverse_complement(const char* seq, char* ret, int len);
/* one window of mkmh/mkmh.hpp:515-528 */
uint64_t rko_calc_hash(const char* kmer, int k);
/* mkmh/mkmh.hpp:505-532 : writes max(len-k,0) hashes, returns that count */
int64_t rko_calc_hashes(const char* seq, int64_t len, int k, uint64_t* out);
/* mkmh/mkmh.hpp:689-711 : per-k arrays concatenated in list order; returns the total */
int64_t rko_calc_hashes_multi(const char* seq, int64_t len, const int* ks, int nk, uint64_t* out);
/* number of hashes rko_calc_hashes_multi will write */
int64_t rko_num_hashes(int64_t len, const int* ks, int nk);

/* mkmh/HASHTCounter.cpp:46-49 (++counts[key % size]); valid_only=1 follows mkmh/mkmh.hpp:729-737
 * (only valid windows are counted), valid_only=0 follows src/rkmh.cpp:331-336 (every hash, zeros too) */
void rko_count_increment(int32_t* counts, uint64_t size, const uint64_t* h, int64_t n, int valid_only);
/* mkmh/mkmh.hpp:898-908 with max_occ = UINT32_MAX */
void rko_mask_by_frequency(uint64_t* h, int64_t n, const int32_t* counts, uint64_t size, int min_occ);
/* -I side of mkmh/mkmh.hpp:954-976 expressed as a mask: h := 0 when counts[h % size] > max_occ */
void rko_mask_max_frequency(uint64_t* h, int64_t n, const int32_t* counts, uint64_t size, int max_occ);

/* mkmh/mkmh.hpp:272-279 + 931-950 : sorts `hashes` in place, copies the first s non-zero values
 * (duplicates kept) into ret (capacity >= min(n,s)); returns the sketch length */
int rko_minhashes(uint64_t* hashes, int64_t n, int s, uint64_t* ret);
/* mkmh/mkmh.hpp:822-840 */
int rko_intersection_size(const uint64_t* a, int na, const uint64_t* b, int nb);
/* multiset merge of two sorted sketches keeping the first s (sharded `hash`) */
int rko_merge_sketches(const uint64_t* a, int na, const uint64_t* b, int nb, int s, uint64_t* out);

/* src/rkmh.cpp:874-883 (and 925-934): argmax over shared counts with the running-prefix diff */
void rko_argmax(const int32_t* counts, int nrefs, int sketch_len, rko_hit* out);

/* ---- batch drivers: sequences are a concatenated `bases` buffer + nseq+1 offsets -------------- */

/* sketches of nseq sequences; sketches is [nseq * s], lens is [nseq].  Sequences are upper-cased
 * on a private copy first (src/rkmh.cpp:252,818).  counts != NULL applies rko_mask_max_frequency
 * with max_occ (the -I path, src/rkmh.cpp:829-837) when mode==1 or rko_mask_by_frequency with
 * min_occ (the -M path, src/rkmh.cpp:924) when mode==2; mode==0 ignores counts. */
void rko_sketch_batch(const char* bases, const int64_t* offs, int64_t nseq, const int* ks, int nk, int s,
                      const int32_t* counts, uint64_t counts_size, int mode, int occ,
                      uint64_t* sketches, int32_t* lens);
/* pass 1 of -M / -I: hash every sequence and count (valid_only as above; per_sample_set=1 follows
 * src/rkmh.cpp:347-355, one increment per distinct hash per sequence, zero included) */
void rko_count_batch(const char* bases, const int64_t* offs, int64_t nseq, const int* ks, int nk,
                     int32_t* counts, uint64_t counts_size, int valid_only, int per_sample_set);
/* classification of nreads reads against nrefs reference sketches (ref_sketches [nrefs * s]). */
void rko_classify_batch(const char* bases, const int64_t* offs, int64_t nreads, const int* ks, int nk, int s,
                        const uint64_t* ref_sketches, const int32_t* ref_lens, int nrefs,
                        const int32_t* counts, uint64_t counts_size, int min_occ,
                        rko_hit* out);

/* ---- hpv16 -------------------------------------------------------------------- */
/* mkmh/mkmh.hpp:842-863 */
int rko_set_intersection_size(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb);
/* src/rkmh.cpp:2643-2706: out[r] = {first type with the most shared distinct hashes, that count, -1, number of
 * hashes of the read}; extra [nreads x (nsets - ntypes)] = hash_intersection_size with the remaining sets */
void rko_classify_sets(const char* bases, const int64_t* offs, int64_t nreads, const int* ks, int nk,
                       const uint64_t* set_hashes, const int64_t* set_offs, int nsets, int ntypes,
                       const int32_t* counts, uint64_t counts_size, int min_occ, rko_hit* out, int32_t* extra);

#ifdef __cplusplus
}
#endif
#endif
