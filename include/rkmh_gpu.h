/* include/rkmh_gpu.h -- C ABI of the B200-native rkmh MinHash read-classification hot path.
 *
 * The reference (edawson/rkmh @ ea3d2e6) has no plugin/FFI interface: its hot path is the set of mkmh
 * library calls made from the per-read loops of main_stream / main_filter / main_hash.  This header is
 * the boundary a maintainer would bind instead of those loops (see INTEGRATION.md for the exact patch
 * to src/rkmh.cpp).  Each entry point names the reference code it replaces (paths relative to the
 * reference root).
 *
 * Conventions
 *   - plain C: opaque context, plain pointers and sizes, int status (0 = ok, <0 = error; the text is in
 *     rkg_last_error()).  No C++ or torch types cross this line.
 *   - a context owns ONE GPU (one process per GPU; a multi-GPU host opens one context per device) and
 *     all device memory; calls on one context must be serialised by the caller.
 *   - sequences travel as a concatenated byte buffer `bases` plus `n + 1` int64 offsets; record i is
 *     bases[offs[i] .. offs[i+1]).  Bytes are taken exactly as parsed from FASTA/FASTQ: upper-casing
 *     (mkmh::to_upper, mkmh/mkmh.hpp:231-236) happens on the device.  Names never cross the boundary.
 *   - host buffers may be pageable; buffers obtained from rkg_host_alloc() are pinned and make the
 *     host<->device copies asynchronous.
 *   - there is NO CPU fallback: every entry point fails with RKG_ERR_CUDA when no sm_100 device is usable.
 *   - k-mer sizes 1..32 are supported (the reference accepts any int; its data sets use 10..21).
 *   - sketch sizes 1..2^26; nrefs x sketch_size <= 2^29.  The reference table is DENSE in the references (one bit per
 *     reference per distinct sketch hash): its size is about 2 nrefs s x (1 + ceil(nrefs / 64)) x 8 bytes, i.e.
 *     quadratic in nrefs -- 182 references at s = 1000 take 16 MB, 5,000 about 11 GB, 20,000 do not fit a B200.  The
 *     reference's own cost is O(nrefs x s) memory and O(nrefs x s) time PER READ; nothing on this path needs more
 *     than a few thousand references (rkg_load_refs fails with RKG_ERR_NOMEM when the table does not fit).
 */
#ifndef RKMH_GPU_H
#define RKMH_GPU_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RKG_OK            0
#define RKG_ERR_ARG      -1   /* bad argument / unsupported parameter        */
#define RKG_ERR_CUDA     -2   /* CUDA runtime error or no usable device      */
#define RKG_ERR_STATE    -3   /* call order (e.g. classify before load_refs) */
#define RKG_ERR_NOMEM    -4

#define RKG_MAX_K        32
#define RKG_MAX_NK       8

typedef struct rkg_ctx rkg_ctx;
typedef struct rkg_batch rkg_batch;   /* a read batch resident in device memory */

/* Result of classifying one read: the argmax block of src/rkmh.cpp:874-883 (stream) and
 * src/equiv.hpp:328-345 (filter), in a form both can be derived from:
 *   stream : max_shared = best, max_id = ref_id, diff = best - prev
 *   filter : max_shared = max(best,0), prev_best = max(prev,0), sample = (best > 0 ? ref_id : none) */
typedef struct {
    int32_t ref_id;      /* first reference index attaining the maximum shared count (0 if all are 0) */
    int32_t best;        /* maximum over references of hash_intersection_size (mkmh/mkmh.hpp:822-840)  */
    int32_t prev;        /* max of the counts at indices < ref_id; -1 when ref_id == 0                 */
    int32_t sketch_len;  /* length of the read's sketch = min(s, #non-zero hashes) (mkmh.hpp:931-950) */
} rkg_hit;

/* ---- context ------------------------------------------------------------------------------------- */
/* Replaces omp_set_num_threads() + the implicit process state (src/rkmh.cpp:734). */
int  rkg_create(int device, rkg_ctx** ctx);
void rkg_destroy(rkg_ctx* ctx);
/* Error text of the last failing call on ctx (ctx == NULL: of the last failing rkg_create). */
const char* rkg_last_error(const rkg_ctx* ctx);

/* k-mer sizes in command-line order (-k repeatable, src/rkmh.cpp:675-677; hashes of each k are
 * concatenated, mkmh/mkmh.hpp:689-711), sketch size -s, and the sizes of the lossy count tables
 * (HASHTCounter, mkmh/HASHTCounter.cpp:12-15): 200000000 for stream (src/rkmh.cpp:739,742),
 * 10000000 for filter (src/rkmh.cpp:1187-1188).  A size of 0 disables that table. */
int rkg_set_params(rkg_ctx* ctx, const int* ks, int nk, int sketch_size,
                   uint64_t read_table_size, uint64_t ref_table_size);

/* Pinned host memory for batch buffers (replaces the per-record `new char[]` of parse_fastas,
 * src/rkmh.cpp:238-263).  These two calls are THREAD-SAFE: they may run concurrently with any other call on the
 * same context (a parser thread pins chunks while a worker thread classifies); on failure the text is in
 * rkg_last_error(NULL) of the calling thread, the context's own error slot is not touched. */
int rkg_host_alloc(rkg_ctx* ctx, size_t bytes, void** ptr);
int rkg_host_free(rkg_ctx* ctx, void* ptr);

/* ---- references ---------------------------------------------------------------------------------- */
/* Hash + sketch every reference and build the HBM-resident (hash,rank) -> reference-bitset table.
 * Replaces src/rkmh.cpp:815-838 (stream) and 1192-1246 (filter).
 *   max_ref_occ <  0 : plain bottom-s sketches (no -I)
 *   max_ref_occ >= 0 : the -I path: count reference hashes in the ref table first, then keep only
 *                      hashes whose count is <= max_ref_occ (minhashes_frequency_filter,
 *                      mkmh/mkmh.hpp:954-976).  per_sample_set = 0 counts every valid window
 *                      occurrence (stream, mkmh.hpp:729-737); 1 counts each distinct hash once per
 *                      reference, zero included (filter, src/rkmh.cpp:347-355). */
int rkg_load_refs(rkg_ctx* ctx, const char* bases, const int64_t* offs, int nrefs,
                  int max_ref_occ, int per_sample_set);
/* Pre-hashed references (the reference's --pre-references / -R option, src/rkmh.cpp:640, 660-662, parsed and never
 * acted on): build the table from sketches computed earlier -- by rkg_get_ref_sketches, `rkmh hash -o`, or
 * any minhashes() -- instead of from sequences.  sketches is [nrefs * s] with each row ascending, non-zero, lens[r] <= s
 * long (checked), for the k list and sketch size of rkg_set_params.  Classification results are those of rkg_load_refs
 * on the sequences the sketches were made from. */
int rkg_load_ref_sketches(rkg_ctx* ctx, const uint64_t* sketches, const int32_t* lens, int nrefs);
/* Copy the reference sketches out (sketches is [nrefs * s], lens is [nrefs]). */
int rkg_get_ref_sketches(rkg_ctx* ctx, uint64_t* sketches, int32_t* lens);

/* ---- -M : read-set k-mer count table ------------------------------------------------------------- */
/* Pass 1 of src/rkmh.cpp:904-910: hash every read and ++counts[hash % read_table_size] for every valid
 * window (count_zeros = 0, mkmh/mkmh.hpp:729-737) or for every hash including the zeros of invalid
 * windows (count_zeros = 1, filter: src/rkmh.cpp:331-336). */
int rkg_count_reset(rkg_ctx* ctx);
int rkg_count_batch(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nreads, int count_zeros);
/* Device pointer and length (in int32 slots) of the read count table, so that a multi-GPU caller can
 * all-reduce it (ncclAllReduce / torch.distributed.all_reduce, sum, int32) between pass 1 and pass 2:
 * the `#pragma omp for` barrier at src/rkmh.cpp:910 is the only global point of the path. */
int rkg_count_table(rkg_ctx* ctx, void** device_ptr, uint64_t* nslots);
/* The same exchange with 12.5x fewer bytes when the memoised path is active (single k <= 12, see below): pass 1 then
 * keeps its increments per k-mer CODE (4^k int32 slots, 64 MB at k = 12, + a few slots for the windows with an invalid
 * base) and folds them into counts[hash % size] afterwards.  The fold is linear, so summing the per-code arrays of all
 * ranks BEFORE the fold gives every rank the table the reference builds (src/rkmh.cpp:904-910).  Call after this
 * rank's rkg_count_batch calls (count_zeros = the flag they used; a rank without reads calls it all the same),
 * all-reduce (sum, int32) device_ptr[0 .. nslots) in place, then go on with rkg_classify_batch / rkg_count_table: the
 * summed array is folded at its next use.  RKG_ERR_STATE when the memoised path is not active (use rkg_count_table). */
int rkg_count_partial(rkg_ctx* ctx, int count_zeros, void** device_ptr, uint64_t* nslots);
/* Persistence of a count table (the reference's --read-kmer-map-file / -p and --ref-kmer-map-file / -q options,
 * src/rkmh.cpp:641-642, 744-761: parsed, loader commented out; docs/todo.md:1 "hash-counter serialization"): the
 * non-zero slots of table `which` (0 = read table, 1 = reference table) as (slot, count) pairs.  rkg_count_export with
 * slots == NULL only returns their number in *n; rkg_count_import clears the table and sets the given slots. */
int rkg_count_export(rkg_ctx* ctx, int which, uint32_t* slots, int32_t* counts, uint64_t cap, uint64_t* n);
int rkg_count_import(rkg_ctx* ctx, int which, const uint32_t* slots, const int32_t* counts, uint64_t n);
/* Parity/debug: counts[keys[i] % read_table_size] for n keys. */
int rkg_count_lookup(rkg_ctx* ctx, const uint64_t* keys, int64_t n, int32_t* out);

/* ---- classification ------------------------------------------------------------------------------ */
/* hash -> (mask_by_frequency with min_occ when min_occ >= 0, mkmh/mkmh.hpp:898-908) -> bottom-s sketch
 * -> intersection with every reference -> argmax.  Replaces the bodies of the per-read loops at
 * src/rkmh.cpp:845-898, 911-948 and 1251-1298.  Host buffers in, host results out. */
int rkg_classify_batch(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nreads,
                       int min_occ, rkg_hit* out);

/* The same for a batch of FIXED-LENGTH reads (Illumina: every record of a chunk has read_len bases, read i is
 * bases[i * read_len .. (i + 1) * read_len)): no offsets array crosses the PCIe link -- the path is bound by the
 * host-to-device copy, and 8 bytes of offset per 150-base read are 5 % of it.  Same results as rkg_classify_batch /
 * rkg_count_batch on the equivalent offsets. */
int rkg_classify_fixed(rkg_ctx* ctx, const char* bases, int64_t read_len, int64_t nreads, int min_occ, rkg_hit* out);
int rkg_count_fixed(rkg_ctx* ctx, const char* bases, int64_t read_len, int64_t nreads, int count_zeros);

/* Packed transfer.  The path is bound by the host-to-device copy, so the library can make the
 * 2-bit codes and the invalid-base mask -- what its ingest kernel makes on the device -- on the HOST, with `threads` host
 * threads (0 = as many as the machine has, at most 16), chunk by chunk inside rkg_classify_batch / rkg_classify_fixed /
 * rkg_count_batch / rkg_count_fixed and the other host-buffer calls, and ship 0.375 bytes per base instead of 1.  Same
 * buffers in, same results out (upper-casing and the validity of a base are all the kernels take from a byte).
 * enable: 0 = off (the default), 1 = every chunk is packed (pays when the link is much slower than the caller's cores),
 * 2 = adaptive: rkg_classify_batch / _fixed and rkg_count_batch / _fixed send a chunk as it is when the link would
 * otherwise idle and pack it while the link is busy with earlier chunks, so that link and cores both stay busy whatever
 * their speeds are (measured on the context's first chunks); 64 MB chunks, four in flight. */
int rkg_set_host_pack(rkg_ctx* ctx, int enable, int threads);
/* What the host-buffer calls sent since the context was created or `reset` was last passed non-zero:
 * stats[0] = host-to-device bytes of chunks sent as they are, [1] = bytes of packed chunks (0.375 per base),
 * [2] = chunks sent as they are, [3] = packed chunks. */
int rkg_transfer_stats(rkg_ctx* ctx, int64_t stats[4], int reset);

/* The same in two steps for callers that keep batches resident in HBM (and for kernel-only timing):
 * upload once, classify any number of times, fetch results. */
int rkg_batch_upload(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nreads, rkg_batch** batch);
/* Run the ingest kernel (upper-case / 2-bit pack / invalid-base mask) on the resident ASCII bases.  The
 * other batch calls run it on demand; on the memoised path (below) the classification and count kernels
 * read the ASCII bases themselves and this call launches nothing.  Asynchronous. */
int rkg_batch_ingest(rkg_ctx* ctx, rkg_batch* batch);
/* Issue this batch's work on the caller's CUDA stream (a cudaStream_t passed as void*; NULL restores the
 * batch's own stream), so that the caller's events bracket the kernels. */
int rkg_batch_set_stream(rkg_ctx* ctx, rkg_batch* batch, void* cuda_stream);
int rkg_batch_classify(rkg_ctx* ctx, rkg_batch* batch, int min_occ);        /* asynchronous */
int rkg_batch_count(rkg_ctx* ctx, rkg_batch* batch, int count_zeros);       /* asynchronous */
int rkg_batch_results(rkg_ctx* ctx, rkg_batch* batch, rkg_hit* out);        /* synchronises  */
int rkg_batch_free(rkg_ctx* ctx, rkg_batch* batch);
int rkg_synchronize(rkg_ctx* ctx);

/* ---- memoised hashing (single k <= 12) ----------------------------------------------------------- */
/* calc_hash (mkmh/mkmh.hpp:505-532) is a pure function of the k-mer; for a single k <= 12 its domain has
 * 4^k <= 16.7 M points, so rkg_load_refs runs the hashing kernel once over every k-mer and the read-side
 * kernels (rkg_classify_batch / rkg_count_batch and their rkg_batch_* forms) look hashes up in an
 * L2-resident table instead of recomputing 2 x MurmurHash3_x64_128 per window.  Results are bit-identical;
 * the read count table (rkg_count_table) keeps the reference's layout.  On by default (environment
 * RKMH_NO_MEMO=1 or rkg_set_memo(ctx, 0) force the Murmur path for every k); a change takes effect at
 * the next rkg_load_refs.  After all-reducing the table obtained from rkg_count_table, call
 * rkg_count_table again before any later all-reduce: the call marks the table as changed.
 * Short reads (every k-mer is in the sketch), at most 256 references and no min_occ take the sieve kernel (sketch
 * membership from a shared-memory Bloom filter, hits through an L2-resident code map; RKMH_NO_SIEVE=1 turns it off);
 * every variant returns the same records. */
int rkg_set_memo(rkg_ctx* ctx, int enable);
int rkg_memo_active(const rkg_ctx* ctx);

/* ---- hpv16: full hash sets instead of sketches ------------------------------------------------------- */
/* The tiered HPV16 classifier (src/rkmh.cpp:2366-2723) compares every hash of a read -- no bottom-s -- with FULL hash
 * sets: the 182 type references (hash_set_intersection_size, mkmh/mkmh.hpp:842-863: distinct non-zero values in
 * common) and the lineage / sublineage specific k-mer tables (sort_by_similarity -> hash_intersection_size against
 * sets of distinct values: the same number).  rkg_load_ref_sets builds the HBM table hash -> set bitset from nsets
 * hash sets given as one array plus nsets + 1 offsets (any order; zeros and duplicates are ignored); it replaces the
 * table of rkg_load_refs, and rkg_classify_batch is refused until rkg_load_refs is called again. */
int rkg_load_ref_sets(rkg_ctx* ctx, const uint64_t* hashes, const int64_t* set_offs, int nsets);
/* Per read: calc_hashes over every k of the list, mask_by_frequency when min_occ >= 0, then count[j] = number of
 * distinct non-zero hashes of the read that set j holds.  out[r] = {first index in [0, ntypes) of the largest count,
 * that count, prev, number of distinct-or-not non-zero hashes}: src/rkmh.cpp:2655-2664; extra [nreads x (nsets -
 * ntypes)] = the counts of the remaining sets, in set order (src/rkmh.cpp:2672-2683). */
int rkg_classify_sets(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nreads, int min_occ, int ntypes,
                      rkg_hit* out, int32_t* extra);

/* ---- call: alignment-free variant calling from k-mer depth (src/rkmh.cpp:1455-1904) ------------------------ */
/* The reference keeps the exact multiplicity of every hash of every read in an unordered_map (read_hash_to_depth,
 * src/rkmh.cpp:1570, 1617-1622; the hash 0 of invalid windows included) and then walks the reference sequence.  Here:
 *   rkg_depth_begin        an empty exact hash -> depth table in HBM with room for expected_windows windows (one k only)
 *   rkg_depth_add          calc_hashes of every read, ++depth[hash]
 *   rkg_depth_of_sequence  depth[hash] of each of the len - k windows of a reference sequence (src/rkmh.cpp:1777)
 *   rkg_depth_rescue       for each listed window start: the depths of its 3k single-base substitutions (position p, the
 *                          three other bases in rotate_snps order, src/rkmh.cpp:1646-1667: out[w][3p + i]) and of the k
 *                          one-base deletions of the (k + 1)-mer that starts one base before it (base d = 1..k removed:
 *                          out[w][3k + d - 1], src/rkmh.cpp:1840-1848); -1 where the reference tries nothing (window 0 has
 *                          no deletions; a base outside ACGT has no substitutions).  out is [nwin x 4k].
 * The sliding-window average, the thresholds and the VCF lines are the caller's (sequential and tiny). */
int rkg_depth_begin(rkg_ctx* ctx, uint64_t expected_windows);
int rkg_depth_add(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nreads);
int rkg_depth_of_sequence(rkg_ctx* ctx, const char* seq, int64_t len, int32_t* depth);
int rkg_depth_rescue(rkg_ctx* ctx, const char* seq, int64_t len, const int64_t* windows, int64_t nwin, int32_t* out);

/* ---- hash command -------------------------------------------------------------------------------- */
/* Bottom-s sketches of nseq records (sorted ascending, duplicates kept, zeros dropped): calc_hashes +
 * minhashes (mkmh/mkmh.hpp:689-711, 931-950), the intended behaviour of `rkmh hash` (src/rkmh.cpp:
 * 2093-2105, README.md:114-117).  sketches is [nseq * s], lens is [nseq]. */
int rkg_sketch_batch(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nseq,
                     int min_occ, uint64_t* sketches, int32_t* lens);
/* Raw calc_hashes output (mkmh/mkmh.hpp:505-532, 689-711) of every record, concatenated; hash_offs
 * receives nseq + 1 offsets into `hashes` (capacity: sum over records and k of max(len - k, 0)). */
int rkg_hash_batch(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nseq,
                   uint64_t* hashes, int64_t* hash_offs);
/* One large record whose k-mer start range [kmer_begin, kmer_end) of the whole record is handled by this
 * context (sequence sharding for `hash` on a multi-GPU box; one k-mer size): `bases` holds record bytes
 * [kmer_begin, min(kmer_end, record_len - k) + k), record_len is the length of the WHOLE record (so
 * that the reference's len-k window count applies only at the true end).  Emits this shard's bottom-s
 * multiset; the caller all-gathers the shards and calls rkg_merge_sketches.  The shard is hashed on the device as
 * one record (two passes: histogram of the top hash bits, then the hashes below the s-th smallest are kept and
 * sorted); a pinned `bases` buffer (rkg_host_alloc) makes the upload run at link speed. */
int rkg_sketch_shard(rkg_ctx* ctx, const char* bases, int64_t nbytes, int64_t kmer_begin, int64_t kmer_end,
                     int64_t record_len, uint64_t* sketch, int32_t* len);
/* Multiset merge of `nparts` sorted sketches (parts is [nparts * s]) keeping the first s. */
int rkg_merge_sketches(rkg_ctx* ctx, const uint64_t* parts, const int32_t* part_lens, int nparts,
                       uint64_t* sketch, int32_t* len);

/* ---- measurement --------------------------------------------------------------------------------- */
/* The host<->device copies of rkg_classify_batch (offs != NULL) or rkg_classify_fixed (offs == NULL, read_len) on the
 * same chunking and streams, WITHOUT the kernels: the ceiling the end-to-end rate can reach on this host.  `out`
 * receives whatever the device hit buffer holds. */
int rkg_copy_probe(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t read_len, int64_t nreads, rkg_hit* out);
/* Number of kernels this context has launched so far. */
int64_t rkg_launch_count(const rkg_ctx* ctx);
/* Per-kernel device time: when enabled every launch is bracketed by CUDA events on its stream. */
int rkg_profile_enable(rkg_ctx* ctx, int on);
/* Fills up to cap entries: kernel name, launches and total milliseconds since the last reset; returns
 * the number of distinct kernels (synchronises the context). */
typedef struct { const char* name; int64_t launches; double ms; } rkg_kernel_time;
int rkg_profile_read(rkg_ctx* ctx, rkg_kernel_time* out, int cap, int reset);

#ifdef __cplusplus
}
#endif
#endif /* RKMH_GPU_H */
