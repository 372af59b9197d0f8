// rk_kernels.cuh -- the sm_100a kernels of the rkmh hot path.
//
//   ingest_kernel            ASCII -> 2-bit codes + invalid-base mask         (mkmh.hpp:231-236, 137-152)
//   hash_kernel              raw calc_hashes output                           (mkmh.hpp:505-532)
//   count_kernel             -M / -I pass 1: ++counts[hash % size]            (mkmh.hpp:713-741, HASHTCounter.cpp:46)
//   sketch_kernel            bottom-s sorted multiset per sequence            (mkmh.hpp:931-976)
//   merge_kernel             multiset merge of sorted sketches                (sharded `hash`)
//   table_* kernels          (hash, rank) -> reference-bitset open-addressing table
//   classify_block_kernel    block per read: hash, select, probe, argmax      (rkmh.cpp:845-898, 911-948)
//   classify_group_kernel    sub-warp group per short read, same result
//
// No tensor cores anywhere: the path is 64-bit integer hashing (INT pipe) plus table probes (L2/HBM).
// The file is also compiled by g++ against tests/emu/cuda_emu.h (RK_EMU) to debug the logic on CPU.
#pragma once
#include "rk_common.cuh"
#include "rk_hash.cuh"

namespace rk {

struct Hit { int ref_id, best, prev, sketch_len; };   // == rkg_hit
template <int V> struct KcTag { static constexpr int value = V; };

// One batch of sequences in device memory.
struct SeqBatch {
    const unsigned char* bases;   // ASCII as parsed, concatenated
    const i64* offs;              // nseq + 1
    const u32* codes;             // 2-bit codes, 16 bases per word (base b at bits 2*(b&15))
    const u32* nmask;             // 1 bit per base, 1 = not one of ACGTacgt, 32 bases per word
    i64 nseq;
    i64 total;                    // number of bases in the batch
    i64 base0;                    // offs[] are offsets into the caller's buffer; the batch starts at base0
};

struct KList { int nk; int k[MAX_NK]; int kv[MAX_NK]; int kmax; int ksum; const u64* lut; };   // kv = kvariant_of(k); lut: rk_hash.cuh tables (global memory)

struct CountTable {
    int* counts;      // nullptr = absent
    FastMod mod;
};

// (hash, rank) -> reference bitset.  Rank-0 keys live in an open-addressing table of `cap` entries of
// `stride` = 1 + nb 64-bit words {key, bits[nb]}; key 0 = empty (0 is never a sketch member).
// Hashes that occur more than once inside one reference sketch own maxmult[slot]-1 extra
// bitset rows starting at dup_base[slot]: row r-1 holds the references in which the hash occurs > r times.
struct RefTable {
    u64* entries;
    u32* maxmult;
    u32* dup_base;
    u64* dup_bits;
    u32 cap_mask;
    u32 stride;
    u32 nb;
    u32 nrefs;
    const u32* filter;     // one-hash Bloom prefilter over the rank-0 keys, 2^filter_log2 bits
    u32 filter_log2;
};

// ------------------------------------------------------------------------------------------------
// ingest: 32 bases per thread.  to_upper (mkmh.hpp:231-236) only matters for a..t here: a base is
// valid iff it is one of ACGTacgt (valid_dna[], mkmh.hpp:137-152) and both cases share a code.
__global__ void __launch_bounds__(256)
ingest_kernel(const unsigned char* __restrict__ bases, i64 total, u32* __restrict__ codes, u32* __restrict__ nmask,
              i64 nwords32) {
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 w = t; w < nwords32; w += stride) {
        i64 b0 = w * 32;
        u32 lo = 0, hi = 0, m = 0;
        if (b0 + 32 <= total) {
            const uint4* p = reinterpret_cast<const uint4*>(bases + b0);
            uint4 v0 = p[0], v1 = p[1];
            u32 wd[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    u32 c = (wd[q] >> (8 * j)) & 0xFFu;
                    int idx = q * 4 + j;
                    u32 code = base_code(c);
                    if (idx < 16) lo |= code << (2 * idx); else hi |= code << (2 * (idx - 16));
                    m |= (base_valid(c) ? 0u : 1u) << idx;
                }
            }
        } else {
            for (int idx = 0; idx < 32; ++idx) {
                u32 c = (b0 + idx < total) ? bases[b0 + idx] : 0u;
                u32 code = base_code(c);
                if (idx < 16) lo |= code << (2 * idx); else hi |= code << (2 * (idx - 16));
                m |= (base_valid(c) ? 0u : 1u) << idx;
            }
        }
        codes[2 * w] = lo;
        codes[2 * w + 1] = hi;
        nmask[w] = m;
    }
}

// unpack: the inverse of ingest for batches that crossed the link packed (rkg_set_host_pack): 32 bases per thread, the
// letter of each code (A C T G), 'N' for every base the mask flags.  What the memoised kernels read is then the
// upper-cased sequence with every invalid byte as N: same codes, same validity, hence the same hashes.
__global__ void __launch_bounds__(256)
unpack_kernel(const u32* __restrict__ codes, const u32* __restrict__ nmask, unsigned char* __restrict__ bases, i64 nwords32) {
    for (i64 w = (i64)blockIdx.x * blockDim.x + threadIdx.x; w < nwords32; w += (i64)gridDim.x * blockDim.x) {
        const u64 c = ((u64)codes[2 * w + 1] << 32) | codes[2 * w];
        const u32 m = nmask[w];
        u32 out[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            u32 word = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = 4 * q + j;
                const u32 letter = ((m >> i) & 1u) ? (u32)'N' : ((CODE_TO_ASCII >> (8 * (u32)((c >> (2 * i)) & 3ULL))) & 0xFFu);
                word |= letter << (8 * j);
            }
            out[q] = word;
        }
        uint4* dst = reinterpret_cast<uint4*>(bases + w * 32);
        dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
        dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
    }
}

// ------------------------------------------------------------------------------------------------
// Sequential reader of packed bases for one thread: a 64-bit shift register of 2-bit codes and a 32-bit
// one of invalid-base flags, refilled every 32 bases.  Threads that walk their chunks in lock step refill
// on the same iteration, so the refill branch does not diverge.
struct BaseReader {
    const u32* codes; const u32* nmask;
    i64 pos;            // position of the first base held in the shift registers
    u64 cw; u32 mw; int left;
    RK_D void refill() {
        const u32* c = codes + (pos >> 4);
        const u32 sh = 2u * ((u32)pos & 15u);
        u32 c0 = c[0], c1 = c[1], c2 = c[2];
        u32 lo = __funnelshift_r(c0, c1, sh), hi = __funnelshift_r(c1, c2, sh);
        cw = ((u64)hi << 32) | lo;
        const u32* m = nmask + (pos >> 5);
        mw = __funnelshift_r(m[0], m[1], (u32)pos & 31u);
        left = 32;
    }
    RK_D void init(const u32* c, const u32* m, i64 p) { codes = c; nmask = m; pos = p; refill(); }
    RK_D void next(u32& code, bool& invalid) {
        if (left == 0) { pos += 32; refill(); }
        code = (u32)cw & 3u;
        invalid = mw & 1u;
        cw >>= 2; mw >>= 1;
        --left;
    }
};

RK_D u64 apply_min_occ(u64 h, const CountTable& ct, int min_occ) {
    // mask_by_frequency, mkmh.hpp:898-908 (max_occ = UINT32_MAX): keep iff min_occ <= counts[h % size]
    if (ct.counts == nullptr || h == 0) return h;
    int freq = ct.counts[fastmod(h, ct.mod)];
    return (min_occ <= freq) ? h : 0;
}

// last loadable word of a batch's codes[] / nmask[] (batch_fill pads 8 words): the readers clamp their loads to it
RK_D i64 codes_last_word(const SeqBatch& sb) { return 2 * ((sb.total + 31) / 32) + 7; }
RK_D i64 nmask_last_word(const SeqBatch& sb) { return (sb.total + 31) / 32 + 7; }

RK_D i64 nkmers(i64 len, int k) { return len - k > 0 ? len - k : 0; }   // len - k windows, SURVEY F7
RK_D i64 nkmers_total(i64 len, const KList& kl) {
    if (len >= kl.kmax) return (i64)kl.nk * len - kl.ksum;
    i64 n = 0;
    for (int j = 0; j < kl.nk; ++j) n += nkmers(len, kl.k[j]);
    return n;
}

// ------------------------------------------------------------------------------------------------
// hash_kernel: one warp per sequence (grid-stride), raw calc_hashes output for ONE k of the list
// (launched once per k); out_offs[seq] + koff(seq) is where this k's hashes of seq start.
template <int KC>
__global__ void __launch_bounds__(256)
hash_kernel(SeqBatch sb, const __grid_constant__ KList kl, int kidx, u64* __restrict__ out, const i64* __restrict__ out_offs) {
    const u64* lut = hash_lut_load(kl.lut);
    int lane = threadIdx.x & 31;
    i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const int k = kl.k[kidx];
    for (i64 s = warp; s < sb.nseq; s += nwarps) {
        i64 beg = sb.offs[s] - sb.base0, len = sb.offs[s + 1] - sb.offs[s];
        i64 koff = 0;
        for (int j = 0; j < kidx; ++j) koff += nkmers(len, kl.k[j]);
        u64* dst = out + out_offs[s] + koff;
        team_hash_k<KC>(kl.kv[kidx], lut, sb.codes, sb.nmask, codes_last_word(sb), nmask_last_word(sb), beg, nkmers(len, k), k, lane, 32,
                        [&](i64 i, u64 h) { dst[i] = h; });
    }
}

// ------------------------------------------------------------------------------------------------
// count_kernel: pass 1 of -M (and of -I on references).  One warp per sequence, grid-stride.
//   count_zeros = 0: only valid windows are counted (mkmh.hpp:729-737)
//   count_zeros = 1: every window, an invalid one counts slot 0 % size (rkmh.cpp:331-336)
template <int KCSEL>
__global__ void __launch_bounds__(256)
count_kernel(SeqBatch sb, const __grid_constant__ KList kl, CountTable ct, int count_zeros) {
    const u64* lut = hash_lut_load(kl.lut);
    int lane = threadIdx.x & 31;
    i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 s = warp; s < sb.nseq; s += nwarps) {
        i64 beg = sb.offs[s] - sb.base0, len = sb.offs[s + 1] - sb.offs[s];
        for (int j = 0; j < kl.nk; ++j) {
            const int k = kl.k[j];
            team_hash_k<KCSEL>(kl.kv[j], lut, sb.codes, sb.nmask, codes_last_word(sb), nmask_last_word(sb), beg, nkmers(len, k), k, lane, 32,
                               [&](i64, u64 h) {
                                   if (h != 0 || count_zeros) atomicAdd(&ct.counts[fastmod(h, ct.mod)], 1);
                               });
        }
    }
}

// Variant for src/rkmh.cpp:347-355: one increment per DISTINCT hash per sequence (zero included).  The
// input is the per-sequence sorted full hash array (sketch_kernel with s = n keeps zeros out, so this
// runs on the raw hashes sorted by sort_kernel).  Only reference sets take this path (tiny inputs).
__global__ void __launch_bounds__(256)
count_distinct_kernel(const u64* __restrict__ sorted, const i64* __restrict__ offs, i64 nseq, CountTable ct) {
    for (i64 s = blockIdx.x; s < nseq; s += gridDim.x) {
        const u64* h = sorted + offs[s];
        i64 n = offs[s + 1] - offs[s];
        for (i64 i = threadIdx.x; i < n; i += blockDim.x)
            if (i == 0 || h[i] != h[i - 1]) atomicAdd(&ct.counts[fastmod(h[i], ct.mod)], 1);
    }
}

// Persistence of a lossy count table: its non-zero slots as (slot, count) pairs, and back.
__global__ void __launch_bounds__(256)
count_export_kernel(const int* __restrict__ counts, u64 nslots, u32* __restrict__ slots, int* __restrict__ vals, u64 cap,
                    unsigned long long* __restrict__ n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += (u64)gridDim.x * blockDim.x) {
        const int c = counts[i];
        if (c == 0) continue;
        const unsigned long long at = atomicAdd(n, 1ULL);
        if (at < cap) { slots[at] = (u32)i; vals[at] = c; }
    }
}
__global__ void __launch_bounds__(256)
count_import_kernel(int* __restrict__ counts, u64 nslots, const u32* __restrict__ slots, const int* __restrict__ vals, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        if ((u64)slots[i] < nslots) counts[slots[i]] = vals[i];
}

// ------------------------------------------------------------------------------------------------
// Block-wide helpers.

// In-place ascending bitonic sort of n = 2^m values by the whole block (a may be shared or global).
RK_D void block_bitonic_sort(u64* a, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                int lo = ((t / stride) * (stride << 1)) + (t % stride);
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                u64 x = a[lo], y = a[hi];
                if ((x > y) == up) { a[lo] = y; a[hi] = x; }
            }
        }
    }
    __syncthreads();
}

// Bitonic merge: a[0..n) is bitonic (n = 2^m), result ascending.
RK_D void block_bitonic_merge(u64* a, int n) {
    for (int stride = n >> 1; stride > 0; stride >>= 1) {
        __syncthreads();
        for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
            int lo = ((t / stride) * (stride << 1)) + (t % stride);
            int hi = lo + stride;
            u64 x = a[lo], y = a[hi];
            if (x > y) { a[lo] = y; a[hi] = x; }
        }
    }
    __syncthreads();
}

// Result of selecting the `want` smallest of the non-zero values of an array:
//   member(h) = hp < prefix || (hp == prefix && (whole || one of the first `need` takers)), hp = h >> shift
struct Select { u64 prefix; int shift; int whole; int need; };

struct SelectScratch { u32 hist[256]; Select sel; int take; };

// MSB-first radix select over the non-zero entries of h[0..n).  Requires want < #non-zero.
RK_D void block_radix_select(const u64* h, i64 n, int want, SelectScratch* ss) {
    u64 prefix = 0; int need = want;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int b = threadIdx.x; b < 256; b += blockDim.x) ss->hist[b] = 0;
        __syncthreads();
        for (i64 i = threadIdx.x; i < n; i += blockDim.x) {
            u64 v = h[i];
            if (v == 0) continue;
            if (shift == 56 || (v >> (shift + 8)) == prefix) atomicAdd(&ss->hist[(u32)(v >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {                       // warp 0: locate the bin where the running count reaches `need`
            int lane = threadIdx.x;
            u32 part = 0;
            for (int b = 0; b < 8; ++b) part += ss->hist[lane * 8 + b];
            u32 incl = part;
            for (int d = 1; d < 32; d <<= 1) {
                u32 o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            u32 excl = incl - part;
            if (excl < (u32)need && (u32)need <= incl) {
                u32 run = excl;
                for (int b = 0; b < 8; ++b) {
                    u32 c = ss->hist[lane * 8 + b];
                    if (run + c >= (u32)need) {
                        ss->sel.prefix = (prefix << 8) | (u64)(lane * 8 + b);
                        ss->sel.shift = shift;
                        ss->sel.whole = (run + c == (u32)need);
                        ss->sel.need = need - (int)run;
                        break;
                    }
                    run += c;
                }
            }
        }
        __syncthreads();
        prefix = ss->sel.prefix;
        need = ss->sel.need;
        if (ss->sel.whole) break;
        // shift == 0 and not whole: `need` copies of the value `prefix` are taken
    }
    if (threadIdx.x == 0) ss->take = 0;
    __syncthreads();
}

RK_D bool select_member(u64 h, const Select& sel, int* take) {
    u64 hp = h >> sel.shift;
    if (hp < sel.prefix) return true;
    if (hp > sel.prefix) return false;
    if (sel.whole) return true;
    return atomicAdd(take, 1) < sel.need;
}

// ------------------------------------------------------------------------------------------------
// sketch_kernel: one block per sequence (grid-stride).  hashes -> (count filter) -> bottom-s sorted.
//   filter_mode 0: none; 1: keep counts[h % size] <= occ (-I, mkmh.hpp:954-976);
//               2: keep counts[h % size] >= occ (-M, mkmh.hpp:898-908)
// scratch: per block `scratch_stride` u64 for the hashes, then `sort_cap` u64 for the sort when it does
// not fit in shared memory.  Dynamic smem: sort_cap_smem u64 (0 = sort in global scratch).
template <int KCSEL>
__global__ void __launch_bounds__(512)
sketch_kernel(SeqBatch sb, const __grid_constant__ KList kl, int s, CountTable ct, int filter_mode, int occ,
              u64* __restrict__ scratch, i64 scratch_stride, int sort_cap, int sort_in_smem,
              u64* __restrict__ sketches, int* __restrict__ lens, u64* __restrict__ sorted_all, const i64* __restrict__ all_offs) {
    const u64* lut = hash_lut_load(kl.lut);
#ifdef RK_EMU
    u64* dsm = reinterpret_cast<u64*>(EMU_DYN_SMEM());
#else
    extern __shared__ __align__(16) unsigned char dyn_smem_raw[];
    u64* dsm = reinterpret_cast<u64*>(dyn_smem_raw);
#endif
    __shared__ SelectScratch ss;
    __shared__ int s_nz, s_out;
    u64* H = scratch + (i64)blockIdx.x * (scratch_stride + sort_cap);
    u64* S = sort_in_smem ? dsm : (H + scratch_stride);

    for (i64 q = blockIdx.x; q < sb.nseq; q += gridDim.x) {
        i64 beg = sb.offs[q] - sb.base0, len = sb.offs[q + 1] - sb.offs[q];
        i64 n = nkmers_total(len, kl);
        if (threadIdx.x == 0) { s_nz = 0; s_out = 0; }
        __syncthreads();
        int local_nz = 0;
        i64 base = 0;
        for (int j = 0; j < kl.nk; ++j) {
            const int k = kl.k[j];
            i64 nj = nkmers(len, k);
            team_hash_k<KCSEL>(kl.kv[j], lut, sb.codes, sb.nmask, codes_last_word(sb), nmask_last_word(sb), beg, nj, k, threadIdx.x, blockDim.x,
                               [&](i64 i, u64 h) {
                                   if (h != 0 && filter_mode != 0) {
                                       int freq = ct.counts[fastmod(h, ct.mod)];
                                       bool keep = (filter_mode == 1) ? (freq >= 0 && (u32)freq <= (u32)occ) : (occ <= freq);
                                       if (!keep) h = 0;
                                   }
                                   H[base + i] = h;
                                   local_nz += (h != 0);
                               });
            base += nj;
        }
        if (local_nz) atomicAdd(&s_nz, local_nz);
        __syncthreads();
        if (sorted_all != nullptr) {
            // full ascending sort of every hash, zeros included (input of count_distinct_kernel);
            // only used for reference sets, where n is small: sort in global scratch
            int cap = 1; while (cap < n) cap <<= 1;
            u64* A = H;  // scratch_stride >= cap is guaranteed by the host for this mode
            for (i64 i = n + threadIdx.x; i < cap; i += blockDim.x) A[i] = ~0ULL;
            block_bitonic_sort(A, cap);
            for (i64 i = threadIdx.x; i < n; i += blockDim.x) sorted_all[all_offs[q] + i] = A[i];
            __syncthreads();
            continue;
        }
        const int nz = s_nz;
        const int L = nz < s ? nz : s;
        bool all = (nz <= s);
        if (!all) block_radix_select(H, n, s, &ss);
        // compact the members into S, pad, sort
        int cap = 1; while (cap < L) cap <<= 1;
        for (i64 i = threadIdx.x; i < n; i += blockDim.x) {
            u64 h = H[i];
            if (h == 0) continue;
            if (all || select_member(h, ss.sel, &ss.take)) S[atomicAdd(&s_out, 1)] = h;
        }
        __syncthreads();
        for (int i = L + threadIdx.x; i < cap; i += blockDim.x) S[i] = ~0ULL;
        block_bitonic_sort(S, cap);
        for (int i = threadIdx.x; i < L; i += blockDim.x) sketches[q * (i64)s + i] = S[i];
        if (threadIdx.x == 0) lens[q] = L;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Bottom-s sketch of ONE long record (a shard of it) without materialising its hashes: `rkmh hash` on a
// chromosome-scale sequence (minhashes(calc_hashes()), mkmh.hpp:505-532, 931-950).
//   shard_scan_kernel   ONE hashing pass when the hashes behave like hashes: keep the non-zero hashes at or under a limit
//                       guessed from s / n (1.5 x the expected s-th smallest), count the non-zero ones.  At least s kept
//                       (and room for them): they hold the sketch.  Otherwise (repeats, mostly invalid windows) the exact
//                       two-pass route below runs on the same device-resident codes:
//   shard_hist_kernel   hash every window, histogram of the top 16 hash bits (non-zero hashes only)
//   shard_pick_kernel   the bin in which the running count reaches s, and the count through that bin
//   shard_emit_kernel   hash again, keep the hashes whose top 16 bits are <= that bin (s plus one bin's worth)
//   shard_sort_kernel   bitonic sort of the kept hashes; the first min(s, kept) are the sketch (duplicates kept)
constexpr int SHARD_BINS = 65536;
constexpr i64 SHARD_SEG = 4096;                 // windows per warp task

template <int KCSEL>
__global__ void __launch_bounds__(256)
shard_hist_kernel(const u32* __restrict__ codes, const u32* __restrict__ nmask, i64 last_word, i64 last_mword, i64 n, const __grid_constant__ KList kl,
                  u32* __restrict__ hist) {
    const u64* lut = hash_lut_load(kl.lut);
    const int lane = threadIdx.x & 31, k = kl.k[0];
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 w0 = warp * SHARD_SEG; w0 < n; w0 += nwarps * SHARD_SEG) {
        const i64 cnt = min(SHARD_SEG, n - w0);
        team_hash_k<KCSEL>(kl.kv[0], lut, codes, nmask, last_word, last_mword, w0, cnt, k, lane, 32,
                           [&](i64, u64 h) { if (h) atomicAdd(&hist[(u32)(h >> 48)], 1u); });
    }
}

// windows [w_begin, w_end) of the record; cnt[0] += kept hashes (cand[] holds the first `cap` of them), cnt[1] += non-zero hashes
template <int KCSEL>
__global__ void __launch_bounds__(256)
shard_scan_kernel(const u32* __restrict__ codes, const u32* __restrict__ nmask, i64 last_word, i64 last_mword, i64 w_begin, i64 w_end,
                  const __grid_constant__ KList kl, u64 limit, u64* __restrict__ cand, unsigned long long* __restrict__ cnt, u64 cap) {
    const u64* lut = hash_lut_load(kl.lut);
    const int lane = threadIdx.x & 31, k = kl.k[0];
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    unsigned long long nz = 0;
    for (i64 w0 = w_begin + warp * SHARD_SEG; w0 < w_end; w0 += nwarps * SHARD_SEG) {
        const i64 c = min(SHARD_SEG, w_end - w0);
        team_hash_k<KCSEL>(kl.kv[0], lut, codes, nmask, last_word, last_mword, w0, c, k, lane, 32, [&](i64, u64 h) {
            if (h) {
                ++nz;
                if (h <= limit) {
                    const unsigned long long at = atomicAdd(&cnt[0], 1ULL);
                    if (at < cap) cand[at] = h;
                }
            }
        });
    }
    if (nz) atomicAdd(&cnt[1], nz);
}

// out[0] = bin, out[1] = number of non-zero hashes with top 16 bits <= bin (all of them when fewer than s exist)
__global__ void __launch_bounds__(1024)
shard_pick_kernel(const u32* __restrict__ hist, u64 s, u64* __restrict__ out) {
    __shared__ u64 part[1024];
    const int per = SHARD_BINS / 1024, b0 = (int)threadIdx.x * per;
    u64 sum = 0;
    for (int i = 0; i < per; ++i) sum += hist[b0 + i];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const u64 v = ((int)threadIdx.x >= d) ? part[threadIdx.x - d] : 0ULL;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    const u64 incl = part[threadIdx.x], excl = incl - sum, total = part[1023];
    if (total < s) { if (threadIdx.x == 0) { out[0] = SHARD_BINS - 1; out[1] = total; } return; }
    if (excl < s && s <= incl) {
        u64 run = excl;
        for (int i = 0; i < per; ++i) {
            run += hist[b0 + i];
            if (run >= s) { out[0] = (u64)(b0 + i); out[1] = run; break; }
        }
    }
}

template <int KCSEL>
__global__ void __launch_bounds__(256)
shard_emit_kernel(const u32* __restrict__ codes, const u32* __restrict__ nmask, i64 last_word, i64 last_mword, i64 n, const __grid_constant__ KList kl,
                  const u64* __restrict__ pick,
                  u64* __restrict__ cand, unsigned long long* __restrict__ ncand, u64 cap) {
    const u64* lut = hash_lut_load(kl.lut);
    const int lane = threadIdx.x & 31, k = kl.k[0];
    const u32 bin = (u32)pick[0];
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 w0 = warp * SHARD_SEG; w0 < n; w0 += nwarps * SHARD_SEG) {
        const i64 cnt = min(SHARD_SEG, n - w0);
        team_hash_k<KCSEL>(kl.kv[0], lut, codes, nmask, last_word, last_mword, w0, cnt, k, lane, 32, [&](i64, u64 h) {
            if (h && (u32)(h >> 48) <= bin) {
                const unsigned long long at = atomicAdd(ncand, 1ULL);
                if (at < cap) cand[at] = h;
            }
        });
    }
}

// a[0, m) are the kept hashes, a has room for P = 2^j >= m values.  in_smem: the block has P * 8 bytes of dynamic shared
// memory and sorts there (a 16 k-entry sort in global memory took 1.8 ms, a fifth of a 400 Mbp shard's time).
__global__ void __launch_bounds__(1024)
shard_sort_kernel(u64* __restrict__ a, int m, int P, int in_smem) {
#ifdef RK_EMU
    u64* dsm = reinterpret_cast<u64*>(EMU_DYN_SMEM());
#else
    extern __shared__ __align__(16) unsigned char dyn_smem_raw[];
    u64* dsm = reinterpret_cast<u64*>(dyn_smem_raw);
#endif
    u64* w = in_smem ? dsm : a;
    for (int i = threadIdx.x; i < P; i += blockDim.x) w[i] = i < m ? a[i] : ~0ULL;
    __syncthreads();
    block_bitonic_sort(w, P);
    if (in_smem) for (int i = threadIdx.x; i < m; i += blockDim.x) a[i] = w[i];
}

// merge_kernel: one block merges `nparts` sorted sketches (rows of `parts`, row stride s) into the
// first s of their multiset union.  P = 2^m >= s; smem holds P u64.
__global__ void __launch_bounds__(512)
merge_kernel(const u64* __restrict__ parts, const int* __restrict__ part_lens, int nparts, int s, int P,
             u64* __restrict__ out, int* __restrict__ out_len, u64* __restrict__ gscratch, int in_smem) {
#ifdef RK_EMU
    u64* dsm = reinterpret_cast<u64*>(EMU_DYN_SMEM());
#else
    extern __shared__ __align__(16) unsigned char dyn_smem_raw[];
    u64* dsm = reinterpret_cast<u64*>(dyn_smem_raw);
#endif
    u64* A = in_smem ? dsm : gscratch;
    int total = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) A[i] = (i < part_lens[0]) ? parts[i] : ~0ULL;
    total = part_lens[0];
    __syncthreads();
    for (int p = 1; p < nparts; ++p) {
        // A ascending (padded with +inf), B = part p ascending: min(A[i], B[P-1-i]) is bitonic and holds
        // the P smallest of the union
        const u64* B = parts + (i64)p * s;
        int lb = part_lens[p];
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            int j = P - 1 - i;
            u64 b = (j < lb) ? B[j] : ~0ULL;
            u64 a = A[i];
            A[i] = a < b ? a : b;
        }
        block_bitonic_merge(A, P);
        total += lb;
    }
    int L = total < s ? total : s;
    for (int i = threadIdx.x; i < L; i += blockDim.x) out[i] = A[i];
    if (threadIdx.x == 0) *out_len = L;
}

// ------------------------------------------------------------------------------------------------
// Reference table construction from the sorted reference sketches.
RK_D u32 table_slot(u64 h, u32 cap_mask) { return (u32)h & cap_mask; }
RK_D u32 filter_bit(u64 h, u32 filter_log2) { return (u32)(h >> (64 - filter_log2)); }

// pass A: insert rank-0 keys, OR the reference bit, record the maximum multiplicity per key
__global__ void __launch_bounds__(256)
table_insert_kernel(const u64* __restrict__ sketches, const int* __restrict__ lens, int nrefs, int s, RefTable tb,
                    u32* __restrict__ filter_rw) {
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    i64 total = (i64)nrefs * s;
    for (; t < total; t += (i64)gridDim.x * blockDim.x) {
        int j = (int)(t / s), i = (int)(t % s);
        if (i >= lens[j]) continue;
        const u64* sk = sketches + (i64)j * s;
        u64 v = sk[i];
        int rank = 0;
        while (rank < i && sk[i - 1 - rank] == v) ++rank;
        u32 slot = table_slot(v, tb.cap_mask);
        for (;;) {
            u64* e = tb.entries + (i64)slot * tb.stride;
            u64 old = atomicCAS(e, 0ULL, v);
            if (old == 0ULL || old == v) break;
            slot = (slot + 1) & tb.cap_mask;
        }
        if (rank == 0) {
            atomicOr(tb.entries + (i64)slot * tb.stride + 1 + (j >> 6), 1ULL << (j & 63));
            u32 fb = filter_bit(v, tb.filter_log2);
            atomicOr(&filter_rw[fb >> 5], 1u << (fb & 31));
        }
        atomicMax(&tb.maxmult[slot], (u32)rank + 1u);
    }
}

// pass B: allocate maxmult-1 bitset rows for keys with duplicates
__global__ void __launch_bounds__(256)
table_alloc_dup_kernel(RefTable tb, u32* __restrict__ cursor) {
    u32 cap = tb.cap_mask + 1;
    for (u32 slot = blockIdx.x * blockDim.x + threadIdx.x; slot < cap; slot += gridDim.x * blockDim.x) {
        u32 m = tb.maxmult[slot];
        if (m >= 2) tb.dup_base[slot] = atomicAdd(cursor, m - 1);
    }
}

// pass C: fill the duplicate rows
__global__ void __launch_bounds__(256)
table_fill_dup_kernel(const u64* __restrict__ sketches, const int* __restrict__ lens, int nrefs, int s, RefTable tb) {
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    i64 total = (i64)nrefs * s;
    for (; t < total; t += (i64)gridDim.x * blockDim.x) {
        int j = (int)(t / s), i = (int)(t % s);
        if (i >= lens[j]) continue;
        const u64* sk = sketches + (i64)j * s;
        u64 v = sk[i];
        int rank = 0;
        while (rank < i && sk[i - 1 - rank] == v) ++rank;
        if (rank == 0) continue;
        u32 slot = table_slot(v, tb.cap_mask);
        while (tb.entries[(i64)slot * tb.stride] != v) slot = (slot + 1) & tb.cap_mask;
        u64* row = tb.dup_bits + ((i64)tb.dup_base[slot] + rank - 1) * tb.nb;
        atomicOr(row + (j >> 6), 1ULL << (j & 63));
    }
}

// Set semantics (hpv16: hash_set_intersection_size, mkmh/mkmh.hpp:842-863, against FULL hash sets): every non-zero
// element of set j, in any order, duplicates allowed, gets bit j; no duplicate rows (a hash counts once per set).
__global__ void __launch_bounds__(256)
table_insert_set_kernel(const u64* __restrict__ hashes, const i64* __restrict__ set_offs, int nsets, RefTable tb,
                        u32* __restrict__ filter_rw) {
    const i64 total = set_offs[nsets];
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
        const u64 v = hashes[t];
        if (v == 0) continue;
        int lo = 0, hi = nsets;                                  // the set this element belongs to
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (set_offs[mid] <= t) lo = mid; else hi = mid; }
        const int j = lo;
        u32 slot = table_slot(v, tb.cap_mask);
        for (;;) {
            u64* e = tb.entries + (i64)slot * tb.stride;
            const u64 old = atomicCAS(e, 0ULL, v);
            if (old == 0ULL || old == v) break;
            slot = (slot + 1) & tb.cap_mask;
        }
        atomicOr(tb.entries + (i64)slot * tb.stride + 1 + (j >> 6), 1ULL << (j & 63));
        const u32 fb = filter_bit(v, tb.filter_log2);
        atomicOr(&filter_rw[fb >> 5], 1u << (fb & 31));
        tb.maxmult[slot] = 1u;
    }
}

// Probe: returns the slot of h or -1.
RK_D int table_find(const RefTable& tb, u64 h) {
    u32 slot = table_slot(h, tb.cap_mask);
    for (;;) {
        u64 key = tb.entries[(i64)slot * tb.stride];
        if (key == h) return (int)slot;
        if (key == 0) return -1;
        slot = (slot + 1) & tb.cap_mask;
    }
}

// The bitset row for (slot, rank), or nullptr when the hash occurs <= rank times in every reference.
RK_D const u64* table_row(const RefTable& tb, int slot, u32 rank) {
    if (rank == 0) return tb.entries + (i64)slot * tb.stride + 1;
    if (rank >= tb.maxmult[slot]) return nullptr;
    return tb.dup_bits + ((i64)tb.dup_base[slot] + rank - 1) * tb.nb;
}

// ------------------------------------------------------------------------------------------------
// Block-wide argmax of src/rkmh.cpp:874-883 over counts[0..R): first index of the maximum, and the
// maximum of the counts before it (-1 when the winner is index 0).
struct ArgmaxScratch { u64 key; int prev; };

RK_D void block_argmax(const u32* counts, int R, ArgmaxScratch* as, Hit* out, int sketch_len) {
    if (threadIdx.x == 0) { as->key = 0; as->prev = -1; }
    __syncthreads();
    u64 best = 0;
    for (int j = threadIdx.x; j < R; j += blockDim.x) {
        u64 key = ((u64)counts[j] << 32) | (u64)(0xFFFFFFFFu - (u32)j);
        best = key > best ? key : best;
    }
    if (best) atomicMax(&as->key, best);
    __syncthreads();
    u64 key = as->key;
    int id = (int)(0xFFFFFFFFu - (u32)(key & 0xFFFFFFFFu));
    int prev = -1;
    for (int j = threadIdx.x; j < id; j += blockDim.x) prev = max(prev, (int)counts[j]);
    if (prev >= 0) atomicMax(&as->prev, prev);
    __syncthreads();
    if (threadIdx.x == 0) {
        out->ref_id = id; out->best = (int)(key >> 32); out->prev = as->prev; out->sketch_len = sketch_len;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// classify_block_kernel: one block per read (grid-stride over the *nlist_ptr entries of `list`, or over all
// nseq reads when list is null).  Exact for every input: any length, duplicates inside the read's sketch, s smaller or larger
// than the number of k-mers.
//
// Dynamic shared memory layout (u64 units first): [hbuf_cap u64 hashes][hs_cap u32 keys][hs_cap u32 counts]
// [rpad u32 per-reference counts].  When hbuf_cap == 0 or the read has more k-mers than hbuf_cap the hashes
// go to `scratch` (scratch_stride u64 per block).  hs_in_smem == 0 puts the hit set in scratch as well.
template <int KCSEL>
__global__ void __launch_bounds__(512)
classify_block_kernel(SeqBatch sb, const __grid_constant__ KList kl, int s, CountTable ct, int min_occ, RefTable tb,
                      const int* __restrict__ list, const int* __restrict__ nlist_ptr, i64 nseq,
                      int hbuf_cap, int hs_cap, int hs_in_smem, int rpad,
                      u64* __restrict__ scratch, i64 scratch_stride, Hit* __restrict__ out,
                      int set_mode = 0, int nargmax = 0, int* __restrict__ extra = nullptr) {
    const u64* lut = hash_lut_load(kl.lut);
#ifdef RK_EMU
    unsigned char* dsm = reinterpret_cast<unsigned char*>(EMU_DYN_SMEM());
#else
    extern __shared__ __align__(16) unsigned char dsm[];
#endif
    __shared__ SelectScratch ss;
    __shared__ ArgmaxScratch as;
    __shared__ int s_nz;
    u64* Hs = reinterpret_cast<u64*>(dsm);
    u32* hs_smem = reinterpret_cast<u32*>(Hs + hbuf_cap);
    u32* counts = hs_smem + (hs_in_smem ? 2 * hs_cap : 0);
    u64* my_scratch = scratch + (i64)blockIdx.x * (scratch_stride + hs_cap);   // hashes, then hit set (u64-sized slots)
    u32* hs_keys = hs_in_smem ? hs_smem : reinterpret_cast<u32*>(my_scratch + scratch_stride);
    u32* hs_cnt = hs_keys + hs_cap;
    const u32 hs_mask = (u32)hs_cap - 1;

    const i64 nlist = list ? (i64)*nlist_ptr : nseq;
    for (i64 q = blockIdx.x; q < nlist; q += gridDim.x) {
        const i64 rid = list ? (i64)list[q] : q;
        i64 beg = sb.offs[rid] - sb.base0, len = sb.offs[rid + 1] - sb.offs[rid];
        i64 n = nkmers_total(len, kl);
        u64* H = (n <= hbuf_cap) ? Hs : my_scratch;
        if (threadIdx.x == 0) s_nz = 0;
        for (int i = threadIdx.x; i < rpad; i += blockDim.x) counts[i] = 0;
        for (int i = threadIdx.x; i < 2 * hs_cap; i += blockDim.x) hs_keys[i] = 0;
        __syncthreads();
        int local_nz = 0;
        i64 base = 0;
        for (int j = 0; j < kl.nk; ++j) {
            const int k = kl.k[j];
            i64 nj = nkmers(len, k);
            team_hash_k<KCSEL>(kl.kv[j], lut, sb.codes, sb.nmask, codes_last_word(sb), nmask_last_word(sb), beg, nj, k, threadIdx.x, blockDim.x,
                               [&](i64 i, u64 h) {
                                   h = apply_min_occ(h, ct, min_occ);
                                   H[base + i] = h;
                                   local_nz += (h != 0);
                               });
            base += nj;
        }
        if (local_nz) atomicAdd(&s_nz, local_nz);
        __syncthreads();
        const int nz = s_nz;
        const int L = nz < s ? nz : s;
        const bool all = (nz <= s);
        if (!all) block_radix_select(H, n, s, &ss);
        for (i64 i = threadIdx.x; i < n; i += blockDim.x) {
            u64 h = H[i];
            if (h == 0) continue;
            if (!(all || select_member(h, ss.sel, &ss.take))) continue;
            u32 fb = filter_bit(h, tb.filter_log2);
            if (!((tb.filter[fb >> 5] >> (fb & 31)) & 1u)) continue;
            int slot = table_find(tb, h);
            if (slot < 0) continue;
            // multiplicity inside the read's sketch: the r-th copy of a hash counts for the references
            // that hold more than r copies (sum of min(multiplicities), mkmh.hpp:822-840)
            u32 key = (u32)slot + 1u, p = (key * 0x9E3779B1u) & hs_mask;
            for (;;) {
                u32 old = atomicCAS(&hs_keys[p], 0u, key);
                if (old == 0u || old == key) break;
                p = (p + 1) & hs_mask;
            }
            u32 rank = atomicAdd(&hs_cnt[p], 1u);
            if (set_mode && rank > 0) continue;                   // set semantics: a hash of the read counts once
            const u64* row = table_row(tb, slot, rank);
            if (row == nullptr) continue;
            for (u32 w = 0; w < tb.nb; ++w) {
                u64 bits = row[w];
                while (bits) {
                    int b = __ffsll((long long)bits) - 1;
                    bits &= bits - 1;
                    atomicAdd(&counts[w * 64 + b], 1u);
                }
            }
        }
        __syncthreads();
        if (set_mode) {
            // hpv16 (src/rkmh.cpp:2655-2664): the first type with the most shared hashes; the remaining sets' counts as they are
            const int nextra = (int)tb.nrefs - nargmax;
            for (int j = threadIdx.x; j < nextra; j += blockDim.x) extra[rid * (i64)nextra + j] = (int)counts[nargmax + j];
            block_argmax(counts, nargmax, &as, &out[rid], L);
            continue;
        }
        block_argmax(counts, (int)tb.nrefs, &as, &out[rid], L);
    }
}


// ------------------------------------------------------------------------------------------------
// classify_group_kernel: the fast path for reads whose sketch is simply "every non-zero hash" (number
// of k-mers <= s and <= GROUP_MAX_KMERS).  A group of G = 2^glog2 lanes owns one read; each lane rolls
// over a contiguous run of k-mers.  Hashes that pass the shared-memory Bloom prefilter are queued per
// warp and probed 32 at a time, so the L2 probes run with full lanes.  Per-reference counts live in
// shared memory as packed 16-bit counters.
//
// Multiset semantics: a per-warp shared-memory set counts how often each table slot was hit by
// the read, so the r-th copy of a hash is looked up at rank r.  A read is handed to classify_block_kernel
// through `redo` instead when it is not eligible or when its hits overflow that set.  The result is
// therefore exact for every input.
constexpr int GROUP_MAX_KMERS = 2048;
constexpr int GQ = 64;      // per-warp candidate ring
constexpr int GDS = 128;    // per-warp set of the table slots hit so far: key + number of copies seen

struct GroupSmem {           // per-warp scratch
    u64 qh[GQ];
    u32 qg[GQ];
    u32 d[2 * GDS];          // [0, GDS): keys, [GDS, 2 GDS): counts
};

template <int KC>
struct GroupLane {
    BaseReader rd;
    KmerWindow<KC> win;
};

template <int KCSEL>
__global__ void __launch_bounds__(1024, 1)
classify_group_kernel(SeqBatch sb, const __grid_constant__ KList kl, int s, CountTable ct, int min_occ, RefTable tb,
                      int glog2, int rwords, int filter_words,
                      int* __restrict__ redo, int* __restrict__ redo_count, Hit* __restrict__ out) {
#ifdef RK_EMU
    unsigned char* dsm = reinterpret_cast<unsigned char*>(EMU_DYN_SMEM());
#else
    extern __shared__ __align__(16) unsigned char dsm[];
#endif
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    const int G = 1 << glog2, gpw = 32 >> glog2;          // lanes per read, reads per warp
    const int g = lane >> glog2, l = lane & (G - 1);
    const u32 gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (g * G));

    GroupSmem* wsm = reinterpret_cast<GroupSmem*>(dsm) + wib;
    u32* cnt_all = reinterpret_cast<u32*>(reinterpret_cast<GroupSmem*>(dsm) + nwb);
    u32* gflag_all = cnt_all + (size_t)nwb * gpw * rwords;
    u32* filt = gflag_all + nwb * gpw;
    u32* cnt = cnt_all + (size_t)(wib * gpw + g) * rwords;   // packed u16 counters of my group
    u32* gflag = gflag_all + wib * gpw + g;                  // bit0: has hits, bit1: needs the exact path

    for (int i = threadIdx.x; i < filter_words; i += blockDim.x) filt[i] = tb.filter[i];
    for (int i = threadIdx.x; i < nwb * gpw * rwords + nwb * gpw; i += blockDim.x) cnt_all[i] = 0;
    for (int i = lane; i < 2 * GDS; i += 32) wsm->d[i] = 0;
    __syncthreads();

    const i64 gwarp = (i64)blockIdx.x * nwb + wib, nwarps = (i64)gridDim.x * nwb;
    int qhead = 0, qn = 0;

    // Lanes cooperate on the reference bitsets of the hits of one drain: a row of nw32 32-bit words is spread
    // over min(nw32, 32) lanes, so the serial part is the number of set bits of ONE word, not of the whole row.
    const int nw32 = 2 * (int)tb.nb;
    const int lanes_per_hit = nw32 < 32 ? nw32 : 32;
    const int hits_per_round = 32 / lanes_per_hit;
    const int my_sub = lane / lanes_per_hit, my_word = lane - my_sub * lanes_per_hit;

    auto drain = [&](int m) {                // probe the first m (<= 32) queued candidates
        const u64* row = nullptr;
        u32 hg = 0;
        if (lane < m) {
            int qi = (qhead + lane) & (GQ - 1);
            u64 h = wsm->qh[qi];
            hg = wsm->qg[qi];
            int slot = table_find(tb, h);
            if (slot >= 0) {
                // multiplicity inside the read's sketch: the r-th copy of a hash counts for the references that hold
                // more than r copies (sum of min(multiplicities), mkmh.hpp:822-840)
                u32 key = (((u32)slot << 3) | hg) + 1u, p = (key * 0x9E3779B1u) >> 25;   // GDS = 128
                bool found = false;
                for (int tries = 0; tries < GDS; ++tries) {
                    u32 old = atomicCAS(&wsm->d[p], 0u, key);
                    if (old == 0u || old == key) { found = true; break; }
                    p = (p + 1) & (GDS - 1);
                }
                if (!found) {
                    atomicOr(&gflag_all[wib * gpw + hg], 2u);          // set overflow: exact path
                } else {
                    atomicOr(&gflag_all[wib * gpw + hg], 1u);
                    const u32 rank = atomicAdd(&wsm->d[GDS + p], 1u);
                    row = table_row(tb, slot, rank);
                }
            }
        }
        __syncwarp();
        // compact the hit rows into the ring slots this drain has just consumed
        const u32 hitmask = __ballot_sync(0xffffffffu, row != nullptr);
        const int nh = __popc(hitmask);
        if (row != nullptr) {
            int qi = (qhead + __popc(hitmask & ((1u << lane) - 1u))) & (GQ - 1);
            wsm->qh[qi] = (u64)(size_t)row;
            wsm->qg[qi] = hg;
        }
        __syncwarp();
        for (int base = 0; base < nh; base += hits_per_round) {
            const int hi = base + my_sub;
            if (hi < nh && my_sub < hits_per_round) {
                const int qi = (qhead + hi) & (GQ - 1);
                const u32* r32 = reinterpret_cast<const u32*>((size_t)wsm->qh[qi]);
                u32* hc = cnt_all + (size_t)(wib * gpw + wsm->qg[qi]) * rwords;
                for (int w = my_word; w < nw32; w += 32) {
                    u32 bits = r32[w];
                    while (bits) {
                        const int b = __ffs((int)bits) - 1;
                        bits &= bits - 1;
                        atomicAdd(&hc[16 * w + (b >> 1)], 1u << (16 * (b & 1)));   // reference 32 w + b
                    }
                }
            }
        }
        qhead = (qhead + m) & (GQ - 1);
        qn -= m;
        __syncwarp();
    };

    for (i64 r0 = gwarp * gpw; r0 < sb.nseq; r0 += nwarps * gpw) {
        const i64 rid = r0 + g;
        const bool have = rid < sb.nseq;
        i64 beg = 0, len = 0;
        if (have) { beg = sb.offs[rid] - sb.base0; len = sb.offs[rid + 1] - sb.offs[rid]; }
        const i64 ntot = have ? nkmers_total(len, kl) : 0;
        const bool eligible = have && ntot <= s && ntot <= GROUP_MAX_KMERS;
        int nz = 0;
        for (int j = 0; j < kl.nk; ++j) {
            const int k = kl.k[j];
            const int nk = eligible ? (int)nkmers(len, k) : 0;
            const int c = (nk + G - 1) >> glog2;                 // k-mers per lane
            int cmax = c;
            for (int d = 16; d >= 1; d >>= 1) cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, d));
            const int c0 = l * c, c1 = min(c0 + c, nk);
            auto body = [&](auto kc_tag) {
                constexpr int KC = decltype(kc_tag)::value;
                BaseReader rd; KmerWindow<KC> win; win.reset();
                const WinSel ws = make_winsel(k);
                u32 code; bool inv;
                if (c0 < c1) {
                    rd.init(sb.codes, sb.nmask, beg + c0);
                    for (int w = 0; w < k - 1; ++w) { rd.next(code, inv); win.push(code, inv, ws); }
                }
                for (int i = 0; i < cmax; ++i) {
                    u64 h = 0;
                    if (c0 + i < c1) {
                        rd.next(code, inv); win.push(code, inv, ws);
                        h = apply_min_occ(win.hash(k), ct, min_occ);
                    }
                    bool cand = false;
                    if (h != 0) {
                        ++nz;
                        u32 fb = filter_bit(h, tb.filter_log2);
                        cand = (filt[fb >> 5] >> (fb & 31)) & 1u;
                    }
                    u32 bal = __ballot_sync(0xffffffffu, cand);
                    if (bal) {
                        if (cand) {
                            int qi = (qhead + qn + __popc(bal & ((1u << lane) - 1u))) & (GQ - 1);
                            wsm->qh[qi] = h; wsm->qg[qi] = (u32)g;
                        }
                        qn += __popc(bal);
                        __syncwarp();
                        if (qn >= 32) drain(32);
                    }
                }
            };
            if (KCSEL < NUM_KV) { body(KcTag<(KCSEL < NUM_KV ? KCSEL : 0)>()); }
            else switch (kl.kv[j]) {
                case 0: body(KcTag<0>()); break; case 1: body(KcTag<1>()); break; case 2: body(KcTag<2>()); break;
                case 3: body(KcTag<3>()); break; case 4: body(KcTag<4>()); break; case 5: body(KcTag<5>()); break;
                case 6: body(KcTag<6>()); break; case 7: body(KcTag<7>()); break; case 8: body(KcTag<8>()); break;
                default: body(KcTag<9>()); break;
            }
        }
        while (qn > 0) drain(min(qn, 32));
        // ---- per-read epilogue --------------------------------------------------------------------
        for (int d = G >> 1; d >= 1; d >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, d);
        const u32 flag = *gflag;
        __syncwarp();
        if (have) {
            if (!eligible || (flag & 2u)) {
                if (l == 0) redo[atomicAdd(redo_count, 1)] = (int)rid;
            } else if (!(flag & 1u)) {
                if (l == 0) { Hit hh; hh.ref_id = 0; hh.best = 0; hh.prev = -1; hh.sketch_len = nz; out[rid] = hh; }
            } else {
                // argmax of src/rkmh.cpp:874-883 on 32-bit keys (count << 16 | 0xFFFF - index): the largest key is the
                // first reference with the highest count
                u32 best = 0;
                for (int w = l; w < rwords; w += G) {
                    const u32 pair = cnt[w];
                    const u32 j = 2u * (u32)w;
                    best = max(best, max((pair << 16) | (0xFFFFu - j), (pair & 0xFFFF0000u) | (0xFFFEu - j)));
                }
                for (int d = G >> 1; d >= 1; d >>= 1) best = max(best, __shfl_xor_sync(gmask, best, d));
                const int id = (int)(0xFFFFu - (best & 0xFFFFu));
                int prev = -1;
                for (int w = l; 2 * w < id; w += G) {
                    const u32 pair = cnt[w];
                    prev = max(prev, (int)(pair & 0xFFFFu));
                    if (2 * w + 1 < id) prev = max(prev, (int)(pair >> 16));
                }
                for (int d = G >> 1; d >= 1; d >>= 1) prev = max(prev, __shfl_xor_sync(gmask, prev, d));
                if (l == 0) { Hit hh; hh.ref_id = id; hh.best = (int)(best >> 16); hh.prev = prev; hh.sketch_len = nz; out[rid] = hh; }
            }
        }
        __syncwarp();
        const u32 anyflag = __ballot_sync(0xffffffffu, flag != 0);
        if (anyflag) {                        // some group of this warp touched counters / the dedupe set
            if (flag) { for (int i = l; i < rwords; i += G) cnt[i] = 0; if (l == 0) *gflag = 0; }
            for (int i = lane; i < 2 * GDS; i += 32) wsm->d[i] = 0;
        }
        __syncwarp();
    }
}

}  // namespace rk
