// rk_memo.cuh -- the memoised path of the rkmh hot path for a single k <= 12.
//
// calc_hash (mkmh.hpp:505-532) is a pure function of the k-mer, and for k <= 12 its whole domain has
// 4^k <= 16,777,216 points.  The hashing kernel (murmur_kmer, rk_common.cuh) is therefore run ONCE over every
// k-mer when the references are loaded, and the per-read kernels look results up in a table that stays resident
// in the 126 MB L2 instead of recomputing 2 x MurmurHash3_x64_128 per window:
//
//   hash[c]  u64  canonical hash of k-mer code c                              (mkmh.hpp:517-524)
//   tab[c]   u32  (rank << 8) | flags; rank = number of codes whose hash is smaller, so that "the s smallest
//                 hashes, duplicates kept" (minhashes, mkmh.hpp:931-950) is "the s smallest ranks"
//                 REF   the hash is a member of some reference sketch          (hash_intersection_size)
//                 ZERO  the hash is 0, the reference's "invalid" sentinel      
//                 FAILM counts[hash % size] < min_occ                          (mask_by_frequency, mkmh.hpp:898)
//   ref[rank] u32  slot + 1 of the hash in the (hash, rank) -> bitset table     (equal hash <=> equal rank)
//   code_counts[c] int32  -M pass 1 counters per code; folded into counts[hash % size] (HASHTCounter.cpp:46)
//
// A k-mer code is its 2-bit base codes, base j at bits 2 j (first base lowest).  The per-read kernels read the ASCII
// bases as parsed (no ingest pass): aligned 32-bit loads, four bases decoded per word (sv_decode), the 16 bases that
// follow a lane's word assembled with three shuffles; the code of a window is one more shuffle, a shift and a mask.
// Results are bit-identical to the Murmur path for every input (tests: memo on/off).
#pragma once
#include "rk_kernels.cuh"

namespace rk {

constexpr int MEMO_MAX_K = 12;
constexpr u32 MEMO_REF = 1u, MEMO_ZERO = 2u, MEMO_FAILM = 4u;
constexpr u32 MEMO_SKIP = 0xFFFFFFFFu;          // "window does not count": never a real entry (5 flag bits unused)
constexpr int MEMO_BUCKET_CAP = 1024;           // rank build: largest hash bucket that is sorted in shared memory
constexpr int MEMO_BINS = 4096;
// The histogram is stored with one word of padding per 16 bins: memo_find_bin gives each thread 16 CONSECUTIVE bins, and
// without the padding the 32 lanes of a warp read addresses 16 words apart -- two banks, a 16-way conflict on every load
// (ncu, round 2: 64 % of this kernel's shared-memory wavefronts were such replays; 6.6 % with the padding).
constexpr int MEMO_HIST_WORDS = MEMO_BINS + MEMO_BINS / 16;
RK_HD u32 memo_hidx(u32 bin) { return bin + (bin >> 4); }
constexpr int MEMO_PARTIAL_PAD = 32;             // int32 slots after code_counts[ncodes): [0] = invalid windows, rest zero (keeps the array 128-byte sized)
constexpr int MEMO_RING = 64;                   // block kernel: per-warp ring of sketch members waiting for their rows                 // radix-select histogram (12 bits of the 24-bit rank per level)

struct Memo {
    u32* tab;                       // indexed by the k-mer code
    u32* ref;
    const u64* hash;
    int* code_counts;               // nullptr when the context has no read count table
    u32* inval;                     // = code_counts + ncodes: windows with an invalid base seen by the count pass (count_zeros mode;
                                    //   modulo 2^32 like the int32 slot it is added to), so that ONE all-reduce covers both
    u32 ncodes;
    u32 kmask;                      // (1 << k) - 1
    int k;
    // sieve kernel (classify_memo_sieve_kernel): shared-memory Bloom filter image and the code -> slot + 1 map
    const u32* sieve;               // nullptr: not built / too full to be useful
    const unsigned long long* cmap;
    u32 sieve_words, cmap_mask;
    int cmap_shift;
    u32 zero_codes;                 // codes whose hash is 0 (the sieve kernel needs none)
};

#ifdef RK_EMU
RK_D u32 warp_sum(u32 v) { for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d); return v; }
RK_D u32 warp_max(u32 v) { for (int d = 16; d >= 1; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d)); return v; }
RK_D u32 warp_or(u32 v) { for (int d = 16; d >= 1; d >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, d); return v; }
#else
RK_D u32 warp_sum(u32 v) { return __reduce_add_sync(0xffffffffu, v); }
RK_D u32 warp_max(u32 v) { return __reduce_max_sync(0xffffffffu, v); }
RK_D u32 warp_or(u32 v) { return __reduce_or_sync(0xffffffffu, v); }
#endif

// Exactly the bytes canonical() accepts (valid_dna[], mkmh.hpp:137-152) as one bit test: A C G T a c g t.
RK_D bool memo_base_invalid(u32 ch) {
    const u32 u = (ch & 0xDFu) - 65u;                       // 'A' -> 0 ... 'Z' -> 25 for either case
    return u > 19u || !((0x00080045u >> u) & 1u);           // bits 0 (A), 2 (C), 6 (G), 19 (T)
}

// Four bases of the read as one aligned 32-bit load: `piece` = their four 2-bit codes (first base in the low bits),
// `bad` has a non-zero byte for every base that is not one of ACGTacgt (canonical(), mkmh.hpp:137-152): the letter is
// rebuilt from its code (A 0x41, C 0x43, T 0x54, G 0x47 for codes 0, 1, 2, 3) and compared with the upper-cased byte.
RK_D void sv_decode(u32 w, u32& piece, u32& bad) {
    const u32 x = (w >> 1) & 0x03030303u;
    piece = (x * 0x01041040u) >> 24;                         // byte i -> bits 2 i (the partial products do not overlap)
    const u32 lo = x & 0x01010101u, hi = (x >> 1) & 0x01010101u;
    const u32 e = 0x41414141u + 2u * lo + 0x13u * hi - 0xFu * (lo & hi);
    bad = (w & 0xDFDFDFDFu) ^ e;
}
RK_D u32 sv_bad4(u32 bad) {
    return ((bad & 0xFFu) ? 1u : 0u) | ((bad & 0xFF00u) ? 2u : 0u) | ((bad & 0xFF0000u) ? 4u : 0u) | ((bad & 0xFF000000u) ? 8u : 0u);
}

// ------------------------------------------------------------------------------------------------
// One warp walks the k-mer starts [s0, s1) of the sequence p[0, len), 128 starts per step, and calls
// f(code, in_range, valid) four times per step on every lane (warp-convergent, straight-line: no branch between the
// calls; in_range = this lane's window starts inside [s0, s1), valid = in_range and no invalid base in the window).
// k <= 12, s1 - s0 < 2^31 - 256.  The order of the windows is not the order of the calls' lanes: stream position 0 is the
// word-aligned byte at or before p + s0, the window of lane l in round r of a step starts at stream base 32 r + l, so the
// first (address & 3) windows reported lie before s0 (out of range), and so do those of the last step beyond s1.
//
// The bytes are read as aligned 32-bit words, 128 bases of the stream per lane-wide load, three loads in flight (a warp
// streams a long read straight from HBM) and never beyond the last byte a window of the range can touch.
template <typename F>
RK_D void memo_stream(const unsigned char* __restrict__ p, i64 len, int k, i64 s0, i64 s1, int lane, F&& f) {
    if (s0 >= s1) return;
    const unsigned char* q = p + s0;
    const int sh = (int)(reinterpret_cast<size_t>(q) & 3u);
    const u32* __restrict__ wp = reinterpret_cast<const u32*>(q - sh);
    const int nrel = (int)(s1 - s0);
    const int pend = sh + nrel;                             // stream positions [sh, pend) start a window of the range
    const int nbytes = (int)min((i64)sh + (len - s0), (i64)pend + k);      // bytes from wp that such a window can touch
    // words that hold no such byte are not loaded and decode as valid 'AAAA'
    auto load_word = [&](int widx) -> u32 { return 4 * widx < nbytes ? __ldg(wp + widx) : 0x41414141u; };
    const u32 cmask2 = (1u << (2 * k)) - 1u, kmask = (1u << k) - 1u;
    // window at stream position 32 r + lane of a step: pieces 8 r + (lane >> 2) .. + 3, from base (lane & 3) on
    const int src0 = lane >> 2;
    const u32 osh = 2u * (u32)(lane & 3), ish = (u32)(lane & 3);
    // piece j + i of lane j comes from the next step when j + i >= 32 (byte 1 of the shuffled pair instead of byte 0)
    const u32 sel1 = lane >= 31 ? 0x3250u : 0x3240u, sel2 = lane >= 30 ? 0x3510u : 0x3410u, sel3 = lane >= 29 ? 0x5210u : 0x4210u;
    u32 wb = load_word(32 + lane), wc = load_word(64 + lane);
    u32 pc, bc;
    sv_decode(load_word(lane), pc, bc);
    for (int c0 = 0; c0 < pend; c0 += 128) {
        u32 pn, bn;
        sv_decode(wb, pn, bn);
        wb = wc; wc = load_word((c0 >> 2) + 96 + lane);
        const bool slow = __any_sync(0xffffffffu, (bc | bn) != 0u);
        const u32 xq = pc | (pn << 8);
        u32 Q = xq;                                          // the 16 bases from piece `lane` on
        Q = __byte_perm(Q, __shfl_sync(0xffffffffu, xq, (lane + 1) & 31), sel1);
        Q = __byte_perm(Q, __shfl_sync(0xffffffffu, xq, (lane + 2) & 31), sel2);
        Q = __byte_perm(Q, __shfl_sync(0xffffffffu, xq, (lane + 3) & 31), sel3);
        u32 code[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) code[r] = (__shfl_sync(0xffffffffu, Q, src0 + 8 * r) >> osh) & cmask2;
        bool bad[4] = {false, false, false, false};          // the window of round r holds an invalid base
        if (slow) {
            const u32 xi = sv_bad4(bc) | (sv_bad4(bn) << 4);
            u32 IQ = xi & 15u;                               // the invalid flags of the same 16 bases
#pragma unroll
            for (int i = 1; i < 4; ++i) {
                const u32 t = __shfl_sync(0xffffffffu, xi, (lane + i) & 31);
                IQ |= ((lane + i < 32) ? (t & 15u) : (t >> 4)) << (4 * i);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) bad[r] = ((__shfl_sync(0xffffffffu, IQ, src0 + 8 * r) >> ish) & kmask) != 0u;
        }
        const int rel0 = c0 + lane - sh;                     // start of this lane's round-0 window, relative to s0
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const bool in = (u32)(rel0 + 32 * r) < (u32)nrel;
            f(code[r], in, in && !bad[r]);
        }
        pc = pn; bc = bn;
    }
}

// ------------------------------------------------------------------------------------------------
// Table construction (once per rkg_load_refs).

// canonical hash of k-mer code c through the hashing kernel's Murmur (mkmh.hpp:515-524)
template <int KV>
RK_D u64 memo_code_hash(u32 c, int k) {
    constexpr int NW = KVar<KV>::NW;
    u32 f[NW], r[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) { f[i] = 0; r[i] = 0; }
#pragma unroll
    for (int j = 0; j < 4 * NW; ++j) {
        if (j < k) {
            const u32 cf = (c >> (2 * j)) & 3u;
            f[j >> 2] |= (code_letters(cf) & 0xFFu) << (8 * (j & 3));
            const int src = k - 1 - j;                          // reverse complement: byte j = complement of base k-1-j
            const u32 cr = (c >> (2 * src)) & 3u;
            r[j >> 2] |= ((code_letters(cr) >> 8) & 0xFFu) << (8 * (j & 3));
        }
    }
    const u64 hf = murmur_kmer<KV>(f, k), hr = murmur_kmer<KV>(r, k);
    return hf < hr ? hf : hr;
}

RK_D u32 memo_bucket(u64 h, int bucket_bits) { return bucket_bits ? (u32)(h >> (64 - bucket_bits)) : 0u; }

template <int KV>
__global__ void __launch_bounds__(256)
memo_hash_kernel(int k, u32 ncodes, u64* __restrict__ hash, u32* __restrict__ bucket_count, int bucket_bits) {
    for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < ncodes; c += gridDim.x * blockDim.x) {
        const u64 h = memo_code_hash<KV>(c, k);
        hash[c] = h;
        atomicAdd(&bucket_count[memo_bucket(h, bucket_bits)], 1u);
    }
}

// exclusive scan of nb (<= 65536) bucket counts by one block of 1024 threads; also resets the fill cursors
__global__ void __launch_bounds__(1024)
memo_scan_kernel(const u32* __restrict__ count, u32* __restrict__ start, u32* __restrict__ fill, u32 nb, int* __restrict__ err) {
    __shared__ u32 part[1024];
    const u32 per = (nb + 1023u) / 1024u, b0 = threadIdx.x * per;
    u32 sum = 0, mx = 0;
    for (u32 i = 0; i < per; ++i) if (b0 + i < nb) { sum += count[b0 + i]; mx = max(mx, count[b0 + i]); }
    if (mx > (u32)MEMO_BUCKET_CAP) *err = 1;
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        u32 v = ((int)threadIdx.x >= d) ? part[threadIdx.x - d] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    u32 run = part[threadIdx.x] - sum;
    for (u32 i = 0; i < per; ++i) if (b0 + i < nb) { start[b0 + i] = run; fill[b0 + i] = 0; run += count[b0 + i]; }
}

__global__ void __launch_bounds__(256)
memo_scatter_kernel(u32 ncodes, const u64* __restrict__ hash, const u32* __restrict__ start, u32* __restrict__ fill,
                    u32* __restrict__ codes, int bucket_bits) {
    for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < ncodes; c += gridDim.x * blockDim.x) {
        const u32 b = memo_bucket(hash[c], bucket_bits);
        codes[start[b] + atomicAdd(&fill[b], 1u)] = c;
    }
}

// One block per hash bucket: sort the bucket by hash in shared memory; rank = start + index of the first equal hash.
__global__ void __launch_bounds__(256)
memo_rank_kernel(const u64* __restrict__ hash, const u32* __restrict__ count, const u32* __restrict__ start,
                 const u32* __restrict__ codes, u32 nb, RefTable tb, u32* __restrict__ tab, u32* __restrict__ ref) {
    __shared__ u64 key[MEMO_BUCKET_CAP];
    __shared__ u32 val[MEMO_BUCKET_CAP];
    for (u32 b = blockIdx.x; b < nb; b += gridDim.x) {
        const int n = (int)min(count[b], (u32)MEMO_BUCKET_CAP);
        const u32 st = start[b];
        int cap = 2; while (cap < n) cap <<= 1;
        for (int i = threadIdx.x; i < cap; i += blockDim.x) {
            if (i < n) { const u32 c = codes[st + i]; val[i] = c; key[i] = hash[c]; } else { val[i] = 0; key[i] = ~0ULL; }
        }
        for (int size = 2; size <= cap; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                __syncthreads();
                for (int t = threadIdx.x; t < (cap >> 1); t += blockDim.x) {
                    const int lo = ((t / stride) * (stride << 1)) + (t % stride), hi = lo + stride;
                    const bool up = ((lo & size) == 0);
                    const u64 x = key[lo], y = key[hi];
                    if ((x > y) == up) { key[lo] = y; key[hi] = x; const u32 v = val[lo]; val[lo] = val[hi]; val[hi] = v; }
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const u64 h = key[i];
            int first = i;
            while (first > 0 && key[first - 1] == h) --first;
            const u32 rank = st + (u32)first;
            u32 flags = 0;
            if (h == 0) flags |= MEMO_ZERO;
            else if (tb.entries != nullptr) {
                const int slot = table_find(tb, h);
                if (slot >= 0) { flags |= MEMO_REF; ref[rank] = (u32)slot + 1u; }
            }
            tab[val[i]] = (rank << 8) | flags;
        }
        __syncthreads();
    }
}

// -M pass 1 epilogue: counts[hash % size] += code_counts[code]  (HASHTCounter::increment, HASHTCounter.cpp:46-58)
__global__ void __launch_bounds__(256)
memo_fold_kernel(Memo mm, CountTable ct, int count_zeros) {
    for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < mm.ncodes; c += gridDim.x * blockDim.x) {
        const int n = mm.code_counts[c];
        if (n == 0) continue;
        mm.code_counts[c] = 0;
        const u64 h = mm.hash[c];
        if (h != 0 || count_zeros) atomicAdd(&ct.counts[fastmod(h, ct.mod)], n);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const u32 inv = *mm.inval;
        *mm.inval = 0;
        if (count_zeros && inv) atomicAdd(&ct.counts[0], (int)inv);     // an invalid window hashes to 0 (rkmh.cpp:331-336)
    }
}

// bake mask_by_frequency (mkmh.hpp:898-908) for the current table and min_occ into the FAILM flag (min_occ < 0: clear)
__global__ void __launch_bounds__(256)
memo_apply_occ_kernel(Memo mm, CountTable ct, int min_occ) {
    for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < mm.ncodes; c += gridDim.x * blockDim.x) {
        const u32 t = mm.tab[c];
        u32 fail = 0;
        if (min_occ >= 0 && ct.counts != nullptr) {
            const u64 h = mm.hash[c];
            if (h != 0 && !(min_occ <= ct.counts[fastmod(h, ct.mod)])) fail = MEMO_FAILM;
        }
        const u32 nt = (t & ~MEMO_FAILM) | fail;
        if (nt != t) mm.tab[c] = nt;
    }
}

// ------------------------------------------------------------------------------------------------
// count_memo_kernel: -M pass 1.  `wpr` warps share one read (1 for short reads, a whole block for long ones).
__global__ void __launch_bounds__(256)
count_memo_kernel(SeqBatch sb, Memo mm, int count_zeros, int wpr) {
    const int lane = threadIdx.x & 31;
    const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((i64)gridDim.x * blockDim.x) >> 5;
    const i64 team = gw / wpr, nteams = nw / wpr;
    const int wt = (int)(gw % wpr);
    u32 inv = 0;
    for (i64 s = team; s < sb.nseq; s += nteams) {
        const i64 beg = sb.offs[s] - sb.base0, len = sb.offs[s + 1] - sb.offs[s];
        const i64 n = nkmers(len, mm.k);
        const i64 per = ((((n + 31) >> 5) + wpr - 1) / wpr) << 5;          // k-mer starts per warp, whole rounds
        const i64 s0 = (i64)wt * per, s1 = min(n, s0 + per);
        memo_stream(sb.bases + beg, len, mm.k, s0, s1, lane, [&](u32 code, bool in, bool valid) {
            if (valid) atomicAdd(&mm.code_counts[code], 1);
            else if (in) ++inv;
        });
    }
    if (count_zeros) {
        inv = warp_sum(inv);
        if (lane == 0 && inv) atomicAdd(mm.inval, inv);
    }
}

// ------------------------------------------------------------------------------------------------
// classify_memo_wide_kernel: the warp-per-read kernel for MORE than 256 references (see classify_memo_warp_kernel
// below for the common case).  Members of a reference sketch (REF) are queued per warp and resolved 32 at a time:
// table slot from ref[rank], copy number inside the read from a per-warp set (multiset semantics),
// bitset rows spread over the lanes into packed 16-bit per-reference counters.  Reads with more k-mers than s, or
// whose hits overflow the set, go to the block kernel through `redo`.
constexpr int MQ = 64;      // per-warp candidate ring
constexpr int MDS = 128;    // per-warp set of the table slots hit so far: key + number of copies seen

struct MemoWarpSmem {
    u64 q[MQ];
    u32 d[2 * MDS];          // [0, MDS): keys, [MDS, 2 MDS): counts
};

__global__ void __launch_bounds__(1024, 1)
classify_memo_wide_kernel(SeqBatch sb, int s, Memo mm, RefTable tb, int rwords,
                          int* __restrict__ redo, int* __restrict__ redo_count, Hit* __restrict__ out) {
#ifdef RK_EMU
    unsigned char* dsm = reinterpret_cast<unsigned char*>(EMU_DYN_SMEM());
#else
    extern __shared__ __align__(16) unsigned char dsm[];
#endif
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    MemoWarpSmem* wsm = reinterpret_cast<MemoWarpSmem*>(dsm) + wib;
    u32* cnt = reinterpret_cast<u32*>(reinterpret_cast<MemoWarpSmem*>(dsm) + nwb) + (size_t)wib * rwords;

    for (int i = lane; i < rwords; i += 32) cnt[i] = 0;
    for (int i = lane; i < 2 * MDS; i += 32) wsm->d[i] = 0;
    __syncwarp();

    const int nw32 = 2 * (int)tb.nb;
    const int lanes_per_hit = nw32 < 32 ? nw32 : 32;
    const int hits_per_round = 32 / lanes_per_hit;
    const int my_sub = lane / lanes_per_hit, my_word = lane - my_sub * lanes_per_hit;
    const int k = mm.k;

    const i64 gwarp = (i64)blockIdx.x * nwb + wib, nwarps = (i64)gridDim.x * nwb;
    for (i64 rid = gwarp; rid < sb.nseq; rid += nwarps) {
        const i64 beg = sb.offs[rid] - sb.base0, len = sb.offs[rid + 1] - sb.offs[rid];
        const i64 n = nkmers(len, k);
        if (n > s) {                                         // the sketch is a strict subset: block kernel
            if (lane == 0) redo[atomicAdd(redo_count, 1)] = (int)rid;
            continue;
        }
        int qhead = 0, qn = 0;
        bool dirty = false, overflow = false;
        u32 nz = 0;

        auto drain = [&](int m) {                            // resolve the first m (<= 32) queued candidates
            const u64* row = nullptr;
            bool over = false;
            if (lane < m) {
                const u32 rank = (u32)wsm->q[(qhead + lane) & (MQ - 1)];
                const u32 key = mm.ref[rank];                // slot + 1
                u32 p = (key * 0x9E3779B1u) >> 25;           // MDS = 128
                bool found = false;
                for (int tries = 0; tries < MDS; ++tries) {
                    const u32 old = atomicCAS(&wsm->d[p], 0u, key);
                    if (old == 0u || old == key) { found = true; break; }
                    p = (p + 1) & (MDS - 1);
                }
                if (!found) over = true;
                else row = table_row(tb, (int)key - 1, atomicAdd(&wsm->d[MDS + p], 1u));
            }
            __syncwarp();
            if (__ballot_sync(0xffffffffu, over)) overflow = true;
            const u32 hitmask = __ballot_sync(0xffffffffu, row != nullptr);
            const int nh = __popc(hitmask);
            if (row != nullptr) wsm->q[(qhead + __popc(hitmask & ((1u << lane) - 1u))) & (MQ - 1)] = (u64)(size_t)row;
            __syncwarp();
            for (int base = 0; base < nh; base += hits_per_round) {
                const int hi = base + my_sub;
                if (hi < nh && my_sub < hits_per_round) {
                    const u32* r32 = reinterpret_cast<const u32*>((size_t)wsm->q[(qhead + hi) & (MQ - 1)]);
                    for (int w = my_word; w < nw32; w += 32) {
                        u32 bits = r32[w];
                        while (bits) {
                            const int b = __ffs((int)bits) - 1;
                            bits &= bits - 1;
                            atomicAdd(&cnt[16 * w + (b >> 1)], 1u << (16 * (b & 1)));   // reference 32 w + b
                        }
                    }
                }
            }
            qhead = (qhead + m) & (MQ - 1);
            qn -= m;
            dirty = true;
            __syncwarp();
        };

        memo_stream(sb.bases + beg, len, k, 0, n, lane, [&](u32 code, bool, bool valid) {
            u32 m = MEMO_ZERO;
            if (valid) m = __ldg(&mm.tab[code]);
            const bool counted = !(m & (MEMO_ZERO | MEMO_FAILM));
            nz += counted;
            const bool cand = counted && (m & MEMO_REF);
            const u32 bal = __ballot_sync(0xffffffffu, cand);
            if (bal) {
                if (cand) wsm->q[(qhead + qn + __popc(bal & ((1u << lane) - 1u))) & (MQ - 1)] = (u64)(m >> 8);
                qn += __popc(bal);
                __syncwarp();
                if (qn >= 32) drain(32);
            }
        });
        while (qn > 0) drain(min(qn, 32));

        nz = warp_sum(nz);
        if (overflow) {
            if (lane == 0) redo[atomicAdd(redo_count, 1)] = (int)rid;
        } else if (!dirty) {
            if (lane == 0) { Hit hh; hh.ref_id = 0; hh.best = 0; hh.prev = -1; hh.sketch_len = (int)nz; out[rid] = hh; }
        } else {
            // argmax of src/rkmh.cpp:874-883 on 32-bit keys (count << 16 | 0xFFFF - index): the largest key is the
            // first reference with the highest count
            u32 best = 0;
            for (int w = lane; w < rwords; w += 32) {
                const u32 pair = cnt[w], j = 2u * (u32)w;
                best = max(best, max((pair << 16) | (0xFFFFu - j), (pair & 0xFFFF0000u) | (0xFFFEu - j)));
            }
            best = warp_max(best);
            const int id = (int)(0xFFFFu - (best & 0xFFFFu));
            u32 prev1 = 0;                                   // max of (count + 1) before the winner, 0 = none
            for (int w = lane; 2 * w < id; w += 32) {
                const u32 pair = cnt[w];
                prev1 = max(prev1, (pair & 0xFFFFu) + 1u);
                if (2 * w + 1 < id) prev1 = max(prev1, (pair >> 16) + 1u);
            }
            prev1 = warp_max(prev1);
            if (lane == 0) { Hit hh; hh.ref_id = id; hh.best = (int)(best >> 16); hh.prev = (int)prev1 - 1; hh.sketch_len = (int)nz; out[rid] = hh; }
        }
        if (dirty) {
            __syncwarp();
            for (int i = lane; i < rwords; i += 32) cnt[i] = 0;
            for (int i = lane; i < 2 * MDS; i += 32) wsm->d[i] = 0;
            __syncwarp();
        }
    }
}

// Column sums of the 32 x (32 NW) bit matrix held one row per lane, reduced on the fly to what the scan of
// src/rkmh.cpp:874-883 reports: best = largest column sum, id = first column attaining it, prev = largest sum at a
// smaller index (0 when there is none but id > 0, -1 when id == 0).  Columns are visited in index order, and only
// those that can win: when some column is set in two rows (OR over lanes of row & exclusive-prefix-OR of rows), only
// such columns are counted (popc of a ballot); the columns with a single row matter through `prev` alone.
template <int NW>
RK_D void memo_scan_rows(const u32 (&x)[NW], int lane, int& best, int& id, int& prev) {
    u32 any[NW], two[NW];
    u32 any_all = 0, two_all = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        any[w] = warp_or(x[w]);
        two[w] = 0;
        if (any[w] & (any[w] - 1u)) {                        // at least two columns or rows could overlap: exact test
            u32 v = __shfl_up_sync(0xffffffffu, x[w], 1);
            if (lane < 1) v = 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v |= t; }
            two[w] = warp_or(x[w] & v);
        } else if (any[w]) {                                 // one column: it is shared iff two lanes hold it
            two[w] = __popc(__ballot_sync(0xffffffffu, x[w] != 0u)) >= 2 ? any[w] : 0u;
        }
        any_all |= any[w]; two_all |= two[w];
    }
    best = 0; id = 0; prev = -1;
    if (!any_all) return;
    if (!two_all) {                                          // every column sum is 0 or 1
#pragma unroll
        for (int w = NW - 1; w >= 0; --w) if (any[w]) id = 32 * w + __ffs((int)any[w]) - 1;
        best = 1; prev = id == 0 ? -1 : 0;
        return;
    }
    int idw = 0; u32 idbit = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        u32 u = two[w];
        while (u) {
            const u32 low = u & (0u - u);
            u ^= low;
            const int c = __popc(__ballot_sync(0xffffffffu, (x[w] & low) != 0u));
            if (c > best) { prev = best; best = c; idw = w; idbit = low; }
        }
    }
    id = 32 * idw + __ffs((int)idbit) - 1;
    // columns before the winner that were not counted hold 1 (any) or 0
    bool single_before = false;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        if (w < idw) single_before |= any[w] != 0u;
        else if (w == idw) single_before |= (any[w] & (idbit - 1u)) != 0u;
    }
    prev = id == 0 ? -1 : max(prev, single_before ? 1 : 0);
}

// ------------------------------------------------------------------------------------------------
// classify_memo_warp_kernel: one warp per read whose sketch is "every counted k-mer" (number of k-mers <= s), for
// at most 256 references (NW = number of 32-bit words of a reference bitset row, 2..8).
//
//   1. the warp streams the read 32 windows at a time (memo_stream) and looks every code up in the L2-resident table;
//   2. the ranks of the REF windows are compacted into one lane each (reads with more than 32 of them take the block
//      kernel: 0.1 % of the config-2 reads);
//   3. each lane fetches its hit's reference bitset row (table slot from ref[rank]; the copy number of a hash that
//      occurs more than once in the read is its position among equal keys, __match_any_sync -- multiset semantics,
//      SURVEY F10) into registers;
//   4. for every reference some hit belongs to, in index order, count = popc(ballot(bit)) -- and the running
//      best / previous-best scan of src/rkmh.cpp:874-883 is applied on the fly.  No counters, no shared-memory
//      atomics, nothing to clear between reads.
constexpr int MEMO_WARP_MAX_KMERS = 1 << 20;

template <int NW>
__global__ void __launch_bounds__(1024, 1)
classify_memo_warp_kernel(SeqBatch sb, int s, Memo mm, RefTable tb,
                          int* __restrict__ redo, int* __restrict__ redo_count, Hit* __restrict__ out) {
    __shared__ u32 qs[32][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    u32* q = qs[wib];
    const int k = mm.k;
    const u32 lt = (1u << lane) - 1u;
    const u32* __restrict__ tab = mm.tab;

    const i64 gwarp = (i64)blockIdx.x * nwb + wib, nwarps = (i64)gridDim.x * nwb;
    for (i64 rid = gwarp; rid < sb.nseq; rid += nwarps) {
        const i64 beg = sb.offs[rid] - sb.base0, len64 = sb.offs[rid + 1] - sb.offs[rid];
        const i64 n64 = nkmers(len64, k);
        if (n64 > s || n64 > MEMO_WARP_MAX_KMERS) {          // the sketch is a strict subset: block kernel
            if (lane == 0) redo[atomicAdd(redo_count, 1)] = (int)rid;
            continue;
        }
        const unsigned char* __restrict__ p = sb.bases + beg;
        const int n = (int)n64;
        // ---- 1 + 2: stream the windows, compact the REF ranks -------------------------------------------------
        int nc = 0;
        u32 nz = 0;
        if (n > 0) {
            memo_stream(p, len64, k, 0, n, lane, [&](u32 code, bool, bool valid) {
                u32 m = MEMO_ZERO;
                if (valid) m = __ldg(&tab[code]);
                nz += !(m & (MEMO_ZERO | MEMO_FAILM));
                const bool cand = (m & (MEMO_REF | MEMO_ZERO | MEMO_FAILM)) == MEMO_REF;
                const u32 bal = __ballot_sync(0xffffffffu, cand);
                if (bal) {
                    const int at = nc + __popc(bal & lt);
                    if (cand && at < 32) q[at] = m >> 8;
                    nc += __popc(bal);
                }
            });
        }
        if (nc > 32) {
            if (lane == 0) redo[atomicAdd(redo_count, 1)] = (int)rid;
            __syncwarp();
            continue;
        }
        nz = warp_sum(nz);
        int best = 0, id = 0, prev = -1;
        if (nc > 0) {
            __syncwarp();
            // ---- 3: one hit per lane: its reference bitset row ----------------------------------------------------
            u32 x[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) x[w] = 0;
            const bool has = lane < nc;
            const u32 active = __ballot_sync(0xffffffffu, has);
            if (has) {
                const u32 key = __ldg(&mm.ref[q[lane]]);             // slot + 1
                const u32 same = __match_any_sync(active, key);
                const u64* row = table_row(tb, (int)key - 1, (u32)__popc(same & lt));
                if (row != nullptr) {
#pragma unroll
                    for (int w = 0; w < NW / 2; ++w) { const u64 v = __ldg(&row[w]); x[2 * w] = (u32)v; x[2 * w + 1] = (u32)(v >> 32); }
                }
            }
            // ---- 4: per-reference counts + the scan of src/rkmh.cpp:874-883 -----------------------------------------
            memo_scan_rows<NW>(x, lane, best, id, prev);
            __syncwarp();
        }
        if (lane == 0) { Hit hh; hh.ref_id = id; hh.best = best; hh.prev = prev; hh.sketch_len = (int)nz; out[rid] = hh; }
    }
}

// ------------------------------------------------------------------------------------------------
// classify_memo_sieve_kernel: the warp-per-read kernel for the common case (no -M mask baked, no k-mer whose hash
// is 0, at most 256 references).  Same results as classify_memo_warp_kernel; what changes is where the memory
// traffic goes:
//
//   * "is this window a member of some reference sketch?" is answered in SHARED memory by a blocked Bloom filter
//     (the sieve: one 32-bit word per window, two bits per code, built once per rkg_load_refs and copied into each
//     block at launch).  Only the windows that pass (the true members plus ~7 % of the others at 182 references)
//     touch global memory at all; the 64 MB rank table is not read (it only half fits a die's L2: ncu showed 37 % of
//     its lookups going to HBM);
//   * those windows are looked up AFTER the read has been streamed, all lanes at once, in a compact open-addressing
//     map code -> table slot + 1 (8 B entries, load 1/16: 32 MB at 182 references, L2 resident), which replaces the two dependent
//     lookups tab[code] -> ref[rank]; the bitset rows follow as before.  A read costs two dependent L2 round trips
//     instead of seven;
//   * the number of counted windows (sketch_len) is the number of valid windows: no table value is needed for it
//     because no code hashes to 0 and no FAILM flag is baked (checked by the host before this kernel is chosen);
//   * per-reference counts: the 32 x (32 NW) bit matrix (a hit per lane) is transposed in registers (five
//     shuffle steps per word: PRMT for the 16/8-bit steps, rotate + LOP3 for 4/2/1), after which lane l holds, for
//     each word w, the rows that contain reference 32 w + l: its count is one POPC.  The scan of
//     src/rkmh.cpp:874-883 is a warp max over keys (count << 8 | 255 - index) and a second max over the counts in
//     front of the winner.  The cost does not depend on how many references the hits share;
//   * the first 64 bases of a warp's NEXT read are requested before the current read's hits are resolved.
constexpr u32 SIEVE_MUL = 0x9E3779B1u;
constexpr int SIEVE_QCAP = 64;                  // sieve-positive windows per read the kernel resolves itself
constexpr u32 SIEVE_MAX_WORDS = 55040;          // 215 KB of the 227 KB a block may use

RK_D u32 sieve_bits(u32 h) { return (1u << ((h >> 3) & 31u)) | (1u << ((h >> 11) & 31u)); }

// stats[0] = codes that are members of a reference sketch, stats[1] = codes whose hash is 0
__global__ void __launch_bounds__(256)
memo_sieve_kernel(Memo mm, u32* __restrict__ sieve, u32 nwords, u32* __restrict__ stats) {
    for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < mm.ncodes; c += gridDim.x * blockDim.x) {
        const u32 t = mm.tab[c];
        if (t & MEMO_ZERO) atomicAdd(&stats[1], 1u);
        else if (t & MEMO_REF) {
            const u32 h = c * SIEVE_MUL;
            atomicOr(&sieve[__umulhi(h, nwords)], sieve_bits(h));
            atomicAdd(&stats[0], 1u);
        }
    }
}

// code -> slot + 1 of its hash in the reference table; entry = slot + 1 in the high word, code in the low word
__global__ void __launch_bounds__(256)
memo_cmap_kernel(Memo mm, unsigned long long* __restrict__ cmap, u32 mask, int shift) {
    for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < mm.ncodes; c += gridDim.x * blockDim.x) {
        const u32 t = mm.tab[c];
        if ((t & (MEMO_REF | MEMO_ZERO)) != MEMO_REF) continue;
        const u32 il = c;
        const unsigned long long e = ((unsigned long long)mm.ref[t >> 8] << 32) | il;
        u32 i = (il * SIEVE_MUL) >> shift;
        for (;;) {
            if (atomicCAS(&cmap[i], ~0ULL, e) == ~0ULL) break;
            i = (i + 1) & mask;
        }
    }
}

// Per-lane constants of the 32 x 32 bit-matrix transpose (lane = row, bit = column).  Step j swaps, inside every
// 2j x 2j block, the block (rows without bit j, columns with bit j) with (rows with bit j, columns without it).
struct TrLane { u32 sel16, sel8, rot4, rot2, rot1, keep4, keep2, keep1; };

RK_D TrLane tr_lane(int lane) {
    TrLane t;
    t.sel16 = (lane & 16) ? 0x3276u : 0x5410u;
    t.sel8 = (lane & 8) ? 0x3715u : 0x6240u;
    t.rot4 = (lane & 4) ? 28u : 4u;  t.keep4 = (lane & 4) ? 0xF0F0F0F0u : 0x0F0F0F0Fu;
    t.rot2 = (lane & 2) ? 30u : 2u;  t.keep2 = (lane & 2) ? 0xCCCCCCCCu : 0x33333333u;
    t.rot1 = (lane & 1) ? 31u : 1u;  t.keep1 = (lane & 1) ? 0xAAAAAAAAu : 0x55555555u;
    return t;
}

RK_D u32 tr32(u32 x, const TrLane& t) {
    u32 y = __shfl_xor_sync(0xffffffffu, x, 16); x = __byte_perm(x, y, t.sel16);
    y = __shfl_xor_sync(0xffffffffu, x, 8);      x = __byte_perm(x, y, t.sel8);
    y = __shfl_xor_sync(0xffffffffu, x, 4); y = __funnelshift_l(y, y, t.rot4); x = (x & t.keep4) | (y & ~t.keep4);
    y = __shfl_xor_sync(0xffffffffu, x, 2); y = __funnelshift_l(y, y, t.rot2); x = (x & t.keep2) | (y & ~t.keep2);
    y = __shfl_xor_sync(0xffffffffu, x, 1); y = __funnelshift_l(y, y, t.rot1); x = (x & t.keep1) | (y & ~t.keep1);
    return x;
}

template <int NW>
__global__ void __launch_bounds__(1024, 1)                   // measured: 768 threads 9.1 ms, 512 threads 12.5 ms, 1024 threads 7.9 ms
classify_memo_sieve_kernel(SeqBatch sb, int s, Memo mm, RefTable tb,
                           int* __restrict__ redo, int* __restrict__ redo_count, Hit* __restrict__ out) {
#ifdef RK_EMU
    unsigned char* dsm = reinterpret_cast<unsigned char*>(EMU_DYN_SMEM());
#else
    extern __shared__ __align__(16) unsigned char dsm[];
#endif
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    u32* sv = reinterpret_cast<u32*>(dsm);
    const u32 nwords = mm.sieve_words;
    u32* q = sv + nwords + wib * SIEVE_QCAP;
    for (u32 i = threadIdx.x; i < nwords; i += blockDim.x) sv[i] = mm.sieve[i];
    __syncthreads();

    const TrLane tl = tr_lane(lane);
    const int k = mm.k;
    const u32 kmask = mm.kmask;
    const u32 cmask2 = (k >= 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    const u32 lt = (1u << lane) - 1u;
    const unsigned long long* __restrict__ cmap = mm.cmap;
    const u32 cmask = mm.cmap_mask;
    const int cshift = mm.cmap_shift;
    const unsigned char* __restrict__ bases = sb.bases;
    // window at position 32 r + lane of a 128-base chunk: pieces 8 r + (lane >> 2) .. + 3, from base (lane & 3) on
    const int src0 = lane >> 2;
    const u32 osh = 2u * (u32)(lane & 3), ish = (u32)(lane & 3);
    // piece j + i of lane j comes from the next chunk when j + i >= 32 (byte 1 of the shuffled pair instead of byte 0)
    const u32 sel1 = lane >= 31 ? 0x3250u : 0x3240u, sel2 = lane >= 30 ? 0x3510u : 0x3410u, sel3 = lane >= 29 ? 0x5210u : 0x4210u;

    auto cmap_find = [&](u32 code) -> u32 {                  // slot + 1, or 0 when the code is no sketch member
        u32 i = (code * SIEVE_MUL) >> cshift;
        for (;;) {
            const unsigned long long e = __ldg(&cmap[i]);
            if ((u32)e == code) return (u32)(e >> 32);
            if ((u32)e == 0xFFFFFFFFu) return 0u;
            i = (i + 1) & cmask;
        }
    };
    // word `widx` of the read's 4-byte-aligned stream (wp = first word, nbytes = bytes from wp to the end of the read);
    // words that hold no base of the read are not loaded and decode as valid 'AAAA'
    auto load_word = [&](const u32* wp, i64 nbytes, i64 widx) -> u32 { return 4 * widx < nbytes ? __ldg(wp + widx) : 0x41414141u; };

    const i64 nwarps = (i64)gridDim.x * nwb;
    i64 rid = (i64)blockIdx.x * nwb + wib;
    i64 beg = 0, len64 = 0;
    u32 wa = 0x41414141u, wb = 0x41414141u;
    if (rid < sb.nseq) {
        beg = sb.offs[rid] - sb.base0; len64 = sb.offs[rid + 1] - sb.offs[rid];
        const u32* wp = reinterpret_cast<const u32*>(bases + (beg & ~(i64)3));
        const i64 nbytes = (beg & 3) + len64;
        wa = load_word(wp, nbytes, lane); wb = load_word(wp, nbytes, 32 + lane);
    }
    while (rid < sb.nseq) {
        const i64 nrid = rid + nwarps;
        i64 nbeg = 0, nlen = 0;
        if (nrid < sb.nseq) { nbeg = sb.offs[nrid] - sb.base0; nlen = sb.offs[nrid + 1] - sb.offs[nrid]; }
        const i64 n64 = nkmers(len64, k);
        const bool skip = n64 > s || n64 > MEMO_WARP_MAX_KMERS;      // the sketch is a strict subset: block kernel
        // ---- stream the windows through the sieve; queue the codes that pass --------------------------------------
        int nc = 0;
        u32 ninv = 0;                                        // in-range windows that contain an invalid base
        if (!skip && n64 > 0) {
            const int sh = (int)(beg & 3), n = (int)n64;
            const u32* __restrict__ wp = reinterpret_cast<const u32*>(bases + (beg & ~(i64)3));
            const i64 nbytes = sh + len64;
            const int pend = sh + n;                         // stream positions [sh, pend) start a window
            u32 pc, bc, wn = wb;
            sv_decode(wa, pc, bc);
            for (int c0 = 0; c0 < pend; c0 += 128) {
                u32 pn, bn;
                sv_decode(wn, pn, bn);
                wn = load_word(wp, nbytes, (i64)(c0 >> 2) + 64 + lane);       // in flight while this chunk is processed
                const bool slow = __any_sync(0xffffffffu, (bc | bn) != 0u);
                // Q: the 16 bases from piece `lane` on
                const u32 xq = pc | (pn << 8);
                u32 Q = xq;
                Q = __byte_perm(Q, __shfl_sync(0xffffffffu, xq, (lane + 1) & 31), sel1);
                Q = __byte_perm(Q, __shfl_sync(0xffffffffu, xq, (lane + 2) & 31), sel2);
                Q = __byte_perm(Q, __shfl_sync(0xffffffffu, xq, (lane + 3) & 31), sel3);
                u32 IQ = 0;                                  // the same for the invalid flags (one bit per base)
                if (slow) {
                    const u32 xi = sv_bad4(bc) | (sv_bad4(bn) << 4);
                    IQ = xi & 15u;
#pragma unroll
                    for (int i = 1; i < 4; ++i) {
                        const u32 t = __shfl_sync(0xffffffffu, xi, (lane + i) & 31);
                        IQ |= ((lane + i < 32) ? (t & 15u) : (t >> 4)) << (4 * i);
                    }
                }
                const int wi0 = c0 + lane - sh;              // window index of this lane in round 0
                auto round = [&](int r) {
                    const u32 code = (__shfl_sync(0xffffffffu, Q, src0 + 8 * r) >> osh) & cmask2;
                    bool ok = (u32)(wi0 + 32 * r) < (u32)n;
                    if (slow) {
                        const bool inv = ((__shfl_sync(0xffffffffu, IQ, src0 + 8 * r) >> ish) & kmask) != 0u;
                        ninv += ok && inv;
                        ok = ok && !inv;
                    }
                    const u32 h = code * SIEVE_MUL;
                    const u32 bits = sieve_bits(h);
                    const u32 word = sv[__umulhi(h, nwords)];                  // unconditional: no divergent branch in the loop
                    const bool cand = ok & ((word & bits) == bits);
                    const u32 bal = __ballot_sync(0xffffffffu, cand);
                    const int at = nc + __popc(bal & lt);
                    if (cand && at < SIEVE_QCAP) q[at] = code;
                    nc += __popc(bal);
                };
                if (c0 + 96 < pend) {                        // four full rounds, straight-line: their chains interleave
                    round(0); round(1); round(2); round(3);
                } else {
                    for (int r = 0; c0 + 32 * r < pend; ++r) round(r);
                }
                pc = pn; bc = bn;
            }
        }
        // ---- the next read's first 256 bases: in flight while this read's hits are resolved ------------------------
        u32 nwa = 0x41414141u, nwb2 = 0x41414141u;
        if (nrid < sb.nseq) {
            const u32* nwp = reinterpret_cast<const u32*>(bases + (nbeg & ~(i64)3));
            const i64 nnb = (nbeg & 3) + nlen;
            nwa = load_word(nwp, nnb, lane); nwb2 = load_word(nwp, nnb, 32 + lane);
        }
        bool give_up = skip || nc > SIEVE_QCAP;
        int best = 0, id = 0, prev = -1;
        if (!give_up && nc > 0) {
            __syncwarp();
            // ---- codes -> table slots (false positives of the sieve drop out here) --------------------------------
            u32 key = 0, key1 = 0;
            if (lane < nc) key = cmap_find(q[lane]);
            if (32 + lane < nc) key1 = cmap_find(q[32 + lane]);
            const u32 b1 = __ballot_sync(0xffffffffu, key1 != 0u);
            if (b1) {                                        // more than 32 candidates: compact the hits into one per lane
                const u32 b0 = __ballot_sync(0xffffffffu, key != 0u);
                const int nh0 = __popc(b0), nh = nh0 + __popc(b1);
                if (nh > 32) give_up = true;
                else {
                    __syncwarp();
                    if (key) q[__popc(b0 & lt)] = key;
                    if (key1) q[nh0 + __popc(b1 & lt)] = key1;
                    __syncwarp();
                    key = lane < nh ? q[lane] : 0u;
                }
                __syncwarp();
            }
            const u32 active = __ballot_sync(0xffffffffu, key != 0u);
            if (!give_up && active) {
                // ---- one hit per lane: its reference bitset row; the copy number of a hash that occurs more than once
                //      in the read is its position among equal keys (multiset semantics) ---------------------
                u32 x[NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) x[w] = 0;
                if (key) {
                    const u32 same = __match_any_sync(active, key);
                    const u64* row = table_row(tb, (int)key - 1, (u32)__popc(same & lt));
                    if (row != nullptr) {
#pragma unroll
                        for (int w = 0; w < NW / 2; ++w) { const u64 v = __ldg(&row[w]); x[2 * w] = (u32)v; x[2 * w + 1] = (u32)(v >> 32); }
                    }
                }
                // ---- counts per reference (lane l, word w: reference 32 w + l) and the scan of src/rkmh.cpp:874-883 -----
                u32 cnt[NW];
                u32 kbest = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    cnt[w] = (u32)__popc(tr32(x[w], tl));
                    kbest = max(kbest, (cnt[w] << 8) | (255u - (u32)(32 * w + lane)));
                }
                kbest = warp_max(kbest);
                best = (int)(kbest >> 8);
                id = 255 - (int)(kbest & 255u);
                u32 p1 = 0;                                  // max of (count + 1) in front of the winner, 0 = none
#pragma unroll
                for (int w = 0; w < NW; ++w) if (32 * w + lane < id) p1 = max(p1, cnt[w] + 1u);
                prev = (int)warp_max(p1) - 1;
            }
        }
        if (give_up) {
            if (lane == 0) redo[atomicAdd(redo_count, 1)] = (int)rid;
        } else {
            // every window counts unless it holds an invalid base: no code hashes to 0, no FAILM flag is baked
            const int nz = (int)n64 - (int)warp_sum(ninv);
            if (lane == 0) { Hit hh; hh.ref_id = id; hh.best = best; hh.prev = prev; hh.sketch_len = nz; out[rid] = hh; }
        }
        __syncwarp();
        rid = nrid; beg = nbeg; len64 = nlen; wa = nwa; wb = nwb2;
    }
}

// ------------------------------------------------------------------------------------------------
// classify_memo_block_kernel: one block per read (grid-stride over `list`, or over every read when list is null).
// Exact for every input.
//
//   pass 1   every warp streams its share of the read and looks the table value (rank | flags) of every window up, the
//            lookups MEMO_DELAY rounds ahead of their use.  The bottom-s multiset only needs the windows at or under its
//            s-th smallest rank T, so a warp KEEPS (appends to its own list: no atomics, the order is irrelevant) only the
//            counted windows whose rank is under a bound B, a guess of 2 T made from the read's length and the share of
//            counted windows in the block's previous read; the kept values also enter the histogram of the top 12 rank
//            bits.  A 10 kb read at s = 1000 keeps a fifth of its windows: the lists are `lcap` words per warp of shared
//            memory, and everything after pass 1 touches the lists only.
//   check    the guess holds when no list overflowed and the lists hold every counted window, or at least s of them
//            (then the s smallest are among them, with all their copies).  Otherwise pass 1 runs again with B = infinity
//            into lists in global memory that are sized for every window (also the first choice of a read of at most s
//            windows whose share per warp exceeds lcap).
//   select   two-level radix select of T on the 24-bit rank over the lists; `need` copies of rank T belong to the sketch.
//   resolve  the REF members are compacted into a per-warp ring and resolved against the reference table 32 at a time,
//            exactly like the Murmur path does (copy number of a repeated hash from a per-block set of table slots).
//
// Dynamic shared memory (u32 units): [nwarps x (lcap + 32) lists][MEMO_HIST_WORDS histogram][hs_cap keys][hs_cap counts][rpad counters]
// ++ of a 32-bit word of shared memory at shared-window address `saddr` (the address conversion of a generic pointer
// is not repeated at every use)
RK_D void smem_inc(u32* base, u32 saddr_base, u32 idx) {
#ifdef RK_EMU
    (void)saddr_base;
    atomicAdd(&base[idx], 1u);
#else
    (void)base;
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(saddr_base + 4u * idx) : "memory");
#endif
}
RK_D u32 smem_addr(const void* p) {
#ifdef RK_EMU
    (void)p;
    return 0u;
#else
    return (u32)__cvta_generic_to_shared(p);
#endif
}
// zero `words` (a multiple of 4) 32-bit words at the 16-byte aligned p, the whole block
RK_D void block_zero(u32* p, int words) {
    uint4* v = reinterpret_cast<uint4*>(p);
    for (int i = threadIdx.x; i < words / 4; i += blockDim.x) v[i] = make_uint4(0u, 0u, 0u, 0u);
}

struct MemoBlockScratch { u32 part[32]; u32 bin, before; int nz, take, kept, over; ArgmaxScratch as; };
#ifndef MEMO_DELAY_ROUNDS
#define MEMO_DELAY_ROUNDS 4
#endif
#ifndef MEMO_BLOCK_MIN_BLOCKS
#define MEMO_BLOCK_MIN_BLOCKS 3
#endif
constexpr int MEMO_DELAY = MEMO_DELAY_ROUNDS;

// the bin in which the running count of hist[0, MEMO_BINS) reaches `target` (1-based), and the count before it
RK_D void memo_find_bin(const u32* hist, u32 target, MemoBlockScratch* bs) {
    const int per = MEMO_BINS / (int)blockDim.x, b0 = (int)threadIdx.x * per;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    u32 sum = 0;
    for (int i = 0; i < per; ++i) sum += hist[memo_hidx((u32)(b0 + i))];
    u32 incl = sum;
    for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) bs->part[wib] = incl;
    __syncthreads();
    u32 woff = 0;
    for (int w = 0; w < nwb; ++w) if (w < wib) woff += bs->part[w];
    const u32 excl = woff + incl - sum;
    if (excl < target && target <= excl + sum) {
        u32 run = excl;
        for (int i = 0; i < per; ++i) {
            const u32 c = hist[memo_hidx((u32)(b0 + i))];
            if (run + c >= target) { bs->bin = (u32)(b0 + i); bs->before = run; break; }
            run += c;
        }
    }
    __syncthreads();
}

// Pass 1 of classify_memo_block_kernel for one warp: the table value of every window of [s0, s1), kept when it counts and
// lies under `bound` (appended at L[cnt], or at L[cap + lane] once the list is full: cnt > cap tells), kept values into
// the histogram of the top 12 rank bits.  SHARED: L is a shared-memory list (32-bit addressing, predicated stores: the loop
// body is branch-free and the warp stays converged).  The table value of a window is consumed MEMO_DELAY rounds after its
// lookup was issued (the 64 MB table does not fit one die's L2: a tenth of the lookups go to HBM; with two rounds ncu
// showed 2.3 long-scoreboard stalls per issued instruction).
template <bool SHARED>
RK_D void memo_block_pass1(const unsigned char* __restrict__ p, i64 len, int k, i64 s0, i64 s1, int lane,
                           const u32* __restrict__ tab, u32 bound, u32* L, int cap, u32* hist, u32 hist_s, int& cnt_out,
                           u32& nz_out) {
    const u32 ltm = (1u << lane) - 1u;
    const u32 L_s = smem_addr(L);
    u32 ring[MEMO_DELAY];
#pragma unroll
    for (int d = 0; d < MEMO_DELAY; ++d) ring[d] = MEMO_SKIP;
    u32 local_nz = 0;
    int cnt = 0;
    auto consume = [&](u32 m) {
        const bool ok = !(m & (MEMO_ZERO | MEMO_FAILM));            // MEMO_SKIP has every flag set
        local_nz += ok;
        const bool keep = ok && m < bound;
        const u32 bal = __ballot_sync(0xffffffffu, keep);
        const int at = cnt + __popc(bal & ltm);
        const int idx = at < cap ? at : cap + lane;
        const u32 hidx = memo_hidx(m >> 20);
#ifndef RK_EMU
        if (SHARED) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.u32 [%1], %2;\n\t@p red.shared.add.u32 [%3], 1;\n\t}"
                         ::"r"((u32)keep), "r"(L_s + 4u * (u32)idx), "r"(m), "r"(hist_s + 4u * hidx) : "memory");
        } else
#endif
        if (keep) {
            L[idx] = m;
            smem_inc(hist, hist_s, hidx);
        }
        cnt += __popc(bal);
    };
    memo_stream(p, len, k, s0, s1, lane, [&](u32 code, bool, bool valid) {
        u32 m = MEMO_SKIP;
        if (valid) m = __ldg(&tab[code]);
        consume(ring[MEMO_DELAY - 1]);
#pragma unroll
        for (int d = MEMO_DELAY - 1; d > 0; --d) ring[d] = ring[d - 1];
        ring[0] = m;
    });
#pragma unroll
    for (int d = MEMO_DELAY - 1; d >= 0; --d) consume(ring[d]);
    cnt_out = cnt;
    nz_out = warp_sum(local_nz);
}

__global__ void __launch_bounds__(256, MEMO_BLOCK_MIN_BLOCKS)
classify_memo_block_kernel(SeqBatch sb, int s, Memo mm, RefTable tb, const int* __restrict__ list,
                           const int* __restrict__ nlist_ptr, i64 nseq, int lcap, int hs_cap, int rpad,
                           u32* __restrict__ scratch, i64 scratch_stride, Hit* __restrict__ out) {
#ifdef RK_EMU
    unsigned char* dsm = reinterpret_cast<unsigned char*>(EMU_DYN_SMEM());
#else
    extern __shared__ __align__(16) unsigned char dsm[];
#endif
    __shared__ MemoBlockScratch bs;
    __shared__ u32 wq[8][MEMO_RING];                         // per-warp ring of hit ranks (blocks of at most 256 threads)
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    u32* lists = reinterpret_cast<u32*>(dsm);                // per warp lcap values + 32 that catch an overflowing list's stores
    u32* hist = lists + (size_t)nwb * (lcap + 32);
    u32* hs_keys = hist + MEMO_HIST_WORDS;
    u32* hs_cnt = hs_keys + hs_cap;
    u32* counts = hs_cnt + hs_cap;
    const u32 hist_s = smem_addr(hist);
    const int zero_words = MEMO_HIST_WORDS + 2 * hs_cap + rpad;            // histogram, hit set and counters are contiguous
    const u32 hs_mask = (u32)hs_cap - 1;
    const u32 ltm = (1u << lane) - 1u;
    const int k = mm.k;
    float frac = 1.0f;                                       // counted windows / windows of this block's previous read

    const i64 nlist = list ? (i64)*nlist_ptr : nseq;
    for (i64 q = blockIdx.x; q < nlist; q += gridDim.x) {
        const i64 rid = list ? (i64)list[q] : q;
        const i64 beg = sb.offs[rid] - sb.base0, len = sb.offs[rid + 1] - sb.offs[rid];
        const i64 n = nkmers(len, k);
        const i64 per = ((((n + 31) >> 5) + nwb - 1) / nwb) << 5;          // k-mer starts per warp, whole rounds
        const i64 s0 = (i64)wib * per, s1 = min(n, s0 + per);
        // the bound on (rank << 8 | flags): twice the expected s-th smallest rank of the read's counted windows
        u32 bound = 0xFFFFFFFFu;
        if (n > s) {
            const float est = frac * (float)n;                   // counted windows expected
            const float cut = 2.0f * (float)s / (est > 1.0f ? est : 1.0f) * (float)mm.ncodes;
            if (cut < (float)mm.ncodes) bound = (u32)cut << 8;
        }
        bool wide = bound == 0xFFFFFFFFu && per > (i64)lcap;
        u32* L;
        int cnt, nz;
        for (;;) {
            L = wide ? scratch + (i64)blockIdx.x * scratch_stride + (i64)wib * (per + 32) : lists + (size_t)wib * (lcap + 32);
            const int cap = wide ? (int)per : lcap;
            if (threadIdx.x == 0) { bs.nz = 0; bs.take = 0; bs.kept = 0; bs.over = 0; }
            block_zero(hist, zero_words);
            __syncthreads();
            // ---- pass 1 ---------------------------------------------------------------------------------------------
            u32 local_nz;
            if (wide) memo_block_pass1<false>(sb.bases + beg, len, k, s0, s1, lane, mm.tab, bound, L, cap, hist, hist_s, cnt, local_nz);
            else memo_block_pass1<true>(sb.bases + beg, len, k, s0, s1, lane, mm.tab, bound, L, cap, hist, hist_s, cnt, local_nz);
            if (lane == 0) {
                if (local_nz) atomicAdd(&bs.nz, (int)local_nz);
                if (cnt > cap) bs.over = 1; else if (cnt) atomicAdd(&bs.kept, cnt);
            }
            __syncthreads();
            nz = bs.nz;
            const bool good = wide || (!bs.over && (bs.kept == nz || bs.kept >= s));
            if (good) break;
            __syncthreads();                                 // every thread has read the verdict before it is cleared
            wide = true;
            bound = 0xFFFFFFFFu;
        }
        frac = (float)nz / (float)(n > 0 ? n : 1);
        const int Lr = nz < s ? nz : s;
        // ---- bottom-s threshold: rank T, of which `need` copies belong to the sketch -----------------------
        u32 T = 0xFFFFFFFFu; int need = 0;
        if (nz > s) {
            memo_find_bin(hist, (u32)s, &bs);
            const u32 b1 = bs.bin, before1 = bs.before;
            __syncthreads();
            block_zero(hist, MEMO_HIST_WORDS);
            __syncthreads();
            for (int i = lane; i < cnt; i += 32) {
                const u32 m = L[i];
                if ((m >> 20) == b1) smem_inc(hist, hist_s, memo_hidx((m >> 8) & 0xFFFu));
            }
            __syncthreads();
            memo_find_bin(hist, (u32)s - before1, &bs);
            T = (b1 << 12) | bs.bin;
            need = s - (int)before1 - (int)bs.before;
            __syncthreads();
        }
        // ---- resolve the REF members of the sketch: 32 windows per warp step ---------------------------------
        const bool small = tb.nb == 1;
        u32 vp[5][2];
#pragma unroll
        for (int pl = 0; pl < 5; ++pl) { vp[pl][0] = 0; vp[pl][1] = 0; }
        int vsteps = 0;
        auto flush_planes = [&]() {
            for (u32 j = 0; j < tb.nrefs; ++j) {
                u32 c = 0;
#pragma unroll
                for (int pl = 0; pl < 5; ++pl) c |= (((j < 32 ? vp[pl][0] : vp[pl][1]) >> (j & 31)) & 1u) << pl;
                c = warp_sum(c);
                if (lane == 0 && c) atomicAdd(&counts[j], c);
            }
#pragma unroll
            for (int pl = 0; pl < 5; ++pl) { vp[pl][0] = 0; vp[pl][1] = 0; }
        };
        // The scan over the read's windows only picks the REF members of the sketch (a few per cent of the windows of a
        // long read) and compacts their ranks into a per-warp ring; rows are resolved 32 hits at a time, all lanes busy.
        u32* ring = wq[wib];
        int qhead = 0, qn = 0;
        auto drain = [&](int cnt) {                          // resolve the first cnt (<= 32) queued ranks
            const u64* row = nullptr;
            if (lane < cnt) {
                const u32 key = mm.ref[ring[(qhead + lane) & (MEMO_RING - 1)]];          // slot + 1
                u32 p = (key * 0x9E3779B1u) & hs_mask;
                for (;;) {
                    const u32 old = atomicCAS(&hs_keys[p], 0u, key);
                    if (old == 0u || old == key) break;
                    p = (p + 1) & hs_mask;
                }
                row = table_row(tb, (int)key - 1, atomicAdd(&hs_cnt[p], 1u));
            }
            qhead = (qhead + cnt) & (MEMO_RING - 1);
            qn -= cnt;
            if (small) {
                // <= 64 references: every lane adds its row into private bit-sliced counters (plane p = bit p of the
                // per-reference counts), flushed to the shared counters every 31 batches and at the end of the read
                if (row != nullptr) {
                    const u64 v = row[0];
                    u32 cy0 = (u32)v, cy1 = (u32)(v >> 32);
#pragma unroll
                    for (int pl = 0; pl < 5; ++pl) {
                        const u32 t0 = vp[pl][0] & cy0, t1 = vp[pl][1] & cy1;
                        vp[pl][0] ^= cy0; vp[pl][1] ^= cy1;
                        cy0 = t0; cy1 = t1;
                    }
                }
                if (++vsteps == 31) { flush_planes(); vsteps = 0; }
                return;
            }
            if (!__any_sync(0xffffffffu, row != nullptr)) return;
            // column sums of this batch's rows: one shared-memory add per reference that occurs, not per set bit
            const u32* r32 = reinterpret_cast<const u32*>(row);
            for (u32 w = 0; w < 2u * tb.nb; ++w) {
                const u32 xw = row ? r32[w] : 0u;
                u32 u = warp_or(xw);
                while (u) {
                    const u32 low = u & (0u - u);
                    u ^= low;
                    const u32 c = (u32)__popc(__ballot_sync(0xffffffffu, (xw & low) != 0u));
                    if (lane == 0) atomicAdd(&counts[32u * w + (u32)__ffs((int)low) - 1u], c);
                }
            }
        };
        for (int i0 = 0; i0 < cnt; i0 += 32) {               // the warp's own list
            const int i = i0 + lane;
            bool hit = false;
            u32 m = 0;
            if (i < cnt) {
                m = L[i];
                const u32 rank = m >> 8;
                const bool member = rank < T || (rank == T && atomicAdd(&bs.take, 1) < need);
                hit = member && (m & MEMO_REF);
            }
            const u32 bal = __ballot_sync(0xffffffffu, hit);
            if (bal) {
                if (hit) ring[(qhead + qn + __popc(bal & ltm)) & (MEMO_RING - 1)] = m >> 8;
                qn += __popc(bal);
                __syncwarp();
                if (qn >= 32) { drain(32); __syncwarp(); }
            }
        }
        if (qn > 0) { __syncwarp(); drain(qn); }
        if (small && vsteps) flush_planes();
        __syncthreads();
        block_argmax(counts, (int)tb.nrefs, &bs.as, &out[rid], Lr);
    }
}

}  // namespace rk
