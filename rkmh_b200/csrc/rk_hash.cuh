// rk_hash.cuh -- warp-cooperative canonical k-mer hashing for every k 1..32: the Murmur path of reference
// mkmh/mkmh.hpp:505-532 (calc_hashes) without any per-lane window priming.  Used by every kernel that hashes.
//
// A warp owns a run of consecutive windows of one packed sequence and walks it 256 windows at a time:
//   * per 256-window chunk every lane loads ONE 32-bit word of 2-bit codes (lanes 0..18 cover the chunk's 256 + 31
//     bases plus the start's misalignment) and ONE word of the invalid-base mask;
//   * per round of 32 windows a lane gets the 32 bases that start at ITS window with two or three shuffles and one or
//     two funnel shifts -- the source lanes advance by a constant per round, the shift never changes inside a chunk;
//     the reverse complement is one bit reversal (`brev` reverses the order of the 2-bit codes; the bit swap inside
//     each code and the complement are folded into the letter table of that strand);
//   * the reference hashes the ASCII letters of the k-mer; here the codes index shared-memory tables of
//     (letters * Murmur constant), see "table-driven key * constant" below, and the two MurmurHash3_x64_128 run with all
//     32 lanes on DIFFERENT windows.
// Validity (canonical(), mkmh.hpp:182-188) is a funnel shift of the mask words, skipped for chunks without an invalid
// base.  Per k-mer this is 2 x Murmur + ~45 instructions, against ~320 of overhead for the rolling per-lane windows of
// round 1 (k - 1 priming pushes per lane run, one PRMT roll per register and base).
#pragma once
#include "rk_common.cuh"

namespace rk {

typedef long long i64;

// ---- table-driven key * constant -------------------------------------------------------------------------------
// Murmur's first step on every 64-bit lane K of the key is K * c (c1 for the k1 lanes, c2 for the k2 lanes,
// murmur3.cpp:250-256, 277-289).  K is eight ASCII letters, i.e. two groups of four 2-bit codes, and multiplication
// distributes over the sum of the two 32-bit halves: K * c = T[c][lo][g_lo] + T[c][hi][g_hi] (mod 2^64) with
// T[c][half][g] = (letters(g) << 32 half) * c.  Eight 256-entry tables (2 strands x 2 constants x 2 halves, 16 KB of
// shared memory) therefore replace the expansion of the codes to ASCII bytes AND the first multiply of every lane:
// two LDS.64 + one 64-bit add per lane instead of ~14 ALU instructions + 3 IMAD.  The reverse strand's tables take the
// bit-reversed codes brev() leaves (value b0 b1 for code b1 b0) and hold the complement's letter.  A last word with
// fewer than four letters (k mod 4 != 0) is looked up with its missing codes as 0 and the constant contribution of
// those phantom letters is subtracted.  The result is exactly MurmurHash3_x64_128 of the ASCII k-mer.
constexpr u32 LETTERS_FWD = 0x47544341u;      // code 0 1 2 3 -> A C T G
// reverse strand: index 0 -> code 0 -> complement T, 1 -> code 2 -> A, 2 -> code 1 -> G, 3 -> code 3 -> C
constexpr u32 LETTERS_REV = 0x43474154u;
constexpr int HASH_LUT_ENTRIES = 8 * 256;

RK_HD u32 letters_of(u32 g, u32 table) {                     // the four letters of the four 2-bit codes in g
    u32 w = 0;
    for (int i = 0; i < 4; ++i) w |= ((table >> (8 * ((g >> (2 * i)) & 3u))) & 0xFFu) << (8 * i);
    return w;
}
RK_HD int hash_lut_index(int strand, int cidx, int half) { return ((strand * 2 + cidx) * 2 + half) * 256; }

// host: fill the 2048-entry table (copied to the device once per context)
inline void build_hash_lut(u64* out) {
    for (int strand = 0; strand < 2; ++strand)
        for (int cidx = 0; cidx < 2; ++cidx)
            for (int half = 0; half < 2; ++half)
                for (u32 g = 0; g < 256; ++g) {
                    const u64 letters = letters_of(g, strand ? LETTERS_REV : LETTERS_FWD);
                    out[hash_lut_index(strand, cidx, half) + g] = (letters << (32 * half)) * (cidx ? MM_C2 : MM_C1);
                }
}

// block-wide: copy the table into shared memory (call once at kernel start, by every thread)
RK_D const u64* hash_lut_load(const u64* __restrict__ g) {
    __shared__ __align__(16) u64 s_hash_lut[HASH_LUT_ENTRIES];
    for (int i = threadIdx.x; i < HASH_LUT_ENTRIES; i += blockDim.x) s_hash_lut[i] = g[i];
    __syncthreads();
    return s_hash_lut;
}

RK_D W64 lut64(const u64* t, u32 idx) { const u64 v = t[idx]; return w64((u32)v, (u32)(v >> 32)); }
RK_D W64 sub64(W64 a, W64 b) {
    const u64 v = (((u64)a.hi << 32) | a.lo) - (((u64)b.hi << 32) | b.lo);
    return w64((u32)v, (u32)(v >> 32));
}

// what has to be subtracted from the product of the lane that holds the key's last word when k mod 4 != 0
struct HashCorr { W64 f, r; };
template <int KV>
RK_D HashCorr make_hash_corr(int k) {
    constexpr int NW = KVar<KV>::NW;
    HashCorr c; c.f = w64(0, 0); c.r = w64(0, 0);
    const int r4 = k & 3;
    if (r4) {
        const u32 phantom = 0x01010101u & ~((1u << (8 * r4)) - 1u);           // one in every missing byte
        const int half = (NW - 1) & 1;
        const u64 mul = (((NW - 1) >> 1) & 1) ? MM_C2 : MM_C1;
        const u64 pf = ((u64)(phantom * 0x41u) << (32 * half)) * mul;        // code 0 -> 'A'
        const u64 pr = ((u64)(phantom * 0x54u) << (32 * half)) * mul;        // reverse table index 0 -> 'T'
        c.f = w64((u32)pf, (u32)(pf >> 32)); c.r = w64((u32)pr, (u32)(pr >> 32));
    }
    return c;
}

// MurmurHash3_x64_128 (rk_common.cuh: murmur_kmer) from the products p[L] = K_L * c of the key's 64-bit lanes
template <int KV>
RK_D u64 murmur_from_products(const W64* p, int k) {
    W64 h1 = w64(MURMUR_SEED, 0u), h2 = w64(MURMUR_SEED, 0u);
    if constexpr (KVar<KV>::BODIES >= 1) {           // murmur3.cpp:248-259
        h1 = xor64(h1, mul_c<MM_C2>(rotl_c<31>(p[0])));
        h1 = rotl_c<27>(h1); h1 = add64(h1, h2); h1 = mul5_add<0x52dce729u>(h1);
        h2 = xor64(h2, mul_c<MM_C1>(rotl_c<33>(p[1])));
        h2 = rotl_c<31>(h2); h2 = add64(h2, h1); h2 = mul5_add<0x38495ab5u>(h2);
    }
    if constexpr (KVar<KV>::BODIES == 2) {
        h1 = xor64(h1, mul_c<MM_C2>(rotl_c<31>(p[2])));
        h1 = rotl_c<27>(h1); h1 = add64(h1, h2); h1 = mul5_add<0x52dce729u>(h1);
        h2 = xor64(h2, mul_c<MM_C1>(rotl_c<33>(p[3])));
        h2 = rotl_c<31>(h2); h2 = add64(h2, h1); h2 = mul5_add<0x38495ab5u>(h2);
    }
    if constexpr (KVar<KV>::TAIL) {                  // murmur3.cpp:269-289
        constexpr int T = KVar<KV>::NW - 4 * KVar<KV>::BODIES;   // 1..4 tail words
        constexpr int B = 2 * KVar<KV>::BODIES;
        if constexpr (T >= 3) h2 = xor64(h2, mul_c<MM_C1>(rotl_c<33>(p[B + 1])));
        h1 = xor64(h1, mul_c<MM_C2>(rotl_c<31>(p[B])));
    }
    h1.lo ^= (u32)k; h2.lo ^= (u32)k;                // murmur3.cpp:294-305 (len < 2^32)
    h1 = add64(h1, h2); h2 = add64(h2, h1);
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 = add64(h1, h2);
    return ((u64)h1.lo << 32) | (u64)h1.hi;          // halves swapped
}

RK_D u32 brev32(u32 x) { return __brev(x); }

// Canonical hash (mkmh.hpp:515-528) of the window whose 2-bit codes are lo (bases 0..15) and hi (bases 16..31).
template <int KV>
RK_D u64 window_hash(u32 lo, u32 hi, int k, const u64* __restrict__ lut, const HashCorr& hc) {
    constexpr int NW = KVar<KV>::NW, NP = (NW + 1) / 2;
    // reverse complement in code space: the last base first (the complement is in the reverse strand's tables)
    u32 rlo, rhi = 0;
    if (k <= 16) {
        rlo = brev32(lo) >> (32 - 2 * k);
    } else {
        const u32 bl = brev32(lo), bh = brev32(hi);           // 64-bit reversal = (brev(lo) : brev(hi)), then >> (64 - 2k)
        const u32 sh = (u32)(64 - 2 * k);                     // 0..30
        rlo = sh ? __funnelshift_r(bh, bl, sh) : bh;
        rhi = bl >> sh;
    }
    const u32 lastmask = (k & 3) ? ((1u << (2 * (k & 3))) - 1u) : 0xFFu;      // codes of the forward key's last word
    W64 pf[NP], pr[NP];
#pragma unroll
    for (int L = 0; L < NP; ++L) {
        const int cidx = L & 1;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int w = 2 * L + half;                       // word of the key
            if (w >= NW) continue;
            u32 gf = ((w < 4 ? lo : hi) >> (8 * (w & 3))) & 0xFFu;
            const u32 gr = ((w < 4 ? rlo : rhi) >> (8 * (w & 3))) & 0xFFu;     // zero beyond k already
            if (w == NW - 1) gf &= lastmask;
            const W64 tf = lut64(lut, (u32)hash_lut_index(0, cidx, half) + gf);
            const W64 tr = lut64(lut, (u32)hash_lut_index(1, cidx, half) + gr);
            if (half == 0) { pf[L] = tf; pr[L] = tr; } else { pf[L] = add64(pf[L], tf); pr[L] = add64(pr[L], tr); }
        }
    }
    if (k & 3) { pf[NP - 1] = sub64(pf[NP - 1], hc.f); pr[NP - 1] = sub64(pr[NP - 1], hc.r); }
    const u64 hf = murmur_from_products<KV>(pf, k), hr = murmur_from_products<KV>(pr, k);
    return hf < hr ? hf : hr;
}

constexpr int WARP_CHUNK = 256;               // windows per chunk: 8 rounds of 32

// One warp hashes windows [w0, w1) of the packed sequence that starts at base index seq_begin and calls
// f(window index, hash, in_range) on every lane each round (warp-convergent; hash = 0 for an invalid or out-of-range
// window).  `last_word` / `last_mword` are the last loadable indices of codes[] / nmask[] (padding included).
template <int KV, typename F>
RK_D void warp_hash(const u64* __restrict__ lut, const u32* __restrict__ codes, const u32* __restrict__ nmask, i64 last_word,
                    i64 last_mword, i64 seq_begin, i64 w0, i64 w1, int k, int lane, F&& f) {
    const u32 kmask = k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u);
    const HashCorr hc = make_hash_corr<KV>(k);
    for (i64 c0 = w0; c0 < w1; c0 += WARP_CHUNK) {
        const i64 g0 = seq_begin + c0;
        const i64 cw = (g0 >> 4) + lane, mw = (g0 >> 5) + lane;
        const u32 W = __ldg(codes + (cw < last_word ? cw : last_word));
        const u32 M = __ldg(nmask + (mw < last_mword ? mw : last_mword));
        const u32 q = (u32)(g0 & 15) + (u32)lane, qm = (u32)(g0 & 31) + (u32)lane;
        const int wb = (int)(q >> 4), mb = (int)(qm >> 5);
        const u32 sh = 2u * (q & 15u), msh = qm & 31u;
        const bool dirty = __any_sync(0xffffffffu, lane < 10 && M != 0u);
        const int rounds = (int)((((w1 - c0) < WARP_CHUNK ? (w1 - c0) : WARP_CHUNK) + 31) >> 5);
        for (int r = 0; r < rounds; ++r) {
            const u32 a0 = __shfl_sync(0xffffffffu, W, wb + 2 * r), a1 = __shfl_sync(0xffffffffu, W, wb + 2 * r + 1);
            const u32 lo = __funnelshift_r(a0, a1, sh);
            u32 hi = 0;
            if (k > 16) { const u32 a2 = __shfl_sync(0xffffffffu, W, wb + 2 * r + 2); hi = __funnelshift_r(a1, a2, sh); }
            const i64 wi = c0 + 32 * r + lane;
            bool ok = wi < w1;
            if (dirty) {
                const u32 m0 = __shfl_sync(0xffffffffu, M, mb + r), m1 = __shfl_sync(0xffffffffu, M, mb + r + 1);
                ok = ok && (__funnelshift_r(m0, m1, msh) & kmask) == 0u;
            }
            u64 h = 0;
            if (ok) h = window_hash<KV>(lo, hi, k, lut, hc);
            f(wi, h, wi < w1);
        }
    }
}

template <typename F>
RK_D void warp_hash_kv(int kv, const u64* __restrict__ lut, const u32* __restrict__ codes, const u32* __restrict__ nmask, i64 last_word, i64 last_mword,
                       i64 seq_begin, i64 w0, i64 w1, int k, int lane, F&& f) {
    switch (kv) {
        case 0: warp_hash<0>(lut, codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 1: warp_hash<1>(lut, codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 2: warp_hash<2>(lut, codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 3: warp_hash<3>(lut, codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 4: warp_hash<4>(lut, codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 5: warp_hash<5>(lut, codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 6: warp_hash<6>(lut, codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 7: warp_hash<7>(lut, codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 8: warp_hash<8>(lut, codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        default: warp_hash<9>(lut, codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
    }
}
// KCSEL < NUM_KV: every k of the list has that key-length variant (no switch in the instantiation)
template <int KCSEL, typename F>
RK_D void warp_hash_sel(int kv, const u64* __restrict__ lut, const u32* __restrict__ codes, const u32* __restrict__ nmask, i64 last_word, i64 last_mword,
                        i64 seq_begin, i64 w0, i64 w1, int k, int lane, F&& f) {
    if constexpr (KCSEL < NUM_KV) warp_hash<KCSEL>(lut, codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f);
    else warp_hash_kv(kv, lut, codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f);
}

// A team of nthreads threads (whole warps; thread tid) covers windows [0, n) of one sequence, a chunk of 256 windows per
// warp at a time, and hands every window to emit(i, hash).  KCSEL as above.
template <int KCSEL, typename F>
RK_D void team_hash_k(int kv, const u64* __restrict__ lut, const u32* __restrict__ codes, const u32* __restrict__ nmask, i64 last_word, i64 last_mword,
                      i64 seq_begin, i64 n, int k, int tid, int nthreads, F&& emit) {
    const int lane = tid & 31, w = tid >> 5, nw = nthreads >> 5;
    for (i64 c0 = (i64)w * WARP_CHUNK; c0 < n; c0 += (i64)nw * WARP_CHUNK) {
        const i64 c1 = c0 + WARP_CHUNK < n ? c0 + WARP_CHUNK : n;
        warp_hash_sel<KCSEL>(kv, lut, codes, nmask, last_word, last_mword, seq_begin, c0, c1, k, lane,
                             [&](i64 i, u64 h, bool in) { if (in) emit(i, h); });
    }
}

}  // namespace rk
