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
//   * only now are the codes expanded to the ASCII bytes Murmur reads (the reference hashes the letters), four bases
//     per PRMT, and the two MurmurHash3_x64_128 of rk_common.cuh run with all 32 lanes on DIFFERENT windows.
// Validity (canonical(), mkmh.hpp:182-188) is a funnel shift of the mask words, skipped for chunks without an invalid
// base.  Per k-mer this is 2 x Murmur + ~45 instructions, against ~320 of overhead for the rolling per-lane windows of
// round 1 (k - 1 priming pushes per lane run, one PRMT roll per register and base).
#pragma once
#include "rk_common.cuh"

namespace rk {

typedef long long i64;

// codes -> ASCII: byte i of the result = table byte (code of group i); `g` holds four 2-bit codes in its low byte
constexpr u32 LETTERS_FWD = 0x47544341u;      // code 0 1 2 3 -> A C T G
// reverse strand: the codes arrive bit-reversed (value b0 b1 for code b1 b0) and must be complemented (code ^ 2):
// index 0 -> code 0 -> T, 1 -> code 2 -> A, 2 -> code 1 -> G, 3 -> code 3 -> C
constexpr u32 LETTERS_REV = 0x43474154u;

RK_D u32 letters4(u32 x, int byte, u32 table) {
    u32 g = byte_perm(x, 0u, 0x4440u + (u32)byte);            // the four codes of this group in bits 0..7
    g = (g | (g << 4)) & 0x0F0Fu;
    g = (g | (g << 2)) & 0x3333u;                             // code i in nibble i
    return byte_perm(table, 0u, g);
}

RK_D u32 brev32(u32 x) { return __brev(x); }

// Canonical hash (mkmh.hpp:515-528) of the window whose 2-bit codes are lo (bases 0..15) and hi (bases 16..31).
template <int KV>
RK_D u64 window_hash(u32 lo, u32 hi, int k) {
    constexpr int NW = KVar<KV>::NW;
    u32 f[NW], r[NW];
    // reverse complement in code space: the last base first
    u32 rlo, rhi = 0;
    if (k <= 16) {
        rlo = brev32(lo) >> (32 - 2 * k);
    } else {
        const u32 bl = brev32(lo), bh = brev32(hi);           // 64-bit reversal = (brev(lo) : brev(hi)), then >> (64 - 2k)
        const u32 sh = (u32)(64 - 2 * k);                     // 0..30
        rlo = sh ? __funnelshift_r(bh, bl, sh) : bh;
        rhi = bl >> sh;
    }
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        f[j] = letters4(j < 4 ? lo : hi, j & 3, LETTERS_FWD);
        r[j] = letters4(j < 4 ? rlo : rhi, j & 3, LETTERS_REV);
    }
    if (k & 3) {                                              // bytes beyond k are zero in the key Murmur reads
        const u32 tail = (1u << (8 * (k & 3))) - 1u;
        f[NW - 1] &= tail; r[NW - 1] &= tail;
    }
    const u64 hf = murmur_kmer<KV>(f, k), hr = murmur_kmer<KV>(r, k);
    return hf < hr ? hf : hr;
}

constexpr int WARP_CHUNK = 256;               // windows per chunk: 8 rounds of 32

// One warp hashes windows [w0, w1) of the packed sequence that starts at base index seq_begin and calls
// f(window index, hash, in_range) on every lane each round (warp-convergent; hash = 0 for an invalid or out-of-range
// window).  `last_word` / `last_mword` are the last loadable indices of codes[] / nmask[] (padding included).
template <int KV, typename F>
RK_D void warp_hash(const u32* __restrict__ codes, const u32* __restrict__ nmask, i64 last_word, i64 last_mword,
                    i64 seq_begin, i64 w0, i64 w1, int k, int lane, F&& f) {
    const u32 kmask = k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u);
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
            if (ok) h = window_hash<KV>(lo, hi, k);
            f(wi, h, wi < w1);
        }
    }
}

template <typename F>
RK_D void warp_hash_kv(int kv, const u32* __restrict__ codes, const u32* __restrict__ nmask, i64 last_word, i64 last_mword,
                       i64 seq_begin, i64 w0, i64 w1, int k, int lane, F&& f) {
    switch (kv) {
        case 0: warp_hash<0>(codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 1: warp_hash<1>(codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 2: warp_hash<2>(codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 3: warp_hash<3>(codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 4: warp_hash<4>(codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 5: warp_hash<5>(codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 6: warp_hash<6>(codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 7: warp_hash<7>(codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        case 8: warp_hash<8>(codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
        default: warp_hash<9>(codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f); break;
    }
}
// KCSEL < NUM_KV: every k of the list has that key-length variant (no switch in the instantiation)
template <int KCSEL, typename F>
RK_D void warp_hash_sel(int kv, const u32* __restrict__ codes, const u32* __restrict__ nmask, i64 last_word, i64 last_mword,
                        i64 seq_begin, i64 w0, i64 w1, int k, int lane, F&& f) {
    if constexpr (KCSEL < NUM_KV) warp_hash<KCSEL>(codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f);
    else warp_hash_kv(kv, codes, nmask, last_word, last_mword, seq_begin, w0, w1, k, lane, f);
}

// A team of nthreads threads (whole warps; thread tid) covers windows [0, n) of one sequence, a chunk of 256 windows per
// warp at a time, and hands every window to emit(i, hash).  KCSEL as above.
template <int KCSEL, typename F>
RK_D void team_hash_k(int kv, const u32* __restrict__ codes, const u32* __restrict__ nmask, i64 last_word, i64 last_mword,
                      i64 seq_begin, i64 n, int k, int tid, int nthreads, F&& emit) {
    const int lane = tid & 31, w = tid >> 5, nw = nthreads >> 5;
    for (i64 c0 = (i64)w * WARP_CHUNK; c0 < n; c0 += (i64)nw * WARP_CHUNK) {
        const i64 c1 = c0 + WARP_CHUNK < n ? c0 + WARP_CHUNK : n;
        warp_hash_sel<KCSEL>(kv, codes, nmask, last_word, last_mword, seq_begin, c0, c1, k, lane,
                             [&](i64 i, u64 h, bool in) { if (in) emit(i, h); });
    }
}

}  // namespace rk
