// rk_call.cuh -- the device side of `rkmh call` (reference src/rkmh.cpp:1455-1904): alignment-free variant
// calling from k-mer depth.
//
//   depth_add_kernel      every window of every read (calc_hashes, one k) into an EXACT hash -> count table: the
//                         reference's unordered_map<hash_t, int> read_hash_to_depth (src/rkmh.cpp:1570, 1617-1622); the
//                         invalid windows' hash 0 is a key like any other there, here it has its own counter
//   depth_lookup_kernel   the depth of every window of a reference sequence (src/rkmh.cpp:1777)
//   depth_rescue_kernel   for the windows the host found under-covered: the depth of every single-base substitution of the
//                         window (k positions x 3 bases, in rotate_snps order, src/rkmh.cpp:1651-1667, 1796-1831) and of
//                         every one-base deletion of the (k + 1)-mer that starts one base earlier (src/rkmh.cpp:1840-1858)
// The sliding-window average, the thresholds and the VCF lines stay on the host (rkmh_main.cpp): they are sequential in
// the reference position and tiny.
#pragma once
#include "rk_kernels.cuh"

namespace rk {

struct DepthTable {
    u64* keys;                     // 0 = empty
    u32* counts;
    unsigned long long* zero;      // occurrences of hash 0 (invalid windows)
    u32 cap_mask;
};

RK_D void depth_add(const DepthTable& dt, u64 h) {
    if (h == 0) { atomicAdd(dt.zero, 1ULL); return; }
    u32 slot = (u32)(h ^ (h >> 32)) & dt.cap_mask;
    for (;;) {
        const u64 old = atomicCAS(&dt.keys[slot], 0ULL, h);
        if (old == 0ULL || old == h) break;
        slot = (slot + 1) & dt.cap_mask;
    }
    atomicAdd(&dt.counts[slot], 1u);
}

RK_D u32 depth_get(const DepthTable& dt, u64 h) {
    if (h == 0) return (u32)*dt.zero;
    u32 slot = (u32)(h ^ (h >> 32)) & dt.cap_mask;
    for (;;) {
        const u64 key = dt.keys[slot];
        if (key == h) return dt.counts[slot];
        if (key == 0) return 0u;
        slot = (slot + 1) & dt.cap_mask;
    }
}

template <int KCSEL>
__global__ void __launch_bounds__(256)
depth_add_kernel(SeqBatch sb, const __grid_constant__ KList kl, DepthTable dt) {
    const u64* lut = hash_lut_load(kl.lut);
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const int k = kl.k[0];
    for (i64 s = warp; s < sb.nseq; s += nwarps) {
        const i64 beg = sb.offs[s] - sb.base0, len = sb.offs[s + 1] - sb.offs[s];
        team_hash_k<KCSEL>(kl.kv[0], lut, sb.codes, sb.nmask, codes_last_word(sb), nmask_last_word(sb), beg, nkmers(len, k), k, lane, 32,
                           [&](i64, u64 h) { depth_add(dt, h); });
    }
}

// depth of windows [0, n) of the single sequence of the batch
template <int KCSEL>
__global__ void __launch_bounds__(256)
depth_lookup_kernel(SeqBatch sb, const __grid_constant__ KList kl, DepthTable dt, i64 n, int* __restrict__ depth) {
    const u64* lut = hash_lut_load(kl.lut);
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const int k = kl.k[0];
    for (i64 c0 = warp * WARP_CHUNK; c0 < n; c0 += nwarps * WARP_CHUNK) {
        const i64 c1 = c0 + WARP_CHUNK < n ? c0 + WARP_CHUNK : n;
        warp_hash_sel<KCSEL>(kl.kv[0], lut, sb.codes, sb.nmask, codes_last_word(sb), nmask_last_word(sb), 0, c0, c1, k, lane,
                             [&](i64 i, u64 h, bool in) { if (in) depth[i] = (int)depth_get(dt, h); });
    }
}

// to_upper of the reference (mkmh.hpp:231-236), as the sequences are after parse_fastas
RK_D u32 call_upper(u32 c) { return ((int)(signed char)c - 91 > 0) ? c - 32u : c; }

// rotate_snps (src/rkmh.cpp:1646-1667): the three bases tried at a position, in the reference's order
RK_D u32 rotate_snp(u32 orig, int which) {
    switch (orig) {
        case 'A': return which == 0 ? 'C' : which == 1 ? 'T' : 'G';
        case 'C': return which == 0 ? 'T' : which == 1 ? 'G' : 'A';
        case 'T': return which == 0 ? 'C' : which == 1 ? 'G' : 'A';
        case 'G': return which == 0 ? 'A' : which == 1 ? 'C' : 'T';
        default: return 0;                               // not one of ACGT: no substitution is tried (the reference's lambda
    }                                                    // returns nothing there: undefined behaviour)
}

// One thread per (window, alternative).  out[w * 4k + a]: a < 3k -> substitution (position a / 3, base a % 3),
// a >= 3k -> deletion of base (a - 3k + 1) of the (k + 1)-mer starting at the window's position - 1.  -1 = not tried.
template <int KV>
__global__ void __launch_bounds__(256)
depth_rescue_kernel(const unsigned char* __restrict__ seq, i64 len, const __grid_constant__ KList kl, DepthTable dt,
                    const i64* __restrict__ windows, i64 nwin, int* __restrict__ out) {
    const u64* lut = hash_lut_load(kl.lut);
    const int k = kl.k[0];
    const HashCorr hc = make_hash_corr<KV>(k);
    const i64 total = nwin * 4 * k;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
        const i64 w = t / (4 * k);
        const int a = (int)(t - w * 4 * k);
        const i64 j = windows[w];
        u32 lo = 0, hi = 0;
        bool valid = true, tried = true;
        if (a < 3 * k) {
            const int pos = a / 3;
            const u32 alt = rotate_snp(call_upper(seq[j + pos]), a - 3 * pos);
            tried = alt != 0;
            for (int i = 0; i < k; ++i) {
                const u32 c = (i == pos) ? alt : call_upper(seq[j + i]);
                valid = valid && base_valid(c) && c < 'a';
                const u32 code = base_code(c);
                if (i < 16) lo |= code << (2 * i); else hi |= code << (2 * (i - 16));
            }
        } else {
            const int del = a - 3 * k + 1;                   // 1..k: index inside the (k + 1)-mer that starts at j - 1
            tried = j > 0;
            if (tried)
                for (int i = 0, src = 0; i < k; ++i, ++src) {
                    if (src == del) ++src;
                    const u32 c = call_upper(seq[j - 1 + src]);
                    valid = valid && base_valid(c) && c < 'a';
                    const u32 code = base_code(c);
                    if (i < 16) lo |= code << (2 * i); else hi |= code << (2 * (i - 16));
                }
        }
        int d = -1;
        if (tried) d = (int)depth_get(dt, valid ? window_hash<KV>(lo, hi, k, lut, hc) : 0ULL);
        out[t] = d;
    }
}

}  // namespace rk
