// rk_common.cuh -- device/host building blocks of the rkmh hot path for sm_100a.
//
//   * exact `x % d` for the lossy count tables (reference mkmh/HASHTCounter.cpp:46-58)
//   * MurmurHash3_x64_128 (seed 42) specialised to k-mers held in registers
//     (reference mkmh/murmur3/murmur3.cpp:234-312, called from mkmh/mkmh.hpp:517-522)
//   * a rolling window that keeps the ASCII forward k-mer and its reverse complement in registers
//     (reference mkmh/mkmh.hpp:182-188 canonical(), :204-215 reverse_complement())
//
// Everything here is __host__ __device__ so the arithmetic is unit-tested on the CPU as well
// (tests/emu/), where the file is compiled by g++ with tests/emu/cuda_emu.h standing in for CUDA.
#pragma once
#include <stdint.h>

#ifndef RK_EMU
#include <cuda_runtime.h>
#define RK_HD __host__ __device__ __forceinline__
#define RK_D __device__ __forceinline__
#else
#define RK_HD inline __attribute__((always_inline))
#define RK_D inline __attribute__((always_inline))
#endif

namespace rk {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr u32 MURMUR_SEED = 42;  // mkmh/mkmh.hpp:517
constexpr int MAX_K = 32;
constexpr int MAX_NK = 8;

// ------------------------------------------------------------------------------------------------
// Exact n % d for a run-time divisor d (2 <= d < 2^63), by multiplication with a precomputed 65-bit
// reciprocal (Granlund & Montgomery 1994, round-up method).  The count-table sizes of the reference
// (200,000,000 and 10,000,000) are not powers of two and collisions are part of its results
// (SURVEY F11), so the modulus must be exact.
struct FastMod {
    u64 d;       // divisor
    u64 magic;   // low 64 bits of the 65-bit multiplier
    u32 shift;   // ceil(log2 d) - 1
    u32 pow2;    // 1 when d is a power of two (then n & (d - 1))
};

inline FastMod make_fastmod(u64 d) {
    FastMod f;
    f.d = d; f.magic = 0; f.shift = 0; f.pow2 = 0;
    if (d == 0) return f;
    if ((d & (d - 1)) == 0) { f.pow2 = 1; return f; }
    u32 l = 64 - (u32)__builtin_clzll(d);                  // ceil(log2 d) for non powers of two
    unsigned __int128 num = ((unsigned __int128)(((unsigned __int128)1 << l) - d)) << 64;
    f.magic = (u64)(num / d) + 1;
    f.shift = l - 1;
    return f;
}

RK_HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

RK_HD u64 fastmod(u64 n, const FastMod& f) {
    if (f.pow2) return n & (f.d - 1);
    u64 t = mulhi64(f.magic, n);
    u64 q = (t + ((n - t) >> 1)) >> f.shift;
    return n - q * f.d;
}

// ------------------------------------------------------------------------------------------------
// MurmurHash3_x64_128, first output word only, over a key of k <= 32 bytes held little-endian in
// w[0..3] (bytes beyond k are zero).  Returns the value mkmh compares and stores:
// (uint64)out32[0] << 32 | out32[1], i.e. the first 64-bit word with its halves swapped.
RK_HD u64 rotl64(u64 x, int r) { return (x << r) | (x >> (64 - r)); }

RK_HD u64 fmix64(u64 v) {  // murmur3.cpp:58-67
    v ^= v >> 33; v *= 0xff51afd7ed558ccdULL;
    v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL;
    v ^= v >> 33;
    return v;
}

constexpr u64 MM_C1 = 0x87c37b91114253d5ULL;
constexpr u64 MM_C2 = 0x4cf5ad432745937fULL;

RK_HD u64 mm_mix_k1(u64 k1) { k1 *= MM_C1; k1 = rotl64(k1, 31); k1 *= MM_C2; return k1; }
RK_HD u64 mm_mix_k2(u64 k2) { k2 *= MM_C2; k2 = rotl64(k2, 33); k2 *= MM_C1; return k2; }

// Key-length classes: which of body blocks / tail words exist is fixed per class, so each kernel
// instantiation carries no length-dependent branches.
//   KC 0: 1..8   tail k1            KC 3: 17..24  body + tail k1
//   KC 1: 9..15  tail k1,k2         KC 4: 25..31  body + tail k1,k2
//   KC 2: 16     body               KC 5: 32      body, body
inline int kclass_of(int k) { return k <= 8 ? 0 : k <= 15 ? 1 : k == 16 ? 2 : k <= 24 ? 3 : k <= 31 ? 4 : 5; }
template <int KC> struct KWords { static constexpr int N = (KC == 0) ? 1 : (KC <= 2) ? 2 : (KC == 3) ? 3 : 4; };

template <int KC>
RK_HD u64 murmur_kmer(const u64* w, int k) {
    u64 h1 = MURMUR_SEED, h2 = MURMUR_SEED;
    if (KC >= 2) {                                   // murmur3.cpp:248-259, first block
        h1 ^= mm_mix_k1(w[0]);
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        h2 ^= mm_mix_k2(w[1]);
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    if (KC == 5) {                                   // second block (k = 32)
        h1 ^= mm_mix_k1(w[2]);
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        h2 ^= mm_mix_k2(w[3]);
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    if (KC == 1) h2 ^= mm_mix_k2(w[1]);              // tail, murmur3.cpp:269-289
    if (KC == 4) h2 ^= mm_mix_k2(w[3]);
    if (KC == 0 || KC == 1) h1 ^= mm_mix_k1(w[0]);
    if (KC == 3 || KC == 4) h1 ^= mm_mix_k1(w[2]);
    h1 ^= (u64)k; h2 ^= (u64)k;                      // murmur3.cpp:294-305
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return (h1 << 32) | (h1 >> 32);
}

// ------------------------------------------------------------------------------------------------
// 2-bit base codes.  code = (ascii >> 1) & 3 gives A=0 C=1 T=2 G=3 for either case, the complement is
// code ^ 2, and the ASCII letter is byte `code` of 0x47544341 ("ACTG" little-endian).
constexpr u32 CODE_TO_ASCII = 0x47544341u;

RK_HD u32 base_code(u32 ascii) { return (ascii >> 1) & 3u; }
// Exactly the bytes canonical() accepts (valid_dna[], mkmh.hpp:137-152): A C G T a c g t.
RK_HD bool base_valid(u32 ascii) {
    u32 u = ascii & 0xDFu;
    return u == 'A' || u == 'C' || u == 'G' || u == 'T';
}
RK_HD u32 code_ascii(u32 code) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(CODE_TO_ASCII, 0u, 0x4440u | code);
#else
    return (CODE_TO_ASCII >> (8 * code)) & 0xFFu;
#endif
}

// Rolling window over one strand pair.  f[] holds the forward k-mer as the k ASCII bytes Murmur reads
// (byte 0 = first base); r[] holds the reverse complement (byte 0 = complement of the LAST base).
// `run` counts the valid bases seen since the last invalid one; a window is hashable iff run >= k.
template <int KC>
struct KmerWindow {
    static constexpr int NW = KWords<KC>::N;
    u64 f[NW], r[NW];
    int run;

    RK_HD void reset() {
#pragma unroll
        for (int i = 0; i < NW; ++i) { f[i] = 0; r[i] = 0; }
        run = 0;
    }

    // k is the k-mer size, top_shift = 8 * ((k - 1) & 7), top_mask keeps the low ((k-1)&7)+1 bytes.
    RK_HD void push(u32 code, bool invalid, int k, int top_shift, u64 top_mask) {
        (void)k;
        u64 a = code_ascii(code), c = code_ascii(code ^ 2u);
#pragma unroll
        for (int i = 0; i < NW - 1; ++i) f[i] = (f[i] >> 8) | (f[i + 1] << 56);
        f[NW - 1] = (f[NW - 1] >> 8) | (a << top_shift);
#pragma unroll
        for (int i = NW - 1; i > 0; --i) r[i] = (r[i] << 8) | (r[i - 1] >> 56);
        r[0] = (r[0] << 8) | c;
        r[NW - 1] &= top_mask;
        run = invalid ? 0 : run + 1;
    }

    // canonical hash of the window ending at the base just pushed; 0 when it holds an invalid base
    // (mkmh.hpp:515-528).
    RK_HD u64 hash(int k) const {
        if (run < k) return 0;
        u64 hf = murmur_kmer<KC>(f, k), hr = murmur_kmer<KC>(r, k);
        return hf < hr ? hf : hr;
    }
};

RK_HD int win_top_shift(int k) { return 8 * ((k - 1) & 7); }
RK_HD u64 win_top_mask(int k) { int b = ((k - 1) & 7) + 1; return b == 8 ? ~0ULL : ((1ULL << (8 * b)) - 1); }

}  // namespace rk
