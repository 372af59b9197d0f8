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
// MurmurHash3_x64_128 (murmur3.cpp:234-312), first output word only, over a key of k <= 32 bytes held
// little-endian in 32-bit registers (bytes beyond k are zero).  Returns the value mkmh compares and stores:
// (uint64)out32[0] << 32 | out32[1], i.e. the first 64-bit word with its halves swapped.
//
// All 64-bit arithmetic is spelled on 32-bit halves: a 64 x 64 -> 64 multiply by a constant is exactly three
// integer multiply-adds (IMAD.WIDE + 2 IMAD) and a rotate is two funnel shifts.  Left to itself the compiler
// turns rotl(x * c, r) into two full multiplies and adds an extra add per product (~40% more instructions).
struct W64 { u32 lo, hi; };

RK_HD W64 w64(u32 lo, u32 hi) { W64 r; r.lo = lo; r.hi = hi; return r; }

template <u64 C>
RK_HD W64 mul_c(W64 a) {
    constexpr u32 clo = (u32)C, chi = (u32)(C >> 32);
    W64 r;
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(r.lo), "=r"(r.hi) : "r"(a.lo), "r"(clo));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r.hi) : "r"(a.lo), "r"(chi));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r.hi) : "r"(a.hi), "r"(clo));
#else
    u64 v = (((u64)a.hi << 32) | a.lo) * C;
    r.lo = (u32)v; r.hi = (u32)(v >> 32);
#endif
    return r;
}
// the same for an operand whose upper half is known to be zero (a 1..4 byte tail word)
template <u64 C>
RK_HD W64 mul_c_lo(u32 alo) {
    constexpr u32 clo = (u32)C, chi = (u32)(C >> 32);
    W64 r;
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(r.lo), "=r"(r.hi) : "r"(alo), "r"(clo));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r.hi) : "r"(alo), "r"(chi));
#else
    u64 v = (u64)alo * C;
    r.lo = (u32)v; r.hi = (u32)(v >> 32);
#endif
    return r;
}

RK_HD u32 fshl(u32 lo, u32 hi, u32 n) {   // upper word of (hi:lo) << n, 0 < n < 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, n);
#else
    return (hi << n) | (lo >> (32 - n));
#endif
}
template <int R>
RK_HD W64 rotl_c(W64 a) {
    if (R == 32) return w64(a.hi, a.lo);
    if (R < 32) return w64(fshl(a.hi, a.lo, R), fshl(a.lo, a.hi, R));
    return w64(fshl(a.lo, a.hi, R - 32), fshl(a.hi, a.lo, R - 32));
}
RK_HD W64 add64(W64 a, W64 b) {
    W64 r;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;" : "=r"(r.lo), "=r"(r.hi) : "r"(a.lo), "r"(a.hi), "r"(b.lo), "r"(b.hi));
#else
    u64 v = (((u64)a.hi << 32) | a.lo) + (((u64)b.hi << 32) | b.lo);
    r.lo = (u32)v; r.hi = (u32)(v >> 32);
#endif
    return r;
}
RK_HD W64 xor64(W64 a, W64 b) { return w64(a.lo ^ b.lo, a.hi ^ b.hi); }
RK_HD W64 xorshr33(W64 a) { return w64(a.lo ^ (a.hi >> 1), a.hi); }   // v ^= v >> 33

RK_HD W64 fmix64(W64 v) {  // murmur3.cpp:58-67
    v = xorshr33(v); v = mul_c<0xff51afd7ed558ccdULL>(v);
    v = xorshr33(v); v = mul_c<0xc4ceb9fe1a85ec53ULL>(v);
    return xorshr33(v);
}

constexpr u64 MM_C1 = 0x87c37b91114253d5ULL;
constexpr u64 MM_C2 = 0x4cf5ad432745937fULL;

RK_HD W64 mm_mix_k1(W64 k1) { return mul_c<MM_C2>(rotl_c<31>(mul_c<MM_C1>(k1))); }
RK_HD W64 mm_mix_k2(W64 k2) { return mul_c<MM_C1>(rotl_c<33>(mul_c<MM_C2>(k2))); }
RK_HD W64 mm_mix_k1_lo(u32 k1) { return mul_c<MM_C2>(rotl_c<31>(mul_c_lo<MM_C1>(k1))); }
RK_HD W64 mm_mix_k2_lo(u32 k2) { return mul_c<MM_C1>(rotl_c<33>(mul_c_lo<MM_C2>(k2))); }
// h = h * 5 + c  (murmur3.cpp:255,259)
template <u32 ADD>
RK_HD W64 mul5_add(W64 h) { return add64(mul_c<5ULL>(h), w64(ADD, 0u)); }

// Key-length variants.  A k-mer of k bytes is held in NW = ceil(k/4) 32-bit registers; which Murmur body
// blocks / tail words exist is fixed per variant, so each kernel instantiation carries no length-dependent
// branches:
//   KV  k       NW  Murmur                      KV  k       NW  Murmur
//   0   1..4    1   tail k1                     5   17..20   5   body + tail k1
//   1   5..8    2   tail k1                     6   21..24   6   body + tail k1
//   2   9..12   3   tail k1,k2(32 bit)          7   25..28   7   body + tail k1,k2(32 bit)
//   3   13..15  4   tail k1,k2                  8   29..31   8   body + tail k1,k2
//   4   16      4   body                        9   32       8   body, body
constexpr int NUM_KV = 10;
inline int kvariant_of(int k) {
    return k <= 4 ? 0 : k <= 8 ? 1 : k <= 12 ? 2 : k <= 15 ? 3 : k == 16 ? 4 : k <= 20 ? 5 : k <= 24 ? 6 : k <= 28 ? 7 : k <= 31 ? 8 : 9;
}
template <int KV> struct KVar {
    static constexpr int NW = (KV == 0) ? 1 : (KV == 1) ? 2 : (KV == 2) ? 3 : (KV <= 4) ? 4 : (KV == 5) ? 5 : (KV == 6) ? 6 : (KV == 7) ? 7 : 8;
    static constexpr int BODIES = (KV <= 3) ? 0 : (KV <= 8) ? 1 : 2;
    static constexpr bool TAIL = (KV != 4 && KV != 9);
};

// w[0..NW): the k key bytes little-endian, bytes beyond k zero.
template <int KV>
RK_HD u64 murmur_kmer(const u32* w, int k) {
    W64 h1 = w64(MURMUR_SEED, 0u), h2 = w64(MURMUR_SEED, 0u);
    if constexpr (KVar<KV>::BODIES >= 1) {           // murmur3.cpp:248-259
        h1 = xor64(h1, mm_mix_k1(w64(w[0], w[1])));
        h1 = rotl_c<27>(h1); h1 = add64(h1, h2); h1 = mul5_add<0x52dce729u>(h1);
        h2 = xor64(h2, mm_mix_k2(w64(w[2], w[3])));
        h2 = rotl_c<31>(h2); h2 = add64(h2, h1); h2 = mul5_add<0x38495ab5u>(h2);
    }
    if constexpr (KVar<KV>::BODIES == 2) {
        h1 = xor64(h1, mm_mix_k1(w64(w[4], w[5])));
        h1 = rotl_c<27>(h1); h1 = add64(h1, h2); h1 = mul5_add<0x52dce729u>(h1);
        h2 = xor64(h2, mm_mix_k2(w64(w[6], w[7])));
        h2 = rotl_c<31>(h2); h2 = add64(h2, h1); h2 = mul5_add<0x38495ab5u>(h2);
    }
    if constexpr (KVar<KV>::TAIL) {                  // murmur3.cpp:269-289
        constexpr int T = KVar<KV>::NW - 4 * KVar<KV>::BODIES;   // 1..4 tail registers
        constexpr int B = 4 * KVar<KV>::BODIES;
        if constexpr (T == 4) h2 = xor64(h2, mm_mix_k2(w64(w[B + 2], w[B + 3])));
        if constexpr (T == 3) h2 = xor64(h2, mm_mix_k2_lo(w[B + 2]));
        if constexpr (T >= 2) h1 = xor64(h1, mm_mix_k1(w64(w[B], w[B + 1])));
        if constexpr (T == 1) h1 = xor64(h1, mm_mix_k1_lo(w[B]));
    }
    h1.lo ^= (u32)k; h2.lo ^= (u32)k;                // murmur3.cpp:294-305 (len < 2^32)
    h1 = add64(h1, h2); h2 = add64(h2, h1);
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 = add64(h1, h2);
    return ((u64)h1.lo << 32) | (u64)h1.hi;          // halves swapped, see above
}

// ------------------------------------------------------------------------------------------------
// 2-bit base codes.  code = (ascii >> 1) & 3 gives A=0 C=1 T=2 G=3 for either case and the complement is
// code ^ 2.  One byte-permute turns a code into both ASCII letters Murmur needs: byte 0 = the base,
// byte 1 = its complement (bytes 2,3 are don't-care).
constexpr u32 CODE_TO_ASCII = 0x47544341u;   // "ACTG"
constexpr u32 CODE_TO_COMP = 0x43414754u;    // "TGAC"

RK_HD u32 base_code(u32 ascii) { return (ascii >> 1) & 3u; }
// Exactly the bytes canonical() accepts (valid_dna[], mkmh.hpp:137-152): A C G T a c g t.
RK_HD bool base_valid(u32 ascii) {
    u32 u = ascii & 0xDFu;
    return u == 'A' || u == 'C' || u == 'G' || u == 'T';
}

RK_HD u32 byte_perm(u32 a, u32 b, u32 sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    u64 src = ((u64)b << 32) | a;
    u32 r = 0;
    for (int i = 0; i < 4; ++i) r |= (u32)((src >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
#endif
}
RK_HD u32 code_letters(u32 code) { return byte_perm(CODE_TO_ASCII, CODE_TO_COMP, code * 0x11u + 0x40u); }

// Byte-permute selectors of the rolling window for a given k (uniform per launch).
struct WinSel { u32 f_last, r_last, r_first; };
RK_HD WinSel make_winsel(int k) {
    const int p = (k - 1) & 3;                       // index of the newest base's byte in the last register
    // one 16-bit selector per value of p, packed in a 64-bit constant (no local-memory table)
    const int sh = 16 * p;
    WinSel s;
    s.f_last = (u32)(0x4321342133413334ULL >> sh) & 0xFFFFu;
    s.r_last = (u32)(0x6543754377437773ULL >> sh) & 0xFFFFu;
    s.r_first = (k <= 4) ? ((u32)(0x6541754177417771ULL >> sh) & 0xFFFFu) : 0x6541u;   // single-register window
    return s;
}

// Rolling window over one strand pair.  f[] holds the forward k-mer as the k ASCII bytes Murmur reads
// (byte 0 = first base); r[] holds the reverse complement (byte 0 = complement of the LAST base).  Bytes
// beyond k stay zero.  `run` counts the valid bases seen since the last invalid one; a window is hashable
// iff run >= k (canonical(), mkmh.hpp:182-188).
template <int KV>
struct KmerWindow {
    static constexpr int NW = KVar<KV>::NW;
    u32 f[NW], r[NW];
    int run;

    RK_HD void reset() {
#pragma unroll
        for (int i = 0; i < NW; ++i) { f[i] = 0; r[i] = 0; }
        run = 0;
    }

    RK_HD void push(u32 code, bool invalid, const WinSel& ws) {
        const u32 ac = code_letters(code);
#pragma unroll
        for (int i = 0; i < NW - 1; ++i) f[i] = byte_perm(f[i], f[i + 1], 0x4321u);
        f[NW - 1] = byte_perm(f[NW - 1], ac, ws.f_last);
        if constexpr (NW > 1) {
            r[NW - 1] = byte_perm(r[NW - 2], r[NW - 1], ws.r_last);
#pragma unroll
            for (int i = NW - 2; i > 0; --i) r[i] = byte_perm(r[i - 1], r[i], 0x6543u);
        }
        r[0] = byte_perm(ac, r[0], ws.r_first);
        run = invalid ? 0 : run + 1;
    }

    // canonical hash of the window ending at the base just pushed; 0 when it holds an invalid base
    // (mkmh.hpp:515-528).
    RK_HD u64 hash(int k) const {
        if (run < k) return 0;
        u64 hf = murmur_kmer<KV>(f, k), hr = murmur_kmer<KV>(r, k);
        return hf < hr ? hf : hr;
    }
};

}  // namespace rk
