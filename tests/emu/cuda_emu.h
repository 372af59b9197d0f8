// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A tiny CPU stand-in for the CUDA execution model, used to run the product's kernel SOURCE
// (rkmh_b200/csrc/*.cuh) under g++ in the build container, which has no GPU.  One OS thread per CUDA
// thread, blocks run one after another; __syncthreads is a barrier, warp collectives rendezvous on
// the participating mask.  It exists to catch logic errors (and, under -fsanitize=thread, shared-
// memory races) before GPU minutes are spent.  It says nothing about performance, and a kernel that
// passes here is still checked on the B200 by `pytest -m gpu`.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <chrono>
#include <sched.h>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

#define RK_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int4 { int x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }

namespace emu {

struct Barrier {  // reusable counting barrier: futex waits (C++20 atomic wait), no mutex for the woken threads to fight over
    std::atomic<unsigned> count{0}, gen{0};
    unsigned need;
    explicit Barrier(unsigned n) : need(n) {}
    void wait() {
        const unsigned g = gen.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == need) {
            count.store(0, std::memory_order_relaxed);
            gen.store(g + 1, std::memory_order_release);
            gen.notify_all();
        } else {
            for (int spin = 0; spin < 8 && gen.load(std::memory_order_acquire) == g; ++spin) sched_yield();
            while (gen.load(std::memory_order_acquire) == g) gen.wait(g, std::memory_order_acquire);
        }
    }
};

struct Warp {
    std::mutex m;
    std::map<unsigned, std::unique_ptr<Barrier>> bars;  // one barrier per participating mask
    uint64_t slot[32];
    unsigned pred[32];
    Barrier& bar(unsigned mask) {
        std::lock_guard<std::mutex> lk(m);
        auto& p = bars[mask];
        if (!p) p.reset(new Barrier(__builtin_popcount(mask)));
        return *p;
    }
};

struct Block {
    std::unique_ptr<Barrier> sync;
    std::vector<std::unique_ptr<Warp>> warps;
    unsigned nthreads;
};

extern thread_local Block* t_block;
extern thread_local unsigned t_lane, t_warp;
extern char* g_dyn_smem;

}  // namespace emu

extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
#define EMU_DEFINE_STATE                                                        \
    thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;                   \
    namespace emu { thread_local Block* t_block; thread_local unsigned t_lane, t_warp; char* g_dyn_smem; }

static inline void __syncthreads() { emu::t_block->sync->wait(); }

static inline unsigned emu_warp_mask_all() {
    unsigned first = emu::t_warp * 32, n = emu::t_block->nthreads - first;
    return n >= 32 ? 0xffffffffu : ((1u << n) - 1);
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
    mask &= emu_warp_mask_all();
    emu::t_block->warps[emu::t_warp]->bar(mask).wait();
}
static inline unsigned __activemask() { return emu_warp_mask_all(); }

template <typename T> static inline T emu_exchange(unsigned mask, T v, int src) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    mask &= emu_warp_mask_all();
    emu::Warp& w = *emu::t_block->warps[emu::t_warp];
    uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(T));
    w.slot[emu::t_lane] = raw;
    w.bar(mask).wait();
    uint64_t got = ((unsigned)src < 32 && (mask >> src & 1)) ? w.slot[src] : raw;
    w.bar(mask).wait();
    T out; std::memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    int base = (int)emu::t_lane & ~(width - 1);
    return emu_exchange(mask, v, base + (src & (width - 1)));
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    int base = (int)emu::t_lane & ~(width - 1), t = (int)emu::t_lane ^ x;
    return emu_exchange(mask, v, (t >= base && t < base + width) ? t : (int)emu::t_lane);
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    int base = (int)emu::t_lane & ~(width - 1), t = (int)emu::t_lane + (int)d;
    return emu_exchange(mask, v, t < base + width ? t : (int)emu::t_lane);
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    int base = (int)emu::t_lane & ~(width - 1), t = (int)emu::t_lane - (int)d;
    return emu_exchange(mask, v, t >= base ? t : (int)emu::t_lane);
}
static inline unsigned __ballot_sync(unsigned mask, int p) {
    mask &= emu_warp_mask_all();
    emu::Warp& w = *emu::t_block->warps[emu::t_warp];
    w.pred[emu::t_lane] = p ? 1u : 0u;
    w.bar(mask).wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if ((mask >> i & 1) && w.pred[i]) r |= 1u << i;
    w.bar(mask).wait();
    return r;
}
static inline unsigned __match_any_sync(unsigned mask, unsigned v) {
    mask &= emu_warp_mask_all();
    emu::Warp& w = *emu::t_block->warps[emu::t_warp];
    w.slot[emu::t_lane] = v;
    w.bar(mask).wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if ((mask >> i & 1) && (unsigned)w.slot[i] == v) r |= 1u << i;
    w.bar(mask).wait();
    return r;
}
static inline int __any_sync(unsigned mask, int p) { return __ballot_sync(mask, p) != 0; }
static inline int __all_sync(unsigned mask, int p) { return __ballot_sync(mask, p) == (mask & emu_warp_mask_all()); }

// ---- atomics ---------------------------------------------------------------------------------------
template <typename T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* p, int v) { return __atomic_fetch_add(p, (unsigned)v, __ATOMIC_RELAXED); }
template <typename T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <typename T> static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <typename T> static inline T atomicCAS(T* p, T cmp, T v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return cmp;
}
template <typename T> static inline T atomicMax(T* p, T v) {
    T cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (cur < v && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return cur;
}
template <typename T> static inline T atomicMin(T* p, T v) {
    T cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (cur > v && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return cur;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- integer intrinsics ----------------------------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline unsigned __brev(unsigned x) {
    x = (x >> 16) | (x << 16); x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
    x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4); x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
    return ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    uint64_t src = ((uint64_t)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned sel = (s >> (4 * i)) & 0xf, byte = (unsigned)(src >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
    return (unsigned)((((uint64_t)hi << 32) | lo) >> (s & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
    return (unsigned)(((((uint64_t)hi << 32) | lo) << (s & 31)) >> 32);
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> static inline T max(T a, T b) { return a > b ? a : b; }

// ---- launch ----------------------------------------------------------------------------------------
namespace emu {
template <typename F> void launch(dim3 grid, dim3 block, size_t dyn_smem, F&& body) {
    struct Trace {                                    // EMU_TRACE=1: wall time of every launch
        std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        dim3 g, b;
        ~Trace() {
            if (getenv("EMU_TRACE"))
                fprintf(stderr, "emu launch grid %u block %u: %.3f s\n", g.x, b.x,
                        std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
    } trace;
    trace.g = grid; trace.b = block;
    std::vector<char> smem(dyn_smem + 256);
    g_dyn_smem = (char*)(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127);
    for (unsigned b = 0; b < grid.x; ++b) {
        Block blk;
        blk.nthreads = block.x;
        blk.sync.reset(new Barrier(block.x));
        for (unsigned w = 0; w < (block.x + 31) / 32; ++w) blk.warps.emplace_back(new Warp);
        std::vector<std::thread> ts;
        for (unsigned t = 0; t < block.x; ++t)
            ts.emplace_back([&, t, b] {
                threadIdx = dim3(t); blockIdx = dim3(b); blockDim = block; gridDim = grid;
                t_block = &blk; t_lane = t & 31; t_warp = t >> 5;
                body();
            });
        for (auto& th : ts) th.join();
    }
}
}  // namespace emu
#define EMU_DYN_SMEM() (emu::g_dyn_smem)
