// rk_pack.cpp -- host side of the packed transfer (include/rkmh_gpu.h: rkg_set_host_pack): the 2-bit codes and the
// invalid-base mask that ingest_kernel would make on the device (rk_kernels.cuh: code = (ascii >> 1) & 3, 16 bases per code
// word; one mask bit per base, set when the base is not one of ACGTacgt = valid_dna[], mkmh/mkmh.hpp:137-152) are made on
// the HOST, so that 0.375 bytes per base cross the PCIe link instead of 1.  The link, not the GPU, bounds the end-to-end
// rate of the path; the packing runs on the caller's cores inside the timed call.
//
// AVX-512BW or AVX2 (+ BMI2) when the CPU has them (64 / 32 bases per step: two byte-to-mask moves give the two code bit
// planes, PDEP interleaves them, four compares give the validity mask), a scalar loop otherwise.  Per-function target
// attributes and run-time dispatch: the file needs no special compiler flags.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

inline void pack_scalar(const unsigned char* b, int64_t n32, int64_t tail, uint32_t* codes, uint32_t* nmask) {
    // n32 full groups of 32 bases, then `tail` (< 32) bases padded as invalid 'A' codes
    for (int64_t w = 0; w <= n32; ++w) {
        const int cnt = w < n32 ? 32 : (int)tail;
        if (w == n32 && tail == 0) break;
        uint64_t c = 0; uint32_t m = 0;
        for (int i = 0; i < 32; ++i) {
            const unsigned ch = i < cnt ? b[w * 32 + i] : 0u;
            c |= (uint64_t)((ch >> 1) & 3u) << (2 * i);
            const unsigned u = ch & 0xDFu;
            m |= (uint32_t)!(u == 'A' || u == 'C' || u == 'G' || u == 'T') << i;
        }
        codes[2 * w] = (uint32_t)c; codes[2 * w + 1] = (uint32_t)(c >> 32);
        nmask[w] = m;
    }
}

#if defined(__x86_64__)
__attribute__((target("avx2,bmi2"))) inline void pack_group_avx2(const unsigned char* p, uint64_t& c, uint32_t& m) {
    const __m256i fold = _mm256_set1_epi8((char)0xDF);
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T');
    const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
    const uint32_t p0 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 6));     // bit 1 of every byte
    const uint32_t p1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 5));     // bit 2 of every byte
    c = _pdep_u64(p0, 0x5555555555555555ULL) | _pdep_u64(p1, 0xAAAAAAAAAAAAAAAAULL);
    const __m256i u = _mm256_and_si256(x, fold);
    const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, cA), _mm256_cmpeq_epi8(u, cC)),
                                       _mm256_or_si256(_mm256_cmpeq_epi8(u, cG), _mm256_cmpeq_epi8(u, cT)));
    m = ~(uint32_t)_mm256_movemask_epi8(ok);
}

__attribute__((target("avx2,bmi2")))
void pack_avx2(const unsigned char* b, int64_t n32, uint32_t* codes, uint32_t* nmask) {
    int64_t w = 0;
    // the staging buffer is written once and read by the DMA engine only: non-temporal stores (no read-for-ownership of
    // the destination lines), eight groups = 64 B of codes + 32 B of mask at a time when the destination is aligned
    if ((reinterpret_cast<uintptr_t>(codes) & 31) == 0 && (reinterpret_cast<uintptr_t>(nmask) & 31) == 0) {
        for (; w + 8 <= n32; w += 8) {
            uint64_t c[8]; uint32_t m[8];
            for (int i = 0; i < 8; ++i) pack_group_avx2(b + (w + i) * 32, c[i], m[i]);
            _mm256_stream_si256(reinterpret_cast<__m256i*>(codes + 2 * w), _mm256_set_epi64x((long long)c[3], (long long)c[2], (long long)c[1], (long long)c[0]));
            _mm256_stream_si256(reinterpret_cast<__m256i*>(codes + 2 * w + 8), _mm256_set_epi64x((long long)c[7], (long long)c[6], (long long)c[5], (long long)c[4]));
            _mm256_stream_si256(reinterpret_cast<__m256i*>(nmask + w), _mm256_set_epi32((int)m[7], (int)m[6], (int)m[5], (int)m[4], (int)m[3], (int)m[2], (int)m[1], (int)m[0]));
        }
        _mm_sfence();
    }
    for (; w < n32; ++w) {
        uint64_t c; uint32_t m;
        pack_group_avx2(b + w * 32, c, m);
        codes[2 * w] = (uint32_t)c; codes[2 * w + 1] = (uint32_t)(c >> 32);
        nmask[w] = m;
    }
}
#endif

// 64 bases per step with AVX-512BW: the byte-to-mask moves and compares produce the bit planes and the validity mask
// directly in mask registers
__attribute__((target("avx512f,avx512bw,bmi2")))
void pack_avx512(const unsigned char* b, int64_t n64, uint32_t* codes, uint32_t* nmask) {
    const __m512i fold = _mm512_set1_epi8((char)0xDF);
    const __m512i cA = _mm512_set1_epi8('A'), cC = _mm512_set1_epi8('C'), cG = _mm512_set1_epi8('G'), cT = _mm512_set1_epi8('T');
    const bool aligned = (reinterpret_cast<uintptr_t>(codes) & 31) == 0 && (reinterpret_cast<uintptr_t>(nmask) & 31) == 0;
    int64_t g = 0;
    for (; g + 4 <= n64 && aligned; g += 4) {              // 256 bases: 64 B of codes, 32 B of mask
        uint64_t c[8], m[4];
        for (int i = 0; i < 4; ++i) {
            const __m512i x = _mm512_loadu_si512(b + (g + i) * 64);
            const uint64_t p0 = _mm512_movepi8_mask(_mm512_slli_epi16(x, 6));
            const uint64_t p1 = _mm512_movepi8_mask(_mm512_slli_epi16(x, 5));
            c[2 * i] = _pdep_u64(p0, 0x5555555555555555ULL) | _pdep_u64(p1, 0xAAAAAAAAAAAAAAAAULL);
            c[2 * i + 1] = _pdep_u64(p0 >> 32, 0x5555555555555555ULL) | _pdep_u64(p1 >> 32, 0xAAAAAAAAAAAAAAAAULL);
            const __m512i u = _mm512_and_si512(x, fold);
            m[i] = ~(uint64_t)(_mm512_cmpeq_epi8_mask(u, cA) | _mm512_cmpeq_epi8_mask(u, cC) | _mm512_cmpeq_epi8_mask(u, cG) |
                               _mm512_cmpeq_epi8_mask(u, cT));
        }
        _mm512_stream_si512(reinterpret_cast<__m512i*>(codes + 4 * g), _mm512_set_epi64((long long)c[7], (long long)c[6], (long long)c[5], (long long)c[4],
                                                                                         (long long)c[3], (long long)c[2], (long long)c[1], (long long)c[0]));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(nmask + 2 * g), _mm256_set_epi64x((long long)m[3], (long long)m[2], (long long)m[1], (long long)m[0]));
    }
    _mm_sfence();
    for (; g < n64; ++g) {
        const __m512i x = _mm512_loadu_si512(b + g * 64);
        const uint64_t p0 = _mm512_movepi8_mask(_mm512_slli_epi16(x, 6));
        const uint64_t p1 = _mm512_movepi8_mask(_mm512_slli_epi16(x, 5));
        const uint64_t c0 = _pdep_u64(p0, 0x5555555555555555ULL) | _pdep_u64(p1, 0xAAAAAAAAAAAAAAAAULL);
        const uint64_t c1 = _pdep_u64(p0 >> 32, 0x5555555555555555ULL) | _pdep_u64(p1 >> 32, 0xAAAAAAAAAAAAAAAAULL);
        const __m512i u = _mm512_and_si512(x, fold);
        const uint64_t m = ~(uint64_t)(_mm512_cmpeq_epi8_mask(u, cA) | _mm512_cmpeq_epi8_mask(u, cC) | _mm512_cmpeq_epi8_mask(u, cG) |
                                       _mm512_cmpeq_epi8_mask(u, cT));
        codes[4 * g] = (uint32_t)c0; codes[4 * g + 1] = (uint32_t)(c0 >> 32); codes[4 * g + 2] = (uint32_t)c1; codes[4 * g + 3] = (uint32_t)(c1 >> 32);
        nmask[2 * g] = (uint32_t)m; nmask[2 * g + 1] = (uint32_t)(m >> 32);
    }
}

bool have_avx512() {
    static const bool ok = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("bmi2");
    return ok;
}

bool have_avx2() {
#if defined(__x86_64__)
    static const bool ok = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
    return ok;
#else
    return false;
#endif
}

void pack_range(const unsigned char* b, int64_t n32, int64_t tail, uint32_t* codes, uint32_t* nmask) {
#if defined(__x86_64__)
    if (have_avx512()) {
        const int64_t n64 = n32 / 2;
        pack_avx512(b, n64, codes, nmask);
        if (n32 & 1) pack_avx2(b + n64 * 64, 1, codes + 4 * n64, nmask + 2 * n64);
        if (tail) pack_scalar(b + n32 * 32, 0, tail, codes + 2 * n32, nmask + n32);
        return;
    }
    if (have_avx2()) {
        pack_avx2(b, n32, codes, nmask);
        if (tail) pack_scalar(b + n32 * 32, 0, tail, codes + 2 * n32, nmask + n32);
        return;
    }
#endif
    pack_scalar(b, n32, tail, codes, nmask);
}

// The calling thread's helpers.  A chunk is packed in about a millisecond and the calls come every millisecond or two:
// starting sixteen threads per call would cost a fifth of that, so the helpers persist (one pool per calling thread, i.e.
// per GPU context's pipeline) and spin for a short while after a job before they go to sleep.
class PackPool {
  public:
    ~PackPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; gen_.fetch_add(1, std::memory_order_release); }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // job(i) for i in [0, n): job(0) on the caller, the rest on helpers; returns when all are done
    void run(int n, const std::function<void(int)>& job) {
        while ((int)workers_.size() < n - 1) {
            const int idx = (int)workers_.size() + 1;
            workers_.emplace_back([this, idx] { work(idx); });
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = &job; njobs_ = n;
            pending_.store(n - 1, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        job(0);
        for (int spin = 0; pending_.load(std::memory_order_acquire) != 0; ++spin)
            if (spin > 2000) std::this_thread::yield();
    }

  private:
    void work(int idx) {
        uint64_t seen = 0;
        for (;;) {
            // spin ~100 us for the next job, then block
            const auto t0 = std::chrono::steady_clock::now();
            while (gen_.load(std::memory_order_acquire) == seen) {
                if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(100)) {
                    std::unique_lock<std::mutex> lk(m_);
                    cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                    break;
                }
            }
            const std::function<void(int)>* job;
            int n;
            {
                std::lock_guard<std::mutex> lk(m_);           // job_ / njobs_ were written under the lock before gen_ moved
                seen = gen_.load(std::memory_order_acquire);
                if (stop_) return;
                job = job_; n = njobs_;
            }
            if (idx < n) {
                (*job)(idx);
                pending_.fetch_sub(1, std::memory_order_acq_rel);
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> pending_{0};
    const std::function<void(int)>* job_ = nullptr;
    int njobs_ = 0;
    bool stop_ = false;
};

}  // namespace

// codes: 2 * ceil(n / 32) words, nmask: ceil(n / 32) words (bases beyond n read as invalid).  `threads` host threads.
extern "C" void rk_pack_bases(const char* bases, int64_t n, uint32_t* codes, uint32_t* nmask, int threads) {
    const unsigned char* b = reinterpret_cast<const unsigned char*>(bases);
    const int64_t n32 = n / 32, tail = n - n32 * 32;
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n32 / 4096 + 1));
    if (threads == 1) { pack_range(b, n32, tail, codes, nmask); return; }
    thread_local PackPool pool;
    pool.run(threads, [=](int t) {
        // ranges start at multiples of 8 groups: the streaming stores want 32-byte aligned destinations
        const int64_t lo = (n32 * t / threads) & ~(int64_t)7, hi = t == threads - 1 ? n32 : ((n32 * (t + 1) / threads) & ~(int64_t)7);
        pack_range(b + lo * 32, hi - lo, t == threads - 1 ? tail : 0, codes + 2 * lo, nmask + lo);
    });
}
