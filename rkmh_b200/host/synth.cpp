// synth.cpp -- seeded synthetic workloads for the BASELINE.json configurations, as a small
// C library for bench.py and the tools (ctypes: rkmh_b200/synth.py).  Measurement infrastructure, not product path.
//
// Generator = std::mt19937_64 as SURVEY 8(d) specifies.  Every stream is cut into fixed BLOCKS of units (reads or
// bases); block b is produced by its own engine seeded with seed_seq{seed, stream tag, b}.  A workload is therefore a
// pure function of (seed, unit index): rank r of N generates units [r n, (r + 1) n) of the same global stream that a
// single process would generate whole, which is what the N > 1 versus N = 1 parity checks of bench.py rely on, and the
// blocks are filled by several host threads.
//
//   syn_sample_reads   config 2 / 4 (viral part): fixed-length reads sampled from a uniformly chosen reference, strand
//                      and offset, with substitutions and N bases; a fraction uniform random instead
//   syn_ont_reads      config 3: circular walk on one of the genomes with ONT-like substitutions / insertions / deletions
//   syn_background     config 4: order-0 random reads with a given GC fraction
//   syn_genome         config 5: uniform ACGT with runs of N
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

namespace {

typedef std::mt19937_64 Rng;

const char ACGT[4] = {'A', 'C', 'G', 'T'};

Rng make_rng(uint64_t seed, uint64_t tag, uint64_t block) {
    std::seed_seq sq{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)tag, (uint32_t)block, (uint32_t)(block >> 32)};
    return Rng(sq);
}

inline double unit(Rng& g) { return (double)(g() >> 11) * (1.0 / 9007199254740992.0); }   // [0, 1)
inline uint64_t below(Rng& g, uint64_t n) { return (uint64_t)(((unsigned __int128)g() * n) >> 64); }

// distance to the next event of a Bernoulli(p) process (>= 0 failures before it)
inline int64_t gap(Rng& g, double inv_log1mp) {
    const double u = 1.0 - unit(g);                       // (0, 1]
    const double v = std::log(u) * inv_log1mp;
    return v > 4e18 ? (int64_t)4e18 : (int64_t)v;
}

inline char complement(char c) {
    switch (c) {
        case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
        default: return c;
    }
}
inline char upper(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }

template <typename F>
void parallel_blocks(int64_t first, int64_t n, int64_t block, int threads, F&& f) {
    // blocks are aligned to the GLOBAL unit index, so a slice of the stream equals the same units of the whole
    const int64_t b0 = first / block, b1 = (first + n + block - 1) / block;
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, b1 - b0));
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
        th.emplace_back([&, t] { for (int64_t b = b0 + t; b < b1; b += threads) f(b); });
    for (auto& x : th) x.join();
}

constexpr int64_t READ_BLOCK = 4096;
constexpr int64_t BASE_BLOCK = 1 << 22;

}  // namespace

extern "C" {

// refs: nrefs sequences concatenated in `cat` (any case; upper-cased on the fly), ref_off[nrefs + 1].  References not
// longer than `length` are never sampled.  Writes reads [first, first + n) of the stream into out[n * length].
int syn_sample_reads(const char* cat, const int64_t* ref_off, int nrefs, int64_t first, int64_t n, int length,
                     double sub, double nrate, double random_frac, uint64_t seed, char* out, int threads) {
    std::vector<int> usable;
    for (int r = 0; r < nrefs; ++r) if (ref_off[r + 1] - ref_off[r] > length) usable.push_back(r);
    if (usable.empty() && random_frac < 1.0) return -1;
    const double ils = sub > 0 ? 1.0 / std::log1p(-sub) : 0, iln = nrate > 0 ? 1.0 / std::log1p(-nrate) : 0;
    parallel_blocks(first, n, READ_BLOCK, threads, [&](int64_t b) {
        Rng g = make_rng(seed, 2, (uint64_t)b);
        for (int64_t i = b * READ_BLOCK; i < (b + 1) * READ_BLOCK; ++i) {
            // every read of the block draws from the engine even when it is outside [first, first + n): the stream
            // must not depend on the slice
            char buf[4096];
            char* dst = (i >= first && i < first + n) ? out + (i - first) * length : nullptr;
            char* w = length <= 4096 ? buf : nullptr;
            std::vector<char> big;
            if (!w) { big.resize((size_t)length); w = big.data(); }
            if (unit(g) < random_frac) {
                for (int j = 0; j < length; j += 32) {
                    uint64_t x = g();
                    for (int q = 0; q < 32 && j + q < length; ++q, x >>= 2) w[j + q] = ACGT[x & 3];
                }
            } else {
                const int r = usable[below(g, usable.size())];
                const int64_t size = ref_off[r + 1] - ref_off[r];
                const int64_t off = (int64_t)below(g, (uint64_t)(size - length));
                const char* src = cat + ref_off[r] + off;
                if (g() & 1) for (int j = 0; j < length; ++j) w[j] = complement(upper(src[length - 1 - j]));
                else for (int j = 0; j < length; ++j) w[j] = upper(src[j]);
            }
            if (sub > 0) for (int64_t p = gap(g, ils); p < length; p += 1 + gap(g, ils)) w[p] = ACGT[g() & 3];
            if (nrate > 0) for (int64_t p = gap(g, iln); p < length; p += 1 + gap(g, iln)) w[p] = 'N';
            if (dst) memcpy(dst, w, (size_t)length);
        }
    });
    return 0;
}

// Circular walk on a uniformly chosen genome (random start and strand): per output base, with probability `ins` a
// random base is inserted (the walk does not advance), with probability `dele` one genome base is skipped first, and
// with probability `sub` the emitted base is replaced by a random one.  Fixed output length.
int syn_ont_reads(const char* cat, const int64_t* gen_off, int ngen, int64_t first, int64_t n, int length,
                  double sub, double ins, double dele, uint64_t seed, char* out, int threads) {
    if (ngen < 1) return -1;
    const double pe = sub + ins + dele;
    const double ile = pe > 0 ? 1.0 / std::log1p(-pe) : 0;
    parallel_blocks(first, n, 256, threads, [&](int64_t b) {
        Rng g = make_rng(seed, 3, (uint64_t)b);
        std::vector<char> tmp((size_t)length);
        for (int64_t i = b * 256; i < (b + 1) * 256; ++i) {
            const bool keep = i >= first && i < first + n;
            const int gi = (int)below(g, (uint64_t)ngen);
            const char* G = cat + gen_off[gi];
            const int64_t gs = gen_off[gi + 1] - gen_off[gi];
            int64_t pos = (int64_t)below(g, (uint64_t)gs);
            const bool rev = g() & 1;
            int64_t next_err = pe > 0 ? gap(g, ile) : (int64_t)4e18;
            char* w = tmp.data();
            for (int j = 0; j < length; ++j) {
                char c;
                if (j == next_err) {
                    const double u = unit(g) * pe;
                    next_err = j + 1 + gap(g, ile);
                    if (u < ins) { w[j] = ACGT[g() & 3]; continue; }                       // inserted base
                    if (u < ins + dele) { pos = pos + 1 == gs ? 0 : pos + 1; c = upper(G[pos]); }   // one base skipped
                    else { c = ACGT[g() & 3]; }                                               // substitution
                } else c = upper(G[pos]);
                w[j] = c;
                pos = pos + 1 == gs ? 0 : pos + 1;
            }
            if (!keep) continue;
            char* dst = out + (i - first) * length;
            if (rev) for (int j = 0; j < length; ++j) dst[j] = complement(w[length - 1 - j]);
            else memcpy(dst, w, (size_t)length);
        }
    });
    return 0;
}

// bases [first, first + n) of an order-0 random stream with GC fraction `gc`
int syn_background(int64_t first, int64_t n, double gc, uint64_t seed, char* out, int threads) {
    const uint32_t tA = (uint32_t)((1.0 - gc) / 2 * 65536.0), tC = tA + (uint32_t)(gc / 2 * 65536.0), tG = tC + (uint32_t)(gc / 2 * 65536.0);
    parallel_blocks(first, n, BASE_BLOCK, threads, [&](int64_t b) {
        Rng g = make_rng(seed, 4, (uint64_t)b);
        const int64_t lo = b * BASE_BLOCK;
        for (int64_t j = 0; j < BASE_BLOCK; j += 4) {
            uint64_t x = g();
            for (int q = 0; q < 4; ++q, x >>= 16) {
                const int64_t at = lo + j + q;
                if (at < first || at >= first + n) continue;
                const uint32_t v = (uint32_t)(x & 0xFFFF);
                out[at - first] = v < tA ? 'A' : v < tC ? 'C' : v < tG ? 'G' : 'T';
            }
        }
    });
    return 0;
}

// bases [first, first + n) of one uniform ACGT sequence in which every block of 2^22 bases holds round(2^22 x
// n_frac / n_run_len) runs of n_run_len N bases at uniform positions inside the block
int syn_genome(int64_t first, int64_t n, double n_frac, int n_run_len, uint64_t seed, char* out, int threads) {
    const int nruns = n_run_len > 0 ? (int)std::llround((double)BASE_BLOCK * n_frac / n_run_len) : 0;
    parallel_blocks(first, n, BASE_BLOCK, threads, [&](int64_t b) {
        Rng g = make_rng(seed, 5, (uint64_t)b);
        const int64_t lo = b * BASE_BLOCK;
        const int64_t w0 = std::max(lo, first), w1 = std::min(lo + BASE_BLOCK, first + n);
        for (int64_t j = 0; j < BASE_BLOCK; j += 32) {
            uint64_t x = g();
            const int64_t at = lo + j;
            if (at + 32 <= w0 || at >= w1) continue;
            for (int q = 0; q < 32; ++q, x >>= 2) if (at + q >= w0 && at + q < w1) out[at + q - first] = ACGT[x & 3];
        }
        for (int r = 0; r < nruns; ++r) {
            const int64_t st = lo + (int64_t)below(g, (uint64_t)(BASE_BLOCK - n_run_len));
            for (int64_t p = std::max(st, w0); p < std::min(st + n_run_len, w1); ++p) out[p - first] = 'N';
        }
    });
    return 0;
}

}  // extern "C"
