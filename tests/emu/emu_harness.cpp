// tests/emu/emu_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the product's kernel source (rkmh_b200/csrc/rk_kernels.cuh) with g++ against the CPU stand-in
// of tests/emu/cuda_emu.h and drives the kernels the way rk_capi.cu does, on plain host memory.  Used by
// tests/test_emu.py to check the kernel LOGIC against the oracle in the GPU-less build container.
#include "cuda_emu.h"
EMU_DEFINE_STATE

#include <algorithm>
#include <vector>

#include "../../rkmh_b200/csrc/rk_kernels.cuh"
#include "../../rkmh_b200/csrc/rk_memo.cuh"
#include "../../rkmh_b200/csrc/rk_warp.cuh"
#include "../../rkmh_b200/csrc/rk_call.cuh"

using namespace rk;

namespace {

struct HostBatch {
    std::vector<unsigned char> bases;
    std::vector<i64> offs;
    std::vector<u32> codes, nmask;
    i64 nseq = 0, total = 0, max_ntotal = 0;
    SeqBatch view() const {
        SeqBatch sb;
        sb.bases = bases.data(); sb.offs = offs.data(); sb.codes = codes.data(); sb.nmask = nmask.data();
        sb.nseq = nseq; sb.total = total; sb.base0 = offs.empty() ? 0 : offs[0];
        return sb;
    }
};

const u64* host_hash_lut() {
    static std::vector<u64> lut;
    if (lut.empty()) { lut.resize(HASH_LUT_ENTRIES); build_hash_lut(lut.data()); }
    return lut.data();
}

KList make_klist(const int* ks, int nk, int* kcsel) {
    KList kl{};
    kl.lut = host_hash_lut();
    kl.nk = nk;
    int sel = -1;
    for (int j = 0; j < nk; ++j) {
        kl.k[j] = ks[j]; kl.kv[j] = kvariant_of(ks[j]);
        kl.ksum += ks[j]; kl.kmax = std::max(kl.kmax, ks[j]);
        sel = (sel == -1 || sel == kl.kv[j]) ? kl.kv[j] : NUM_KV;
    }
    *kcsel = sel;
    return kl;
}

void fill_batch(HostBatch& b, const char* bases, const i64* offs, i64 n, const KList& kl, int block) {
    b.nseq = n; b.total = offs[n] - offs[0];
    i64 nwords32 = (b.total + 31) / 32;
    b.bases.assign((size_t)nwords32 * 32 + 64, 0);
    std::copy(bases + offs[0], bases + offs[n], b.bases.begin());
    b.offs.assign(offs, offs + n + 1);
    b.codes.assign((size_t)2 * nwords32 + 8, 0);
    b.nmask.assign((size_t)nwords32 + 8, 0xFFFFFFFFu);
    i64 max_len = 0;
    for (i64 r = 0; r < n; ++r) max_len = std::max(max_len, offs[r + 1] - offs[r]);
    b.max_ntotal = 0;
    for (int j = 0; j < kl.nk; ++j) b.max_ntotal += std::max<i64>(max_len - kl.k[j], 0);
    if (nwords32 > 0) {
        unsigned char* pb = b.bases.data(); u32* pc = b.codes.data(); u32* pm = b.nmask.data(); i64 tot = b.total;
        emu::launch(dim3(2), dim3(block), 0, [=] { ingest_kernel(pb, tot, pc, pm, nwords32); });
    }
}

int next_pow2(i64 v) { int p = 1; while (p < v) p <<= 1; return p; }

template <typename F> void dispatch(int sel, F&& f) {
    switch (sel) {
        case 0: f(KcTag<0>()); break; case 1: f(KcTag<1>()); break; case 2: f(KcTag<2>()); break;
        case 3: f(KcTag<3>()); break; case 4: f(KcTag<4>()); break; case 5: f(KcTag<5>()); break;
        case 6: f(KcTag<6>()); break; case 7: f(KcTag<7>()); break; case 8: f(KcTag<8>()); break;
        case 9: f(KcTag<9>()); break;
        default: f(KcTag<NUM_KV>()); break;
    }
}

struct HostTable {
    std::vector<u64> entries, dup_bits;
    std::vector<u32> maxmult, dup_base, filter, cursor;
    RefTable tb{};
};

void run_sketch(const HostBatch& b, const KList& kl, int kcsel, int s, CountTable ct, int filter_mode, int occ, int block,
                int grid, int force_global_sort, u64* sketches, int* lens) {
    const int sort_cap = next_pow2(std::max<i64>(std::min<i64>(s, std::max<i64>(b.max_ntotal, 1)), 2));
    const int sort_in_smem = force_global_sort ? 0 : 1;
    const i64 stride = std::max<i64>(b.max_ntotal, 1);
    std::vector<u64> scratch((size_t)grid * (size_t)(stride + sort_cap) + 8);
    SeqBatch sb = b.view();
    u64* ps = scratch.data();
    dispatch(kcsel, [&](auto tag) {
        constexpr int KS = decltype(tag)::value;
        emu::launch(dim3(grid), dim3(block), (size_t)sort_cap * 8, [=] {
            sketch_kernel<KS>(sb, kl, s, ct, filter_mode, occ, ps, stride, sort_cap, sort_in_smem, sketches, lens, nullptr, nullptr);
        });
    });
}

void build_table(HostTable& ht, const u64* sketches, const int* lens, int nrefs, int s, int block) {
    const i64 nent = (i64)nrefs * s;
    const u32 cap = (u32)next_pow2(std::max<i64>(2 * nent, 1024));
    const u32 nb = (u32)((nrefs + 63) / 64);
    RefTable& tb = ht.tb;
    tb.cap_mask = cap - 1; tb.nb = nb; tb.stride = 1 + nb; tb.nrefs = (u32)nrefs;
    int flog2 = 10;
    while (flog2 < 20 && (1LL << flog2) < 16 * nent) ++flog2;
    tb.filter_log2 = (u32)flog2;
    ht.entries.assign((size_t)cap * tb.stride, 0); ht.maxmult.assign(cap, 0); ht.dup_base.assign(cap, 0);
    ht.dup_bits.assign((size_t)std::max<i64>(nent, 1) * nb, 0); ht.filter.assign(((size_t)1 << flog2) / 32, 0);
    ht.cursor.assign(64, 0);
    tb.entries = ht.entries.data(); tb.maxmult = ht.maxmult.data(); tb.dup_base = ht.dup_base.data();
    tb.dup_bits = ht.dup_bits.data(); tb.filter = ht.filter.data();
    RefTable t = tb; u32* f = ht.filter.data(); u32* cur = ht.cursor.data();
    emu::launch(dim3(2), dim3(block), 0, [=] { table_insert_kernel(sketches, lens, nrefs, s, t, f); });
    emu::launch(dim3(2), dim3(block), 0, [=] { table_alloc_dup_kernel(t, cur); });
    emu::launch(dim3(2), dim3(block), 0, [=] { table_fill_dup_kernel(sketches, lens, nrefs, s, t); });
}

}  // namespace

extern "C" {

// raw calc_hashes of every record (concatenated, per-k order), out_offs has nseq + 1 entries
int emu_hash(const char* bases, const i64* offs, i64 nseq, const int* ks, int nk, u64* out, i64* out_offs, int block) {
    int kcsel; KList kl = make_klist(ks, nk, &kcsel);
    HostBatch b; fill_batch(b, bases, offs, nseq, kl, block);
    out_offs[0] = 0;
    for (i64 r = 0; r < nseq; ++r) {
        i64 len = offs[r + 1] - offs[r], n = 0;
        for (int j = 0; j < nk; ++j) n += std::max<i64>(len - ks[j], 0);
        out_offs[r + 1] = out_offs[r] + n;
    }
    SeqBatch sb = b.view();
    for (int kidx = 0; kidx < nk; ++kidx)
        dispatch(kl.kv[kidx], [&](auto tag) {
            constexpr int KS = decltype(tag)::value;
            emu::launch(dim3(2), dim3(block), 0, [=] { hash_kernel<(KS < NUM_KV ? KS : 0)>(sb, kl, kidx, out, out_offs); });
        });
    return 0;
}

int emu_count(const char* bases, const i64* offs, i64 nseq, const int* ks, int nk, int* counts, u64 size, int count_zeros,
              int block) {
    int kcsel; KList kl = make_klist(ks, nk, &kcsel);
    HostBatch b; fill_batch(b, bases, offs, nseq, kl, block);
    CountTable ct{counts, make_fastmod(size)};
    SeqBatch sb = b.view();
    dispatch(kcsel, [&](auto tag) {
        constexpr int KS = decltype(tag)::value;
        emu::launch(dim3(2), dim3(block), 0, [=] { count_kernel<KS>(sb, kl, ct, count_zeros); });
    });
    return 0;
}

int emu_sketch(const char* bases, const i64* offs, i64 nseq, const int* ks, int nk, int s, const int* counts, u64 size,
               int filter_mode, int occ, u64* sketches, int* lens, int block, int force_global_sort) {
    int kcsel; KList kl = make_klist(ks, nk, &kcsel);
    HostBatch b; fill_batch(b, bases, offs, nseq, kl, block);
    CountTable ct{const_cast<int*>(counts), make_fastmod(size ? size : 1)};
    run_sketch(b, kl, kcsel, s, ct, counts ? filter_mode : 0, occ, block, 2, force_global_sort, sketches, lens);
    return 0;
}

// rkg_sketch_shard's kernels on the first n windows of one record.  scan_first != 0: the one-pass route with the guessed
// limit (in `pieces` launches over consecutive window ranges, as the pieces of a shard arrive), *route = 1 when its result
// was accepted; otherwise / else the exact route (histogram of the top 16 hash bits, pick, emit, sort), *route = 2.
int emu_sketch_shard(const char* bases, i64 nbytes, i64 n, int k, int s, u64* out, int* out_len, int block, int scan_first,
                     int pieces, int* route) {
    int ks[1] = {k}; int kcsel; KList kl = make_klist(ks, 1, &kcsel);
    const i64 offs[2] = {0, nbytes};
    HostBatch b; fill_batch(b, bases, offs, 1, kl, block);
    std::vector<u32> hist(SHARD_BINS, 0);
    u64 pick[2] = {0, 0};
    const u32* pc = b.codes.data(); const u32* pm = b.nmask.data(); u32* ph = hist.data(); u64* pp = pick;
    const i64 lastw = (i64)b.codes.size() - 1, lastm = (i64)b.nmask.size() - 1;
    *route = 2;
    if (scan_first) {
        u64 limit = ~0ULL;
        if ((i64)s < n) {
            const long double frac = 1.5L * (long double)s / (long double)n;
            if (frac < 1.0L) limit = (u64)(frac * 18446744073709551615.0L);
        }
        const u64 cap = std::max<u64>(4ULL * (u64)s, 1ULL << 10);
        std::vector<u64> sc(next_pow2((i64)cap), 0);
        unsigned long long cnt[2] = {0, 0};
        u64* psc = sc.data(); unsigned long long* pcnt = cnt;
        pieces = std::max(pieces, 1);
        for (int p = 0; p < pieces; ++p) {
            const i64 lo = n * p / pieces, hi = n * (p + 1) / pieces;
            if (hi <= lo) continue;
            dispatch(kcsel, [&](auto tag) {
                constexpr int KS = decltype(tag)::value;
                emu::launch(dim3(2), dim3(block), 0, [=] { shard_scan_kernel<KS>(pc, pm, lastw, lastm, lo, hi, kl, limit, psc, pcnt, cap); });
            });
        }
        const u64 kept = cnt[0];
        if (kept <= cap && (kept >= (u64)s || kept == cnt[1])) {
            *route = 1;
            *out_len = 0;
            if (kept == 0) return 0;
            const int P = next_pow2((i64)std::max<u64>(kept, 2));
            emu::launch(dim3(1), dim3(block), (size_t)P * 8, [=] { shard_sort_kernel(psc, (int)kept, P, 1); });
            const int m = (int)std::min<u64>(kept, (u64)s);
            std::copy(sc.begin(), sc.begin() + m, out);
            *out_len = m;
            return 0;
        }
    }
    dispatch(kcsel, [&](auto tag) {
        constexpr int KS = decltype(tag)::value;
        emu::launch(dim3(2), dim3(block), 0, [=] { shard_hist_kernel<KS>(pc, pm, lastw, lastm, n, kl, ph); });
    });
    emu::launch(dim3(1), dim3(1024), 0, [=] { shard_pick_kernel(ph, (u64)s, pp); });
    const u64 kept = pick[1];
    *out_len = 0;
    if (kept == 0) return 0;
    const int P = next_pow2((i64)std::max<u64>(kept, 2));
    std::vector<u64> cand(P, 0);
    unsigned long long ncand = 0;
    u64* pcand = cand.data(); unsigned long long* pn = &ncand;
    dispatch(kcsel, [&](auto tag) {
        constexpr int KS = decltype(tag)::value;
        emu::launch(dim3(2), dim3(block), 0, [=] { shard_emit_kernel<KS>(pc, pm, lastw, lastm, n, kl, pp, pcand, pn, kept); });
    });
    if (ncand != kept) return -2;
    emu::launch(dim3(1), dim3(block), 0, [=] { shard_sort_kernel(pcand, (int)kept, P, 0); });
    const int m = (int)std::min<u64>(kept, (u64)s);
    std::copy(cand.begin(), cand.begin() + m, out);
    *out_len = m;
    return 0;
}

// `rkmh call` on the device (rk_call.cuh): the exact hash -> depth table of the reads' windows, the depth of every window of
// `ref`, and for the listed windows of `ref` the depth of every single-base substitution / deletion (4 k values per window)
int emu_call_depth(const char* bases, const i64* offs, i64 nreads, const char* ref, i64 ref_len, int k, int* depth,
                   const i64* windows, i64 nwin, int* rescue, int block) {
    int ks[1] = {k}; int kcsel; KList kl = make_klist(ks, 1, &kcsel);
    HostBatch b; fill_batch(b, bases, offs, nreads, kl, block);
    const u32 cap = (u32)next_pow2(std::max<i64>(2 * b.max_ntotal * std::max<i64>(nreads, 1), 64));
    std::vector<u64> keys(cap, 0); std::vector<u32> counts(cap, 0); unsigned long long zero = 0;
    DepthTable dt{keys.data(), counts.data(), &zero, cap - 1};
    SeqBatch sb = b.view();
    dispatch(kcsel, [&](auto tag) {
        constexpr int KS = decltype(tag)::value;
        emu::launch(dim3(2), dim3(block), 0, [=] { depth_add_kernel<KS>(sb, kl, dt); });
    });
    const i64 roffs[2] = {0, ref_len};
    HostBatch rb; fill_batch(rb, ref, roffs, 1, kl, block);
    SeqBatch rsb = rb.view();
    const i64 n = std::max<i64>(ref_len - k, 0);
    dispatch(kcsel, [&](auto tag) {
        constexpr int KS = decltype(tag)::value;
        emu::launch(dim3(2), dim3(block), 0, [=] { depth_lookup_kernel<KS>(rsb, kl, dt, n, depth); });
    });
    const unsigned char* rseq = reinterpret_cast<const unsigned char*>(ref);
    dispatch(kl.kv[0], [&](auto tag) {
        constexpr int KV = decltype(tag)::value < NUM_KV ? decltype(tag)::value : 0;
        emu::launch(dim3(2), dim3(block), 0, [=] { depth_rescue_kernel<KV>(rseq, ref_len, kl, dt, windows, nwin, rescue); });
    });
    return 0;
}

// unpack_kernel (what the memoised kernels read after a packed transfer) on host-packed codes / mask
int emu_unpack(const u32* codes, const u32* nmask, unsigned char* bases, i64 nwords32, int block) {
    emu::launch(dim3(2), dim3(block), 0, [=] { unpack_kernel(codes, nmask, bases, nwords32); });
    return 0;
}

int emu_merge(const u64* parts, const int* part_lens, int nparts, int s, u64* out, int* out_len, int block, int in_smem) {
    const int P = next_pow2(std::max(s, 2));
    std::vector<u64> scr(P);
    u64* ps = scr.data();
    emu::launch(dim3(1), dim3(block), (size_t)P * 8, [=] { merge_kernel(parts, part_lens, nparts, s, P, out, out_len, ps, in_smem); });
    return 0;
}

// Full classification: refs -> sketches -> table; reads -> group kernel (optional) -> block kernel.
// path: 0 = block kernel for every read, 1 = group kernel + block kernel on its redo list.
// hbuf_mode: 0 = hashes in shared memory when they fit, 1 = force global scratch; hs_smem likewise.
int emu_classify(const char* ref_bases, const i64* ref_offs, int nrefs, const char* bases, const i64* offs, i64 nreads,
                 const int* ks, int nk, int s, const int* counts, u64 size, int min_occ, int path, int glog2, int block,
                 int hbuf_global, int hs_global, Hit* out, u64* ref_sketches_out, int* ref_lens_out, int* redo_count_out) {
    int kcsel; KList kl = make_klist(ks, nk, &kcsel);
    HostBatch rb; fill_batch(rb, ref_bases, ref_offs, nrefs, kl, block);
    std::vector<u64> rsk((size_t)nrefs * s, 0); std::vector<int> rlen(nrefs, 0);
    run_sketch(rb, kl, kcsel, s, CountTable{nullptr, {}}, 0, 0, block, 2, 0, rsk.data(), rlen.data());
    if (ref_sketches_out) std::copy(rsk.begin(), rsk.end(), ref_sketches_out);
    if (ref_lens_out) std::copy(rlen.begin(), rlen.end(), ref_lens_out);
    HostTable ht; build_table(ht, rsk.data(), rlen.data(), nrefs, s, block);

    HostBatch b; fill_batch(b, bases, offs, nreads, kl, block);
    CountTable ct{const_cast<int*>(counts), make_fastmod(size ? size : 1)};
    if (!counts) min_occ = -1;
    SeqBatch sb = b.view();
    RefTable tb = ht.tb;
    std::vector<int> redo((size_t)std::max<i64>(nreads, 1)); int redo_count = 0;
    int* predo = redo.data(); int* pcount = &redo_count;
    const int grid = 2;
    if (path == 1) {
        const int nwb = block / 32, gpw = 32 >> glog2, rwords = (nrefs + 1) / 2;
        const int filter_words = (int)((1u << tb.filter_log2) / 32u);
        size_t smem = (size_t)nwb * sizeof(GroupSmem) + (size_t)nwb * gpw * rwords * 4 + (size_t)nwb * gpw * 4 + (size_t)filter_words * 4;
        dispatch(kcsel, [&](auto tag) {
            constexpr int KS = decltype(tag)::value;
            emu::launch(dim3(grid), dim3(block), smem, [=] {
                classify_group_kernel<KS>(sb, kl, s, ct, min_occ, tb, glog2, rwords, filter_words, predo, pcount, out);
            });
        });
    }
    if (path == 2) {                                    // warp-cooperative Murmur kernel (<= 256 references)
        const int nwb = block / 32;
        const int filter_words = (int)((1u << tb.filter_log2) / 32u);
        size_t smem = (size_t)filter_words * 4 + (size_t)nwb * WARP_QCAP * 8;
        dispatch(kcsel, [&](auto tag) {
            constexpr int KS = decltype(tag)::value;
            emu::launch(dim3(grid), dim3(block), smem, [=] {
                if (tb.nb <= 1) classify_warp_kernel<KS, 2>(sb, kl, s, ct, min_occ, tb, filter_words, predo, pcount, out);
                else classify_warp_kernel<KS, 8>(sb, kl, s, ct, min_occ, tb, filter_words, predo, pcount, out);
            });
        });
    }
    if (redo_count_out) *redo_count_out = redo_count;
    {
        const int hs_cap = next_pow2(2 * (i64)std::min<i64>(s, std::max<i64>(b.max_ntotal, 1)));
        const int rpad = (nrefs + 63) / 64 * 64;
        const int hs_in_smem = hs_global ? 0 : 1;
        const i64 hbuf_cap = hbuf_global ? 0 : b.max_ntotal;
        const i64 stride = hbuf_global ? std::max<i64>(b.max_ntotal, 1) : 0;
        size_t smem = (size_t)hbuf_cap * 8 + (hs_in_smem ? (size_t)hs_cap * 8 : 0) + (size_t)rpad * 4;
        std::vector<u64> scratch((size_t)grid * (size_t)(stride + hs_cap) + 8);
        u64* ps = scratch.data();
        const int* list = (path == 1 || path == 2) ? predo : nullptr;
        dispatch(kcsel, [&](auto tag) {
            constexpr int KS = decltype(tag)::value;
            emu::launch(dim3(grid), dim3(block), smem, [=] {
                classify_block_kernel<KS>(sb, kl, s, ct, min_occ, tb, list, pcount, nreads, (int)hbuf_cap, hs_cap, hs_in_smem,
                                          rpad, ps, stride, out);
            });
        });
    }
    return 0;
}

// `rkmh hpv16` on the device (rkg_load_ref_sets + rkg_classify_sets): a SET table from full hash sets (table_insert_set_kernel)
// and the exact block kernel in set mode with an unbounded sketch size: hits over sets [0, ntypes), per-read counts of the rest
extern "C" int emu_classify_sets(const u64* hashes, const i64* set_offs, int nsets, const char* bases, const i64* offs, i64 nreads,
                                 int k, const int* counts, u64 size, int min_occ, int ntypes, int block, int hbuf_global, Hit* out,
                                 int* extra) {
    int ks[1] = {k}; int kcsel; KList kl = make_klist(ks, 1, &kcsel);
    const i64 nent = set_offs[nsets] - set_offs[0];
    const u32 cap = (u32)next_pow2(std::max<i64>(2 * nent, 1024));
    RefTable tb{};
    tb.cap_mask = cap - 1; tb.nb = (u32)((nsets + 63) / 64); tb.stride = 1 + tb.nb; tb.nrefs = (u32)nsets;
    int flog2 = 10;
    while (flog2 < 20 && (1LL << flog2) < 16 * nent) ++flog2;
    tb.filter_log2 = (u32)flog2;
    std::vector<u64> entries((size_t)cap * tb.stride, 0), dupbits(8, 0);
    std::vector<u32> maxmult(cap, 0), dupbase(cap, 0), filter(((size_t)1 << flog2) / 32, 0);
    tb.entries = entries.data(); tb.maxmult = maxmult.data(); tb.dup_base = dupbase.data(); tb.dup_bits = dupbits.data(); tb.filter = filter.data();
    std::vector<i64> rel((size_t)nsets + 1);
    for (int i = 0; i <= nsets; ++i) rel[i] = set_offs[i] - set_offs[0];
    const u64* ph = hashes + set_offs[0]; const i64* prel = rel.data(); u32* pf = filter.data();
    emu::launch(dim3(2), dim3(block), 0, [=] { table_insert_set_kernel(ph, prel, nsets, tb, pf); });
    HostBatch b; fill_batch(b, bases, offs, nreads, kl, block);
    CountTable ct{const_cast<int*>(counts), make_fastmod(size ? size : 1)};
    if (!counts) { ct.counts = nullptr; min_occ = -1; }
    SeqBatch sb = b.view();
    const int s_all = 1 << 30, grid = 2;
    const int hs_cap = next_pow2(2 * std::max<i64>(b.max_ntotal, 1));
    const int rpad = (nsets + 63) / 64 * 64;
    const i64 hbuf_cap = hbuf_global ? 0 : b.max_ntotal;
    const i64 stride = hbuf_global ? std::max<i64>(b.max_ntotal, 1) : 0;
    const size_t smem = (size_t)hbuf_cap * 8 + (size_t)hs_cap * 8 + (size_t)rpad * 4;
    std::vector<u64> scratch((size_t)grid * (size_t)(stride + hs_cap) + 8);
    u64* ps = scratch.data();
    dispatch(kcsel, [&](auto tag) {
        constexpr int KS = decltype(tag)::value;
        emu::launch(dim3(grid), dim3(block), smem, [=] {
            classify_block_kernel<KS>(sb, kl, s_all, ct, min_occ, tb, nullptr, nullptr, nreads, (int)hbuf_cap, hs_cap, 1, rpad, ps, stride,
                                      out, 1, ntypes, extra);
        });
    });
    return 0;
}

// ---- memoised path (rk_memo.cuh) -------------------------------------------------------------------------
struct HostMemo {
    std::vector<u32> tab, ref, tmp, sieve;
    std::vector<unsigned long long> cmap;
    u32 stats[2] = {0, 0};
    std::vector<u64> hash;
    std::vector<int> code_counts;
    u32 inval = 0;
    int err = 0;
    Memo mm{};
};

static void build_memo(HostMemo& hm, int k, const RefTable& tb, int block) {
    const u32 ncodes = 1u << (2 * k);
    const int bucket_bits = std::max(2 * k - 8, 0);
    const u32 nb = 1u << bucket_bits;
    hm.tab.assign(ncodes, 0); hm.ref.assign(ncodes, 0); hm.hash.assign(ncodes, 0); hm.code_counts.assign(ncodes, 0);
    hm.tmp.assign((size_t)3 * nb + ncodes, 0);
    u32* bcount = hm.tmp.data(); u32* bstart = bcount + nb; u32* bfill = bstart + nb; u32* bcodes = bfill + nb;
    u64* hash = hm.hash.data(); u32* tab = hm.tab.data(); u32* ref = hm.ref.data(); int* err = &hm.err;
    dispatch(kvariant_of(k), [&](auto tag) {
        constexpr int KV = decltype(tag)::value <= 2 ? decltype(tag)::value : 2;
        emu::launch(dim3(2), dim3(block), 0, [=] { memo_hash_kernel<KV>(k, ncodes, hash, bcount, bucket_bits); });
    });
    emu::launch(dim3(1), dim3(1024), 0, [=] { memo_scan_kernel(bcount, bstart, bfill, nb, err); });
    emu::launch(dim3(2), dim3(block), 0, [=] { memo_scatter_kernel(ncodes, hash, bstart, bfill, bcodes, bucket_bits); });
    emu::launch(dim3(2), dim3(block), 0, [=] { memo_rank_kernel(hash, bcount, bstart, bcodes, nb, tb, tab, ref); });
    hm.mm.tab = tab; hm.mm.ref = ref; hm.mm.hash = hash; hm.mm.code_counts = hm.code_counts.data();
    hm.mm.inval = &hm.inval; hm.mm.ncodes = ncodes; hm.mm.kmask = (1u << k) - 1u; hm.mm.k = k;
    Memo mm = hm.mm;
    // sieve + code map (classify_memo_sieve_kernel); a small sieve so that false positives do occur in the tests
    if (tb.entries != nullptr) {
        const u32 nwords = 257;
        hm.sieve.assign(nwords, 0);
        u32* sieve = hm.sieve.data(); u32* stats = hm.stats;
        emu::launch(dim3(2), dim3(block), 0, [=] { memo_sieve_kernel(mm, sieve, nwords, stats); });
        const u32 cap = (u32)next_pow2(std::max<i64>(2 * (i64)hm.stats[0], 16));
        int log2cap = 0; while ((1u << log2cap) < cap) ++log2cap;
        hm.cmap.assign(cap, ~0ULL);
        unsigned long long* cmap = hm.cmap.data();
        emu::launch(dim3(2), dim3(block), 0, [=] { memo_cmap_kernel(mm, cmap, cap - 1, 32 - log2cap); });
        hm.mm.sieve = sieve; hm.mm.sieve_words = nwords; hm.mm.cmap = cmap; hm.mm.cmap_mask = cap - 1; hm.mm.cmap_shift = 32 - log2cap;
        hm.mm.zero_codes = hm.stats[1];
    }
}

extern "C" {

// -M pass 1 through the memoised count kernel + fold; counts must be zero-initialised by the caller
int emu_count_memo(const char* bases, const i64* offs, i64 nseq, int k, int* counts, u64 size, int count_zeros, int wpr,
                   int block) {
    int ks[1] = {k}; int kcsel; KList kl = make_klist(ks, 1, &kcsel);
    HostBatch b; fill_batch(b, bases, offs, nseq, kl, block);
    HostMemo hm; RefTable none{}; build_memo(hm, k, none, block);
    if (hm.err) return -1;
    CountTable ct{counts, make_fastmod(size)};
    SeqBatch sb = b.view(); Memo mm = hm.mm;
    emu::launch(dim3(2), dim3(block), 0, [=] { count_memo_kernel(sb, mm, count_zeros, wpr); });
    emu::launch(dim3(2), dim3(block), 0, [=] { memo_fold_kernel(mm, ct, count_zeros); });
    return 0;
}

// path: 0 = block kernel for every read, 1 = warp kernel + block kernel on its redo list.  vcap_limit > 0 caps the
// per-warp shared-memory lists of the block kernel (forces its second attempt with global lists).
int emu_classify_memo(const char* ref_bases, const i64* ref_offs, int nrefs, const char* bases, const i64* offs, i64 nreads,
                      int k, int s, const int* counts, u64 size, int min_occ, int path, int block, int vcap_limit,
                      Hit* out, int* redo_count_out) {
    int ks[1] = {k}; int kcsel; KList kl = make_klist(ks, 1, &kcsel);
    HostBatch rb; fill_batch(rb, ref_bases, ref_offs, nrefs, kl, block);
    std::vector<u64> rsk((size_t)nrefs * s, 0); std::vector<int> rlen(nrefs, 0);
    run_sketch(rb, kl, kcsel, s, CountTable{nullptr, {}}, 0, 0, block, 2, 0, rsk.data(), rlen.data());
    HostTable ht; build_table(ht, rsk.data(), rlen.data(), nrefs, s, block);
    HostMemo hm; build_memo(hm, k, ht.tb, block);
    if (hm.err) return -1;
    HostBatch b; fill_batch(b, bases, offs, nreads, kl, block);
    CountTable ct{const_cast<int*>(counts), make_fastmod(size ? size : 1)};
    if (!counts) min_occ = -1;
    Memo mm = hm.mm;
    emu::launch(dim3(2), dim3(block), 0, [=] { memo_apply_occ_kernel(mm, ct, min_occ); });
    SeqBatch sb = b.view();
    RefTable tb = ht.tb;
    std::vector<int> redo((size_t)std::max<i64>(nreads, 1)); int redo_count = 0;
    int* predo = redo.data(); int* pcount = &redo_count;
    const int grid = 2;
    if (path == 4 && tb.nb <= 4 && mm.sieve != nullptr && mm.zero_codes == 0 && min_occ < 0) {
        // (path 4: the sieve kernel; falls through to the warp kernel when its preconditions do not hold)
        const size_t smem = ((size_t)mm.sieve_words + (size_t)(block / 32) * SIEVE_QCAP) * 4;
        switch (tb.nb) {
            case 1: emu::launch(dim3(grid), dim3(block), smem, [=] { classify_memo_sieve_kernel<2>(sb, s, mm, tb, predo, pcount, out); }); break;
            case 2: emu::launch(dim3(grid), dim3(block), smem, [=] { classify_memo_sieve_kernel<4>(sb, s, mm, tb, predo, pcount, out); }); break;
            case 3: emu::launch(dim3(grid), dim3(block), smem, [=] { classify_memo_sieve_kernel<6>(sb, s, mm, tb, predo, pcount, out); }); break;
            default: emu::launch(dim3(grid), dim3(block), smem, [=] { classify_memo_sieve_kernel<8>(sb, s, mm, tb, predo, pcount, out); }); break;
        }
    } else if ((path == 1 || path == 4) && tb.nb <= 4) {
        // (path 1: the register-row warp kernel for <= 256 references)
        switch (tb.nb) {
            case 1: emu::launch(dim3(grid), dim3(block), 0, [=] { classify_memo_warp_kernel<2>(sb, s, mm, tb, predo, pcount, out); }); break;
            case 2: emu::launch(dim3(grid), dim3(block), 0, [=] { classify_memo_warp_kernel<4>(sb, s, mm, tb, predo, pcount, out); }); break;
            case 3: emu::launch(dim3(grid), dim3(block), 0, [=] { classify_memo_warp_kernel<6>(sb, s, mm, tb, predo, pcount, out); }); break;
            default: emu::launch(dim3(grid), dim3(block), 0, [=] { classify_memo_warp_kernel<8>(sb, s, mm, tb, predo, pcount, out); }); break;
        }
    } else if (path >= 1) {       // path 2 forces the wide (shared-memory counter) kernel
        const int rwords = (nrefs + 1) / 2;
        size_t smem = (size_t)(block / 32) * (sizeof(MemoWarpSmem) + (size_t)rwords * 4);
        emu::launch(dim3(grid), dim3(block), smem, [=] { classify_memo_wide_kernel(sb, s, mm, tb, rwords, predo, pcount, out); });
    }
    if (redo_count_out) *redo_count_out = redo_count;
    {
        const int hs_cap = next_pow2(2 * (i64)std::min<i64>(s, std::max<i64>(b.max_ntotal, 1)));
        const int rpad = (nrefs + 63) / 64 * 64;
        const int nwb = block / 32;
        const i64 most = ((std::max<i64>(b.max_ntotal, 1) + 31) / 32 + nwb - 1) / nwb * 32;
        int lcap = (int)std::min<i64>(most, std::max<i64>(128, ((i64)(0.4 * s) + 31) / 32 * 32));
        if (vcap_limit > 0) lcap = std::min(lcap, vcap_limit);
        const i64 stride = (std::max<i64>(b.max_ntotal, 1) + 64 * nwb + 31) / 32 * 32;
        size_t smem = 4 * ((size_t)nwb * (lcap + 32) + MEMO_HIST_WORDS + 2 * (size_t)hs_cap + rpad);
        std::vector<u32> scratch((size_t)grid * (size_t)stride + 8);
        u32* ps = scratch.data();
        const int* list = (path >= 1) ? predo : nullptr;
        emu::launch(dim3(grid), dim3(block), smem, [=] {
            classify_memo_block_kernel(sb, s, mm, tb, list, pcount, nreads, lcap, hs_cap, rpad, ps, stride, out);
        });
    }
    return 0;
}

}  // extern "C"

// host-side arithmetic checks
u64 emu_fastmod(u64 n, u64 d) { FastMod f = make_fastmod(d); return fastmod(n, f); }

u64 emu_window_hash(const char* kmer, int k) {   // canonical hash of one window through window_hash (rk_hash.cuh)
    u32 lo = 0, hi = 0;
    for (int i = 0; i < 32; ++i) {                 // junk codes beyond k must not matter
        const u32 c = i < k ? base_code((u32)(unsigned char)kmer[i]) : (u32)((i * 7 + 1) & 3);
        if (i < 16) lo |= c << (2 * i); else hi |= c << (2 * (i - 16));
    }
    u64 result = 0;
    dispatch(kvariant_of(k), [&](auto tag) {
        constexpr int KC = decltype(tag)::value < NUM_KV ? decltype(tag)::value : 0;
        const HashCorr hc = make_hash_corr<KC>(k);
        result = window_hash<KC>(lo, hi, k, host_hash_lut(), hc);
    });
    return result;
}

u64 emu_kmer_hash(const char* kmer, int k) {   // canonical hash of one window through the rolling window code
    u64 result = 0;
    dispatch(kvariant_of(k), [&](auto tag) {
        constexpr int KC = decltype(tag)::value < NUM_KV ? decltype(tag)::value : 0;
        KmerWindow<KC> w; w.reset();
        const WinSel ws = make_winsel(k);
        for (int i = 0; i < k + 3; ++i) w.push(base_code((u32)'G'), false, ws);   // junk first: bytes beyond k must stay zero
        for (int i = 0; i < k; ++i) w.push(base_code((u32)(unsigned char)kmer[i]), !base_valid((u32)(unsigned char)kmer[i]), ws);
        result = w.hash(k);
    });
    return result;
}

}  // extern "C"
