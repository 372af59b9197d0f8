// rk_capi.cu -- implementation of include/rkmh_gpu.h: context, device buffers, kernel launches.
// Host side only orchestrates; all arithmetic of the hot path runs in the kernels of rk_kernels.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <chrono>
#include <vector>

#include "../../include/rkmh_gpu.h"
#include "rk_kernels.cuh"
#include "rk_memo.cuh"
#include "rk_warp.cuh"
#include "rk_call.cuh"

using namespace rk;

extern "C" void rk_pack_bases(const char* bases, int64_t n, uint32_t* codes, uint32_t* nmask, int threads);   // rk_pack.cpp

static_assert(sizeof(rkg_hit) == sizeof(Hit), "rkg_hit layout");
static_assert(RKG_MAX_K == MAX_K && RKG_MAX_NK == MAX_NK, "limits");

namespace {

thread_local std::string g_create_error;

enum KernelId { K_INGEST, K_HASH, K_COUNT, K_COUNT_DISTINCT, K_SKETCH, K_MERGE, K_TABLE, K_CLASSIFY_GROUP, K_CLASSIFY_WARP,
                K_CLASSIFY_BLOCK, K_MEMO_BUILD, K_MEMO_MAINT, K_COUNT_MEMO, K_CLASSIFY_MEMO_WARP, K_CLASSIFY_MEMO_BLOCK,
                K_CLASSIFY_MEMO_SIEVE,
                K_NUM };
const char* const kKernelNames[K_NUM] = {"ingest_kernel", "hash_kernel", "count_kernel", "count_distinct_kernel",
                                         "sketch_kernel", "merge_kernel", "table_build_kernels",
                                         "classify_group_kernel", "classify_warp_kernel", "classify_block_kernel", "memo_build_kernels",
                                         "memo_fold_apply_kernels", "count_memo_kernel", "classify_memo_warp_kernel",
                                         "classify_memo_block_kernel", "classify_memo_sieve_kernel"};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct EventPair { cudaEvent_t a, b; int id; };

}  // namespace

struct rkg_batch {
    DevBuf bases, offs, codes, nmask, hits, redo, scratch, misc;   // misc: [0] redo_count
    i64 nseq = 0, total = 0, base0 = 0, max_ntotal = 0, max_len = 0;
    bool packed = false;                 // codes / nmask hold this batch (ingest_kernel has run, or the batch arrived packed)
    bool ascii = true;                   // bases holds this batch's letters (false: it arrived packed and was not unpacked yet)
    cudaStream_t stream = nullptr;       // the stream every operation of this batch is issued on
    cudaStream_t own_stream = nullptr;   // created with the batch; `stream` may be redirected to a caller's stream
    cudaEvent_t done = nullptr;
};

constexpr int NSLOTS = 4;

struct rkg_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    size_t l2_persist_max = 0, l2_window_max = 0;   // persisting-L2 carve-out and largest access-policy window of the device
    bool l2_policy = true;                          // RKMH_NO_L2_POLICY=1 turns the access-policy windows off (A/B timing)
    std::string err;
    KList kl{};
    int kcsel = 0;
    int kmax = 0, kmin = 0;
    int s = 1000;
    bool params_set = false;
    u64 read_tbl_size = 0, ref_tbl_size = 0;
    DevBuf depth_keys, depth_counts, depth_zero;   // `call`: exact hash -> depth table (rk_call.cuh)
    DepthTable depth{};
    i64 depth_room = 0;                            // windows that may still be added before the table could fill
    DevBuf read_counts, ref_counts, hash_lut;   // hash_lut: the tables of rk_hash.cuh (16 KB, built in rkg_create)
    CountTable read_ct{nullptr, {}}, ref_ct{nullptr, {}};
    // references
    bool refs_loaded = false;
    bool set_table = false;              // the table was built by rkg_load_ref_sets (set semantics, hpv16)
    int nrefs = 0;
    DevBuf ref_sketches, ref_lens, tb_entries, tb_maxmult, tb_dupbase, tb_dupbits, tb_filter, tb_cursor;
    RefTable tb{};
    cudaStream_t stream = nullptr;
    rkg_batch* slots[NSLOTS] = {};       // internal buffers of rkg_classify_batch / count (two; four with adaptive packing)
    // packed transfer (rkg_set_host_pack): per slot, pinned staging for the codes and the mask packed on the host
    int host_pack = 0;                   // 0 off, 1 every chunk, 2 adaptive: a chunk is packed when the link is busy anyway
    int pack_threads = 1;
    void* stage[NSLOTS] = {};
    size_t stage_cap[NSLOTS] = {};
    DevBuf shard_scan, shard_cnt;        // rkg_sketch_shard: kept hashes and counters of the one-pass route
    cudaStream_t copy_stream = nullptr;  // rkg_sketch_shard: the shard's pieces arrive here while earlier pieces are hashed
    cudaEvent_t copy_event = nullptr;
    bool shard_two_pass = false;         // RKMH_SHARD_TWO_PASS=1: always take the exact two-pass route (tests, comparison)
    cudaEvent_t h2d_done[NSLOTS] = {};   // adaptive mode: recorded after a chunk's host-to-device copies
    int64_t xfer[4] = {0, 0, 0, 0};       // rkg_transfer_stats
    double pack_max_ratio = 1.25;        // pack only while packing a chunk takes at most this many times its plain copy (RKMH_PACK_MAX_RATIO)
    double pack_lead = 0.75;             // pack when the queued copies last at least this share of the packing time (RKMH_PACK_LEAD)
    double pack_ms_per_byte = 0.0, dma_ms_per_byte = 0.0;      // adaptive mode: measured on this context's first chunks
    // memoised path (single k <= MEMO_MAX_K, rk_memo.cuh)
    bool memo_enabled = true, memo_ready = false, sieve_enabled = true, group_kernel = false;
    DevBuf memo_tab, memo_ref, memo_hash, memo_counts, memo_tmp, memo_misc, memo_sieve, memo_cmap;
    double sieve_fill = 0.0;             // fraction of the sieve's bits that are set (false-positive rate = fill^2)
    Memo memo{};
    bool memo_pending = false;           // code_counts holds increments not yet folded into the read count table
    int memo_pending_zeros = 0;
    bool memo_occ_stale = true;          // the read count table changed since FAILM was baked
    int memo_occ = INT_MIN;              // min_occ currently baked into the FAILM flags (INT_MIN: none)
    // measurement
    int64_t launches = 0;
    bool profiling = false;
    std::vector<EventPair> events;
    double prof_ms[K_NUM] = {0};
    int64_t prof_n[K_NUM] = {0};
};

namespace {

int fail(rkg_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? RKG_ERR_NOMEM : RKG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                 \
    } while (0)

struct Timed {   // brackets one launch with events when profiling is on
    rkg_ctx* c; cudaStream_t st; int id; EventPair ep{};
    Timed(rkg_ctx* c_, cudaStream_t st_, int id_) : c(c_), st(st_), id(id_) {
        ++c->launches;
        if (c->profiling) {
            cudaEventCreate(&ep.a); cudaEventCreate(&ep.b); ep.id = id;
            cudaEventRecord(ep.a, st);
        }
    }
    ~Timed() {
        if (c->profiling) { cudaEventRecord(ep.b, st); c->events.push_back(ep); }
    }
};

void profile_collect(rkg_ctx* c) {
    for (auto& ep : c->events) {
        cudaEventSynchronize(ep.b);
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ep.a, ep.b) == cudaSuccess) { c->prof_ms[ep.id] += ms; c->prof_n[ep.id] += 1; }
        cudaEventDestroy(ep.a); cudaEventDestroy(ep.b);
    }
    c->events.clear();
}

// smallest power of two >= v; 0 when it would not fit the 2^30 the callers can index with (they fail with RKG_ERR_ARG)
int next_pow2(i64 v) {
    if (v > (i64)1 << 30) return 0;
    int p = 1; while (p < v) p <<= 1; return p;
}

// Launch helper over the key-length-variant template parameter (NUM_KV = mixed variants).
#define DISPATCH_KCSEL(sel, ...)                  \
    switch (sel) {                                \
        case 0: { constexpr int KS = 0; __VA_ARGS__; } break; \
        case 1: { constexpr int KS = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int KS = 2; __VA_ARGS__; } break; \
        case 3: { constexpr int KS = 3; __VA_ARGS__; } break; \
        case 4: { constexpr int KS = 4; __VA_ARGS__; } break; \
        case 5: { constexpr int KS = 5; __VA_ARGS__; } break; \
        case 6: { constexpr int KS = 6; __VA_ARGS__; } break; \
        case 7: { constexpr int KS = 7; __VA_ARGS__; } break; \
        case 8: { constexpr int KS = 8; __VA_ARGS__; } break; \
        case 9: { constexpr int KS = 9; __VA_ARGS__; } break; \
        default: { constexpr int KS = NUM_KV; __VA_ARGS__; } break; \
    }

int batch_alloc(rkg_ctx* ctx, rkg_batch** out) {
    rkg_batch* b = new rkg_batch();
    CK(cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking));
    b->stream = b->own_stream;
    CK(cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming));
    *out = b;
    return RKG_OK;
}

void batch_release(rkg_batch* b) {
    if (!b) return;
    if (b->stream) cudaStreamSynchronize(b->stream);
    b->bases.release(); b->offs.release(); b->codes.release(); b->nmask.release(); b->hits.release();
    b->redo.release(); b->scratch.release(); b->misc.release();
    if (b->done) cudaEventDestroy(b->done);
    if (b->own_stream) cudaStreamDestroy(b->own_stream);
    delete b;
}

// 2-bit pack + invalid-base mask of the batch's resident ASCII bases.
int launch_ingest(rkg_ctx* ctx, rkg_batch* b) {
    const i64 nwords32 = (b->total + 31) / 32;
    if (nwords32 > 0) {
        Timed t(ctx, b->stream, K_INGEST);
        int grid = (int)std::min<i64>((nwords32 + 255) / 256, (i64)ctx->sm_count * 16);
        ingest_kernel<<<grid, 256, 0, b->stream>>>(b->bases.as<unsigned char>(), b->total, b->codes.as<u32>(),
                                                   b->nmask.as<u32>(), nwords32);
    }
    CK(cudaGetLastError());
    b->packed = true;
    return RKG_OK;
}

// The Murmur-path kernels read the packed codes; the memoised path reads the ASCII bases and never needs them.
int ensure_packed(rkg_ctx* ctx, rkg_batch* b) { return b->packed ? RKG_OK : launch_ingest(ctx, b); }

// The memoised kernels read letters: a batch that arrived packed is unpacked on the device first (0.3 ms per 10 M reads).
int ensure_ascii(rkg_ctx* ctx, rkg_batch* b) {
    if (b->ascii) return RKG_OK;
    const i64 nwords32 = (b->total + 31) / 32;
    if (nwords32 > 0) {
        Timed t(ctx, b->stream, K_INGEST);
        const int grid = (int)std::min<i64>((nwords32 + 255) / 256, (i64)ctx->sm_count * 16);
        unpack_kernel<<<grid, 256, 0, b->stream>>>(b->codes.as<u32>(), b->nmask.as<u32>(), b->bases.as<unsigned char>(), nwords32);
    }
    CK(cudaGetLastError());
    b->ascii = true;
    return RKG_OK;
}

// offs[i] = base0 + i * len for i in [0, n]: the offsets of a batch of fixed-length reads, made on the device so that
// they never cross the PCIe link (8 bytes per read: 5 % of the host-to-device bytes of 150-base reads)
__global__ void __launch_bounds__(256)
fixed_offs_kernel(i64* __restrict__ offs, i64 n, i64 base0, i64 len) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (i64)gridDim.x * blockDim.x) offs[i] = base0 + i * len;
}

// Copy reads [r0, r1) of the caller's buffers into the batch (asynchronous on b->stream); packing is done on demand.
// offs == nullptr: every read has `fixed_len` bases and read i starts at bases[i * fixed_len].
// Adaptive packed transfer: mixing packed and plain chunks pays when the host packs a chunk in about the time the link
// takes to send it as it is.  When packing is much slower (few threads per rank, or the ranks of a full box sharing the
// host's memory bandwidth with each other's copies) only a few chunks could be packed, and the packer's memory traffic
// slows the copies down by more than those chunks save: measured at N = 4 and N = 8.  Then every
// chunk goes plain.
bool pack_pays(const rkg_ctx* ctx) { return ctx->pack_ms_per_byte <= ctx->pack_max_ratio * ctx->dma_ms_per_byte; }

// pinned staging of a slot for the packed transfer
int ensure_stage(rkg_ctx* ctx, int slot, size_t bytes) {
    if (ctx->stage_cap[slot] >= bytes) return RKG_OK;
    if (ctx->stage[slot]) cudaFreeHost(ctx->stage[slot]);
    ctx->stage[slot] = nullptr; ctx->stage_cap[slot] = 0;
    const size_t want = bytes + bytes / 8 + 4096;
    CK(cudaHostAlloc(&ctx->stage[slot], want, cudaHostAllocPortable));
    ctx->stage_cap[slot] = want;
    return RKG_OK;
}

int batch_fill(rkg_ctx* ctx, rkg_batch* b, const char* bases, const int64_t* offs, i64 r0, i64 r1, i64 fixed_len = 0,
               int pack_mode = -1) {          // pack_mode: -1 = as the context says (host_pack == 1), 0 = plain, 1 = packed
    const i64 n = r1 - r0;
    const i64 base0 = offs ? offs[r0] : r0 * fixed_len, total = offs ? offs[r1] - base0 : n * fixed_len;
    b->nseq = n; b->total = total; b->base0 = base0;
    const i64 nwords32 = (total + 31) / 32;
    CK(b->bases.ensure((size_t)nwords32 * 32 + 64));
    CK(b->offs.ensure(sizeof(i64) * (size_t)(n + 1)));
    CK(b->codes.ensure(sizeof(u32) * (size_t)(2 * nwords32 + 8)));
    CK(b->nmask.ensure(sizeof(u32) * (size_t)(nwords32 + 8)));
    CK(b->hits.ensure(sizeof(Hit) * (size_t)std::max<i64>(n, 1)));
    CK(b->redo.ensure(sizeof(int) * (size_t)std::max<i64>(n, 1)));
    CK(b->misc.ensure(256));
    int slot = -1;
    for (int i = 0; i < NSLOTS; ++i) if (b == ctx->slots[i]) slot = i;
    const bool pack_here = (pack_mode < 0 ? ctx->host_pack == 1 : pack_mode == 1) && slot >= 0 && total > 0;
    if (pack_here) {
        // 0.375 bytes per base cross the link: codes (2 words per 32 bases) then the mask (1 word per 32 bases), packed by
        // the host's threads into this slot's pinned staging buffer (free again: the caller waited for the slot's stream)
        const size_t mask_at = ((size_t)2 * nwords32 + 7) & ~(size_t)7;      // the mask starts 32-byte aligned (streaming stores)
        const size_t words = mask_at + (size_t)nwords32;
        { int rc = ensure_stage(ctx, slot, words * 4); if (rc) return rc; }
        u32* hc = static_cast<u32*>(ctx->stage[slot]);
        u32* hm = hc + mask_at;
        rk_pack_bases(bases + base0, total, hc, hm, ctx->pack_threads);
        ctx->xfer[1] += 12 * nwords32; ++ctx->xfer[3];
        CK(cudaMemcpyAsync(b->codes.p, hc, sizeof(u32) * 2 * (size_t)nwords32, cudaMemcpyHostToDevice, b->stream));
        CK(cudaMemcpyAsync(b->nmask.p, hm, sizeof(u32) * (size_t)nwords32, cudaMemcpyHostToDevice, b->stream));
    } else if (total > 0) {
        CK(cudaMemcpyAsync(b->bases.p, bases + base0, (size_t)total, cudaMemcpyHostToDevice, b->stream));
        ctx->xfer[0] += total; ++ctx->xfer[2];
    }
    if (offs) CK(cudaMemcpyAsync(b->offs.p, offs + r0, sizeof(i64) * (size_t)(n + 1), cudaMemcpyHostToDevice, b->stream));
    else {
        ++ctx->launches;
        fixed_offs_kernel<<<(int)std::min<i64>((n + 256) / 256, (i64)ctx->sm_count * 8), 256, 0, b->stream>>>(b->offs.as<i64>(), n, base0, fixed_len);
        CK(cudaGetLastError());
    }
    // the reader may touch one word past the end: keep the padding defined
    CK(cudaMemsetAsync(b->codes.as<u32>() + 2 * nwords32, 0, sizeof(u32) * 8, b->stream));
    CK(cudaMemsetAsync(b->nmask.as<u32>() + nwords32, 0xFF, sizeof(u32) * 8, b->stream));
    // longest read of the batch (sizes the exact path's scratch); overlaps the copies above
    i64 max_len = offs ? 0 : fixed_len;
    if (offs) for (i64 r = r0; r < r1; ++r) max_len = std::max<i64>(max_len, offs[r + 1] - offs[r]);
    b->max_len = max_len;
    i64 mt = 0;
    for (int j = 0; j < ctx->kl.nk; ++j) mt += std::max<i64>(max_len - ctx->kl.k[j], 0);
    b->max_ntotal = mt;
    b->packed = pack_here;
    b->ascii = !pack_here;
    return RKG_OK;
}

SeqBatch seq_view(const rkg_batch* b) {
    SeqBatch sb;
    sb.bases = b->bases.as<unsigned char>();
    sb.offs = b->offs.as<i64>();
    sb.codes = b->codes.as<u32>();
    sb.nmask = b->nmask.as<u32>();
    sb.nseq = b->nseq;
    sb.total = b->total;
    sb.base0 = b->base0;
    return sb;
}

// ---- memoised path (rk_memo.cuh) ---------------------------------------------------------------------------
// Build hash / rank / REF tables over all 4^k codes.  Called at the end of rkg_load_refs (the reference table exists).
int memo_build(rkg_ctx* ctx, cudaStream_t st) {
    ctx->memo_ready = false;
    if (!ctx->memo_enabled || ctx->kl.nk != 1 || ctx->kl.k[0] > MEMO_MAX_K) return RKG_OK;
    const int k = ctx->kl.k[0];
    const u32 ncodes = 1u << (2 * k);
    const int bucket_bits = std::max(2 * k - 8, 0);
    const u32 nb = 1u << bucket_bits;
    CK(ctx->memo_tab.ensure(sizeof(u32) * (size_t)ncodes));
    CK(ctx->memo_ref.ensure(sizeof(u32) * (size_t)ncodes));
    CK(ctx->memo_hash.ensure(sizeof(u64) * (size_t)ncodes));
    CK(ctx->memo_tmp.ensure(sizeof(u32) * ((size_t)3 * nb + ncodes)));
    CK(ctx->memo_misc.ensure(256));
    u32* bcount = ctx->memo_tmp.as<u32>();
    u32* bstart = bcount + nb; u32* bfill = bstart + nb; u32* bcodes = bfill + nb;
    CK(cudaMemsetAsync(bcount, 0, sizeof(u32) * (size_t)nb, st));
    CK(cudaMemsetAsync(ctx->memo_ref.p, 0, sizeof(u32) * (size_t)ncodes, st));
    CK(cudaMemsetAsync(ctx->memo_misc.p, 0, 256, st));
    Memo mm{};
    mm.tab = ctx->memo_tab.as<u32>(); mm.ref = ctx->memo_ref.as<u32>(); mm.hash = ctx->memo_hash.as<u64>();
    mm.inval = nullptr;
    int* err = ctx->memo_misc.as<int>() + 8;
    mm.ncodes = ncodes; mm.kmask = (1u << k) - 1u; mm.k = k;
    mm.code_counts = nullptr;
    if (ctx->read_ct.counts) {
        CK(ctx->memo_counts.ensure(sizeof(int) * ((size_t)ncodes + MEMO_PARTIAL_PAD)));
        CK(cudaMemsetAsync(ctx->memo_counts.p, 0, sizeof(int) * ((size_t)ncodes + MEMO_PARTIAL_PAD), st));
        mm.code_counts = ctx->memo_counts.as<int>();
        mm.inval = ctx->memo_counts.as<u32>() + ncodes;
    }
    {
        Timed t(ctx, st, K_MEMO_BUILD);
        const int grid = (int)std::min<i64>(((i64)ncodes + 255) / 256, (i64)ctx->sm_count * 16);
        u64* hash = ctx->memo_hash.as<u64>();
        switch (kvariant_of(k)) {
            case 0: memo_hash_kernel<0><<<grid, 256, 0, st>>>(k, ncodes, hash, bcount, bucket_bits); break;
            case 1: memo_hash_kernel<1><<<grid, 256, 0, st>>>(k, ncodes, hash, bcount, bucket_bits); break;
            default: memo_hash_kernel<2><<<grid, 256, 0, st>>>(k, ncodes, hash, bcount, bucket_bits); break;
        }
        memo_scan_kernel<<<1, 1024, 0, st>>>(bcount, bstart, bfill, nb, err);
        memo_scatter_kernel<<<grid, 256, 0, st>>>(ncodes, hash, bstart, bfill, bcodes, bucket_bits);
        memo_rank_kernel<<<(int)std::min<i64>(nb, (i64)ctx->sm_count * 8), 256, 0, st>>>(hash, bcount, bstart, bcodes, nb, ctx->tb,
                                                                                        mm.tab, mm.ref);
        ctx->launches += 3;
    }
    CK(cudaGetLastError());
    int herr = 0;
    CK(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (herr) return RKG_OK;          // a hash bucket larger than the sort capacity: keep the Murmur path (still exact)
    // sieve + code map of classify_memo_sieve_kernel (rk_memo.cuh).  The sieve takes what a 1024-thread block can
    // have beside its candidate queues; it is dropped when it would let more than a quarter of all windows through.
    mm.sieve = nullptr; mm.cmap = nullptr; mm.sieve_words = 0; mm.zero_codes = 0;
    ctx->sieve_fill = 0.0;
    const size_t queues = (size_t)32 * SIEVE_QCAP * 4 + 1024;
    if (ctx->sieve_enabled && ctx->tb.entries != nullptr && ctx->smem_optin > queues + 8192) {
        const u32 nwords = (u32)std::min<size_t>(SIEVE_MAX_WORDS, (ctx->smem_optin - queues) / 4);
        CK(ctx->memo_sieve.ensure(sizeof(u32) * (size_t)nwords));
        CK(cudaMemsetAsync(ctx->memo_sieve.p, 0, sizeof(u32) * (size_t)nwords, st));
        u32* stats = ctx->memo_misc.as<u32>() + 16;
        const int grid = (int)std::min<i64>(((i64)ncodes + 255) / 256, (i64)ctx->sm_count * 16);
        {
            Timed t(ctx, st, K_MEMO_BUILD);
            memo_sieve_kernel<<<grid, 256, 0, st>>>(mm, ctx->memo_sieve.as<u32>(), nwords, stats);
            ctx->launches += 1;
        }
        CK(cudaGetLastError());
        u32 hstats[2] = {0, 0};
        std::vector<u32> image(nwords);
        CK(cudaMemcpyAsync(hstats, stats, sizeof(hstats), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(image.data(), ctx->memo_sieve.p, sizeof(u32) * (size_t)nwords, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        size_t set_bits = 0;
        for (u32 w : image) set_bits += (size_t)__builtin_popcount(w);
        ctx->sieve_fill = (double)set_bits / (32.0 * nwords);
        mm.zero_codes = hstats[1];
        if (ctx->sieve_fill <= 0.5) {
            // load <= 1/16: one probe, nearly always (config 2, sieve kernel per 10 M reads: load 1/2 9.4 ms, 1/4 8.1, 1/8 7.8, 1/16 7.7, 1/32 7.8)
            const u32 cap = (u32)next_pow2(std::max<i64>(16 * (i64)hstats[0], 1024));
            int log2cap = 0; while ((1u << log2cap) < cap) ++log2cap;
            CK(ctx->memo_cmap.ensure(sizeof(u64) * (size_t)cap));
            CK(cudaMemsetAsync(ctx->memo_cmap.p, 0xFF, sizeof(u64) * (size_t)cap, st));
            {
                Timed t(ctx, st, K_MEMO_BUILD);
                memo_cmap_kernel<<<grid, 256, 0, st>>>(mm, ctx->memo_cmap.as<unsigned long long>(), cap - 1, 32 - log2cap);
                ctx->launches += 1;
            }
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(st));
            mm.sieve = ctx->memo_sieve.as<u32>(); mm.sieve_words = nwords;
            mm.cmap = ctx->memo_cmap.as<unsigned long long>(); mm.cmap_mask = cap - 1; mm.cmap_shift = 32 - log2cap;
        }
    }
    ctx->memo = mm;
    ctx->memo_ready = true;
    ctx->memo_pending = false; ctx->memo_occ_stale = true; ctx->memo_occ = INT_MIN;
    return RKG_OK;
}

// Fold the per-code counters of pass 1 into the read count table (the table the reference keeps, rkmh.cpp:738-743).
int memo_flush(rkg_ctx* ctx) {
    if (!ctx->memo_ready || !ctx->memo_pending) return RKG_OK;
    CK(cudaDeviceSynchronize());
    {
        Timed t(ctx, ctx->stream, K_MEMO_MAINT);
        const int grid = (int)std::min<i64>(((i64)ctx->memo.ncodes + 255) / 256, (i64)ctx->sm_count * 16);
        memo_fold_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->memo, ctx->read_ct, ctx->memo_pending_zeros);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->memo_pending = false;
    ctx->memo_occ_stale = true;
    return RKG_OK;
}

// Make the FAILM flags of the table reflect (read count table, min_occ) before a classification pass.
int memo_prepare(rkg_ctx* ctx, int min_occ) {
    int rc = memo_flush(ctx);
    if (rc) return rc;
    const int want = min_occ >= 0 ? min_occ : INT_MIN;
    if (want == ctx->memo_occ && !(want != INT_MIN && ctx->memo_occ_stale)) return RKG_OK;
    CK(cudaDeviceSynchronize());
    {
        Timed t(ctx, ctx->stream, K_MEMO_MAINT);
        const int grid = (int)std::min<i64>(((i64)ctx->memo.ncodes + 255) / 256, (i64)ctx->sm_count * 16);
        memo_apply_occ_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->memo, ctx->read_ct, min_occ);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->memo_occ = want;
    ctx->memo_occ_stale = false;
    return RKG_OK;
}

// Keep `bytes` at `ptr` resident in L2 for the kernels launched on `st` from now on: hits on the window persist, every
// other access of those kernels (the read bases, streamed once) is marked streaming and is evicted first.  Without it
// the 1 byte per base streaming through L2 evicts the randomly accessed table (ncu, round 1: the count kernel wrote
// 319 MB back to DRAM per 200 MB of bases, L2 hit rate 65 %).  bytes = 0 clears the window.
void l2_window(rkg_ctx* ctx, cudaStream_t st, const void* ptr, size_t bytes) {
    if (!ctx->l2_policy || ctx->l2_persist_max == 0 || ctx->l2_window_max == 0) return;
    cudaStreamAttrValue v{};
    if (bytes > 0) {
        const size_t win = std::min(bytes, ctx->l2_window_max);
        v.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
        v.accessPolicyWindow.num_bytes = win;
        // the fraction of the window that fits the carve-out persists, the rest is treated as normal
        v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)ctx->l2_persist_max / (double)win);
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        v.accessPolicyWindow.num_bytes = 0;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
}

int launch_count(rkg_ctx* ctx, rkg_batch* b, CountTable ct, int count_zeros) {
    if (b->nseq == 0) return RKG_OK;
    if (ctx->memo_ready && ctx->memo.code_counts && ct.counts == ctx->read_ct.counts) {
        if (ctx->memo_pending && ctx->memo_pending_zeros != count_zeros) { int rc = memo_flush(ctx); if (rc) return rc; }
        ctx->memo_pending = true; ctx->memo_pending_zeros = count_zeros;
        { int rc = ensure_ascii(ctx, b); if (rc) return rc; }
        const int wpr = (b->total / b->nseq > 2048) ? 8 : 1;            // a block per long read, a warp per short one
        const i64 teams = (b->nseq + 0) * 1;
        const int grid = (int)std::min<i64>((teams * wpr + 7) / 8, (i64)ctx->sm_count * 8);
        l2_window(ctx, b->stream, ctx->memo.code_counts, sizeof(int) * (size_t)ctx->memo.ncodes);
        {
            Timed t(ctx, b->stream, K_COUNT_MEMO);
            count_memo_kernel<<<grid, 256, 0, b->stream>>>(seq_view(b), ctx->memo, count_zeros, wpr);
        }
        l2_window(ctx, b->stream, nullptr, 0);
        CK(cudaGetLastError());
        return RKG_OK;
    }
    { int rc = ensure_packed(ctx, b); if (rc) return rc; }
    Timed t(ctx, b->stream, K_COUNT);
    int grid = (int)std::min<i64>((b->nseq + 7) / 8, (i64)ctx->sm_count * 8);
    DISPATCH_KCSEL(ctx->kcsel, (count_kernel<KS><<<grid, 256, 0, b->stream>>>(seq_view(b), ctx->kl, ct, count_zeros)));
    CK(cudaGetLastError());
    return RKG_OK;
}

// Geometry of classify_group_kernel for this context, or ok = false when it cannot be used.
struct GroupGeom { bool ok; int block, glog2, rwords, filter_words; size_t smem; };

GroupGeom group_geometry(const rkg_ctx* ctx, const rkg_batch* b) {
    GroupGeom g{};
    if (b->nseq == 0) return g;
    const i64 mean_len = b->total / b->nseq;
    i64 n = 0;
    for (int j = 0; j < ctx->kl.nk; ++j) n += std::max<i64>(mean_len - ctx->kl.k[j], 0);
    if (n > ctx->s || n > GROUP_MAX_KMERS || ctx->nrefs > 65000) return g;   // typical read is not eligible: exact path for all
    int glog2 = 5;                                        // ~8-12 k-mers per lane
    while (glog2 > 3 && (n >> (glog2 - 1)) < 10) --glog2;
    g.glog2 = glog2;
    g.rwords = (ctx->nrefs + 1) / 2;
    g.filter_words = (int)((1u << ctx->tb.filter_log2) / 32u);
    for (int block : {1024, 512, 256}) {
        int nwb = block / 32, gpw = 32 >> glog2;
        size_t smem = (size_t)nwb * sizeof(GroupSmem) + (size_t)nwb * gpw * g.rwords * 4 + (size_t)nwb * gpw * 4 +
                      (size_t)g.filter_words * 4;
        if (smem <= ctx->smem_optin) { g.ok = true; g.block = block; g.smem = smem; return g; }
    }
    return g;
}

// Geometry of the two memoised classification kernels for this batch, or ok = false (Murmur path instead).
struct MemoGeom { bool ok, use_warp; int warp_block, rwords; size_t warp_smem; int lcap, hs_cap, rpad, block_per_sm; size_t block_smem; i64 scratch_stride; };

MemoGeom memo_geometry(const rkg_ctx* ctx, const rkg_batch* b) {
    MemoGeom g{};
    // block kernel: [8 x lcap kept values][MEMO_HIST_WORDS histogram][2 hs_cap hit set][rpad counters] in u32.  A warp keeps
    // about 2 s / 8 values of a long read (classify_memo_block_kernel); 1.6 x that leaves room for uneven shares.  As many
    // blocks per SM as that allows, at most four (registers).
    g.hs_cap = next_pow2(2 * (i64)std::min<i64>(ctx->s, std::max<i64>(b->max_ntotal, 1)));
    g.rpad = (ctx->nrefs + 63) / 64 * 64;
    const size_t fixed = 4 * ((size_t)MEMO_HIST_WORDS + 2 * (size_t)g.hs_cap + (size_t)g.rpad);
    const i64 most = ((std::max<i64>(b->max_ntotal, 1) + 31) / 32 + 7) / 8 * 32;      // a warp's share of the longest read
    const i64 want = std::min<i64>(most, std::max<i64>(128, ((i64)(0.4 * (double)ctx->s) + 31) / 32 * 32));
    for (int per_sm = 4; per_sm >= 1 && !g.ok; --per_sm) {
        const size_t budget = ctx->smem_optin / per_sm - 4608;     // static scratch + hit rings + the 1 KB reserve per block
        if (fixed + 4 * 8 * (size_t)64 > budget) continue;
        const i64 room = (i64)((budget - fixed) / (4 * 8)) / 32 * 32 - 32;     // 32 words per warp catch an overflowing list
        if (room < want && per_sm > 1) continue;
        g.lcap = (int)std::min<i64>(want, room);
        g.block_per_sm = per_sm;
        g.ok = true;
    }
    if (!g.ok) return g;
    g.scratch_stride = (std::max<i64>(b->max_ntotal, 1) + 64 * 8 + 31) / 32 * 32;      // the lists of the second attempt: every window
    g.block_smem = fixed + 4 * 8 * (size_t)(g.lcap + 32);
    // warp kernel: the typical read keeps all its k-mers (n <= s)
    const i64 mean_n = std::max<i64>(b->total / b->nseq - ctx->kl.k[0], 0);
    g.rwords = (ctx->nrefs + 1) / 2;
    if (mean_n <= ctx->s && ctx->nrefs <= 65000) {
        for (int block : {1024, 512, 256}) {
            const size_t smem = (size_t)(block / 32) * (sizeof(MemoWarpSmem) + (size_t)g.rwords * 4);
            if (smem <= ctx->smem_optin) { g.use_warp = true; g.warp_block = block; g.warp_smem = smem; break; }
        }
    }
    return g;
}

int launch_classify(rkg_ctx* ctx, rkg_batch* b, int min_occ) {
    if (b->nseq == 0) return RKG_OK;
    CountTable ct = (min_occ >= 0) ? ctx->read_ct : CountTable{nullptr, {}};
    if (min_occ >= 0 && ct.counts == nullptr)
        return fail(ctx, RKG_ERR_STATE, "min_occ >= 0 needs a read count table (rkg_set_params read_table_size)");
    SeqBatch sb = seq_view(b);
    int* redo = b->redo.as<int>();
    int* redo_count = b->misc.as<int>();
    Hit* hits = b->hits.as<Hit>();
    CK(cudaMemsetAsync(redo_count, 0, sizeof(int), b->stream));

    if (ctx->memo_ready) {
        MemoGeom mg = memo_geometry(ctx, b);
        if (mg.ok) {
            int rc = memo_prepare(ctx, min_occ);
            if (rc) return rc;
            if ((rc = ensure_ascii(ctx, b))) return rc;
            if (mg.use_warp && ctx->tb.nb <= 4 && ctx->memo.sieve != nullptr && ctx->memo.zero_codes == 0 &&
                       ctx->memo_occ == INT_MIN) {
                // <= 256 references, no FAILM flags baked, no zero hash: sketch membership from a shared-memory sieve
                Timed t(ctx, b->stream, K_CLASSIFY_MEMO_SIEVE);
                const int grid = ctx->sm_count;
                const size_t smem = ((size_t)ctx->memo.sieve_words + (size_t)32 * SIEVE_QCAP) * 4;
#define RK_SIEVE_LAUNCH(NWV)                                                                                             \
    do {                                                                                                                 \
        CK(cudaFuncSetAttribute(classify_memo_sieve_kernel<NWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        classify_memo_sieve_kernel<NWV><<<grid, 1024, smem, b->stream>>>(sb, ctx->s, ctx->memo, ctx->tb, redo, redo_count, \
                                                                          hits);                                        \
    } while (0)
                switch (ctx->tb.nb) {
                    case 1: RK_SIEVE_LAUNCH(2); break;
                    case 2: RK_SIEVE_LAUNCH(4); break;
                    case 3: RK_SIEVE_LAUNCH(6); break;
                    default: RK_SIEVE_LAUNCH(8); break;
                }
#undef RK_SIEVE_LAUNCH
                CK(cudaGetLastError());
            } else if (mg.use_warp && ctx->tb.nb <= 4) {     // <= 256 references: bitset rows in registers
                Timed t(ctx, b->stream, K_CLASSIFY_MEMO_WARP);
                const int grid = ctx->sm_count;
                switch (ctx->tb.nb) {
                    case 1: classify_memo_warp_kernel<2><<<grid, 1024, 0, b->stream>>>(sb, ctx->s, ctx->memo, ctx->tb, redo, redo_count, hits); break;
                    case 2: classify_memo_warp_kernel<4><<<grid, 1024, 0, b->stream>>>(sb, ctx->s, ctx->memo, ctx->tb, redo, redo_count, hits); break;
                    case 3: classify_memo_warp_kernel<6><<<grid, 1024, 0, b->stream>>>(sb, ctx->s, ctx->memo, ctx->tb, redo, redo_count, hits); break;
                    default: classify_memo_warp_kernel<8><<<grid, 1024, 0, b->stream>>>(sb, ctx->s, ctx->memo, ctx->tb, redo, redo_count, hits); break;
                }
                CK(cudaGetLastError());
            } else if (mg.use_warp) {
                Timed t(ctx, b->stream, K_CLASSIFY_MEMO_WARP);
                CK(cudaFuncSetAttribute(classify_memo_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mg.warp_smem));
                classify_memo_wide_kernel<<<ctx->sm_count, mg.warp_block, mg.warp_smem, b->stream>>>(
                    sb, ctx->s, ctx->memo, ctx->tb, mg.rwords, redo, redo_count, hits);
                CK(cudaGetLastError());
            }
            CK(cudaFuncSetAttribute(classify_memo_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mg.block_smem));
            int per_sm = 2;                                  // as many resident blocks as registers and shared memory allow
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, classify_memo_block_kernel, 256, mg.block_smem));
            int grid = ctx->sm_count * std::max(1, std::min(per_sm, 8));
            // the global lists (a read whose bound was too tight) are sized for the longest read of the batch: fewer blocks
            // when that read is enormous
            const i64 most_blocks = std::max<i64>(1, ((i64)4 << 30) / (4 * mg.scratch_stride));
            grid = (int)std::min<i64>(grid, most_blocks);
            CK(b->scratch.ensure(sizeof(u32) * (size_t)grid * (size_t)mg.scratch_stride + 64));
            // every window of every read looks its rank up in the 4^k-entry table: keep it in L2, stream the bases
            if (!mg.use_warp) l2_window(ctx, b->stream, ctx->memo.tab, sizeof(u32) * (size_t)ctx->memo.ncodes);
            {
                Timed t(ctx, b->stream, K_CLASSIFY_MEMO_BLOCK);
                classify_memo_block_kernel<<<grid, 256, mg.block_smem, b->stream>>>(
                    sb, ctx->s, ctx->memo, ctx->tb, mg.use_warp ? redo : nullptr, mg.use_warp ? redo_count : nullptr, b->nseq,
                    mg.lcap, mg.hs_cap, mg.rpad, b->scratch.as<u32>(), mg.scratch_stride, hits);
            }
            if (!mg.use_warp) l2_window(ctx, b->stream, nullptr, 0);
            CK(cudaGetLastError());
            return RKG_OK;
        }
    }
    { int rc = memo_flush(ctx); if (rc) return rc; }
    { int rc = ensure_packed(ctx, b); if (rc) return rc; }

    GroupGeom gg = group_geometry(ctx, b);
    // <= 256 references: the warp-cooperative kernel (rk_warp.cuh); more: the sub-warp group kernel with shared-memory counters
    const size_t warp_smem = (size_t)gg.filter_words * 4 + (size_t)32 * WARP_QCAP * 8;
    if (gg.ok && ctx->tb.nb <= 4 && !ctx->group_kernel && warp_smem + sizeof(u64) * HASH_LUT_ENTRIES + 1024 <= ctx->smem_optin) {
        Timed t(ctx, b->stream, K_CLASSIFY_WARP);
        const int grid = ctx->sm_count;
        DISPATCH_KCSEL(ctx->kcsel, {
            if (ctx->tb.nb <= 1) {
                CK(cudaFuncSetAttribute(classify_warp_kernel<KS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warp_smem));
                classify_warp_kernel<KS, 2><<<grid, 1024, warp_smem, b->stream>>>(sb, ctx->kl, ctx->s, ct, min_occ, ctx->tb, gg.filter_words,
                                                                                   redo, redo_count, hits);
            } else {
                CK(cudaFuncSetAttribute(classify_warp_kernel<KS, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warp_smem));
                classify_warp_kernel<KS, 8><<<grid, 1024, warp_smem, b->stream>>>(sb, ctx->kl, ctx->s, ct, min_occ, ctx->tb, gg.filter_words,
                                                                                   redo, redo_count, hits);
            }
        });
        CK(cudaGetLastError());
    } else if (gg.ok) {
        Timed t(ctx, b->stream, K_CLASSIFY_GROUP);
        int grid = ctx->sm_count;
        DISPATCH_KCSEL(ctx->kcsel, {
            CK(cudaFuncSetAttribute(classify_group_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gg.smem));
            classify_group_kernel<KS><<<grid, gg.block, gg.smem, b->stream>>>(sb, ctx->kl, ctx->s, ct, min_occ, ctx->tb,
                                                                               gg.glog2, gg.rwords, gg.filter_words,
                                                                               redo, redo_count, hits);
        });
        CK(cudaGetLastError());
    }
    // exact path: the redo list of the group kernel, or every read when the group kernel was skipped
    {
        const int block = 512;
        const int hs_cap = next_pow2(2 * (i64)std::min<i64>(ctx->s, std::max<i64>(b->max_ntotal, 1)));
        const int rpad = (ctx->nrefs + 63) / 64 * 64;
        const size_t budget = ctx->smem_optin / 2 - 1024 - sizeof(u64) * HASH_LUT_ENTRIES;   // two blocks per SM, each with the hashing tables
        size_t fixed = (size_t)rpad * 4;
        int hs_in_smem = ((size_t)hs_cap * 8 + fixed <= budget / 2) ? 1 : 0;
        if (hs_in_smem) fixed += (size_t)hs_cap * 8;
        i64 hbuf_cap = 0;
        if (fixed + (size_t)b->max_ntotal * 8 <= budget) hbuf_cap = b->max_ntotal;
        const size_t smem = fixed + (size_t)hbuf_cap * 8;
        const i64 stride = (hbuf_cap >= b->max_ntotal) ? 0 : b->max_ntotal;
        const int grid = ctx->sm_count * 2;
        CK(b->scratch.ensure(sizeof(u64) * (size_t)grid * (size_t)(stride + hs_cap) + 64));
        Timed t(ctx, b->stream, K_CLASSIFY_BLOCK);
        DISPATCH_KCSEL(ctx->kcsel, {
            CK(cudaFuncSetAttribute(classify_block_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            classify_block_kernel<KS><<<grid, block, smem, b->stream>>>(sb, ctx->kl, ctx->s, ct, min_occ, ctx->tb,
                                                                         gg.ok ? redo : nullptr, gg.ok ? redo_count : nullptr,
                                                                         b->nseq, (int)hbuf_cap, hs_cap, hs_in_smem, rpad,
                                                                         b->scratch.as<u64>(), stride, hits);
        });
        CK(cudaGetLastError());
    }
    return RKG_OK;
}

// Sketch every sequence of a batch into d_sketches [nseq * s] / d_lens [nseq] (device pointers).
int launch_sketch(rkg_ctx* ctx, rkg_batch* b, CountTable ct, int filter_mode, int occ, u64* d_sketches, int* d_lens,
                  u64* d_sorted_all, const i64* d_all_offs) {
    if (b->nseq == 0) return RKG_OK;
    { int rc = memo_flush(ctx); if (rc) return rc; }       // the count filter reads the read count table itself
    { int rc = ensure_packed(ctx, b); if (rc) return rc; }
    const int block = 512;
    const int sort_cap = next_pow2(std::max<i64>(std::min<i64>(ctx->s, std::max<i64>(b->max_ntotal, 1)), 2));
    const int sort_in_smem = ((size_t)sort_cap * 8 + sizeof(u64) * HASH_LUT_ENTRIES <= ctx->smem_optin - 4096) ? 1 : 0;
    const size_t smem = sort_in_smem ? (size_t)sort_cap * 8 : 0;
    i64 stride = std::max<i64>(b->max_ntotal, 1);
    if (d_sorted_all) {
        if (stride > ((i64)1 << 30)) return fail(ctx, RKG_ERR_ARG, "per-sample -I counting: a reference of more than 2^30 k-mers is not supported");
        stride = next_pow2(stride);
    }
    const int grid = (int)std::min<i64>(b->nseq, (i64)ctx->sm_count * 2);
    CK(b->scratch.ensure(sizeof(u64) * (size_t)grid * (size_t)(stride + sort_cap) + 64));
    Timed t(ctx, b->stream, K_SKETCH);
    DISPATCH_KCSEL(ctx->kcsel, {
        CK(cudaFuncSetAttribute(sketch_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sketch_kernel<KS><<<grid, block, smem, b->stream>>>(seq_view(b), ctx->kl, ctx->s, ct, filter_mode, occ,
                                                             b->scratch.as<u64>(), stride, sort_cap, sort_in_smem,
                                                             d_sketches, d_lens, d_sorted_all, d_all_offs);
    });
    CK(cudaGetLastError());
    return RKG_OK;
}

// The (hash, rank) -> reference-bitset table (+ the memoised tables) from the sorted sketches in ctx->ref_sketches / ref_lens.
int build_ref_table(rkg_ctx* ctx, int nrefs, cudaStream_t st) {
    const int s = ctx->s;
    int rc = RKG_OK;
    // table geometry
    // The table is dense in the references: cap x (1 + ceil(nrefs / 64)) words with cap >= 2 nrefs s, i.e. quadratic in
    // the number of references (5,000 references at s = 1000 take about 11 GB; include/rkmh_gpu.h states the limit).
    const i64 nent = (i64)nrefs * s;
    const u32 cap = (u32)next_pow2(std::max<i64>(2 * nent, 1024));
    const u32 nb = (u32)((nrefs + 63) / 64);
    RefTable tb{};
    tb.cap_mask = cap - 1; tb.nb = nb; tb.stride = 1 + nb; tb.nrefs = (u32)nrefs;
    int flog2 = 10;
    while (flog2 < 20 && (1LL << flog2) < 16 * nent) ++flog2;
    tb.filter_log2 = (u32)flog2;
    CK(ctx->tb_entries.ensure(sizeof(u64) * (size_t)cap * tb.stride));
    CK(ctx->tb_maxmult.ensure(sizeof(u32) * (size_t)cap));
    CK(ctx->tb_dupbase.ensure(sizeof(u32) * (size_t)cap));
    CK(ctx->tb_dupbits.ensure(sizeof(u64) * (size_t)std::max<i64>(nent, 1) * nb));
    CK(ctx->tb_filter.ensure(sizeof(u32) * ((size_t)1 << flog2) / 32));
    CK(ctx->tb_cursor.ensure(256));
    CK(cudaMemsetAsync(ctx->tb_entries.p, 0, sizeof(u64) * (size_t)cap * tb.stride, st));
    CK(cudaMemsetAsync(ctx->tb_maxmult.p, 0, sizeof(u32) * (size_t)cap, st));
    CK(cudaMemsetAsync(ctx->tb_dupbase.p, 0, sizeof(u32) * (size_t)cap, st));
    CK(cudaMemsetAsync(ctx->tb_dupbits.p, 0, sizeof(u64) * (size_t)std::max<i64>(nent, 1) * nb, st));
    CK(cudaMemsetAsync(ctx->tb_filter.p, 0, sizeof(u32) * ((size_t)1 << flog2) / 32, st));
    CK(cudaMemsetAsync(ctx->tb_cursor.p, 0, 256, st));
    tb.entries = ctx->tb_entries.as<u64>();
    tb.maxmult = ctx->tb_maxmult.as<u32>();
    tb.dup_base = ctx->tb_dupbase.as<u32>();
    tb.dup_bits = ctx->tb_dupbits.as<u64>();
    tb.filter = ctx->tb_filter.as<u32>();
    {
        Timed t(ctx, st, K_TABLE);
        int grid = (int)std::min<i64>((nent + 255) / 256, (i64)ctx->sm_count * 8);
        table_insert_kernel<<<grid, 256, 0, st>>>(ctx->ref_sketches.as<u64>(), ctx->ref_lens.as<int>(), nrefs, s, tb,
                                                         ctx->tb_filter.as<u32>());
        table_alloc_dup_kernel<<<(int)std::min<i64>((cap + 255) / 256, (i64)ctx->sm_count * 8), 256, 0, st>>>(
            tb, ctx->tb_cursor.as<u32>());
        table_fill_dup_kernel<<<grid, 256, 0, st>>>(ctx->ref_sketches.as<u64>(), ctx->ref_lens.as<int>(), nrefs, s, tb);
        ctx->launches += 2;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    ctx->tb = tb;
    ctx->set_table = false;
    if ((rc = memo_build(ctx, st))) return rc;
    ctx->refs_loaded = true;
    return RKG_OK;
}

int ensure_slots(rkg_ctx* ctx, int n = 2) {
    for (int i = 0; i < n; ++i)
        if (!ctx->slots[i]) { int rc = batch_alloc(ctx, &ctx->slots[i]); if (rc) return rc; }
    return RKG_OK;
}

// Split [0, n) into chunks of bounded bytes / reads for the double-buffered host paths.
std::vector<std::pair<i64, i64>> make_chunks(const int64_t* offs, i64 n, i64 max_bytes, i64 max_reads) {
    std::vector<std::pair<i64, i64>> out;
    i64 r0 = 0;
    while (r0 < n) {
        i64 hi = std::min<i64>(n, r0 + max_reads);
        // largest r1 in (r0, hi] with offs[r1] - offs[r0] <= max_bytes (at least one read)
        i64 lo = r0 + 1, best = r0 + 1;
        while (lo <= hi) {
            i64 mid = (lo + hi) / 2;
            if (offs[mid] - offs[r0] <= max_bytes) { best = mid; lo = mid + 1; } else hi = mid - 1;
        }
        out.emplace_back(r0, best);
        r0 = best;
    }
    return out;
}

constexpr i64 CHUNK_BYTES = 192LL << 20;
constexpr i64 CHUNK_READS = 2LL << 20;

}  // namespace

// =====================================================================================================
extern "C" {

int rkg_create(int device, rkg_ctx** out) {
    rkg_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, RKG_ERR_ARG, "rkg_create: null output pointer");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RKG_ERR_CUDA, "rkg_create: no CUDA device (%s); this library has no CPU fallback",
                    cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, RKG_ERR_ARG, "rkg_create: device %d of %d", device, ndev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, RKG_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, RKG_ERR_CUDA, "rkg_create: device %d is sm_%d%d; this build targets sm_100a (B200) only",
                    device, prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return fail(nullptr, RKG_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    ctx = new rkg_ctx();
    ctx->device = device;
    { const char* nm = getenv("RKMH_NO_MEMO"); ctx->memo_enabled = !(nm && nm[0] && nm[0] != '0'); }
    { const char* nm = getenv("RKMH_NO_SIEVE"); ctx->sieve_enabled = !(nm && nm[0] && nm[0] != '0'); }
    { const char* nm = getenv("RKMH_SHARD_TWO_PASS"); ctx->shard_two_pass = nm && nm[0] && nm[0] != '0'; }
    { const char* nm = getenv("RKMH_GROUP_KERNEL"); ctx->group_kernel = nm && nm[0] && nm[0] != '0'; }   // round-1 Murmur fast path (A/B timing)
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    { const char* nm = getenv("RKMH_NO_L2_POLICY"); ctx->l2_policy = !(nm && nm[0] && nm[0] != '0'); }
    // a persisting-L2 carve-out for the tables the streaming kernels hit at random (access-policy windows below)
    ctx->l2_persist_max = (size_t)std::max(prop.persistingL2CacheMaxSize, 0);
    ctx->l2_window_max = (size_t)std::max(prop.accessPolicyMaxWindowSize, 0);
    if (ctx->l2_policy && ctx->l2_persist_max > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, ctx->l2_persist_max);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, RKG_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    {   // key * constant tables of the Murmur path (rk_hash.cuh), copied into shared memory by every hashing kernel
        std::vector<u64> lut(HASH_LUT_ENTRIES);
        build_hash_lut(lut.data());
        if (ctx->hash_lut.ensure(sizeof(u64) * HASH_LUT_ENTRIES) != cudaSuccess ||
            cudaMemcpy(ctx->hash_lut.p, lut.data(), sizeof(u64) * HASH_LUT_ENTRIES, cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaStreamDestroy(ctx->stream);
            delete ctx;
            return fail(nullptr, RKG_ERR_CUDA, "rkg_create: cannot allocate the hashing tables");
        }
    }
    *out = ctx;
    return RKG_OK;
}

void rkg_destroy(rkg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    profile_collect(ctx);
    for (int i = 0; i < NSLOTS; ++i) batch_release(ctx->slots[i]);
    for (int i = 0; i < NSLOTS; ++i) if (ctx->stage[i]) cudaFreeHost(ctx->stage[i]);
    for (int i = 0; i < NSLOTS; ++i) if (ctx->h2d_done[i]) cudaEventDestroy(ctx->h2d_done[i]);
    if (ctx->copy_event) cudaEventDestroy(ctx->copy_event);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (DevBuf* d : {&ctx->shard_scan, &ctx->shard_cnt, &ctx->depth_keys, &ctx->depth_counts, &ctx->depth_zero, &ctx->hash_lut, &ctx->read_counts, &ctx->ref_counts, &ctx->ref_sketches, &ctx->ref_lens, &ctx->tb_entries,
                      &ctx->tb_maxmult, &ctx->tb_dupbase, &ctx->tb_dupbits, &ctx->tb_filter, &ctx->tb_cursor,
                      &ctx->memo_tab, &ctx->memo_ref, &ctx->memo_hash, &ctx->memo_counts, &ctx->memo_tmp, &ctx->memo_misc,
                      &ctx->memo_sieve, &ctx->memo_cmap})
        d->release();
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* rkg_last_error(const rkg_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rkg_set_params(rkg_ctx* ctx, const int* ks, int nk, int sketch_size, uint64_t read_table_size,
                   uint64_t ref_table_size) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ks || nk < 1 || nk > MAX_NK) return fail(ctx, RKG_ERR_ARG, "rkg_set_params: 1..%d k-mer sizes, got %d", MAX_NK, nk);
    if (sketch_size < 1 || sketch_size > (1 << 26))
        return fail(ctx, RKG_ERR_ARG, "rkg_set_params: sketch size %d (supported: 1..2^26)", sketch_size);
    CK(cudaSetDevice(ctx->device));
    KList kl{};
    kl.nk = nk;
    int kcsel = -1, kmax = 0, kmin = MAX_K;
    for (int j = 0; j < nk; ++j) {
        if (ks[j] < 1 || ks[j] > MAX_K)
            return fail(ctx, RKG_ERR_ARG, "rkg_set_params: k = %d unsupported (1..%d)", ks[j], MAX_K);
        kl.k[j] = ks[j];
        kl.kv[j] = kvariant_of(ks[j]);
        kl.ksum += ks[j];
        kcsel = (kcsel == -1 || kcsel == kl.kv[j]) ? kl.kv[j] : NUM_KV;
        kmax = std::max(kmax, ks[j]); kmin = std::min(kmin, ks[j]);
    }
    kl.kmax = kmax;
    kl.lut = ctx->hash_lut.as<u64>();
    ctx->kl = kl; ctx->kcsel = kcsel; ctx->kmax = kmax; ctx->kmin = kmin; ctx->s = sketch_size;
    ctx->refs_loaded = false;
    ctx->memo_ready = false;
    auto setup = [&](uint64_t size, DevBuf& buf, CountTable& ct) -> cudaError_t {
        ct.counts = nullptr;
        if (size == 0) { buf.release(); return cudaSuccess; }
        cudaError_t e = buf.ensure(sizeof(int) * size);
        if (e != cudaSuccess) return e;
        ct.counts = buf.as<int>();
        ct.mod = make_fastmod(size);
        return cudaMemsetAsync(buf.p, 0, sizeof(int) * size, ctx->stream);
    };
    if (read_table_size >= (1ULL << 62) || ref_table_size >= (1ULL << 62)) return fail(ctx, RKG_ERR_ARG, "table size");
    CK(setup(read_table_size, ctx->read_counts, ctx->read_ct));
    CK(setup(ref_table_size, ctx->ref_counts, ctx->ref_ct));
    ctx->read_tbl_size = read_table_size; ctx->ref_tbl_size = ref_table_size;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->params_set = true;
    return RKG_OK;
}

// Thread-safe (see the header): they read ctx->device only and report errors through the thread-local slot of
// rkg_last_error(NULL), so a parser thread may pin chunks while a worker thread is inside a batch call on the same context.
int rkg_host_alloc(rkg_ctx* ctx, size_t bytes, void** ptr) {
    if (!ctx || !ptr) return RKG_ERR_ARG;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e == cudaSuccess) e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) return fail(nullptr, e == cudaErrorMemoryAllocation ? RKG_ERR_NOMEM : RKG_ERR_CUDA,
                                      "rkg_host_alloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    return RKG_OK;
}

int rkg_host_free(rkg_ctx* ctx, void* ptr) {
    if (!ctx) return RKG_ERR_ARG;
    const cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) return fail(nullptr, RKG_ERR_CUDA, "rkg_host_free: %s", cudaGetErrorString(e));
    return RKG_OK;
}

int rkg_synchronize(rkg_ctx* ctx) {
    if (!ctx) return RKG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    return RKG_OK;
}

// ---- references -------------------------------------------------------------------------------------
int rkg_load_refs(rkg_ctx* ctx, const char* bases, const int64_t* offs, int nrefs, int max_ref_occ, int per_sample_set) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->params_set) return fail(ctx, RKG_ERR_STATE, "rkg_load_refs before rkg_set_params");
    if (!bases || !offs || nrefs < 1) return fail(ctx, RKG_ERR_ARG, "rkg_load_refs: need at least one reference");
    if (max_ref_occ >= 0 && ctx->ref_ct.counts == nullptr)
        return fail(ctx, RKG_ERR_STATE, "max_ref_occ >= 0 needs a reference count table (rkg_set_params ref_table_size)");
    if (2 * (i64)nrefs * ctx->s > ((i64)1 << 30))
        return fail(ctx, RKG_ERR_ARG, "rkg_load_refs: %d references x sketch size %d exceeds the 2^29 table entries supported", nrefs, ctx->s);
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_slots(ctx);
    if (rc) return rc;
    rkg_batch* b = ctx->slots[0];
    CK(cudaStreamSynchronize(b->stream));
    ctx->memo_ready = false;                  // the reference side always goes through the hashing kernels
    ctx->refs_loaded = false;
    if ((rc = batch_fill(ctx, b, bases, offs, 0, nrefs))) return rc;
    const int s = ctx->s;
    ctx->nrefs = nrefs;
    CK(ctx->ref_sketches.ensure(sizeof(u64) * (size_t)nrefs * s));
    CK(ctx->ref_lens.ensure(sizeof(int) * (size_t)nrefs));
    CK(cudaMemsetAsync(ctx->ref_sketches.p, 0, sizeof(u64) * (size_t)nrefs * s, b->stream));

    int filter_mode = 0;
    if (max_ref_occ >= 0) {
        filter_mode = 1;
        CK(cudaMemsetAsync(ctx->ref_ct.counts, 0, sizeof(int) * ctx->ref_tbl_size, b->stream));
        if (!per_sample_set) {
            if ((rc = launch_count(ctx, b, ctx->ref_ct, 0))) return rc;
        } else {
            // distinct hashes per reference (zero included): sort every hash of each reference, count run heads
            std::vector<i64> all_offs(nrefs + 1, 0);
            for (int r = 0; r < nrefs; ++r) {
                i64 len = offs[r + 1] - offs[r], n = 0;
                for (int j = 0; j < ctx->kl.nk; ++j) n += std::max<i64>(len - ctx->kl.k[j], 0);
                all_offs[r + 1] = all_offs[r] + n;
            }
            DevBuf d_all, d_all_offs;
            CK(d_all.ensure(sizeof(u64) * (size_t)std::max<i64>(all_offs[nrefs], 1)));
            CK(d_all_offs.ensure(sizeof(i64) * (size_t)(nrefs + 1)));
            CK(cudaMemcpyAsync(d_all_offs.p, all_offs.data(), sizeof(i64) * (size_t)(nrefs + 1), cudaMemcpyHostToDevice, b->stream));
            rc = launch_sketch(ctx, b, CountTable{nullptr, {}}, 0, 0, ctx->ref_sketches.as<u64>(), ctx->ref_lens.as<int>(),
                               d_all.as<u64>(), d_all_offs.as<i64>());
            if (!rc) {
                Timed t(ctx, b->stream, K_COUNT_DISTINCT);
                count_distinct_kernel<<<std::min(nrefs, ctx->sm_count * 4), 256, 0, b->stream>>>(
                    d_all.as<u64>(), d_all_offs.as<i64>(), nrefs, ctx->ref_ct);
            }
            cudaError_t e = cudaStreamSynchronize(b->stream);
            d_all.release(); d_all_offs.release();
            if (rc) return rc;
            CK(e);
        }
    }
    if ((rc = launch_sketch(ctx, b, ctx->ref_ct, filter_mode, max_ref_occ, ctx->ref_sketches.as<u64>(),
                            ctx->ref_lens.as<int>(), nullptr, nullptr)))
        return rc;

    return build_ref_table(ctx, nrefs, b->stream);
}

int rkg_load_ref_sketches(rkg_ctx* ctx, const uint64_t* sketches, const int32_t* lens, int nrefs) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->params_set) return fail(ctx, RKG_ERR_STATE, "rkg_load_ref_sketches before rkg_set_params");
    if (!sketches || !lens || nrefs < 1) return fail(ctx, RKG_ERR_ARG, "rkg_load_ref_sketches: need at least one reference");
    const int s = ctx->s;
    if (2 * (i64)nrefs * s > ((i64)1 << 30))
        return fail(ctx, RKG_ERR_ARG, "rkg_load_ref_sketches: %d references x sketch size %d exceeds the 2^29 table entries supported", nrefs, s);
    for (int r = 0; r < nrefs; ++r) {                       // what minhashes() guarantees: ascending, non-zero, at most s
        if (lens[r] < 0 || lens[r] > s) return fail(ctx, RKG_ERR_ARG, "rkg_load_ref_sketches: sketch %d has length %d (sketch size %d)", r, lens[r], s);
        const uint64_t* sk = sketches + (size_t)r * s;
        for (int i = 0; i < lens[r]; ++i)
            if (sk[i] == 0 || (i > 0 && sk[i] < sk[i - 1])) return fail(ctx, RKG_ERR_ARG, "rkg_load_ref_sketches: sketch %d is not an ascending list of non-zero hashes", r);
    }
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    ctx->memo_ready = false; ctx->refs_loaded = false;
    ctx->nrefs = nrefs;
    CK(ctx->ref_sketches.ensure(sizeof(u64) * (size_t)nrefs * s));
    CK(ctx->ref_lens.ensure(sizeof(int) * (size_t)nrefs));
    CK(cudaMemcpyAsync(ctx->ref_sketches.p, sketches, sizeof(u64) * (size_t)nrefs * s, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->ref_lens.p, lens, sizeof(int) * (size_t)nrefs, cudaMemcpyHostToDevice, ctx->stream));
    return build_ref_table(ctx, nrefs, ctx->stream);
}

int rkg_get_ref_sketches(rkg_ctx* ctx, uint64_t* sketches, int32_t* lens) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->refs_loaded || ctx->set_table) return fail(ctx, RKG_ERR_STATE, "no references loaded");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(sketches, ctx->ref_sketches.p, sizeof(u64) * (size_t)ctx->nrefs * ctx->s, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lens, ctx->ref_lens.p, sizeof(int) * (size_t)ctx->nrefs, cudaMemcpyDeviceToHost));
    return RKG_OK;
}

// ---- count table ------------------------------------------------------------------------------------
int rkg_count_reset(rkg_ctx* ctx) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->read_ct.counts) return fail(ctx, RKG_ERR_STATE, "no read count table (rkg_set_params read_table_size)");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    // on the context's own stream and waited for: the batch streams are non-blocking, a legacy-stream memset would
    // not be ordered before the count kernels that follow
    CK(cudaMemsetAsync(ctx->read_ct.counts, 0, sizeof(int) * ctx->read_tbl_size, ctx->stream));
    if (ctx->memo_ready && ctx->memo.code_counts) {
        CK(cudaMemsetAsync(ctx->memo.code_counts, 0, sizeof(int) * ((size_t)ctx->memo.ncodes + MEMO_PARTIAL_PAD), ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->memo_pending = false;
    ctx->memo_occ_stale = true;
    return RKG_OK;
}

int rkg_count_table(rkg_ctx* ctx, void** device_ptr, uint64_t* nslots) {
    if (!ctx || !device_ptr || !nslots) return RKG_ERR_ARG;
    if (!ctx->read_ct.counts) return fail(ctx, RKG_ERR_STATE, "no read count table");
    CK(cudaSetDevice(ctx->device));
    { int rc = memo_flush(ctx); if (rc) return rc; }
    ctx->memo_occ_stale = true;               // the caller may all-reduce the table in place
    *device_ptr = ctx->read_ct.counts;
    *nslots = ctx->read_tbl_size;
    return RKG_OK;
}

int rkg_count_partial(rkg_ctx* ctx, int count_zeros, void** device_ptr, uint64_t* nslots) {
    if (!ctx || !device_ptr || !nslots) return RKG_ERR_ARG;
    if (!ctx->read_ct.counts) return fail(ctx, RKG_ERR_STATE, "no read count table");
    if (!ctx->memo_ready || !ctx->memo.code_counts)
        return fail(ctx, RKG_ERR_STATE, "rkg_count_partial: the memoised path is not active (all-reduce rkg_count_table instead)");
    if (ctx->memo_pending && ctx->memo_pending_zeros != (count_zeros != 0))
        return fail(ctx, RKG_ERR_ARG, "rkg_count_partial: count_zeros differs from the pending rkg_count_batch calls");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());              // the count kernels of every batch stream have finished
    // whatever the caller sums into the array is folded into the read count table at its next use, also on a rank
    // that counted no read itself
    ctx->memo_pending = true; ctx->memo_pending_zeros = count_zeros != 0;
    *device_ptr = ctx->memo.code_counts;
    *nslots = (uint64_t)ctx->memo.ncodes + MEMO_PARTIAL_PAD;
    return RKG_OK;
}

// which: 0 = read count table, 1 = reference count table
static CountTable* pick_table(rkg_ctx* ctx, int which, u64* nslots) {
    if (which == 0 && ctx->read_ct.counts) { *nslots = ctx->read_tbl_size; return &ctx->read_ct; }
    if (which == 1 && ctx->ref_ct.counts) { *nslots = ctx->ref_tbl_size; return &ctx->ref_ct; }
    return nullptr;
}

int rkg_count_export(rkg_ctx* ctx, int which, uint32_t* slots, int32_t* counts, uint64_t cap, uint64_t* n) {
    if (!ctx || !n) return RKG_ERR_ARG;
    u64 nslots = 0;
    CountTable* ct = pick_table(ctx, which, &nslots);
    if (!ct) return fail(ctx, RKG_ERR_STATE, "rkg_count_export: table %d does not exist (rkg_set_params)", which);
    if (nslots > 0xFFFFFFFFULL) return fail(ctx, RKG_ERR_ARG, "rkg_count_export: tables of more than 2^32 slots are not supported");
    CK(cudaSetDevice(ctx->device));
    if (which == 0) { int rc = memo_flush(ctx); if (rc) return rc; }
    CK(cudaDeviceSynchronize());
    DevBuf d_s, d_v, d_n;
    int rc = RKG_OK;
    do {
        const u64 dcap = (slots && counts) ? cap : 0;
        if (d_s.ensure(sizeof(u32) * (size_t)std::max<u64>(dcap, 1)) || d_v.ensure(sizeof(int) * (size_t)std::max<u64>(dcap, 1)) || d_n.ensure(64)) {
            rc = fail(ctx, RKG_ERR_NOMEM, "rkg_count_export: device allocation failed"); break;
        }
        cudaMemsetAsync(d_n.p, 0, 64, ctx->stream);
        {
            Timed t(ctx, ctx->stream, K_MEMO_MAINT);
            count_export_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ct->counts, nslots, d_s.as<u32>(), d_v.as<int>(), dcap,
                                                                           d_n.as<unsigned long long>());
        }
        unsigned long long found = 0;
        cudaMemcpyAsync(&found, d_n.p, sizeof(found), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = fail(ctx, RKG_ERR_CUDA, "rkg_count_export: %s", cudaGetErrorString(cudaGetLastError())); break; }
        *n = found;
        if (dcap > 0) {
            if (found > dcap) { rc = fail(ctx, RKG_ERR_ARG, "rkg_count_export: %llu non-zero slots, room for %llu", found, (unsigned long long)dcap); break; }
            cudaMemcpy(slots, d_s.p, sizeof(u32) * (size_t)found, cudaMemcpyDeviceToHost);
            cudaMemcpy(counts, d_v.p, sizeof(int) * (size_t)found, cudaMemcpyDeviceToHost);
        }
    } while (0);
    d_s.release(); d_v.release(); d_n.release();
    return rc;
}

int rkg_count_import(rkg_ctx* ctx, int which, const uint32_t* slots, const int32_t* counts, uint64_t n) {
    if (!ctx || (n > 0 && (!slots || !counts))) return RKG_ERR_ARG;
    u64 nslots = 0;
    CountTable* ct = pick_table(ctx, which, &nslots);
    if (!ct) return fail(ctx, RKG_ERR_STATE, "rkg_count_import: table %d does not exist (rkg_set_params)", which);
    for (u64 i = 0; i < n; ++i)
        if ((u64)slots[i] >= nslots) return fail(ctx, RKG_ERR_ARG, "rkg_count_import: slot %u of a %llu-slot table", slots[i], (unsigned long long)nslots);
    CK(cudaSetDevice(ctx->device));
    if (which == 0) { int rc = rkg_count_reset(ctx); if (rc) return rc; }
    else { CK(cudaDeviceSynchronize()); CK(cudaMemsetAsync(ct->counts, 0, sizeof(int) * nslots, ctx->stream)); }
    if (n == 0) { CK(cudaStreamSynchronize(ctx->stream)); return RKG_OK; }
    DevBuf d_s, d_v;
    int rc = RKG_OK;
    do {
        if (d_s.ensure(sizeof(u32) * (size_t)n) || d_v.ensure(sizeof(int) * (size_t)n)) { rc = fail(ctx, RKG_ERR_NOMEM, "rkg_count_import: device allocation failed"); break; }
        cudaMemcpyAsync(d_s.p, slots, sizeof(u32) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d_v.p, counts, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
        {
            Timed t(ctx, ctx->stream, K_MEMO_MAINT);
            count_import_kernel<<<(int)std::min<u64>((n + 255) / 256, (u64)ctx->sm_count * 8), 256, 0, ctx->stream>>>(ct->counts, nslots, d_s.as<u32>(),
                                                                                                                    d_v.as<int>(), n);
        }
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = fail(ctx, RKG_ERR_CUDA, "rkg_count_import: %s", cudaGetErrorString(cudaGetLastError())); break; }
    } while (0);
    d_s.release(); d_v.release();
    if (which == 0) ctx->memo_occ_stale = true;
    return rc;
}

int rkg_count_lookup(rkg_ctx* ctx, const uint64_t* keys, int64_t n, int32_t* out) {
    if (!ctx || !keys || !out) return RKG_ERR_ARG;
    if (!ctx->read_ct.counts) return fail(ctx, RKG_ERR_STATE, "no read count table");
    CK(cudaSetDevice(ctx->device));
    { int rc = memo_flush(ctx); if (rc) return rc; }
    CK(cudaDeviceSynchronize());
    for (int64_t i = 0; i < n; ++i) {   // parity/debug helper: one small copy per key
        u64 slot = keys[i] % ctx->read_tbl_size;
        CK(cudaMemcpy(out + i, ctx->read_ct.counts + slot, sizeof(int), cudaMemcpyDeviceToHost));
    }
    return RKG_OK;
}

static int host_pipeline(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nreads, int mode, int arg,
                         rkg_hit* out, int64_t fixed_len = 0) {
    // mode 0: count (arg = count_zeros); mode 1: classify (arg = min_occ); mode 2: the copies of mode 1 without its kernels
    // Adaptive packing (rkg_set_host_pack mode 2): a chunk either crosses the link as it is (no host work at all) or is
    // packed to 0.375 bytes per base by the host's threads first.  Packing pays exactly while the link is busy with earlier
    // chunks, so each chunk is packed when the copies still queued will last about as long as packing it takes, and sent
    // plain otherwise: the link never waits for the packer and the packer never waits for the link, whatever their speeds
    // are on this host (both are measured on the context's first chunks; `lead` is re-synchronised with the copy events).
    const bool adaptive = ctx->host_pack == 2 && mode != 2;
    const int ns = adaptive ? NSLOTS : 2;
    int rc = ensure_slots(ctx, ns);
    if (rc) return rc;
    const i64 chunk_bytes = adaptive ? (64LL << 20) : CHUNK_BYTES;
    std::vector<std::pair<i64, i64>> chunks;
    if (offs) chunks = make_chunks(offs, nreads, chunk_bytes, CHUNK_READS);
    else {
        const i64 per = std::max<i64>(1, std::min<i64>(CHUNK_READS, chunk_bytes / std::max<i64>(fixed_len, 1)));
        for (i64 r0 = 0; r0 < nreads; r0 += per) chunks.emplace_back(r0, std::min<i64>(nreads, r0 + per));
    }
    using clk = std::chrono::steady_clock;
    auto ms_since = [](clk::time_point t) { return std::chrono::duration<double, std::milli>(clk::now() - t).count(); };
    double lead = 0.0;                                       // estimated time the queued host-to-device copies still take
    clk::time_point t_last = clk::now();
    int last_copy_slot = -1;
    if (adaptive)
        for (int i = 0; i < ns; ++i)
            if (!ctx->h2d_done[i]) CK(cudaEventCreateWithFlags(&ctx->h2d_done[i], cudaEventDisableTiming));
    for (size_t c = 0; c < chunks.size(); ++c) {
        const int slot = (int)(c % (size_t)ns);
        rkg_batch* b = ctx->slots[slot];
        CK(cudaStreamSynchronize(b->stream));     // the slot's previous chunk (c - ns) is finished
        int pack_mode = -1;
        if (adaptive) {
            const i64 r0 = chunks[c].first, r1 = chunks[c].second;
            const double bytes = (double)(offs ? offs[r1] - offs[r0] : (r1 - r0) * fixed_len);
            lead = std::max(0.0, lead - ms_since(t_last));
            t_last = clk::now();
            if (last_copy_slot >= 0) {
                // the estimate against the truth: copies all done although some were expected to be left -> the link is
                // faster than assumed; none expected but the last one is still under way -> slower
                const bool done = cudaEventQuery(ctx->h2d_done[last_copy_slot]) == cudaSuccess;
                if (done && lead > 0.25 * ctx->dma_ms_per_byte * bytes) ctx->dma_ms_per_byte *= 0.93;
                if (!done && lead == 0.0) { ctx->dma_ms_per_byte *= 1.07; lead = 0.1 * ctx->dma_ms_per_byte * bytes; }
                if (done) lead = 0.0;
            }
            // staging sized before anything is timed (page-locking it takes milliseconds)
            if ((rc = ensure_stage(ctx, slot, (size_t)(bytes * 0.375) + (64 << 10)))) return rc;
            if (ctx->dma_ms_per_byte == 0.0) {
                // first chunk of the context: plain and alone, to time the link
                for (int i = 0; i < ns; ++i) CK(cudaStreamSynchronize(ctx->slots[i]->stream));
                if ((rc = batch_fill(ctx, b, bases, offs, r0, r1, fixed_len, 0))) return rc;     // (first use: allocates the slot's device buffers)
                CK(cudaStreamSynchronize(b->stream));
                const clk::time_point t0 = clk::now();
                if (b->total > 0) CK(cudaMemcpyAsync(b->bases.p, bases + b->base0, (size_t)b->total, cudaMemcpyHostToDevice, b->stream));
                CK(cudaStreamSynchronize(b->stream));
                if (bytes > 0) ctx->dma_ms_per_byte = std::max(ms_since(t0), 1e-3) / bytes;
                pack_mode = 2;                               // filled already
                t_last = clk::now();
            } else if (ctx->pack_ms_per_byte == 0.0) {
                pack_mode = 1;                               // second chunk: packed, to time the packer
            } else {
                pack_mode = (pack_pays(ctx) && lead >= ctx->pack_lead * ctx->pack_ms_per_byte * bytes) ? 1 : 0;
            }
            if (pack_mode != 2) {
                const bool first_use = b->bases.cap == 0;   // the slot's device buffers are allocated inside: not a timing sample
                const clk::time_point t0 = clk::now();
                if ((rc = batch_fill(ctx, b, bases, offs, r0, r1, fixed_len, pack_mode))) return rc;
                if (pack_mode == 1 && bytes > 0 && !first_use) {
                    const double per = ms_since(t0) / bytes;
                    ctx->pack_ms_per_byte = ctx->pack_ms_per_byte == 0.0 ? per : 0.5 * ctx->pack_ms_per_byte + 0.5 * per;
                }
                lead = std::max(0.0, lead - ms_since(t_last));
                t_last = clk::now();
                lead += ctx->dma_ms_per_byte * bytes * (pack_mode == 1 ? 0.375 : 1.0);
                CK(cudaEventRecord(ctx->h2d_done[slot], b->stream));
                last_copy_slot = slot;
            }
        } else if ((rc = batch_fill(ctx, b, bases, offs, chunks[c].first, chunks[c].second, fixed_len))) return rc;
        if (mode == 0) {
            if ((rc = launch_count(ctx, b, ctx->read_ct, arg))) return rc;
        } else {
            if (mode == 1 && (rc = launch_classify(ctx, b, arg))) return rc;
            CK(cudaMemcpyAsync(out + chunks[c].first, b->hits.p, sizeof(Hit) * (size_t)b->nseq, cudaMemcpyDeviceToHost,
                               b->stream));
        }
    }
    for (int i = 0; i < ns; ++i) CK(cudaStreamSynchronize(ctx->slots[i]->stream));
    if (adaptive && getenv("RKMH_PACK_DEBUG"))
        fprintf(stderr, "rkmh: adaptive transfer: link %.3f ms, packer %.3f ms per 64 MB (ratio %.2f, packing %s)\n",
                ctx->dma_ms_per_byte * 67108864.0, ctx->pack_ms_per_byte * 67108864.0,
                ctx->dma_ms_per_byte > 0 ? ctx->pack_ms_per_byte / ctx->dma_ms_per_byte : 0.0, pack_pays(ctx) ? "on" : "off");
    return RKG_OK;
}

int rkg_count_batch(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nreads, int count_zeros) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->params_set) return fail(ctx, RKG_ERR_STATE, "rkg_count_batch before rkg_set_params");
    if (!ctx->read_ct.counts) return fail(ctx, RKG_ERR_STATE, "no read count table (rkg_set_params read_table_size)");
    if (nreads < 0 || !offs || (!bases && nreads > 0)) return fail(ctx, RKG_ERR_ARG, "rkg_count_batch: bad buffers");
    if (nreads == 0) return RKG_OK;
    CK(cudaSetDevice(ctx->device));
    return host_pipeline(ctx, bases, offs, nreads, 0, count_zeros, nullptr);
}

// ---- classification ---------------------------------------------------------------------------------
int rkg_classify_batch(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nreads, int min_occ, rkg_hit* out) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->refs_loaded || ctx->set_table) return fail(ctx, RKG_ERR_STATE, "rkg_classify_batch before rkg_load_refs");
    if (nreads < 0 || !offs || (!bases && nreads > 0) || (!out && nreads > 0))
        return fail(ctx, RKG_ERR_ARG, "rkg_classify_batch: bad buffers");
    if (nreads == 0) return RKG_OK;
    CK(cudaSetDevice(ctx->device));
    return host_pipeline(ctx, bases, offs, nreads, 1, min_occ, out);
}

int rkg_classify_fixed(rkg_ctx* ctx, const char* bases, int64_t read_len, int64_t nreads, int min_occ, rkg_hit* out) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->refs_loaded || ctx->set_table) return fail(ctx, RKG_ERR_STATE, "rkg_classify_fixed before rkg_load_refs");
    if (nreads < 0 || read_len < 0 || (!bases && nreads > 0) || (!out && nreads > 0))
        return fail(ctx, RKG_ERR_ARG, "rkg_classify_fixed: bad buffers");
    if (nreads == 0) return RKG_OK;
    CK(cudaSetDevice(ctx->device));
    return host_pipeline(ctx, bases, nullptr, nreads, 1, min_occ, out, read_len);
}

int rkg_count_fixed(rkg_ctx* ctx, const char* bases, int64_t read_len, int64_t nreads, int count_zeros) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->params_set) return fail(ctx, RKG_ERR_STATE, "rkg_count_fixed before rkg_set_params");
    if (!ctx->read_ct.counts) return fail(ctx, RKG_ERR_STATE, "no read count table (rkg_set_params read_table_size)");
    if (nreads < 0 || read_len < 0 || (!bases && nreads > 0)) return fail(ctx, RKG_ERR_ARG, "rkg_count_fixed: bad buffers");
    if (nreads == 0) return RKG_OK;
    CK(cudaSetDevice(ctx->device));
    return host_pipeline(ctx, bases, nullptr, nreads, 0, count_zeros, nullptr, read_len);
}

int rkg_copy_probe(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t read_len, int64_t nreads, rkg_hit* out) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->params_set) return fail(ctx, RKG_ERR_STATE, "rkg_copy_probe before rkg_set_params");
    if (nreads < 0 || (!bases && nreads > 0) || (!out && nreads > 0)) return fail(ctx, RKG_ERR_ARG, "rkg_copy_probe: bad buffers");
    if (nreads == 0) return RKG_OK;
    CK(cudaSetDevice(ctx->device));
    return host_pipeline(ctx, bases, offs, nreads, 2, 0, out, read_len);
}

int rkg_set_host_pack(rkg_ctx* ctx, int enable, int threads) {
    if (!ctx) return RKG_ERR_ARG;
    if (enable < 0 || enable > 2) return fail(ctx, RKG_ERR_ARG, "rkg_set_host_pack: mode %d (0 off, 1 every chunk, 2 adaptive)", enable);
    ctx->host_pack = enable;
    ctx->pack_threads = threads > 0 ? std::min(threads, 64) : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    ctx->pack_ms_per_byte = 0.0; ctx->dma_ms_per_byte = 0.0;                // measured again on the next chunks
    if (const char* e = getenv("RKMH_PACK_LEAD")) ctx->pack_lead = atof(e);
    if (const char* e = getenv("RKMH_PACK_MAX_RATIO")) ctx->pack_max_ratio = atof(e);
    return RKG_OK;
}

int rkg_transfer_stats(rkg_ctx* ctx, int64_t stats[4], int reset) {
    if (!ctx || !stats) return RKG_ERR_ARG;
    for (int i = 0; i < 4; ++i) { stats[i] = ctx->xfer[i]; if (reset) ctx->xfer[i] = 0; }
    return RKG_OK;
}

int rkg_batch_upload(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nreads, rkg_batch** batch) {
    if (!ctx || !batch) return RKG_ERR_ARG;
    if (!ctx->params_set) return fail(ctx, RKG_ERR_STATE, "rkg_batch_upload before rkg_set_params");
    if (nreads < 0 || !offs || (!bases && nreads > 0)) return fail(ctx, RKG_ERR_ARG, "rkg_batch_upload: bad buffers");
    CK(cudaSetDevice(ctx->device));
    rkg_batch* b = nullptr;
    int rc = batch_alloc(ctx, &b);
    if (rc) return rc;
    if ((rc = batch_fill(ctx, b, bases, offs, 0, nreads))) { batch_release(b); return rc; }
    CK(cudaStreamSynchronize(b->stream));
    *batch = b;
    return RKG_OK;
}

int rkg_batch_classify(rkg_ctx* ctx, rkg_batch* batch, int min_occ) {
    if (!ctx || !batch) return RKG_ERR_ARG;
    if (!ctx->refs_loaded || ctx->set_table) return fail(ctx, RKG_ERR_STATE, "rkg_batch_classify before rkg_load_refs");
    CK(cudaSetDevice(ctx->device));
    return launch_classify(ctx, batch, min_occ);
}

int rkg_batch_count(rkg_ctx* ctx, rkg_batch* batch, int count_zeros) {
    if (!ctx || !batch) return RKG_ERR_ARG;
    if (!ctx->read_ct.counts) return fail(ctx, RKG_ERR_STATE, "no read count table");
    CK(cudaSetDevice(ctx->device));
    return launch_count(ctx, batch, ctx->read_ct, count_zeros);
}

int rkg_batch_ingest(rkg_ctx* ctx, rkg_batch* batch) {
    if (!ctx || !batch) return RKG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->memo_ready) return RKG_OK;       // the memoised kernels read the ASCII bases themselves (fused ingest)
    return launch_ingest(ctx, batch);
}

int rkg_batch_set_stream(rkg_ctx* ctx, rkg_batch* batch, void* cuda_stream) {
    if (!ctx || !batch) return RKG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(batch->stream));
    batch->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : batch->own_stream;
    return RKG_OK;
}

int rkg_batch_results(rkg_ctx* ctx, rkg_batch* batch, rkg_hit* out) {
    if (!ctx || !batch || !out) return RKG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out, batch->hits.p, sizeof(Hit) * (size_t)batch->nseq, cudaMemcpyDeviceToHost, batch->stream));
    CK(cudaStreamSynchronize(batch->stream));
    return RKG_OK;
}

int rkg_batch_free(rkg_ctx* ctx, rkg_batch* batch) {
    if (!ctx) return RKG_ERR_ARG;
    cudaSetDevice(ctx->device);
    batch_release(batch);
    return RKG_OK;
}

// ---- hpv16: full hash sets instead of sketches ------------------------------------------------------------
int rkg_load_ref_sets(rkg_ctx* ctx, const uint64_t* hashes, const int64_t* set_offs, int nsets) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->params_set) return fail(ctx, RKG_ERR_STATE, "rkg_load_ref_sets before rkg_set_params");
    if (!hashes || !set_offs || nsets < 1) return fail(ctx, RKG_ERR_ARG, "rkg_load_ref_sets: need at least one set");
    const i64 nent = set_offs[nsets] - set_offs[0];
    if (nent < 0 || 2 * nent > ((i64)1 << 30)) return fail(ctx, RKG_ERR_ARG, "rkg_load_ref_sets: %lld elements (supported: 2^29)", (long long)nent);
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    ctx->refs_loaded = false; ctx->memo_ready = false;
    const u32 cap = (u32)next_pow2(std::max<i64>(2 * nent, 1024));
    const u32 nb = (u32)((nsets + 63) / 64);
    RefTable tb{};
    tb.cap_mask = cap - 1; tb.nb = nb; tb.stride = 1 + nb; tb.nrefs = (u32)nsets;
    int flog2 = 10;
    while (flog2 < 20 && (1LL << flog2) < 16 * nent) ++flog2;
    tb.filter_log2 = (u32)flog2;
    CK(ctx->tb_entries.ensure(sizeof(u64) * (size_t)cap * tb.stride));
    CK(ctx->tb_maxmult.ensure(sizeof(u32) * (size_t)cap));
    CK(ctx->tb_dupbase.ensure(sizeof(u32) * (size_t)cap));
    CK(ctx->tb_dupbits.ensure(64));
    CK(ctx->tb_filter.ensure(sizeof(u32) * ((size_t)1 << flog2) / 32));
    cudaStream_t st = ctx->stream;
    CK(cudaMemsetAsync(ctx->tb_entries.p, 0, sizeof(u64) * (size_t)cap * tb.stride, st));
    CK(cudaMemsetAsync(ctx->tb_maxmult.p, 0, sizeof(u32) * (size_t)cap, st));
    CK(cudaMemsetAsync(ctx->tb_dupbase.p, 0, sizeof(u32) * (size_t)cap, st));
    CK(cudaMemsetAsync(ctx->tb_filter.p, 0, sizeof(u32) * ((size_t)1 << flog2) / 32, st));
    tb.entries = ctx->tb_entries.as<u64>();
    tb.maxmult = ctx->tb_maxmult.as<u32>();
    tb.dup_base = ctx->tb_dupbase.as<u32>();
    tb.dup_bits = ctx->tb_dupbits.as<u64>();
    tb.filter = ctx->tb_filter.as<u32>();
    DevBuf d_h, d_o;
    int rc = RKG_OK;
    do {
        if (d_h.ensure(sizeof(u64) * (size_t)std::max<i64>(nent, 1)) || d_o.ensure(sizeof(i64) * (size_t)(nsets + 1))) {
            rc = fail(ctx, RKG_ERR_NOMEM, "rkg_load_ref_sets: device allocation failed");
            break;
        }
        std::vector<i64> rel((size_t)nsets + 1);
        for (int i = 0; i <= nsets; ++i) rel[i] = set_offs[i] - set_offs[0];
        cudaMemcpyAsync(d_h.p, hashes + set_offs[0], sizeof(u64) * (size_t)nent, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_o.p, rel.data(), sizeof(i64) * (size_t)(nsets + 1), cudaMemcpyHostToDevice, st);
        {
            Timed t(ctx, st, K_TABLE);
            const int grid = (int)std::min<i64>((nent + 255) / 256 + 1, (i64)ctx->sm_count * 8);
            table_insert_set_kernel<<<grid, 256, 0, st>>>(d_h.as<u64>(), d_o.as<i64>(), nsets, tb, ctx->tb_filter.as<u32>());
        }
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = fail(ctx, RKG_ERR_CUDA, "rkg_load_ref_sets: %s", cudaGetErrorString(cudaGetLastError())); break; }
    } while (0);
    d_h.release(); d_o.release();
    if (rc) return rc;
    ctx->tb = tb;
    ctx->nrefs = nsets;
    ctx->set_table = true;
    ctx->refs_loaded = true;
    return RKG_OK;
}

int rkg_classify_sets(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nreads, int min_occ, int ntypes,
                      rkg_hit* out, int32_t* extra) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->refs_loaded || !ctx->set_table) return fail(ctx, RKG_ERR_STATE, "rkg_classify_sets before rkg_load_ref_sets");
    if (ntypes < 1 || ntypes > ctx->nrefs) return fail(ctx, RKG_ERR_ARG, "rkg_classify_sets: ntypes %d of %d sets", ntypes, ctx->nrefs);
    const int nextra = ctx->nrefs - ntypes;
    if (nreads < 0 || !offs || (!bases && nreads > 0) || (!out && nreads > 0) || (nextra > 0 && !extra && nreads > 0))
        return fail(ctx, RKG_ERR_ARG, "rkg_classify_sets: bad buffers");
    if (min_occ >= 0 && ctx->read_ct.counts == nullptr)
        return fail(ctx, RKG_ERR_STATE, "min_occ >= 0 needs a read count table (rkg_set_params read_table_size)");
    if (nreads == 0) return RKG_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_slots(ctx);
    if (rc) return rc;
    { rc = memo_flush(ctx); if (rc) return rc; }
    const CountTable ct = (min_occ >= 0) ? ctx->read_ct : CountTable{nullptr, {}};
    DevBuf d_extra[2];
    auto chunks = make_chunks(offs, nreads, CHUNK_BYTES, CHUNK_READS);
    for (size_t c = 0; c < chunks.size() && rc == RKG_OK; ++c) {
        rkg_batch* b = ctx->slots[c & 1];
        DevBuf& dx = d_extra[c & 1];
        if (cudaStreamSynchronize(b->stream) != cudaSuccess) { rc = fail(ctx, RKG_ERR_CUDA, "rkg_classify_sets: %s", cudaGetErrorString(cudaGetLastError())); break; }
        if ((rc = batch_fill(ctx, b, bases, offs, chunks[c].first, chunks[c].second))) break;
        if ((rc = ensure_packed(ctx, b))) break;
        if (dx.ensure(sizeof(int) * (size_t)std::max<i64>(b->nseq * std::max(nextra, 1), 1)) != cudaSuccess) { rc = fail(ctx, RKG_ERR_NOMEM, "rkg_classify_sets: device allocation failed"); break; }
        // every hash of the read takes part (no bottom-s): the exact block kernel with an unbounded sketch size
        const int s_all = 1 << 30;
        const int block = 512;
        const int hs_cap = next_pow2(2 * std::max<i64>(b->max_ntotal, 1));
        if (hs_cap == 0) { rc = fail(ctx, RKG_ERR_ARG, "rkg_classify_sets: a read of more than 2^29 k-mers"); break; }
        const int rpad = (ctx->nrefs + 63) / 64 * 64;
        const size_t budget = ctx->smem_optin / 2 - 1024 - sizeof(u64) * HASH_LUT_ENTRIES;
        size_t fixed = (size_t)rpad * 4;
        const int hs_in_smem = ((size_t)hs_cap * 8 + fixed <= budget / 2) ? 1 : 0;
        if (hs_in_smem) fixed += (size_t)hs_cap * 8;
        i64 hbuf_cap = 0;
        if (fixed + (size_t)b->max_ntotal * 8 <= budget) hbuf_cap = b->max_ntotal;
        const size_t smem = fixed + (size_t)hbuf_cap * 8;
        const i64 stride = (hbuf_cap >= b->max_ntotal) ? 0 : b->max_ntotal;
        const int grid = ctx->sm_count * 2;
        if (b->scratch.ensure(sizeof(u64) * (size_t)grid * (size_t)(stride + hs_cap) + 64) != cudaSuccess) { rc = fail(ctx, RKG_ERR_NOMEM, "rkg_classify_sets: device allocation failed"); break; }
        {
            Timed t(ctx, b->stream, K_CLASSIFY_BLOCK);
            DISPATCH_KCSEL(ctx->kcsel, {
                cudaFuncSetAttribute(classify_block_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                classify_block_kernel<KS><<<grid, block, smem, b->stream>>>(seq_view(b), ctx->kl, s_all, ct, min_occ, ctx->tb, nullptr, nullptr,
                                                                             b->nseq, (int)hbuf_cap, hs_cap, hs_in_smem, rpad,
                                                                             b->scratch.as<u64>(), stride, b->hits.as<Hit>(), 1, ntypes, dx.as<int>());
            });
        }
        if (cudaGetLastError() != cudaSuccess) { rc = fail(ctx, RKG_ERR_CUDA, "rkg_classify_sets: launch failed"); break; }
        cudaMemcpyAsync(out + chunks[c].first, b->hits.p, sizeof(Hit) * (size_t)b->nseq, cudaMemcpyDeviceToHost, b->stream);
        if (nextra > 0)
            cudaMemcpyAsync(extra + (size_t)chunks[c].first * nextra, dx.p, sizeof(int) * (size_t)b->nseq * nextra, cudaMemcpyDeviceToHost, b->stream);
    }
    for (int i = 0; i < 2; ++i)
        if (cudaStreamSynchronize(ctx->slots[i]->stream) != cudaSuccess && rc == RKG_OK)
            rc = fail(ctx, RKG_ERR_CUDA, "rkg_classify_sets: %s", cudaGetErrorString(cudaGetLastError()));
    d_extra[0].release(); d_extra[1].release();
    return rc;
}

// ---- call: exact k-mer depth ------------------------------------------------------------------------------
int rkg_depth_begin(rkg_ctx* ctx, uint64_t expected_windows) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->params_set) return fail(ctx, RKG_ERR_STATE, "rkg_depth_begin before rkg_set_params");
    if (ctx->kl.nk != 1) return fail(ctx, RKG_ERR_ARG, "rkg_depth_begin: exactly one k-mer size (src/rkmh.cpp:1545-1554)");
    if (expected_windows > ((uint64_t)1 << 29)) return fail(ctx, RKG_ERR_ARG, "rkg_depth_begin: more than 2^29 windows");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    const u32 cap = (u32)next_pow2(std::max<i64>(2 * (i64)expected_windows, 1024));
    CK(ctx->depth_keys.ensure(sizeof(u64) * (size_t)cap));
    CK(ctx->depth_counts.ensure(sizeof(u32) * (size_t)cap));
    CK(ctx->depth_zero.ensure(64));
    CK(cudaMemsetAsync(ctx->depth_keys.p, 0, sizeof(u64) * (size_t)cap, ctx->stream));
    CK(cudaMemsetAsync(ctx->depth_counts.p, 0, sizeof(u32) * (size_t)cap, ctx->stream));
    CK(cudaMemsetAsync(ctx->depth_zero.p, 0, 64, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->depth.keys = ctx->depth_keys.as<u64>(); ctx->depth.counts = ctx->depth_counts.as<u32>();
    ctx->depth.zero = ctx->depth_zero.as<unsigned long long>(); ctx->depth.cap_mask = cap - 1;
    ctx->depth_room = (i64)expected_windows;
    return RKG_OK;
}

int rkg_depth_add(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nreads) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->depth.keys) return fail(ctx, RKG_ERR_STATE, "rkg_depth_add before rkg_depth_begin");
    if (nreads < 0 || !offs || (!bases && nreads > 0)) return fail(ctx, RKG_ERR_ARG, "rkg_depth_add: bad buffers");
    if (nreads == 0) return RKG_OK;
    const int k = ctx->kl.k[0];
    i64 windows = 0;
    for (int64_t r = 0; r < nreads; ++r) windows += std::max<i64>(offs[r + 1] - offs[r] - k, 0);
    if (windows > ctx->depth_room) return fail(ctx, RKG_ERR_ARG, "rkg_depth_add: more windows than rkg_depth_begin was told to expect");
    ctx->depth_room -= windows;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_slots(ctx);
    if (rc) return rc;
    auto chunks = make_chunks(offs, nreads, CHUNK_BYTES, CHUNK_READS);
    for (size_t c = 0; c < chunks.size(); ++c) {
        rkg_batch* b = ctx->slots[c & 1];
        CK(cudaStreamSynchronize(b->stream));
        if ((rc = batch_fill(ctx, b, bases, offs, chunks[c].first, chunks[c].second))) return rc;
        if ((rc = ensure_packed(ctx, b))) return rc;
        Timed t(ctx, b->stream, K_COUNT);
        const int grid = (int)std::min<i64>((b->nseq + 7) / 8, (i64)ctx->sm_count * 8);
        DISPATCH_KCSEL(ctx->kcsel, (depth_add_kernel<KS><<<grid, 256, 0, b->stream>>>(seq_view(b), ctx->kl, ctx->depth)));
        CK(cudaGetLastError());
    }
    for (int i = 0; i < 2; ++i) CK(cudaStreamSynchronize(ctx->slots[i]->stream));
    return RKG_OK;
}

int rkg_depth_of_sequence(rkg_ctx* ctx, const char* seq, int64_t len, int32_t* depth) {
    if (!ctx || !seq || !depth) return RKG_ERR_ARG;
    if (!ctx->depth.keys) return fail(ctx, RKG_ERR_STATE, "rkg_depth_of_sequence before rkg_depth_begin");
    const int k = ctx->kl.k[0];
    const i64 n = std::max<i64>(len - k, 0);
    if (n == 0) return RKG_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_slots(ctx);
    if (rc) return rc;
    rkg_batch* b = ctx->slots[0];
    CK(cudaStreamSynchronize(b->stream));
    const int64_t one[2] = {0, len};
    if ((rc = batch_fill(ctx, b, seq, one, 0, 1))) return rc;
    if ((rc = ensure_packed(ctx, b))) return rc;
    DevBuf d;
    if (d.ensure(sizeof(int) * (size_t)n) != cudaSuccess) return fail(ctx, RKG_ERR_NOMEM, "rkg_depth_of_sequence: device allocation failed");
    {
        Timed t(ctx, b->stream, K_HASH);
        const int grid = (int)std::min<i64>((n + WARP_CHUNK * 8 - 1) / (WARP_CHUNK * 8), (i64)ctx->sm_count * 8);
        DISPATCH_KCSEL(ctx->kcsel, (depth_lookup_kernel<KS><<<grid, 256, 0, b->stream>>>(seq_view(b), ctx->kl, ctx->depth, n, d.as<int>())));
    }
    cudaMemcpyAsync(depth, d.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, b->stream);
    const cudaError_t e = cudaStreamSynchronize(b->stream);
    d.release();
    CK(e);
    return RKG_OK;
}

int rkg_depth_rescue(rkg_ctx* ctx, const char* seq, int64_t len, const int64_t* windows, int64_t nwin, int32_t* out) {
    if (!ctx || !seq || (nwin > 0 && (!windows || !out))) return RKG_ERR_ARG;
    if (!ctx->depth.keys) return fail(ctx, RKG_ERR_STATE, "rkg_depth_rescue before rkg_depth_begin");
    if (nwin <= 0) return RKG_OK;
    const int k = ctx->kl.k[0];
    for (int64_t i = 0; i < nwin; ++i)
        if (windows[i] < 0 || windows[i] + k > len) return fail(ctx, RKG_ERR_ARG, "rkg_depth_rescue: window %lld outside the sequence", (long long)windows[i]);
    CK(cudaSetDevice(ctx->device));
    DevBuf d_seq, d_win, d_out;
    int rc = RKG_OK;
    do {
        const size_t nout = (size_t)nwin * 4 * (size_t)k;
        if (d_seq.ensure((size_t)len + 64) || d_win.ensure(sizeof(i64) * (size_t)nwin) || d_out.ensure(sizeof(int) * nout)) {
            rc = fail(ctx, RKG_ERR_NOMEM, "rkg_depth_rescue: device allocation failed"); break;
        }
        cudaMemcpyAsync(d_seq.p, seq, (size_t)len, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d_win.p, windows, sizeof(i64) * (size_t)nwin, cudaMemcpyHostToDevice, ctx->stream);
        {
            Timed t(ctx, ctx->stream, K_HASH);
            const int grid = (int)std::min<i64>(((i64)nout + 255) / 256, (i64)ctx->sm_count * 8);
            DISPATCH_KCSEL(ctx->kl.kv[0], (depth_rescue_kernel<(KS < NUM_KV ? KS : 0)><<<grid, 256, 0, ctx->stream>>>(
                                               d_seq.as<unsigned char>(), len, ctx->kl, ctx->depth, d_win.as<i64>(), nwin, d_out.as<int>())));
        }
        cudaMemcpyAsync(out, d_out.p, sizeof(int) * nout, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = fail(ctx, RKG_ERR_CUDA, "rkg_depth_rescue: %s", cudaGetErrorString(cudaGetLastError())); break; }
    } while (0);
    d_seq.release(); d_win.release(); d_out.release();
    return rc;
}

// ---- hash command -----------------------------------------------------------------------------------
int rkg_sketch_batch(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nseq, int min_occ,
                     uint64_t* sketches, int32_t* lens) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->params_set) return fail(ctx, RKG_ERR_STATE, "rkg_sketch_batch before rkg_set_params");
    if (nseq < 0 || !offs || (!bases && nseq > 0) || !sketches || !lens) return fail(ctx, RKG_ERR_ARG, "rkg_sketch_batch: bad buffers");
    if (min_occ >= 0 && !ctx->read_ct.counts) return fail(ctx, RKG_ERR_STATE, "min_occ >= 0 needs a read count table");
    if (nseq == 0) return RKG_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_slots(ctx);
    if (rc) return rc;
    rkg_batch* b = ctx->slots[0];
    DevBuf d_sk, d_len;
    auto chunks = make_chunks(offs, nseq, CHUNK_BYTES, 1 << 16);
    for (auto& ch : chunks) {
        CK(cudaStreamSynchronize(b->stream));
        if ((rc = batch_fill(ctx, b, bases, offs, ch.first, ch.second))) break;
        const size_t n = (size_t)(ch.second - ch.first);
        if (d_sk.ensure(sizeof(u64) * n * ctx->s) != cudaSuccess || d_len.ensure(sizeof(int) * n) != cudaSuccess) {
            rc = fail(ctx, RKG_ERR_NOMEM, "rkg_sketch_batch: device allocation failed");
            break;
        }
        cudaMemsetAsync(d_sk.p, 0, sizeof(u64) * n * ctx->s, b->stream);
        CountTable ct = (min_occ >= 0) ? ctx->read_ct : CountTable{nullptr, {}};
        if ((rc = launch_sketch(ctx, b, ct, min_occ >= 0 ? 2 : 0, min_occ, d_sk.as<u64>(), d_len.as<int>(), nullptr, nullptr))) break;
        cudaMemcpyAsync(sketches + (size_t)ch.first * ctx->s, d_sk.p, sizeof(u64) * n * ctx->s, cudaMemcpyDeviceToHost, b->stream);
        cudaMemcpyAsync(lens + ch.first, d_len.p, sizeof(int) * n, cudaMemcpyDeviceToHost, b->stream);
        if (cudaStreamSynchronize(b->stream) != cudaSuccess) { rc = fail(ctx, RKG_ERR_CUDA, "rkg_sketch_batch: %s", cudaGetErrorString(cudaGetLastError())); break; }
    }
    cudaStreamSynchronize(b->stream);
    d_sk.release(); d_len.release();
    return rc;
}

int rkg_hash_batch(rkg_ctx* ctx, const char* bases, const int64_t* offs, int64_t nseq, uint64_t* hashes, int64_t* hash_offs) {
    if (!ctx) return RKG_ERR_ARG;
    if (!ctx->params_set) return fail(ctx, RKG_ERR_STATE, "rkg_hash_batch before rkg_set_params");
    if (nseq < 0 || !offs || (!bases && nseq > 0) || !hashes || !hash_offs) return fail(ctx, RKG_ERR_ARG, "rkg_hash_batch: bad buffers");
    hash_offs[0] = 0;
    for (int64_t r = 0; r < nseq; ++r) {
        i64 len = offs[r + 1] - offs[r], n = 0;
        for (int j = 0; j < ctx->kl.nk; ++j) n += std::max<i64>(len - ctx->kl.k[j], 0);
        hash_offs[r + 1] = hash_offs[r] + n;
    }
    if (nseq == 0) return RKG_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_slots(ctx);
    if (rc) return rc;
    rkg_batch* b = ctx->slots[0];
    DevBuf d_h, d_ho;
    auto chunks = make_chunks(offs, nseq, 64LL << 20, CHUNK_READS);
    for (auto& ch : chunks) {
        CK(cudaStreamSynchronize(b->stream));
        if ((rc = batch_fill(ctx, b, bases, offs, ch.first, ch.second))) break;
        const i64 n = ch.second - ch.first;
        const i64 h0 = hash_offs[ch.first], nh = hash_offs[ch.second] - h0;
        if (d_h.ensure(sizeof(u64) * (size_t)std::max<i64>(nh, 1)) != cudaSuccess || d_ho.ensure(sizeof(i64) * (size_t)(n + 1)) != cudaSuccess) {
            rc = fail(ctx, RKG_ERR_NOMEM, "rkg_hash_batch: device allocation failed");
            break;
        }
        std::vector<i64> rel(n + 1);
        for (i64 i = 0; i <= n; ++i) rel[i] = hash_offs[ch.first + i] - h0;
        cudaMemcpyAsync(d_ho.p, rel.data(), sizeof(i64) * (size_t)(n + 1), cudaMemcpyHostToDevice, b->stream);
        if ((rc = ensure_packed(ctx, b))) break;
        for (int kidx = 0; kidx < ctx->kl.nk; ++kidx) {
            Timed t(ctx, b->stream, K_HASH);
            int grid = (int)std::min<i64>((n + 7) / 8, (i64)ctx->sm_count * 8);
            DISPATCH_KCSEL(ctx->kl.kv[kidx], (hash_kernel<(KS < NUM_KV ? KS : 0)><<<grid, 256, 0, b->stream>>>(
                                                 seq_view(b), ctx->kl, kidx, d_h.as<u64>(), d_ho.as<i64>())));
        }
        cudaMemcpyAsync(hashes + h0, d_h.p, sizeof(u64) * (size_t)nh, cudaMemcpyDeviceToHost, b->stream);
        if (cudaStreamSynchronize(b->stream) != cudaSuccess) { rc = fail(ctx, RKG_ERR_CUDA, "rkg_hash_batch: %s", cudaGetErrorString(cudaGetLastError())); break; }
    }
    cudaStreamSynchronize(b->stream);
    d_h.release(); d_ho.release();
    return rc;
}

int rkg_merge_sketches(rkg_ctx* ctx, const uint64_t* parts, const int32_t* part_lens, int nparts, uint64_t* sketch, int32_t* len) {
    if (!ctx || !parts || !part_lens || nparts < 1 || !sketch || !len) return RKG_ERR_ARG;
    if (!ctx->params_set) return fail(ctx, RKG_ERR_STATE, "rkg_merge_sketches before rkg_set_params");
    CK(cudaSetDevice(ctx->device));
    const int s = ctx->s, P = next_pow2(std::max(s, 2));
    DevBuf d_parts, d_lens, d_out, d_len, d_scr;
    int rc = RKG_OK;
    do {
        if (d_parts.ensure(sizeof(u64) * (size_t)nparts * s) || d_lens.ensure(sizeof(int) * (size_t)nparts) ||
            d_out.ensure(sizeof(u64) * (size_t)s) || d_len.ensure(sizeof(int)) || d_scr.ensure(sizeof(u64) * (size_t)P)) {
            rc = fail(ctx, RKG_ERR_NOMEM, "rkg_merge_sketches: device allocation failed");
            break;
        }
        cudaMemcpyAsync(d_parts.p, parts, sizeof(u64) * (size_t)nparts * s, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d_lens.p, part_lens, sizeof(int) * (size_t)nparts, cudaMemcpyHostToDevice, ctx->stream);
        const int in_smem = ((size_t)P * 8 <= ctx->smem_optin - 1024) ? 1 : 0;
        const size_t smem = in_smem ? (size_t)P * 8 : 0;
        {
            Timed t(ctx, ctx->stream, K_MERGE);
            cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            merge_kernel<<<1, 512, smem, ctx->stream>>>(d_parts.as<u64>(), d_lens.as<int>(), nparts, s, P, d_out.as<u64>(),
                                                        d_len.as<int>(), d_scr.as<u64>(), in_smem);
        }
        cudaMemcpyAsync(len, d_len.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = fail(ctx, RKG_ERR_CUDA, "merge: %s", cudaGetErrorString(cudaGetLastError())); break; }
        cudaMemcpy(sketch, d_out.p, sizeof(u64) * (size_t)(*len), cudaMemcpyDeviceToHost);
    } while (0);
    d_parts.release(); d_lens.release(); d_out.release(); d_len.release(); d_scr.release();
    return rc;
}

static void launch_shard_sort(rkg_ctx* ctx, cudaStream_t st, u64* a, int m, int P) {
    const size_t bytes = (size_t)P * 8;
    const int in_smem = bytes + 1024 <= ctx->smem_optin ? 1 : 0;
    if (in_smem) cudaFuncSetAttribute(shard_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    shard_sort_kernel<<<1, 1024, in_smem ? bytes : 0, st>>>(a, m, P, in_smem);
}

int rkg_sketch_shard(rkg_ctx* ctx, const char* bases, int64_t nbytes, int64_t kmer_begin, int64_t kmer_end,
                     int64_t record_len, uint64_t* sketch, int32_t* len) {
    if (!ctx || !bases || !sketch || !len) return RKG_ERR_ARG;
    if (!ctx->params_set) return fail(ctx, RKG_ERR_STATE, "rkg_sketch_shard before rkg_set_params");
    if (ctx->kl.nk != 1) return fail(ctx, RKG_ERR_ARG, "rkg_sketch_shard: exactly one k-mer size");
    // The shard is ONE device-resident record: its windows are hashed twice (histogram of the top 16 hash bits, then
    // the hashes at or below the bin that holds the s-th smallest are kept and sorted).  Nothing of size n is ever
    // materialised and nothing but the sketch comes back to the host.
    const int k = ctx->kl.k[0];
    const i64 last_start = record_len - k;            // k-mer starts [0, record_len - k) exist
    kmer_end = std::min<i64>(kmer_end, last_start);
    *len = 0;
    if (kmer_end <= kmer_begin) return RKG_OK;
    const i64 n = kmer_end - kmer_begin;               // windows of this shard
    if (n + k > nbytes) return fail(ctx, RKG_ERR_ARG, "rkg_sketch_shard: shard buffer too short");
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_slots(ctx);
    if (rc) return rc;
    rkg_batch* b = ctx->slots[0];
    CK(cudaStreamSynchronize(b->stream));
    // The shard arrives in 32 MB pieces on a copy stream; behind each piece the compute stream packs it (ingest_kernel) and
    // hashes the windows whose k bases have arrived (shard_scan_kernel), so that copy and hashing overlap.
    const i64 total = n + k;
    const i64 nwords32 = (total + 31) / 32;
    b->nseq = 1; b->total = total; b->base0 = 0; b->max_len = total; b->max_ntotal = n;
    CK(b->bases.ensure((size_t)nwords32 * 32 + 64));
    CK(b->codes.ensure(sizeof(u32) * (size_t)(2 * nwords32 + 8)));
    CK(b->nmask.ensure(sizeof(u32) * (size_t)(nwords32 + 8)));
    CK(b->misc.ensure(256));
    if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->copy_event) CK(cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming));
    CK(cudaMemsetAsync(b->codes.as<u32>() + 2 * nwords32, 0, sizeof(u32) * 8, b->stream));
    CK(cudaMemsetAsync(b->nmask.as<u32>() + nwords32, 0xFF, sizeof(u32) * 8, b->stream));
    const int grid = ctx->sm_count * 8;
    const i64 lastw = 2 * nwords32 + 7, lastm = nwords32 + 7;               // 8 words of padding follow the record
    // the guess: 1.5 x the s-th smallest of n uniform hashes (s = 10 000: 15 000 +- 120 kept, a 16 k-entry sort in shared
    // memory); room for four times s kept hashes
    u64 limit = ~0ULL;
    if ((i64)ctx->s < n) {
        const long double frac = 1.5L * (long double)ctx->s / (long double)n;
        if (frac < 1.0L) limit = (u64)(frac * 18446744073709551615.0L);
    }
    const u64 cap = std::max<u64>(4ULL * (u64)ctx->s, 1ULL << 16);
    DevBuf& d_scan = ctx->shard_scan;                                        // kept between calls (a cudaFree synchronises the device)
    DevBuf& d_cnt = ctx->shard_cnt;
    if (d_scan.ensure(sizeof(u64) * (size_t)next_pow2((i64)cap)) || d_cnt.ensure(64))
        return fail(ctx, RKG_ERR_NOMEM, "rkg_sketch_shard: device allocation failed");
    CK(cudaMemsetAsync(d_cnt.p, 0, 64, b->stream));
    CK(cudaEventRecord(ctx->copy_event, b->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_event, 0));          // the buffers are free (and sized) from here on
    const i64 piece = 32LL << 20;
    i64 w_lo = 0;
    // rkg_set_host_pack(2): a piece is packed on the host while the link is busy with earlier pieces (the policy of
    // host_pipeline; the two rates are measured on the context's first pieces if they are not known yet)
    const bool adaptive = ctx->host_pack == 2;
    using clk = std::chrono::steady_clock;
    auto ms_since = [](clk::time_point t) { return std::chrono::duration<double, std::milli>(clk::now() - t).count(); };
    double lead = 0.0;
    clk::time_point t_last = clk::now();
    int last_stage = -1, npiece = 0;
    if (adaptive)
        for (int i = 0; i < NSLOTS; ++i) {
            if (!ctx->h2d_done[i]) CK(cudaEventCreateWithFlags(&ctx->h2d_done[i], cudaEventDisableTiming));
            if ((rc = ensure_stage(ctx, i, (size_t)(piece * 3 / 8) + (64 << 10)))) return rc;
        }
    for (i64 off = 0; off < total; off += piece, ++npiece) {
        const i64 plen = std::min<i64>(piece, total - off);
        const i64 pw = (plen + 31) / 32;
        bool pack = false, time_link = false;
        if (adaptive) {
            lead = std::max(0.0, lead - ms_since(t_last));
            t_last = clk::now();
            if (last_stage >= 0) {
                const bool done = cudaEventQuery(ctx->h2d_done[last_stage]) == cudaSuccess;
                if (done && lead > 0.25 * ctx->dma_ms_per_byte * (double)plen) ctx->dma_ms_per_byte *= 0.93;
                if (!done && lead == 0.0) { ctx->dma_ms_per_byte *= 1.07; lead = 0.1 * ctx->dma_ms_per_byte * (double)plen; }
                if (done) lead = 0.0;
            }
            if (ctx->dma_ms_per_byte == 0.0) time_link = true;
            else if (ctx->pack_ms_per_byte == 0.0) pack = true;
            else pack = pack_pays(ctx) && lead >= ctx->pack_lead * ctx->pack_ms_per_byte * (double)plen;
        }
        if (pack) {
            const int st = npiece % NSLOTS;
            CK(cudaEventSynchronize(ctx->h2d_done[st]));     // the staging buffer's previous copy has left it
            const size_t mask_at = ((size_t)2 * pw + 7) & ~(size_t)7;
            u32* hc = static_cast<u32*>(ctx->stage[st]);
            u32* hm = hc + mask_at;
            const clk::time_point t0 = clk::now();
            rk_pack_bases(bases + off, plen, hc, hm, ctx->pack_threads);
            const double per = ms_since(t0) / (double)plen;
            ctx->pack_ms_per_byte = ctx->pack_ms_per_byte == 0.0 ? per : 0.5 * ctx->pack_ms_per_byte + 0.5 * per;
            CK(cudaMemcpyAsync(b->codes.as<u32>() + 2 * (off / 32), hc, sizeof(u32) * 2 * (size_t)pw, cudaMemcpyHostToDevice, ctx->copy_stream));
            CK(cudaMemcpyAsync(b->nmask.as<u32>() + off / 32, hm, sizeof(u32) * (size_t)pw, cudaMemcpyHostToDevice, ctx->copy_stream));
            CK(cudaEventRecord(ctx->h2d_done[st], ctx->copy_stream));
            last_stage = st;
            ctx->xfer[1] += 12 * pw; ++ctx->xfer[3];
        } else {
            clk::time_point t0;
            if (time_link) { CK(cudaStreamSynchronize(ctx->copy_stream)); t0 = clk::now(); }
            CK(cudaMemcpyAsync(b->bases.as<unsigned char>() + off, bases + off, (size_t)plen, cudaMemcpyHostToDevice, ctx->copy_stream));
            if (time_link) {
                CK(cudaStreamSynchronize(ctx->copy_stream));
                ctx->dma_ms_per_byte = std::max(ms_since(t0), 1e-3) / (double)plen;
            }
            if (adaptive) {
                const int st = npiece % NSLOTS;
                CK(cudaEventSynchronize(ctx->h2d_done[st]));
                CK(cudaEventRecord(ctx->h2d_done[st], ctx->copy_stream));
                last_stage = st;
            }
            ctx->xfer[0] += plen; ++ctx->xfer[2];
        }
        if (adaptive) {
            lead = std::max(0.0, lead - ms_since(t_last));
            t_last = clk::now();
            if (!time_link) lead += ctx->dma_ms_per_byte * (double)plen * (pack ? 0.375 : 1.0);
        }
        CK(cudaEventRecord(ctx->copy_event, ctx->copy_stream));
        CK(cudaStreamWaitEvent(b->stream, ctx->copy_event, 0));
        if (!pack) {
            Timed t(ctx, b->stream, K_INGEST);
            ingest_kernel<<<(int)std::min<i64>((pw + 255) / 256, (i64)ctx->sm_count * 16), 256, 0, b->stream>>>(
                b->bases.as<unsigned char>() + off, plen, b->codes.as<u32>() + 2 * (off / 32), b->nmask.as<u32>() + off / 32, pw);
            ++ctx->launches;
        }
        const i64 w_hi = std::min<i64>(n, off + plen - k + 1);
        if (w_hi > w_lo) {
            Timed t(ctx, b->stream, K_SKETCH);
            DISPATCH_KCSEL(ctx->kcsel, (shard_scan_kernel<KS><<<grid, 256, 0, b->stream>>>(b->codes.as<u32>(), b->nmask.as<u32>(), lastw, lastm, w_lo, w_hi,
                                                                                         ctx->kl, limit, d_scan.as<u64>(),
                                                                                         d_cnt.as<unsigned long long>(), cap)));
            ++ctx->launches;
            w_lo = w_hi;
        }
    }
    b->packed = true; b->ascii = false;             // (some pieces may have arrived packed: the bytes are not all there)
    {
        unsigned long long hc[2] = {0, 0};
        cudaMemcpyAsync(hc, d_cnt.p, sizeof(hc), cudaMemcpyDeviceToHost, b->stream);
        if (cudaStreamSynchronize(b->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
            return fail(ctx, RKG_ERR_CUDA, "rkg_sketch_shard: %s", cudaGetErrorString(cudaGetLastError()));
        const u64 kept = hc[0], nonzero = hc[1];
        if (kept <= cap && (kept >= (u64)ctx->s || kept == nonzero) && !ctx->shard_two_pass) {
            // the guess held: the kept hashes contain the s smallest non-zero ones (or all of them)
            rc = RKG_OK;
            if (kept > 0) {
                const int P = next_pow2((i64)std::max<u64>(kept, 2));
                {
                    Timed t(ctx, b->stream, K_SKETCH);
                    launch_shard_sort(ctx, b->stream, d_scan.as<u64>(), (int)kept, P);
                    ++ctx->launches;
                }
                const int out_n = (int)std::min<u64>(kept, (u64)ctx->s);
                cudaMemcpyAsync(sketch, d_scan.p, sizeof(u64) * (size_t)out_n, cudaMemcpyDeviceToHost, b->stream);
                if (cudaStreamSynchronize(b->stream) != cudaSuccess) rc = fail(ctx, RKG_ERR_CUDA, "rkg_sketch_shard: %s", cudaGetErrorString(cudaGetLastError()));
                else *len = out_n;
            }
            return rc;
        }
    }
    // the exact two-pass route on the codes that are on the device now
    DevBuf d_hist, d_pick, d_cand;
    rc = RKG_OK;
    do {
        if (d_hist.ensure(sizeof(u32) * SHARD_BINS) || d_pick.ensure(64)) { rc = fail(ctx, RKG_ERR_NOMEM, "rkg_sketch_shard: device allocation failed"); break; }
        cudaMemsetAsync(d_hist.p, 0, sizeof(u32) * SHARD_BINS, b->stream);
        cudaMemsetAsync(d_pick.p, 0, 64, b->stream);
        u64 hpick[2] = {0, 0};
        {
            Timed t(ctx, b->stream, K_SKETCH);
            DISPATCH_KCSEL(ctx->kcsel, (shard_hist_kernel<KS><<<grid, 256, 0, b->stream>>>(b->codes.as<u32>(), b->nmask.as<u32>(), lastw, lastm, n, ctx->kl,
                                                                                         d_hist.as<u32>())));
            shard_pick_kernel<<<1, 1024, 0, b->stream>>>(d_hist.as<u32>(), (u64)ctx->s, d_pick.as<u64>());
            ctx->launches += 2;
        }
        cudaMemcpyAsync(hpick, d_pick.p, sizeof(hpick), cudaMemcpyDeviceToHost, b->stream);
        if (cudaStreamSynchronize(b->stream) != cudaSuccess) { rc = fail(ctx, RKG_ERR_CUDA, "rkg_sketch_shard: %s", cudaGetErrorString(cudaGetLastError())); break; }
        const u64 kept = hpick[1];
        if (kept == 0) break;                          // no valid window
        if (kept > (1ULL << 30)) { rc = fail(ctx, RKG_ERR_ARG, "rkg_sketch_shard: more than 2^30 hashes share the boundary bin"); break; }
        const int P = next_pow2((i64)std::max<u64>(kept, 2));
        if (d_cand.ensure(sizeof(u64) * (size_t)P)) { rc = fail(ctx, RKG_ERR_NOMEM, "rkg_sketch_shard: device allocation failed"); break; }
        unsigned long long* ncand = d_pick.as<unsigned long long>() + 4;
        {
            Timed t(ctx, b->stream, K_SKETCH);
            DISPATCH_KCSEL(ctx->kcsel, (shard_emit_kernel<KS><<<grid, 256, 0, b->stream>>>(b->codes.as<u32>(), b->nmask.as<u32>(), lastw, lastm, n, ctx->kl,
                                                                                         d_pick.as<u64>(), d_cand.as<u64>(), ncand, kept)));
            launch_shard_sort(ctx, b->stream, d_cand.as<u64>(), (int)kept, P);
            ctx->launches += 2;
        }
        const int out_n = (int)std::min<u64>(kept, (u64)ctx->s);
        cudaMemcpyAsync(sketch, d_cand.p, sizeof(u64) * (size_t)out_n, cudaMemcpyDeviceToHost, b->stream);
        if (cudaStreamSynchronize(b->stream) != cudaSuccess) { rc = fail(ctx, RKG_ERR_CUDA, "rkg_sketch_shard: %s", cudaGetErrorString(cudaGetLastError())); break; }
        *len = out_n;
    } while (0);
    d_hist.release(); d_pick.release(); d_cand.release();
    return rc;
}

int rkg_set_memo(rkg_ctx* ctx, int enable) {
    if (!ctx) return RKG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    { int rc = memo_flush(ctx); if (rc) return rc; }
    ctx->memo_enabled = enable != 0;
    if (!enable) ctx->memo_ready = false;
    return RKG_OK;
}

int rkg_memo_active(const rkg_ctx* ctx) { return ctx && ctx->memo_ready ? 1 : 0; }

// ---- measurement ------------------------------------------------------------------------------------
int64_t rkg_launch_count(const rkg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int rkg_profile_enable(rkg_ctx* ctx, int on) {
    if (!ctx) return RKG_ERR_ARG;
    cudaSetDevice(ctx->device);
    profile_collect(ctx);
    ctx->profiling = on != 0;
    return RKG_OK;
}

int rkg_profile_read(rkg_ctx* ctx, rkg_kernel_time* out, int cap, int reset) {
    if (!ctx) return RKG_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    profile_collect(ctx);
    int n = 0;
    for (int i = 0; i < K_NUM; ++i) {
        if (ctx->prof_n[i] == 0) continue;
        if (out && n < cap) { out[n].name = kKernelNames[i]; out[n].launches = ctx->prof_n[i]; out[n].ms = ctx->prof_ms[i]; }
        ++n;
    }
    if (reset) for (int i = 0; i < K_NUM; ++i) { ctx->prof_ms[i] = 0; ctx->prof_n[i] = 0; }
    return n;
}

}  // extern "C"
