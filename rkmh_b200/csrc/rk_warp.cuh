// rk_warp.cuh -- the short-read classification kernel of the Murmur path, built on the warp-cooperative hashing of
// rk_hash.cuh (every k and every k list; the memoised kernels of rk_memo.cuh take over for a single k <= 12).
#pragma once
#include "rk_kernels.cuh"
#include "rk_memo.cuh"

namespace rk {

// ------------------------------------------------------------------------------------------------
// classify_warp_kernel: one warp per read whose sketch is "every non-zero hash" (number of k-mers <= s and
// <= WARP_MAX_KMERS), at most 256 references (NW = 32-bit words of a reference bitset row: 2 or 8).  The Murmur
// counterpart of classify_memo_sieve_kernel, for every k and every k list:
//   1. warp_hash gives each lane the canonical hash of its window (mask_by_frequency, mkmh.hpp:898-908, applied when a
//      count table is given); hashes that pass the shared-memory Bloom prefilter of the reference table are queued;
//   2. the queued hashes are looked up 32 at a time (open addressing, L2 resident; false positives of the prefilter
//      drop out) and the table slots of the true hits are compacted to one per lane; each lane fetches its hit's
//      reference bitset row -- the copy number of a hash that occurs more than once in the read is its position among
//      equal slots, __match_any_sync (multiset semantics) -- and the 32 x (32 NW) bit matrix is transposed
//      in registers, after which a per-reference count is one POPC;
//   3. the scan of src/rkmh.cpp:874-883 is a warp max over (count << 8 | 255 - index) and a second max in front of it.
// Reads that are not eligible, whose candidates overflow the queue or that have more than 32 true hits go to
// classify_block_kernel through `redo`: the result is exact for every input.
constexpr int WARP_MAX_KMERS = 4096;
constexpr int WARP_QCAP = 128;                // queued candidates per read

template <int KCSEL, int NW>
__global__ void __launch_bounds__(1024, 1)
classify_warp_kernel(SeqBatch sb, const __grid_constant__ KList kl, int s, CountTable ct, int min_occ, RefTable tb, int filter_words,
                     int* __restrict__ redo, int* __restrict__ redo_count, Hit* __restrict__ out) {
#ifdef RK_EMU
    unsigned char* dsm = reinterpret_cast<unsigned char*>(EMU_DYN_SMEM());
#else
    extern __shared__ __align__(16) unsigned char dsm[];
#endif
    const u64* lut = hash_lut_load(kl.lut);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    u32* filt = reinterpret_cast<u32*>(dsm);
    u64* q = reinterpret_cast<u64*>(filt + filter_words) + (size_t)wib * WARP_QCAP;
    for (int i = threadIdx.x; i < filter_words; i += blockDim.x) filt[i] = tb.filter[i];
    __syncthreads();

    const TrLane tl = tr_lane(lane);
    const u32 lt = (1u << lane) - 1u;
    const i64 last_word = codes_last_word(sb), last_mword = nmask_last_word(sb);
    const u32 flog2 = tb.filter_log2;

    const i64 gwarp = (i64)blockIdx.x * nwb + wib, nwarps = (i64)gridDim.x * nwb;
    for (i64 rid = gwarp; rid < sb.nseq; rid += nwarps) {
        const i64 beg = sb.offs[rid] - sb.base0, len = sb.offs[rid + 1] - sb.offs[rid];
        const i64 ntot = nkmers_total(len, kl);
        bool give_up = ntot > s || ntot > WARP_MAX_KMERS;
        int nc = 0;
        u32 nz = 0;
        if (!give_up) {
            for (int j = 0; j < kl.nk; ++j) {
                const int k = kl.k[j];
                warp_hash_sel<KCSEL>(kl.kv[j], lut, sb.codes, sb.nmask, last_word, last_mword, beg, 0, nkmers(len, k), k, lane,
                                     [&](i64, u64 h, bool) {
                                         h = apply_min_occ(h, ct, min_occ);
                                         bool cand = false;
                                         if (h != 0) {
                                             ++nz;
                                             const u32 fb = filter_bit(h, flog2);
                                             cand = (filt[fb >> 5] >> (fb & 31)) & 1u;
                                         }
                                         const u32 bal = __ballot_sync(0xffffffffu, cand);
                                         if (bal) {
                                             const int at = nc + __popc(bal & lt);
                                             if (cand && at < WARP_QCAP) q[at] = h;
                                             nc += __popc(bal);
                                         }
                                     });
            }
            if (nc > WARP_QCAP) give_up = true;
        }
        int best = 0, id = 0, prev = -1;
        if (!give_up && nc > 0) {
            __syncwarp();
            // ---- candidates -> table slots (false positives of the prefilter drop out), compacted to one hit per lane ----
            u32* qk = reinterpret_cast<u32*>(q);             // keys overwrite queue entries that have been consumed
            int nh = 0;
            for (int b0 = 0; b0 < nc; b0 += 32) {
                u32 key = 0;                                 // slot + 1 of this lane's candidate, 0 = no reference holds it
                if (b0 + lane < nc) key = (u32)(table_find(tb, q[b0 + lane]) + 1);
                const u32 hit = __ballot_sync(0xffffffffu, key != 0u);
                __syncwarp();
                if (key && nh + __popc(hit & lt) < 32) qk[nh + __popc(hit & lt)] = key;
                nh += __popc(hit);
                __syncwarp();
            }
            if (nh > 32) give_up = true;
            else if (nh > 0) {
                // ---- one hit per lane: its reference bitset row; the copy number of a hash that occurs more than once in
                //      the read is its position among equal keys (multiset semantics) ---------------------------
                const u32 key = lane < nh ? qk[lane] : 0u;
                const u32 active = __ballot_sync(0xffffffffu, key != 0u);
                u32 x[NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) x[w] = 0;
                if (key) {
                    const u32 same = __match_any_sync(active, key);
                    const u64* row = table_row(tb, (int)key - 1, (u32)__popc(same & lt));
                    if (row != nullptr) {
#pragma unroll
                        for (int w = 0; w < NW / 2; ++w)
                            if ((u32)w < tb.nb) { const u64 v = __ldg(&row[w]); x[2 * w] = (u32)v; x[2 * w + 1] = (u32)(v >> 32); }
                    }
                }
                // ---- counts per reference (lane l, word w: reference 32 w + l) and the scan of src/rkmh.cpp:874-883 ---------
                u32 cnt[NW];
                u32 kbest = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    cnt[w] = ((u32)w < 2u * tb.nb) ? (u32)__popc(tr32(x[w], tl)) : 0u;
                    kbest = max(kbest, (cnt[w] << 8) | (255u - (u32)(32 * w + lane)));
                }
                kbest = warp_max(kbest);
                best = (int)(kbest >> 8);
                id = 255 - (int)(kbest & 255u);
                u32 p1 = 0;                                  // max of (count + 1) in front of the winner, 0 = none
#pragma unroll
                for (int w = 0; w < NW; ++w) if (32 * w + lane < id) p1 = max(p1, cnt[w] + 1u);
                prev = (int)warp_max(p1) - 1;
            }
            __syncwarp();
        }
        if (give_up) {
            if (lane == 0) redo[atomicAdd(redo_count, 1)] = (int)rid;
        } else {
            nz = warp_sum(nz);
            if (lane == 0) { Hit hh; hh.ref_id = id; hh.best = best; hh.prev = prev; hh.sketch_len = (int)nz; out[rid] = hh; }
        }
        __syncwarp();
    }
}

}  // namespace rk
