// scan_kernels.cuh — first pass of the two-pass Filter + Select plan: the predicate scan.
//
// Replaces (reference, /root/reference/src) the mask construction:
//   eager     physical_plan/plan.rs:112-130       (per row: series[i] <op> literal through AnyValue::partial_cmp, series.rs:87-117)
//   streaming execution/record_batch.rs:235-240   (mask.value(i) == Some(true) -> index list)
//
// Persistent kernel, one CTA per SM, no dependency between CTAs and none between the warps of a CTA.
// Every warp owns a contiguous range of 2048-row tiles and streams the predicate column through its own ring
// of 8 KB shared-memory slots with TMA 1-D bulk copies (cp.async.bulk + mbarrier complete_tx); the lane that
// finished a slot re-arms it, so ~190 KB per SM stay in flight without a producer warp or empty-barriers.
// Per 32 rows: one conflict-free LDS.64 per lane, the compare, one ballot = one selection word; lane k keeps word k,
// so a 1024-row slot ends in one coalesced 128-byte store of selection words.
//
// Outputs (consumed by compact_kernels.cuh / string_kernels.cuh):
//   sel_out      row-order selection bitmap, whole tiles written
//   tile_info    per tile: (exclusive survivor count inside the owning warp's range << 12) | survivors of the tile
//   chunk_base   per warp range: global exclusive output index of its first tile (base_in included) — written by the last
//                CTA to finish (ticket), which scans the <= 2368 range totals; also total_out = base + all survivors
//   dense_list / sparse_list + counts: tiles for the TMA-streamed compaction and for the gather kernel
// LIMIT: a warp stops once its own range holds `limit` survivors (rows after them cannot be among the first `limit`
// of the batch); tiles whose global prefix is beyond the limit are dropped by the second pass.
//
// HBM roofline: 8 B per row read (+ 1 bit validity), 1 bit per row written.
#pragma once
#include "fused_filter.cuh"

namespace rvl {

constexpr int kScanMaxWarps = 32;
// W warps per CTA, each with its own ring of n_slots slots of R rows (R = 0: 8192 / W rows, i.e. 8 warps: 8 KB slots, 16 warps:
// 4 KB slots, so that a CTA holds n_slots x 64 KB of predicate values).  A slot is re-armed when its last value has been
// consumed: smaller slots keep a larger share of the ring in flight while the warp computes.
template <int W, int R = 0> struct ScanShape {
    static constexpr int kItemRows = R > 0 ? R : 8192 / W;
    static constexpr int kItemWords = kItemRows / 32;
    static constexpr int kItemsPerTile = kTileRows / kItemRows;
    static constexpr uint32_t kItemBytes = kItemRows * 8;
};
constexpr int kInfoShift = 12;                          // tile_info = (prefix << 12) | count, count <= 2048

struct ScanParams {
    int64_t n_rows;
    int64_t n_tiles;
    int64_t tiles_per_warp;
    int64_t limit;
    const uint64_t* pred_values;
    int64_t lit_bits;
    uint64_t range_lo, range_span;
    uint32_t range_neg;
    uint32_t truth;
    uint32_t keep_null;
    int32_t pred_vec_ok;
    BitSrc pred_valid;
    uint32_t pb_a, pb_b;
    BitSrc pb_vals;
    int32_t n_slots;      // ring slots per warp
    uint32_t sparse_max;
    const unsigned long long* base_in;
    uint32_t* sel_out;
    uint64_t* tile_info;
    uint64_t* chunk_base;     // [gridDim.x * kScanWarps]
    uint32_t* dense_list;
    uint32_t* sparse_list;
    uint32_t* list_counts;    // [0] dense tiles, [1] sparse tiles, [2] CTAs finished (all zeroed before the launch)
    unsigned long long* total_out;
    uint32_t debug_skip;      // timing experiments only (RVL_SCAN_DEBUG): 1 = no selection-word stores, 2 = no tile_info / list stores
    uint32_t l2_hints;        // bit 0: predicate column loaded evict_first, bit 1: selection words stored evict_last
};

template <int PRED>
__device__ __forceinline__ bool scan_keep(const ScanParams& p, uint64_t v) {
    if (PRED == kPredI64) return ((v - p.range_lo) <= p.range_span) != (p.range_neg != 0u);
    return (p.truth & cmp_code<PRED>(v, p.lit_bits)) != 0u;
}

template <int PRED, int W, int R = 0>
__global__ void __launch_bounds__(W * 32, 1) predicate_scan_kernel(const __grid_constant__ ScanParams p) {
    constexpr int kScanWarps = W;
    constexpr int kScanItemRows = ScanShape<W, R>::kItemRows;
    constexpr int kItemWords = ScanShape<W, R>::kItemWords;
    constexpr int kItemsPerTile = ScanShape<W, R>::kItemsPerTile;
    constexpr uint32_t kScanItemBytes = ScanShape<W, R>::kItemBytes;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_part[kScanWarps];
    __shared__ uint32_t s_last;
    constexpr bool kNumeric = (PRED == kPredI64 || PRED == kPredF64);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.n_slots;
    uint64_t* const ring = reinterpret_cast<uint64_t*>(smem_raw) + (size_t)warp * D * kScanItemRows;
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kScanWarps * D * kScanItemBytes) + warp * D;

    const int64_t g = (int64_t)blockIdx.x * kScanWarps + warp;  // this warp's range index, in row order
    const int64_t t0 = min(p.n_tiles, g * p.tiles_per_warp);
    const int64_t t1 = min(p.n_tiles, t0 + p.tiles_per_warp);
    const int64_t n_items = (t1 - t0) * kItemsPerTile;
    const int64_t range_row0 = t0 * kTileRows;
    const uint64_t base0 = p.base_in != nullptr ? (uint64_t)*p.base_in : 0ull;
    const bool use_tma = kNumeric && p.pred_vec_ok != 0;

    auto item_tma = [&](int64_t j) { return use_tma && range_row0 + (j + 1) * kScanItemRows <= p.n_rows; };
    // the 8 bytes per row pass through once: evict_first keeps them from flushing the selection bitmap (1 bit per row, stored
    // evict_last) out of L2 before the compaction pass reads it back
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    auto load_item = [&](uint64_t* dst, const uint64_t* src, uint64_t* bar) {
        if (p.l2_hints & 1u) tma_load_1d_hint(dst, src, kScanItemBytes, bar, pol_stream); else tma_load_1d(dst, src, kScanItemBytes, bar);
    };
    if (kNumeric && lane == 0) {
        for (int s = 0; s < D; ++s) mbar_init(&full[s], 1);
        mbar_init_fence();
        for (int64_t j = 0; j < min((int64_t)D, n_items); ++j)
            if (item_tma(j)) load_item(ring + (size_t)j * kScanItemRows, p.pred_values + range_row0 + j * kScanItemRows, &full[j]);
    }
    __syncwarp();

    uint64_t lprefix = 0;      // survivors of this range so far
    uint32_t n_dense = 0, n_sparse = 0, pend_dense = 0, pend_sparse = 0;  // list batching: lane i holds the i-th pending tile
    auto flush = [&](uint32_t* list, uint32_t* counter, uint32_t mine, uint32_t n) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(counter, n);
        b = __shfl_sync(0xFFFFFFFFu, b, 0);
        if ((uint32_t)lane < n) list[b + lane] = mine;
    };

    int slot = 0;
    uint32_t phase = 0;
    int64_t j = 0;  // next item of the range
    // validity / bitmap words of the next item, loaded one item ahead (lane <-> 32-row word)
    uint32_t nx_a = 0, nx_b = 0;
    auto prefetch_words = [&](int64_t jj) {
        if (jj >= n_items) return;
        const uint64_t r = (uint64_t)(range_row0 + jj * kScanItemRows + lane * 32);
        if (kNumeric) { if (p.pred_valid.words != nullptr) nx_a = load_bits32(p.pred_valid, r); }
        else if (PRED == kPredBits) { nx_a = load_bits32(p.pb_vals, r); nx_b = load_bits32(p.pred_valid, r); }
    };
    prefetch_words(0);

#pragma unroll 1
    for (int64_t t = t0; t < t1; ++t) {
        if (p.limit >= 0 && base0 + lprefix >= (uint64_t)p.limit) break;
        uint32_t tcount = 0;
#pragma unroll 1
        for (int h = 0; h < kItemsPerTile; ++h, ++j) {
            const int64_t row0 = t * kTileRows + (int64_t)h * kScanItemRows;
            const bool whole = row0 + kScanItemRows <= p.n_rows;
            const uint32_t wa = nx_a, wb = nx_b;
            prefetch_words(j + 1);
            uint32_t myword = 0;
            if (kNumeric) {
                const bool tma = use_tma && whole;
                const bool has_valid = p.pred_valid.words != nullptr;
                if (tma) mbar_wait(&full[slot], phase);
                const uint64_t* src = ring + (size_t)slot * kScanItemRows + lane;
                // one body, instantiated for the two (warp-uniform) sources so the inner loop carries no branch: the TMA ring
                // (LDS.64, conflict-free) or, for a ragged tail / unaligned view, the column itself
                auto evaluate = [&](auto load) {
#pragma unroll
                    for (int b = 0; b < kItemWords / 8; ++b) {
                        uint64_t v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] = load(b * 8 + k);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int w = b * 8 + k;
                            bool c = scan_keep<PRED>(p, v[k]);
                            if (has_valid) {
                                const uint32_t vw = __shfl_sync(0xFFFFFFFFu, wa, w);
                                c = ((vw >> lane) & 1u) ? c : (p.keep_null != 0u);
                            }
                            const uint32_t m = __ballot_sync(0xFFFFFFFFu, c);
                            if (lane == w) myword = m;
                        }
                    }
                };
                if (tma) evaluate([&](int w) -> uint64_t { return src[w * 32]; });
                else evaluate([&](int w) -> uint64_t {
                    const int64_t row = row0 + w * 32 + lane;
                    return row < p.n_rows ? ld_stream(p.pred_values + row) : 0ull;
                });
                // every value of the slot has been consumed: re-arm it with the item D steps ahead
                if (lane == 0 && j + D < n_items && item_tma(j + D))
                    load_item(ring + (size_t)slot * kScanItemRows, p.pred_values + range_row0 + (j + D) * kScanItemRows, &full[slot]);
                if (++slot == D) { slot = 0; phase ^= 1u; }
            } else if (PRED == kPredBits) {
                const uint32_t a = p.pb_a ? ~0u : 0u, b = p.pb_b ? ~0u : 0u, kn = p.keep_null ? ~0u : 0u;
                myword = (wb & ((wa & a) ^ b)) | (~wb & kn);
            } else {
                myword = ~0u;
            }
            if (!whole) {
                const int64_t rem = p.n_rows - (row0 + lane * 32);
                if (rem < 32) myword = rem <= 0 ? 0u : (myword & ((1u << rem) - 1u));
            }
            if (lane >= kItemWords) myword = 0u;  // lanes beyond the item's words hold nothing
            if (lane < kItemWords && !(p.debug_skip & 1u)) {
                if (p.l2_hints & 2u) st_u32_hint(p.sel_out + (row0 >> 5) + lane, myword, pol_keep); else p.sel_out[(row0 >> 5) + lane] = myword;
            }
            tcount += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(myword));
        }
        if (lane == 0 && !(p.debug_skip & 2u)) p.tile_info[t] = (lprefix << kInfoShift) | (uint64_t)tcount;
        lprefix += tcount;
        if (p.debug_skip & 2u) continue;
        if (tcount > p.sparse_max) {
            if ((uint32_t)lane == n_dense) pend_dense = (uint32_t)t;
            if (++n_dense == 32u) { flush(p.dense_list, p.list_counts, pend_dense, 32u); n_dense = 0; }
        } else if (tcount != 0u) {
            if ((uint32_t)lane == n_sparse) pend_sparse = (uint32_t)t;
            if (++n_sparse == 32u) { flush(p.sparse_list, p.list_counts + 1, pend_sparse, 32u); n_sparse = 0; }
        }
    }
    if (n_dense != 0u) flush(p.dense_list, p.list_counts, pend_dense, n_dense);
    if (n_sparse != 0u) flush(p.sparse_list, p.list_counts + 1, pend_sparse, n_sparse);
    if (kNumeric) {
        // LIMIT stop: drain the copies still in flight before the CTA may exit
        for (int64_t jj = j; jj < min(n_items, j + D); ++jj) {
            if (item_tma(jj)) mbar_wait(&full[slot], phase);
            if (++slot == D) { slot = 0; phase ^= 1u; }
        }
    }
    if (lane == 0) p.chunk_base[g] = lprefix;

    // ---- last CTA to finish turns the range totals into global exclusive bases
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(p.list_counts + 2, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (s_last == 0u) return;
    __threadfence();
    const int G = (int)gridDim.x * kScanWarps;
    const int per = (G + (int)blockDim.x - 1) / (int)blockDim.x;  // <= 10 at 2 CTAs/SM
    uint64_t vals[12];
    uint64_t mine = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const int idx = tid * per + i;
        vals[i] = (i < per && idx < G) ? ld_relaxed_gpu(p.chunk_base + idx) : 0ull;
        mine += vals[i];
    }
    uint64_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    uint64_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kScanWarps; ++w) {
        const uint64_t c = s_part[w];
        woff += (w < warp) ? c : 0ull;
        total += c;
    }
    uint64_t run = base0 + woff + incl - mine;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const int idx = tid * per + i;
        if (i < per && idx < G) { p.chunk_base[idx] = run; run += vals[i]; }
    }
    if (tid == 0) *p.total_out = (unsigned long long)(base0 + total);
}

}  // namespace rvl
