// fused_filter.cuh — K1+K2+K3: predicate -> ballot/popc -> block scan -> decoupled look-back ->
// ordered stream compaction of every projected fixed-width and bit-packed column, in ONE pass.
//
// Replaces (reference, /root/reference/src):
//   eager     physical_plan/plan.rs:112-130 (mask loop) + :132-147 (per-column gather)
//   streaming execution/record_batch.rs:235-240 (mask -> indices) + :131-178 (take_array builders)
//             execution/array/bitmap.rs:142-155 (bit-at-a-time BitmapBuilder::append)
//
// Data layout in HBM: Arrow columns — 8-byte values, LSB-first bitmaps (1 = valid / true), any
// (offset, length) view.
//
// Work decomposition.  A CTA (256 threads) owns one SUPER-TILE of 8192 consecutive rows = 4 sub-tiles of
// 2048 rows.  Inside a sub-tile warp w owns rows [256w, 256w+256) as 4 groups of 64 rows; lane l of a group
// owns rows 2l and 2l+1, i.e. ONE 128-bit access per 8-byte column per group.  Two ballots per group (even
// rows / odd rows) give the group's 64-bit selection word in row order.
//
//   phase A  the predicate column of the whole super-tile (64 KB) is fetched with four TMA 1-D bulk copies
//            (cp.async.bulk + mbarrier::complete_tx) issued by one thread at CTA start — 64 KB per CTA and
//            ~190 KB per SM in flight without holding a single register; each sub-tile is evaluated as its
//            copy lands; the selection words and per-warp counts stay in shared memory.
//   order    ONE decoupled look-back per super-tile (256 descriptors per round, run by warp 0) resolves the
//            global output offset; the other warps meanwhile gather and stage the first sub-tile, which
//            only needs tile-local ranks.  Super-tiles are taken in blockIdx order, so a CTA only waits on
//            CTAs that were dispatched before it (resident or finished).
//   phase B  warp-local: a warp's survivors of a sub-tile form one contiguous output run.  The warp gathers
//            them with predicated 128-bit loads (two columns in flight), stages them in output order in its
//            private 8 KB slice of the (by then dead) predicate buffer and writes the run out contiguously;
//            bit-packed columns are compacted with warp REDUX.OR into staged words and funnel-shifted to the
//            output bit position (boundary words OR-ed atomically).  No block barrier after the look-back
//            join, so the warps of a CTA drift apart and loads, staging and stores overlap.
//
// HBM roofline: per row the kernel must read the predicate value (8 B [+1 bit
// validity]), the 32-B sectors of each projected column that hold at least one survivor, and write
// s*8 B per projected column.  Dead sectors are never requested: gather loads are predicated on the
// lane's own keep bits, and a lane's two rows sit in one sector.
#pragma once
#include "device_utils.cuh"

namespace rvl {

constexpr int kMaxCol8 = 8;      // 8-byte columns compacted per launch
constexpr int kMaxBitCols = 8;   // bit-packed columns compacted per launch
constexpr int kWarpBitWords = 256 / 32 + 2;     // a warp emits at most 256 survivors per sub-tile
constexpr int kSub = 4;                        // sub-tiles per super-tile
constexpr int kSuperRows = kSub * kTileRows;   // 8192 rows per CTA
constexpr int kStageBufs = 4;                  // staging buffers (alias the predicate buffer)
constexpr int kSparseWarp = 16;                // warps with <= this many survivors in the super-tile use the list path

enum PredKind : int { kPredI64 = 0, kPredF64 = 1, kPredBits = 2, kPredTrue = 3 };

struct Col8 {
    const uint64_t* in;  // row 0 of the view
    uint64_t* out;
    BitSrc valid;        // words == nullptr: no nulls
    int32_t vec_ok;      // `in` is 16-byte aligned -> 128-bit loads
    int32_t pad;
};

struct BitCol {
    BitSrc in;    // bits to compact (validity, or Boolean values); nullptr words = all ones
    BitSrc mask;  // optional AND mask (Boolean values are stored 0 under a null: boolean.rs:29-32)
    uint32_t* out;  // zero-initialised, bit 0 = output row 0
};

struct FusedParams {
    int64_t n_rows;
    int64_t n_super;  // super-tiles = gridDim.x
    int64_t limit;    // < 0: none; else survivors whose global output index >= limit are dropped
    // comparison predicate: keep = valid ? truth[cmp(value, literal)] : keep_null
    const uint64_t* pred_values;
    int64_t lit_bits;    // Float64: literal bits for the 3-way compare + truth mask
    // Int64: every comparison is one unsigned range test, keep = ((v - range_lo) <= range_span) != range_neg
    uint64_t range_lo, range_span;
    uint32_t range_neg;
    uint32_t pad1;
    BitSrc pred_valid;
    uint32_t truth;      // bit0: value < lit, bit1: ==, bit2: >, bit3: unordered (NaN)
    uint32_t keep_null;  // 0 / 1: what a null row evaluates to (series.rs:105-107 puts Null below everything)
    int32_t pred_vec_ok; // predicate values are 16-byte aligned: TMA bulk copies + 128-bit accesses
    int32_t n_col8;
    int32_t n_bits;
    // bitmap predicate: sel = (valid & ((vals & a) ^ b)) | (~valid & keep_null)
    uint32_t pb_a, pb_b;
    int32_t pad0;
    BitSrc pb_vals;
    Col8 col8[kMaxCol8];
    BitCol bits[kMaxBitCols];
    uint64_t* tile_status;               // one descriptor per super-tile, zeroed before the launch
    const unsigned long long* base_in;   // rows already emitted by earlier batches of the same query (streaming), or nullptr.
                                         // The limit applies to base + rank; output buffers are indexed by rank alone.
    unsigned long long* total_out;       // base + survivors of this launch (saturates at >= limit once the limit trips)
    uint32_t* done_flag;                 // set once some super-tile's inclusive prefix reaches the limit
    uint32_t* sel_out;                   // optional: row-order selection bitmap (n_rows bits, 8-byte aligned)
    uint64_t* tile_prefix_out;           // optional: exclusive output prefix of every 2048-row tile
    // two-pass plan (compact_kernels.cuh): this launch only scans the predicate; every 2048-row tile with survivors is
    // appended to the dense list (> sparse_max survivors: TMA-streamed compaction) or the sparse list (gathered)
    uint32_t* dense_list;
    uint32_t* sparse_list;
    uint32_t* list_counts;               // [0] dense tiles, [1] sparse tiles (zeroed before the launch)
    uint32_t sparse_max;
    uint32_t pad2;
};

// dynamic shared memory of the fused kernel
struct __align__(128) FusedSmem {
    // phase A: the super-tile's predicate values (TMA destination); phase B: 4 staging buffers of 2048 survivors
    uint64_t buf[kSub][kTileRows];                // 64 KB
    uint32_t sel0[kSub][kTileRows / 64];          // per 64-row group: keep mask of the even rows (lane l <-> row 2l)
    uint32_t sel1[kSub][kTileRows / 64];          //                   keep mask of the odd rows  (lane l <-> row 2l+1)
    uint32_t wbits[kWarps][kMaxBitCols][kWarpBitWords];  // per warp: compacted bit columns of the run being emitted
    uint32_t warp_count[kSub][kWarps];
    uint64_t mbar[kSub];                          // one single-use mbarrier per sub-tile copy
    uint16_t wl_row[kWarps][kSparseWarp];         // sparse-warp path: survivor rows (within the super-tile) ...
    uint16_t wl_out[kWarps][kSparseWarp];         // ... and their output positions relative to the super-tile's offset
    uint64_t excl;
    uint64_t base;
    uint32_t done;
    uint32_t ready;                               // set by warp 0 once excl/base are valid
};

template <int PRED>
__device__ __forceinline__ uint32_t cmp_code(uint64_t bits, int64_t lit_bits) {
    if (PRED == kPredI64) {
        const int64_t a = (int64_t)bits, b = lit_bits;
        return a < b ? 1u : (a == b ? 2u : 4u);
    } else {
        const double a = __longlong_as_double((long long)bits), b = __longlong_as_double(lit_bits);
        return a < b ? 1u : (a == b ? 2u : (a > b ? 4u : 8u));
    }
}

// ---- TMA 1-D bulk copy + mbarrier (sm_90+/sm_100a) ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// L2 cache policies for streams that pass through once (evict_first) and for small intermediates the next kernel re-reads (evict_last)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void st_u32_hint(uint32_t* p, uint32_t v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0u;
}
// non-blocking poll (try_wait may suspend the thread for a system-dependent time: an event loop must not use it)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}


template <int PRED>
__global__ void __launch_bounds__(kBlock, 3) fused_filter_project_kernel(const __grid_constant__ FusedParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    constexpr bool kNumeric = (PRED == kPredI64 || PRED == kPredF64);
    constexpr uint32_t kTileBytes = kTileRows * 8;
    const int64_t super = blockIdx.x;
    const int64_t super_row0 = super * kSuperRows;
    const bool use_tma = kNumeric && p.pred_vec_ok != 0;

    // ---------------------------------------------------------------- start: fetch the whole super-tile's predicate values
    if (tid == 0) {
        sm.done = p.limit >= 0 ? ld_relaxed_gpu_u32(p.done_flag) : 0u;
        sm.ready = 0u;
        if (use_tma) {
#pragma unroll
            for (int s = 0; s < kSub; ++s) mbar_init(&sm.mbar[s], 1);
            mbar_init_fence();
            if (sm.done == 0u) {
#pragma unroll
                for (int s = 0; s < kSub; ++s) {
                    const int64_t row0 = super_row0 + (int64_t)s * kTileRows;
                    if (row0 + kTileRows <= p.n_rows) tma_load_1d(sm.buf[s], p.pred_values + row0, kTileBytes, &sm.mbar[s]);
                }
            }
        }
    }
    __syncthreads();

    // LIMIT early termination (streaming.rs:269-271 / plan.rs:163): an earlier super-tile already reached the
    // limit — publish a saturated prefix and leave without touching HBM.
    if (sm.done != 0u) {
        if (tid == 0) {
            st_relaxed_gpu(p.tile_status + super, kStatusPrefix | (uint64_t)p.limit);
            if (p.tile_prefix_out) {
                for (int s = 0; s < kSub; ++s) {
                    const int64_t t = super * kSub + s;
                    if (t * kTileRows < p.n_rows) p.tile_prefix_out[t] = (uint64_t)p.limit;
                }
            }
            if (super == p.n_super - 1) *p.total_out = (unsigned long long)p.limit;
        }
        return;
    }

    // ---------------------------------------------------------------- phase A: predicate -> selection masks + counts
    const uint32_t lt = lanemask_lt();
#pragma unroll 1
    for (int s = 0; s < kSub; ++s) {
        const int64_t sub_row0 = super_row0 + (int64_t)s * kTileRows;
        const int64_t warp_row0 = sub_row0 + (int64_t)warp * (kGroups * 64);
        const bool full = sub_row0 + kTileRows <= p.n_rows;  // uniform: no ragged tail in this sub-tile
        uint32_t kb[kGroups];
        if (sub_row0 >= p.n_rows) {
#pragma unroll
            for (int g = 0; g < kGroups; ++g) kb[g] = 0u;
        } else if (kNumeric) {
            uint64_t v[kGroups][2];
            if (use_tma && full) {
                mbar_wait(&sm.mbar[s], 0);
#pragma unroll
                for (int g = 0; g < kGroups; ++g) {
                    const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(&sm.buf[s][warp * (kGroups * 64) + g * 64 + lane * 2]);
                    v[g][0] = t.x; v[g][1] = t.y;
                }
            } else {
#pragma unroll
                for (int g = 0; g < kGroups; ++g) {
                    const int64_t row = warp_row0 + g * 64 + lane * 2;
                    v[g][0] = 0; v[g][1] = 0;
                    if (row + 1 < p.n_rows) {
                        if (p.pred_vec_ok) { const ulonglong2 t = ld_stream_v2(p.pred_values + row); v[g][0] = t.x; v[g][1] = t.y; }
                        else { v[g][0] = ld_stream(p.pred_values + row); v[g][1] = ld_stream(p.pred_values + row + 1); }
                    } else if (row < p.n_rows) {
                        v[g][0] = ld_stream(p.pred_values + row);
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                uint32_t c;
                if (PRED == kPredI64) {
                    const uint32_t c0 = ((v[g][0] - p.range_lo) <= p.range_span) ? 1u : 0u;
                    const uint32_t c1 = ((v[g][1] - p.range_lo) <= p.range_span) ? 2u : 0u;
                    c = (c0 | c1) ^ (p.range_neg ? 3u : 0u);
                } else {
                    const uint32_t c0 = (p.truth & cmp_code<PRED>(v[g][0], p.lit_bits)) != 0u ? 1u : 0u;
                    const uint32_t c1 = (p.truth & cmp_code<PRED>(v[g][1], p.lit_bits)) != 0u ? 2u : 0u;
                    c = c0 | c1;
                }
                if (p.pred_valid.words != nullptr) {
                    const uint32_t vb = (uint32_t)(load_bits64(p.pred_valid, (uint64_t)(warp_row0 + g * 64)) >> (2 * lane)) & 3u;
                    c = (c & vb) | ((p.keep_null ? 3u : 0u) & ~vb);
                }
                if (!full) {
                    const int64_t row = warp_row0 + g * 64 + lane * 2;
                    c &= (row < p.n_rows ? 1u : 0u) | (row + 1 < p.n_rows ? 2u : 0u);
                }
                kb[g] = c;
            }
        } else {
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                const int64_t grow = warp_row0 + g * 64;
                uint64_t sel = ~0ull;
                if (PRED == kPredBits) {
                    const uint64_t vals = load_bits64(p.pb_vals, (uint64_t)grow);
                    const uint64_t valid = load_bits64(p.pred_valid, (uint64_t)grow);
                    const uint64_t a = p.pb_a ? ~0ull : 0ull, b = p.pb_b ? ~0ull : 0ull, kn = p.keep_null ? ~0ull : 0ull;
                    sel = (valid & ((vals & a) ^ b)) | (~valid & kn);
                }
                const int64_t rem = p.n_rows - grow;
                if (rem < 64) sel = rem <= 0 ? 0ull : (sel & ((1ull << rem) - 1ull));
                kb[g] = (uint32_t)(sel >> (2 * lane)) & 3u;
            }
        }
        uint32_t wrun = 0;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, (kb[g] & 1u) != 0u);
            const uint32_t m1 = __ballot_sync(0xFFFFFFFFu, (kb[g] & 2u) != 0u);
            wrun += __popc(m0) + __popc(m1);
            if (lane == 0) {
                sm.sel0[s][warp * kGroups + g] = m0;
                sm.sel1[s][warp * kGroups + g] = m1;
            }
            if (p.sel_out != nullptr) {  // row-order selection bitmap for the string kernels / mask output
                const int64_t grow = warp_row0 + g * 64;
                if (lane == 0 && grow < p.n_rows) *reinterpret_cast<uint64_t*>(p.sel_out + (grow >> 5)) = interleave_masks(m0, m1);
            }
        }
        if (lane == 0) sm.warp_count[s][warp] = wrun;
    }
    __syncthreads();  // barrier #1: selection masks and counts of the whole super-tile; predicate buffer is dead

    // per-warp bookkeeping, one sub-tile per lane (lanes 0..kSub-1), broadcast with shuffles when needed
    uint32_t l_cnt = 0, l_woff = 0, l_mine = 0;  // sub-tile survivors / survivors of lower warps / this warp's survivors
    if (lane < kSub) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = sm.warp_count[lane][w];
            l_woff += (w < warp) ? c : 0u;
            l_mine = (w == warp) ? c : l_mine;
            l_cnt += c;
        }
    }
    uint32_t l_soff = l_cnt;  // exclusive prefix of the sub-tile counts
#pragma unroll
    for (int o = 1; o < kSub; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, l_soff, o);
        if (lane >= o) l_soff += n;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, l_soff, kSub - 1);
    l_soff -= l_cnt;
    uint32_t warp_total = l_mine;
    warp_total += __shfl_xor_sync(0xFFFFFFFFu, warp_total, 1);
    warp_total += __shfl_xor_sync(0xFFFFFFFFu, warp_total, 2);
    warp_total = __shfl_sync(0xFFFFFFFFu, warp_total, 0);

    // ---------------------------------------------------------------- global order: one decoupled look-back per super-tile (warp 0)
    if (warp == 0) {
        uint64_t excl;
        const uint64_t base0 = p.base_in != nullptr ? (uint64_t)*p.base_in : 0ull;
        if (super == 0) {
            excl = base0;
            if (lane == 0) st_relaxed_gpu(p.tile_status, kStatusPrefix | (excl + total));
        } else {
            if (lane == 0) st_relaxed_gpu(p.tile_status + super, kStatusAggregate | (uint64_t)total);
            excl = lookback_exclusive_wide(p.tile_status, super, lane);
            if (lane == 0) st_relaxed_gpu(p.tile_status + super, kStatusPrefix | (excl + total));
        }
        if (lane == 0) {
            sm.excl = excl;
            sm.base = base0;
            __threadfence_block();
            *reinterpret_cast<volatile uint32_t*>(&sm.ready) = 1u;  // releases the warps polling for the output offset
            if (super == p.n_super - 1) *p.total_out = (unsigned long long)(excl + total);
            if (p.limit >= 0 && excl + total >= (uint64_t)p.limit) atomicExch(p.done_flag, 1u);
        }
        if (p.tile_prefix_out != nullptr && lane < kSub) {
            const int64_t t = super * kSub + lane;
            if (t * kTileRows < p.n_rows) p.tile_prefix_out[t] = excl + l_soff;
        }
        if (p.list_counts != nullptr) {
            // classify this super-tile's tiles for the second pass; tiles entirely beyond the limit are dropped
            const int64_t t = super * kSub + lane;
            const bool live = lane < kSub && l_cnt != 0u && (p.limit < 0 || excl + l_soff < (uint64_t)p.limit);
            const bool dense = live && l_cnt > p.sparse_max;
            const uint32_t dm = __ballot_sync(0xFFFFFFFFu, dense), sm_ = __ballot_sync(0xFFFFFFFFu, live && !dense);
            uint32_t db = 0, sb = 0;
            if (lane == 0) {
                if (dm != 0u) db = atomicAdd(p.list_counts, (uint32_t)__popc(dm));
                if (sm_ != 0u) sb = atomicAdd(p.list_counts + 1, (uint32_t)__popc(sm_));
            }
            db = __shfl_sync(0xFFFFFFFFu, db, 0);
            sb = __shfl_sync(0xFFFFFFFFu, sb, 0);
            if (dense) p.dense_list[db + __popc(dm & lt)] = (uint32_t)t;
            else if (live) p.sparse_list[sb + __popc(sm_ & lt)] = (uint32_t)t;
        }
    }
    if (p.n_col8 == 0 && p.n_bits == 0) return;  // scan-only launch (mask / two-pass plan / string-only projection)
    if (warp_total == 0u) return;  // this warp has nothing to emit (no block barrier follows)

    // ---------------------------------------------------------------- phase B: every warp emits its own survivors
    // A warp's survivors of one sub-tile occupy the contiguous output range [excl + sub_off + warp_off, + my_cnt).
    // No block barrier after barrier #1: a warp waits for the output offset (sm.ready) only when it is about to
    // store, so its gather loads overlap warp 0's look-back, and the eight warps drift apart.
    auto wait_offset = [&](uint64_t& excl, uint64_t& base0) {
        while (*reinterpret_cast<volatile uint32_t*>(&sm.ready) == 0u) {}
        __threadfence_block();
        __syncwarp();
        excl = *reinterpret_cast<volatile uint64_t*>(&sm.excl);
        base0 = *reinterpret_cast<volatile uint64_t*>(&sm.base);
    };
    uint64_t excl = 0, base0 = 0;

    if (warp_total <= (uint32_t)kSparseWarp) {
        // ---- few survivors in this warp (low selectivity): one (survivor, column) pair per lane, so every gather
        // load of the warp is in flight at once and exactly one memory round trip is exposed
        uint16_t* wrow = sm.wl_row[warp];
        uint16_t* wout = sm.wl_out[warp];
        uint32_t base = 0;
#pragma unroll 1
        for (int s = 0; s < kSub; ++s) {
            const uint32_t mine = __shfl_sync(0xFFFFFFFFu, l_mine, s);
            if (mine == 0u) continue;
            const uint32_t pos0 = __shfl_sync(0xFFFFFFFFu, l_soff, s) + __shfl_sync(0xFFFFFFFFu, l_woff, s);
            uint32_t run = 0;
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                const uint32_t m0 = sm.sel0[s][warp * kGroups + g], m1 = sm.sel1[s][warp * kGroups + g];
                uint32_t r = run + __popc(m0 & lt) + __popc(m1 & lt);
                const uint32_t row = (uint32_t)(s * kTileRows + warp * (kGroups * 64) + g * 64 + lane * 2);
                if ((m0 >> lane) & 1u) { wrow[base + r] = (uint16_t)row; wout[base + r] = (uint16_t)(pos0 + r); ++r; }
                if ((m1 >> lane) & 1u) { wrow[base + r] = (uint16_t)(row + 1); wout[base + r] = (uint16_t)(pos0 + r); }
                run += __popc(m0) + __popc(m1);
            }
            base += mine;
        }
        __syncwarp();
        const uint32_t n_task = warp_total * (uint32_t)p.n_col8;
        bool have = false;
        for (uint32_t t0 = 0; t0 < n_task; t0 += 32) {
            const uint32_t t = t0 + lane;
            uint64_t v = 0;
            uint32_t e = 0, c = 0;
            if (t < n_task) {
                e = t / (uint32_t)p.n_col8; c = t - e * (uint32_t)p.n_col8;
                const Col8& col = p.col8[c];
                const int64_t row = super_row0 + wrow[e];
                bool ok = true;
                if (col.valid.words != nullptr) { const uint64_t bit = col.valid.bit0 + (uint64_t)row; ok = (__ldg(col.valid.words + (bit >> 5)) >> (bit & 31)) & 1u; }
                if (ok) v = ld_gather(col.in + row);  // 64-byte DRAM granule; placeholder 0 under a null (primitive.rs:175-178)
            }
            if (!have) { wait_offset(excl, base0); have = true; }
            if (t < n_task) {
                const uint64_t gpos = excl + wout[e];
                if (p.limit < 0 || gpos < (uint64_t)p.limit) st_stream(p.col8[c].out + (gpos - base0), v);
            }
        }
        const uint32_t n_btask = warp_total * (uint32_t)p.n_bits;
        for (uint32_t t0 = 0; t0 < n_btask; t0 += 32) {
            const uint32_t t = t0 + lane;
            bool set = false;
            uint32_t e = 0, b = 0;
            if (t < n_btask) {
                e = t / (uint32_t)p.n_bits; b = t - e * (uint32_t)p.n_bits;
                const BitCol& bc = p.bits[b];
                const uint64_t row = (uint64_t)(super_row0 + wrow[e]);
                set = true;
                if (bc.in.words != nullptr) { const uint64_t bit = bc.in.bit0 + row; set = (__ldg(bc.in.words + (bit >> 5)) >> (bit & 31)) & 1u; }
                if (set && bc.mask.words != nullptr) { const uint64_t bit = bc.mask.bit0 + row; set = (__ldg(bc.mask.words + (bit >> 5)) >> (bit & 31)) & 1u; }
            }
            if (!have) { wait_offset(excl, base0); have = true; }
            if (set) {
                const uint64_t gpos = excl + wout[e];
                if (p.limit < 0 || gpos < (uint64_t)p.limit) { const uint64_t pos = gpos - base0; atomicOr(p.bits[b].out + (pos >> 5), 1u << (pos & 31)); }
            }
        }
        return;
    }

    // ---- many survivors: gather with predicated 128-bit loads (two columns in flight), stage in the warp's private
    // 8 KB slice of the dead predicate buffer in output order, write each run out contiguously
    uint64_t* const wstage = &sm.buf[0][0] + warp * (kStageBufs * 256);  // [kStageBufs][256] survivors
    uint32_t* const wbits = &sm.wbits[warp][0][0];                       // [kMaxBitCols][kWarpBitWords]
    bool have_excl = false;
#pragma unroll 1
    for (int s = 0; s < kSub; ++s) {
        const uint32_t my_cnt = __shfl_sync(0xFFFFFFFFu, l_mine, s);
        if (my_cnt == 0u) continue;
        const uint32_t pos0 = __shfl_sync(0xFFFFFFFFu, l_soff, s) + __shfl_sync(0xFFFFFFFFu, l_woff, s);
        const int64_t sub_row0 = super_row0 + (int64_t)s * kTileRows;
        const int64_t warp_row0 = sub_row0 + (int64_t)warp * (kGroups * 64);
        const bool full = sub_row0 + kTileRows <= p.n_rows;

        // this lane's keep bits and warp-local ranks, rebuilt from the selection masks
        uint32_t kb[kGroups], rank0[kGroups];
        {
            uint32_t run = 0;
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                const uint32_t m0 = sm.sel0[s][warp * kGroups + g], m1 = sm.sel1[s][warp * kGroups + g];
                kb[g] = ((m0 >> lane) & 1u) | (((m1 >> lane) & 1u) << 1);
                rank0[g] = run + __popc(m0 & lt) + __popc(m1 & lt);
                run += __popc(m0) + __popc(m1);
            }
        }

        // gather one or two columns: all loads are issued before the first one is consumed
        auto stage_columns = [&](const Col8& colA, const Col8& colB, uint64_t* stageA, uint64_t* stageB, const bool two) {
            uint64_t va[kGroups][2], vb2[kGroups][2];
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                va[g][0] = 0; va[g][1] = 0; vb2[g][0] = 0; vb2[g][1] = 0;
                if (kb[g] != 0u) {
                    const int64_t row = warp_row0 + g * 64 + lane * 2;
                    const bool pair_ok = full || row + 1 < p.n_rows;
                    if (colA.vec_ok && pair_ok) { const ulonglong2 t = ld_stream_v2(colA.in + row); va[g][0] = t.x; va[g][1] = t.y; }
                    else {
                        if ((kb[g] & 1u) != 0u) va[g][0] = ld_stream(colA.in + row);
                        if ((kb[g] & 2u) != 0u) va[g][1] = ld_stream(colA.in + row + 1);
                    }
                    if (two) {
                        if (colB.vec_ok && pair_ok) { const ulonglong2 t = ld_stream_v2(colB.in + row); vb2[g][0] = t.x; vb2[g][1] = t.y; }
                        else {
                            if ((kb[g] & 1u) != 0u) vb2[g][0] = ld_stream(colB.in + row);
                            if ((kb[g] & 2u) != 0u) vb2[g][1] = ld_stream(colB.in + row + 1);
                        }
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                if (kb[g] != 0u) {
                    uint32_t ma = 3u, mb = 3u;
                    if (colA.valid.words != nullptr) ma = (uint32_t)(load_bits64(colA.valid, (uint64_t)(warp_row0 + g * 64)) >> (2 * lane)) & 3u;
                    if (two && colB.valid.words != nullptr) mb = (uint32_t)(load_bits64(colB.valid, (uint64_t)(warp_row0 + g * 64)) >> (2 * lane)) & 3u;
                    uint32_t r = rank0[g];
                    if ((kb[g] & 1u) != 0u) {  // placeholder 0 under a null (primitive.rs:175-178)
                        stageA[r] = (ma & 1u) ? va[g][0] : 0ull;
                        if (two) stageB[r] = (mb & 1u) ? vb2[g][0] : 0ull;
                        ++r;
                    }
                    if ((kb[g] & 2u) != 0u) {
                        stageA[r] = (ma & 2u) ? va[g][1] : 0ull;
                        if (two) stageB[r] = (mb & 2u) ? vb2[g][1] : 0ull;
                    }
                }
            }
        };
        auto stage_round = [&](int c0, int nc) {  // stage columns [c0, c0 + nc) into the warp's buffers [0, nc)
            for (int c = 0; c < nc; c += 2) {
                const bool two = c + 1 < nc;
                stage_columns(p.col8[c0 + c], p.col8[c0 + (two ? c + 1 : c)], wstage + c * 256, wstage + (two ? c + 1 : c) * 256, two);
            }
        };
        const int nc_first = p.n_col8 < kStageBufs ? p.n_col8 : kStageBufs;

        // ---- stage the bit-packed columns and the first round of 8-byte columns (warp-local ranks only)
        for (int b = 0; b < p.n_bits; ++b) {
            const BitCol bc = p.bits[b];
            uint32_t* wb = wbits + b * kWarpBitWords;
            if (lane < kWarpBitWords) wb[lane] = 0u;
            __syncwarp();
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                const int64_t grow = warp_row0 + g * 64;
                const uint32_t gbase = __shfl_sync(0xFFFFFFFFu, rank0[g], 0);  // lane 0 has no lower lanes
                const uint64_t w = load_bits64(bc.in, (uint64_t)grow) & load_bits64(bc.mask, (uint64_t)grow);
                const uint32_t lb = (uint32_t)(w >> (2 * lane)) & 3u & kb[g];
                const uint32_t lp = rank0[g] - gbase;  // rank inside the group, < 64
                const uint64_t contrib = ((uint64_t)(lb & 1u) << lp) | ((uint64_t)((lb >> 1) & 1u) << (lp + (kb[g] & 1u)));
                const uint32_t lo = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)contrib);
                const uint32_t hi = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)(contrib >> 32));
                if (lane == 0 && (lo | hi) != 0u) {
                    const uint32_t sh = gbase & 31u;  // gbase = warp-local position of the group's first survivor
                    uint32_t* dst = wb + (gbase >> 5);
                    dst[0] |= lo << sh;
                    dst[1] |= __funnelshift_l(lo, hi, sh);
                    if (sh != 0u) dst[2] |= hi >> (32u - sh);
                }
            }
        }
        stage_round(0, nc_first);
        if (!have_excl) { wait_offset(excl, base0); have_excl = true; }
        __syncwarp();
        const uint64_t my_excl = excl + pos0;  // global index of this warp's first survivor of the sub-tile
        uint32_t my_lim = my_cnt;
        if (p.limit >= 0) my_lim = my_excl >= (uint64_t)p.limit ? 0u : (uint32_t)min((uint64_t)my_cnt, (uint64_t)p.limit - my_excl);
        const uint64_t oexcl = my_excl - base0;  // where this run starts in the output buffers

        // ---- bit-packed columns: shift the staged bits to output bit `oexcl`; words fully owned by this warp's
        // run are stored, the (at most two) boundary words shared with neighbours are OR-ed in
        if (p.n_bits > 0 && my_lim != 0u) {
            const uint32_t sh = (uint32_t)oexcl & 31u;
            const uint64_t first_word = oexcl >> 5;
            const uint32_t n_words = (sh + my_lim + 31u) >> 5;  // <= 9
            const uint64_t end_bit = oexcl + my_lim;
            for (int b = 0; b < p.n_bits; ++b) {
                const uint32_t* wb = wbits + b * kWarpBitWords;
                uint32_t* out = p.bits[b].out;
                if ((uint32_t)lane < n_words) {
                    auto staged = [&](uint32_t i) -> uint32_t {
                        const uint32_t lo_bit = i * 32u;
                        if (lo_bit >= my_lim) return 0u;
                        uint32_t w = wb[i];
                        if (my_lim - lo_bit < 32u) w &= (1u << (my_lim - lo_bit)) - 1u;
                        return w;
                    };
                    const uint32_t t = (uint32_t)lane;
                    const uint32_t cur = staged(t);
                    const uint32_t prev = t > 0u ? staged(t - 1) : 0u;
                    const uint32_t val = __funnelshift_l(prev, cur, sh);
                    const uint64_t k = first_word + t;
                    const bool owned = (t > 0u || sh == 0u) && ((k + 1) * 32ull <= end_bit);
                    if (owned) out[k] = val;
                    else if (val != 0u) atomicOr(out + k, val);
                }
            }
        }
        // ---- 8-byte columns: contiguous copy-out of the staged survivors, kStageBufs columns per round
        for (int c0 = 0; c0 < p.n_col8; c0 += kStageBufs) {
            const int nc = (p.n_col8 - c0) < kStageBufs ? (p.n_col8 - c0) : kStageBufs;
            if (c0 > 0) {
                __syncwarp();  // the warp is done reading the previous round's staging buffers
                stage_round(c0, nc);
                __syncwarp();
            }
            for (int c = 0; c < nc; ++c) {
                const uint64_t* stage = wstage + c * 256;
                uint64_t* out = p.col8[c0 + c].out + oexcl;
                for (uint32_t i = lane; i < my_lim; i += 32) st_stream(out + i, stage[i]);
            }
        }
        __syncwarp();  // staging slices are reused by the warp's next sub-tile
    }
}

}  // namespace rvl
