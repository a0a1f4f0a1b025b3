// fused_filter.cuh — K1+K2+K3: predicate -> ballot/popc -> block scan -> decoupled look-back ->
// ordered stream compaction of every projected fixed-width and bit-packed column, in ONE pass.
//
// Replaces (reference, /root/reference/src):
//   eager     physical_plan/plan.rs:112-130 (mask loop) + :132-147 (per-column gather)
//   streaming execution/record_batch.rs:235-240 (mask -> indices) + :131-178 (take_array builders)
//             execution/array/bitmap.rs:142-155 (bit-at-a-time BitmapBuilder::append)
//
// Data layout in HBM: Arrow columns — 8-byte values, LSB-first bitmaps (1 = valid / true), any
// (offset, length) view.  A tile is 2048 consecutive rows; warp w owns rows [256w, 256w+256) of the
// tile as 4 groups of 64 rows; lane l of a group owns rows 2l and 2l+1, i.e. ONE 128-bit load per
// 8-byte column per group.  Two ballots per group (even rows / odd rows) give the selection masks;
// popc of the lower lanes gives each lane its rank in row order.
//
// HBM roofline: per row the kernel must read the predicate value (8 B [+1 bit
// validity]), the 32-B sectors of each projected column that hold at least one survivor, and write
// s*8 B per projected column.  Dead sectors are never requested: gather loads are predicated on the
// lane's own keep bits, and a lane's two rows sit in one sector.
#pragma once
#include "device_utils.cuh"

namespace rvl {

constexpr int kMaxCol8 = 8;      // 8-byte columns compacted per launch
constexpr int kMaxBitCols = 16;  // bit-packed columns compacted per launch
constexpr int kSparseTile = 32;  // tiles with <= this many survivors scatter straight from registers

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
    int64_t limit;  // < 0: none; else survivors whose global output index >= limit are dropped
    // comparison predicate: keep = valid ? truth[cmp(value, literal)] : keep_null
    const uint64_t* pred_values;
    int64_t lit_bits;
    BitSrc pred_valid;
    uint32_t truth;      // bit0: value < lit, bit1: ==, bit2: >, bit3: unordered (NaN)
    uint32_t keep_null;  // 0 / 1: what a null row evaluates to (series.rs:105-107 puts Null below everything)
    int32_t pred_vec_ok;
    int32_t n_col8;
    int32_t n_bits;
    // bitmap predicate: sel = (valid & ((vals & a) ^ b)) | (~valid & keep_null)
    uint32_t pb_a, pb_b;
    int32_t pad0;
    BitSrc pb_vals;
    Col8 col8[kMaxCol8];
    BitCol bits[kMaxBitCols];
    uint64_t* tile_status;               // one descriptor per tile, zeroed before the launch
    const unsigned long long* base_in;   // rows already emitted by earlier batches of the same query (streaming), or nullptr.
                                         // The limit applies to base + rank; output buffers are indexed by rank alone.
    unsigned long long* total_out;       // base + survivors of this launch (saturates at >= limit once the limit trips)
    uint32_t* done_flag;                 // set once some tile's inclusive prefix reaches the limit
    uint32_t* sel_out;                   // optional: row-order selection bitmap (n_rows bits, 8-byte aligned)
    uint64_t* tile_prefix_out;           // optional: exclusive output prefix of every tile
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

template <int PRED>
__global__ void __launch_bounds__(kBlock, 4) fused_filter_project_kernel(const __grid_constant__ FusedParams p) {
    __shared__ __align__(16) uint64_t s_stage[2][kTileRows];
    __shared__ uint32_t s_bits[kMaxBitCols][kTileWords + 2];
    __shared__ uint32_t s_warp_count[kWarps];
    __shared__ uint64_t s_excl;
    __shared__ uint64_t s_base;
    __shared__ uint32_t s_done;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t warp_row0 = tile * kTileRows + (int64_t)warp * (kGroups * 64);

    // LIMIT early termination (streaming.rs:269-271 / plan.rs:163): once an earlier tile's inclusive
    // prefix reached the limit, later tiles publish a saturated prefix and leave without touching HBM.
    if (p.limit >= 0) {
        if (tid == 0) s_done = ld_relaxed_gpu_u32(p.done_flag);
        __syncthreads();
        if (s_done != 0u) {
            if (tid == 0) {
                st_relaxed_gpu(p.tile_status + tile, kStatusPrefix | (uint64_t)p.limit);
                if (p.tile_prefix_out) p.tile_prefix_out[tile] = (uint64_t)p.limit;
                if (tile == (int64_t)gridDim.x - 1) *p.total_out = (unsigned long long)p.limit;
            }
            return;
        }
    }

    if (p.n_bits > 0) {
        uint32_t* flat = &s_bits[0][0];
        for (int i = tid; i < p.n_bits * (kTileWords + 2); i += kBlock) flat[i] = 0u;
    }

    // ---------------------------------------------------------------- phase 1: predicate + ranks
    uint32_t kb[kGroups];     // keep bits of this lane's two rows
    uint32_t rank0[kGroups];  // rank (within the warp's 256 rows) of the lane's first surviving row
    uint32_t wrun = 0;

    if (PRED == kPredI64 || PRED == kPredF64) {
        uint64_t v[kGroups][2];
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
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
            const int64_t grow = warp_row0 + g * 64;
            const int64_t row = grow + lane * 2;
            const uint32_t exists = (row < p.n_rows ? 1u : 0u) | (row + 1 < p.n_rows ? 2u : 0u);
            uint32_t vb = 3u;
            if (p.pred_valid.words != nullptr) vb = (uint32_t)(load_bits64(p.pred_valid, (uint64_t)grow) >> (2 * lane)) & 3u;
            const uint32_t c0 = (p.truth & cmp_code<PRED>(v[g][0], p.lit_bits)) != 0u ? 1u : 0u;
            const uint32_t c1 = (p.truth & cmp_code<PRED>(v[g][1], p.lit_bits)) != 0u ? 2u : 0u;
            const uint32_t kn = p.keep_null ? 3u : 0u;
            kb[g] = (((c0 | c1) & vb) | (kn & ~vb)) & exists;
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

#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
        const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, (kb[g] & 1u) != 0u);
        const uint32_t m1 = __ballot_sync(0xFFFFFFFFu, (kb[g] & 2u) != 0u);
        const uint32_t lt = lanemask_lt();
        rank0[g] = wrun + __popc(m0 & lt) + __popc(m1 & lt);
        wrun += __popc(m0) + __popc(m1);
        if (p.sel_out != nullptr && lane == 0) {
            const int64_t grow = warp_row0 + g * 64;
            if (grow < p.n_rows) *reinterpret_cast<uint64_t*>(p.sel_out + (grow >> 5)) = interleave_masks(m0, m1);
        }
    }
    if (lane == 0) s_warp_count[warp] = wrun;
    __syncthreads();

    uint32_t warp_off = 0, cnt = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const uint32_t c = s_warp_count[w];
        warp_off += (w < warp) ? c : 0u;
        cnt += c;
    }

    // ---------------------------------------------------------------- global order: decoupled look-back
    if (warp == 0) {
        uint64_t excl;
        const uint64_t base0 = p.base_in != nullptr ? (uint64_t)*p.base_in : 0ull;
        if (tile == 0) {
            excl = base0;
            if (lane == 0) st_relaxed_gpu(p.tile_status, kStatusPrefix | (excl + cnt));
        } else {
            if (lane == 0) st_relaxed_gpu(p.tile_status + tile, kStatusAggregate | (uint64_t)cnt);
            excl = lookback_exclusive(p.tile_status, tile, lane);
            if (lane == 0) st_relaxed_gpu(p.tile_status + tile, kStatusPrefix | (excl + cnt));
        }
        if (lane == 0) {
            s_excl = excl;
            s_base = base0;
            if (p.tile_prefix_out != nullptr) p.tile_prefix_out[tile] = excl;
            if (tile == (int64_t)gridDim.x - 1) *p.total_out = (unsigned long long)(excl + cnt);
            if (p.limit >= 0 && excl + cnt >= (uint64_t)p.limit) atomicExch(p.done_flag, 1u);
        }
    }
    __syncthreads();
    const uint64_t excl = s_excl;
    uint32_t cnt_lim = cnt;
    if (p.limit >= 0) cnt_lim = excl >= (uint64_t)p.limit ? 0u : (uint32_t)min((uint64_t)cnt, (uint64_t)p.limit - excl);
    if (cnt_lim == 0u) return;
    const uint64_t oexcl = excl - s_base;  // where this tile's survivors start in the output buffers

    // ---------------------------------------------------------------- phase 2a: sparse tile — scatter from registers
    if (cnt_lim <= (uint32_t)kSparseTile) {
        for (int b = 0; b < p.n_bits; ++b) {
            const BitCol bc = p.bits[b];
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                if (kb[g] == 0u) continue;
                const int64_t grow = warp_row0 + g * 64;
                const uint64_t w = load_bits64(bc.in, (uint64_t)grow) & load_bits64(bc.mask, (uint64_t)grow);
                const uint32_t lb = (uint32_t)(w >> (2 * lane)) & 3u;
                uint32_t r = warp_off + rank0[g];
                if ((kb[g] & 1u) != 0u) {
                    if ((lb & 1u) != 0u && r < cnt_lim) { const uint64_t pos = oexcl + r; atomicOr(bc.out + (pos >> 5), 1u << (pos & 31)); }
                    ++r;
                }
                if ((kb[g] & 2u) != 0u && (lb & 2u) != 0u && r < cnt_lim) { const uint64_t pos = oexcl + r; atomicOr(bc.out + (pos >> 5), 1u << (pos & 31)); }
            }
        }
        for (int c = 0; c < p.n_col8; ++c) {
            const Col8 col = p.col8[c];
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                if (kb[g] == 0u) continue;
                const int64_t grow = warp_row0 + g * 64;
                const int64_t row = grow + lane * 2;
                uint32_t vb = 3u;
                if (col.valid.words != nullptr) vb = (uint32_t)(load_bits64(col.valid, (uint64_t)grow) >> (2 * lane)) & 3u;
                uint32_t r = warp_off + rank0[g];
                if ((kb[g] & 1u) != 0u) {
                    if (r < cnt_lim) st_stream(col.out + oexcl + r, (vb & 1u) ? __ldg(col.in + row) : 0ull);
                    ++r;
                }
                if ((kb[g] & 2u) != 0u && r < cnt_lim) st_stream(col.out + oexcl + r, (vb & 2u) ? __ldg(col.in + row + 1) : 0ull);
            }
        }
        return;
    }

    // ---------------------------------------------------------------- phase 2b: bit-packed columns (K3)
    if (p.n_bits > 0) {
        for (int b = 0; b < p.n_bits; ++b) {
            const BitCol bc = p.bits[b];
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                const int64_t grow = warp_row0 + g * 64;
                const uint32_t gbase = __shfl_sync(0xFFFFFFFFu, rank0[g], 0);  // lane 0 has no lower lanes
                const uint64_t w = load_bits64(bc.in, (uint64_t)grow) & load_bits64(bc.mask, (uint64_t)grow);
                const uint32_t lb = (uint32_t)(w >> (2 * lane)) & 3u & kb[g];
                const uint32_t lp = rank0[g] - gbase;  // rank inside the group, < 64
                const uint64_t contrib = ((uint64_t)(lb & 1u) << lp) | ((uint64_t)((lb >> 1) & 1u) << (lp + (kb[g] & 1u)));
                uint32_t lo = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)contrib);
                uint32_t hi = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)(contrib >> 32));
                if (lane == 0) {
                    const uint32_t P = warp_off + gbase;  // tile-local output position of the group's first survivor
                    if (P < cnt_lim) {
                        const uint32_t room = cnt_lim - P;
                        if (room < 64u) {
                            const uint64_t keep = (1ull << room) - 1ull;
                            lo &= (uint32_t)keep; hi &= (uint32_t)(keep >> 32);
                        }
                        const uint32_t sh = P & 31u;
                        uint32_t* dst = &s_bits[b][P >> 5];
                        const uint32_t o0 = lo << sh;
                        const uint32_t o1 = __funnelshift_l(lo, hi, sh);
                        const uint32_t o2 = sh != 0u ? (hi >> (32u - sh)) : 0u;
                        if (o0) atomicOr(dst, o0);
                        if (o1) atomicOr(dst + 1, o1);
                        if (o2) atomicOr(dst + 2, o2);
                    }
                }
            }
        }
        __syncthreads();
        // shift the tile's staged bits to the output bit position `oexcl`; words fully owned by this
        // tile are stored, the (at most two) boundary words shared with neighbours are OR-ed in.
        const uint32_t sh = (uint32_t)oexcl & 31u;
        const uint64_t first_word = oexcl >> 5;
        const uint32_t n_words = (sh + cnt_lim + 31u) >> 5;
        const uint64_t end_bit = oexcl + cnt_lim;
        for (int b = 0; b < p.n_bits; ++b) {
            uint32_t* out = p.bits[b].out;
            for (uint32_t t = tid; t < n_words; t += kBlock) {
                const uint32_t cur = s_bits[b][t];
                const uint32_t prev = t > 0u ? s_bits[b][t - 1] : 0u;
                const uint32_t val = __funnelshift_l(prev, cur, sh);
                const uint64_t k = first_word + t;
                const bool owned = (t > 0u || sh == 0u) && ((k + 1) * 32ull <= end_bit);
                if (owned) out[k] = val;
                else if (val != 0u) atomicOr(out + k, val);
            }
        }
    }

    // ---------------------------------------------------------------- phase 2c: 8-byte columns (K2)
    for (int c = 0; c < p.n_col8; ++c) {
        const Col8 col = p.col8[c];
        uint64_t* stage = s_stage[c & 1];
        uint64_t v[kGroups][2];
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
            v[g][0] = 0; v[g][1] = 0;
            if (kb[g] != 0u) {
                const int64_t row = warp_row0 + g * 64 + lane * 2;
                if (col.vec_ok && row + 1 < p.n_rows) {
                    const ulonglong2 t = ld_stream_v2(col.in + row); v[g][0] = t.x; v[g][1] = t.y;
                } else {
                    if ((kb[g] & 1u) != 0u) v[g][0] = ld_stream(col.in + row);
                    if ((kb[g] & 2u) != 0u) v[g][1] = ld_stream(col.in + row + 1);
                }
            }
        }
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
            if (kb[g] != 0u) {
                uint32_t vb = 3u;
                if (col.valid.words != nullptr)
                    vb = (uint32_t)(load_bits64(col.valid, (uint64_t)(warp_row0 + g * 64)) >> (2 * lane)) & 3u;
                uint32_t r = warp_off + rank0[g];
                if ((kb[g] & 1u) != 0u) {
                    if (r < cnt_lim) stage[r] = (vb & 1u) ? v[g][0] : 0ull;  // placeholder 0 under a null (primitive.rs:175-178)
                    ++r;
                }
                if ((kb[g] & 2u) != 0u && r < cnt_lim) stage[r] = (vb & 2u) ? v[g][1] : 0ull;
            }
        }
        __syncthreads();
        uint64_t* out = col.out + oexcl;
        for (uint32_t i = tid; i < cnt_lim; i += kBlock) st_stream(out + i, stage[i]);
        // no barrier here: the next column stages into the other buffer, and the barrier after that
        // staging orders this read against the column after next.
    }
}

}  // namespace rvl
