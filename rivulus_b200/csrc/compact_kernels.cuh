// compact_kernels.cuh — second pass of the two-pass Filter + Select plan used for large batches.
//
// Replaces (reference, /root/reference/src) the per-column gather of the survivors:
//   eager     physical_plan/plan.rs:132-147   (filtered_data.push(col[i].clone()) per kept row, per column)
//   streaming execution/record_batch.rs:108-178 (take -> take_array: builder append per index, per column)
//             execution/array/bitmap.rs:142-155 (bit-at-a-time validity append)
//
// Pass 1 (fused_filter.cuh launched without columns) has already written, for the whole batch: the row-order
// selection bitmap, the exclusive output index of every 2048-row tile, and two tile lists — DENSE tiles (more than
// sparse_max survivors) and SPARSE tiles (1..sparse_max survivors).  No tile of this pass depends on another one.
//
//   compact_dense_kernel   persistent, one CTA per SM, warp-specialised.  A producer lane streams every projected
//                          8-byte column of its tiles through a ring of 16 KB shared-memory slots with TMA 1-D bulk
//                          copies (cp.async.bulk + mbarrier complete_tx): ~190 KB per SM in flight with no registers
//                          held, fully coalesced 128-byte DRAM bursts.  Eight consumer warps pick the survivors out of
//                          each slot (lane <-> row, rank = popc of the selection word below the lane) and store them
//                          at their final position: a warp store covers consecutive output addresses.
//                          Why whole tiles: B200 DRAM moves 128-byte lines, so at >= ~5 % selectivity nearly every
//                          line of a projected column is needed anyway (1 - 0.95^16 = 56 %, 1 - 0.9^16 = 81 %).
//   gather_sparse_kernel   high occupancy, one warp per sparse tile: survivors are listed in shared memory and the
//                          (survivor, column) pairs are spread over the lanes so that every gather load of the tile is
//                          in flight at once; dead lines are never touched.
//
//   compact_bits_kernel    bit-packed columns (validity bitmaps, Boolean values) of ALL tiles: one warp per tile, lane <-> two
//                          32-row words.  Every lane compresses its own words with a branch-free parallel-suffix bit compress
//                          (the five move masks depend only on the selection word, so they are built once per tile and shared
//                          by all bit columns), the 64 variable-length pieces are merged at their bit offsets in a per-warp
//                          shared buffer, interior output words are stored, the two boundary words OR-ed atomically.
#pragma once
#include "scan_kernels.cuh"

namespace rvl {

constexpr int kCompactMaxWarps = 16;                      // consumer warps of the dense kernel: 8 (256 rows each per tile) or 16 (128 rows)
constexpr uint32_t kSlotBytes = kTileRows * 8;            // one column tile
constexpr int kSparseCap = 640;                           // most survivors a "sparse" tile may hold
constexpr int kMaxBitSrc = kMaxCol8;                      // bitmaps a dense launch reads: the validity of each 8-byte column

struct CompactParams {
    int64_t n_rows;
    int64_t limit;                       // < 0 none; survivors whose global index >= limit are dropped
    const uint32_t* sel;                 // row-order selection words, whole tiles readable, zero beyond n_rows
    const uint64_t* tile_info;           // scan_kernels.cuh: (exclusive count inside the owning range << 12) | survivors
    const uint64_t* chunk_base;          // global exclusive output index of every range's first tile
    int64_t tiles_per_chunk;
    const unsigned long long* base_in;   // rows emitted by earlier batches of the same query (streaming) or nullptr
    const uint32_t* list;                // tile ids to process
    const uint32_t* list_count;          // device word: entries in `list`
    int32_t n_col8, n_bits;
    int32_t n_slots;                     // dense kernel: ring depth
    int32_t pad;
    Col8 col8[kMaxCol8];
    BitCol bits[kMaxBitCols];
    // dense kernel: every bitmap it reads, listed once; consumer warps stage the next tile's words of each through cp.async
    int32_t n_bsrc;
    int8_t col8_vsrc[kMaxCol8];          // index into bsrc of col8[c].valid, or -1 (no nulls)
    int32_t pad2;
    BitSrc bsrc[kMaxBitSrc];
};

__device__ __forceinline__ uint64_t tile_prefix_of(const CompactParams& p, int64_t tile) {
    return p.chunk_base[(uint32_t)tile / (uint32_t)p.tiles_per_chunk] + (p.tile_info[tile] >> kInfoShift);
}

__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// 4-byte asynchronous global -> shared copy; src_bytes = 0 zero-fills (words beyond the end of a bitmap)
__device__ __forceinline__ void cp_async_4(uint32_t smem_addr, const void* gmem, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_addr), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(bar_addr), "r"(parity)
            : "memory");
    } while (ok == 0u);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Exclusive ranks of the 64 selection words of a tile, computed redundantly by every warp:
// lane l holds words l (w0) and l + 32 (w1); returns their exclusive survivor offsets inside the tile.
__device__ __forceinline__ void tile_word_scan(uint32_t w0, uint32_t w1, int lane, uint32_t& e0, uint32_t& e1, uint32_t& total) {
    const uint32_t c0 = __popc(w0), c1 = __popc(w1);
    uint32_t i0 = c0, i1 = c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, i0, o), b = __shfl_up_sync(0xFFFFFFFFu, i1, o);
        if (lane >= o) { i0 += a; i1 += b; }
    }
    const uint32_t t0 = __shfl_sync(0xFFFFFFFFu, i0, 31);
    e0 = i0 - c0;
    e1 = t0 + i1 - c1;
    total = t0 + __shfl_sync(0xFFFFFFFFu, i1, 31);
}

template <int CW>
__global__ void __launch_bounds__((CW + 1) * 32, CW == 8 ? 2 : 1) __maxnreg__(CW == 8 ? 112 : 64) compact_dense_kernel(const __grid_constant__ CompactParams p) {
    constexpr int kCompactWarps = CW;
    constexpr int KW = kTileWords / CW;        // selection words (32-row groups) per consumer warp and tile
    constexpr int kWarpRows = kTileRows / CW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* const slots = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)p.n_slots * kSlotBytes);
    uint64_t* const empty = full + p.n_slots;
    // bitmap staging: [consumer warp][2 halves][n_bsrc][KW + 1 words]
    constexpr int kPer = KW + 1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_list = *p.list_count;
    if (blockIdx.x >= n_list) return;
    if (tid == 0) {
        for (int s = 0; s < p.n_slots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kCompactWarps); }
        mbar_init_fence();
    }
    __syncthreads();

    if (warp == kCompactWarps) {
        // ---------------------------------------------------------------- producer: one lane feeds the ring
        if (lane == 0 && p.n_col8 > 0) {
            int slot = 0;
            uint32_t round = 0;
            uint32_t next_id = p.list[blockIdx.x];
            for (uint32_t i = blockIdx.x; i < n_list; i += gridDim.x) {
                const int64_t tile = (int64_t)next_id;
                if (i + gridDim.x < n_list) next_id = p.list[i + gridDim.x];
                if (p.limit >= 0 && tile_prefix_of(p, tile) >= (uint64_t)p.limit) continue;  // consumers skip it too
                const int64_t row0 = tile * kTileRows;
                const bool whole = row0 + kTileRows <= p.n_rows;
                for (int c = 0; c < p.n_col8; ++c) {
                    if (round > 0u) mbar_wait(&empty[slot], (round - 1u) & 1u);  // consumers drained the previous use
                    if (whole && p.col8[c].vec_ok) tma_load_1d(slots + (size_t)slot * kTileRows, p.col8[c].in + row0, kSlotBytes, &full[slot]);
                    else mbar_arrive(&full[slot]);  // ragged tail / unaligned view: consumers read global memory themselves
                    if (++slot == p.n_slots) { slot = 0; ++round; }
                }
            }
        }
        return;
    }

    // -------------------------------------------------------------------- consumers
    const uint32_t lt = lanemask_lt();
    const uint64_t base0 = p.base_in != nullptr ? (uint64_t)*p.base_in : 0ull;
    const uint32_t tpc = (uint32_t)p.tiles_per_chunk;  // tile ids and ranges fit 32 bits (checked by the host)
    // shared-memory addresses as 32-bit offsets: slot s of the ring, this warp's 256 rows, this lane's row
    const uint32_t slot_addr0 = smem_u32(slots) + (uint32_t)warp * (kWarpRows * 8u) + (uint32_t)lane * 8u;
    const uint32_t full_addr0 = smem_u32(full), empty_addr0 = smem_u32(empty);
    int slot = 0;
    uint32_t phase = 0;
    // Bitmap words (validity of the 8-byte columns, values / validity of the bit-packed columns) of this warp's rows are staged
    // one tile ahead with 4-byte cp.async copies into a per-warp double buffer: no register is held and no consumer ever waits
    // on a dependent global load of a bitmap word.  Word j of source s covers bits [32 j, 32 j + 32) from the 32-bit word that
    // holds the warp's first row; a funnel shift by (bit0 & 31) aligns it (the warp's first row is a multiple of 32).
    const uint32_t bst_half = (uint32_t)p.n_bsrc * kPer * 4u;
    const uint32_t bst_addr0 = smem_u32(empty + p.n_slots) + (uint32_t)warp * 2u * bst_half;
    auto stage_bits = [&](uint32_t tile_id, uint32_t half) {
        const uint64_t wr0 = (uint64_t)tile_id * kTileRows + (uint64_t)warp * kWarpRows;
        const int n_task = p.n_bsrc * kPer;
        for (int t = lane; t < n_task; t += 32) {
            const int s = t / kPer, j = t - s * kPer;
            const BitSrc& b = p.bsrc[s];
            const uint64_t w = ((b.bit0 + wr0) >> 5) + (uint64_t)j;
            const bool ok = w < b.nwords;
            cp_async_4(bst_addr0 + half * bst_half + (uint32_t)t * 4u, b.words + (ok ? w : 0ull), ok ? 4u : 0u);
        }
        cp_async_commit();
    };
    auto staged32 = [&](int s, int k, uint32_t half) -> uint32_t {
        if (s < 0) return 0xFFFFFFFFu;
        const uint32_t a = bst_addr0 + half * bst_half + (uint32_t)(s * kPer + k) * 4u;
        return __funnelshift_r(lds32(a), lds32(a + 4u), (uint32_t)p.bsrc[s].bit0 & 31u);
    };
    uint32_t half = 0;
    // two-stage software prefetch: the tile id two iterations ahead, the selection words / prefix words one ahead.
    // Nothing loaded here is touched before the next iteration, so no consumer warp waits on these round trips.
    uint32_t nw0 = 0, nw1 = 0;
    uint64_t ninfo = 0, ncbase = 0;
    uint32_t ntile = p.list[blockIdx.x];
    uint32_t nntile = blockIdx.x + gridDim.x < n_list ? p.list[blockIdx.x + gridDim.x] : 0u;
    {
        const uint32_t* sw = p.sel + (size_t)ntile * kTileWords;
        nw0 = __ldg(sw + lane); nw1 = __ldg(sw + 32 + lane);
        ninfo = p.tile_info[ntile]; ncbase = p.chunk_base[ntile / tpc];
        if (p.n_bsrc > 0) stage_bits(ntile, 0u);
    }
#pragma unroll 1
    for (uint32_t i = blockIdx.x; i < n_list; i += gridDim.x) {
        const int64_t tile = (int64_t)ntile;
        const uint32_t w0 = nw0, w1 = nw1;
        const uint64_t prefix = ncbase + (ninfo >> kInfoShift);
        const uint32_t cur = half;
        if (p.n_bsrc > 0) {
            cp_async_wait_all();   // this tile's bitmap words (issued one iteration ago)
            __syncwarp();          // ... of every lane; and nobody still reads the other half
            half ^= 1u;
        }
        if (i + gridDim.x < n_list) {
            ntile = nntile;
            const uint32_t* sw = p.sel + (size_t)ntile * kTileWords;
            nw0 = __ldg(sw + lane); nw1 = __ldg(sw + 32 + lane);
            ninfo = p.tile_info[ntile]; ncbase = p.chunk_base[ntile / tpc];
            if (i + 2 * gridDim.x < n_list) nntile = p.list[i + 2 * gridDim.x];
            if (p.n_bsrc > 0) stage_bits(ntile, half);
        }
        if (p.limit >= 0 && prefix >= (uint64_t)p.limit) continue;  // tile lies entirely beyond the limit (no slots were filled)
        const int64_t row0 = tile * kTileRows;
        const int64_t wrow0 = row0 + (int64_t)warp * kWarpRows;  // this warp's rows = KW selection words
        const bool whole = row0 + kTileRows <= p.n_rows;

        // survivors of the tile in front of this warp's rows: two warp REDUX.ADDs instead of a 64-word scan
        const uint32_t c0 = __popc(w0), c1 = __popc(w1);
        uint32_t wfirst;
        const int first_word = warp * KW;
        if (first_word < 32) wfirst = __reduce_add_sync(0xFFFFFFFFu, lane < first_word ? c0 : 0u);
        else wfirst = __reduce_add_sync(0xFFFFFFFFu, c0) + __reduce_add_sync(0xFFFFFFFFu, lane < first_word - 32 ? c1 : 0u);
        const uint32_t wsel = first_word < 32 ? w0 : w1;
        // survivors with a tile-local rank >= lim_rel lie beyond the LIMIT
        const uint32_t lim_rel = p.limit < 0 ? 0xFFFFFFFFu : (uint32_t)min((uint64_t)p.limit - prefix, (uint64_t)0xFFFFFFFFu);
        uint32_t selw[KW];
        uint32_t off8[KW];    // byte offset of this lane's row (word k) from the tile's first output element
        bool kp[KW];          // this lane's row of word k survives (and lies below the limit)
        uint32_t wcnt = 0;    // survivors of this warp
#pragma unroll
        for (int k = 0; k < KW; ++k) {
            selw[k] = __shfl_sync(0xFFFFFFFFu, wsel, (first_word + k) & 31);
            const uint32_t r = wfirst + wcnt + __popc(selw[k] & lt);
            wcnt += __popc(selw[k]);
            kp[k] = ((selw[k] >> lane) & 1u) != 0u && r < lim_rel;
            off8[k] = r * 8u;
        }
        const uint64_t obase = prefix - base0;  // output index of the tile's first survivor

        // ---- 8-byte columns: one ring slot each
        for (int c = 0; c < p.n_col8; ++c) {
            const Col8& col = p.col8[c];
            const bool via_tma = whole && col.vec_ok != 0;
            uint64_t v[KW];
            mbar_wait_addr(full_addr0 + (uint32_t)slot * 8u, phase);
            if (via_tma) {
                const uint32_t src = slot_addr0 + (uint32_t)slot * kSlotBytes;
#pragma unroll
                for (int k = 0; k < KW; ++k) v[k] = lds64(src + k * 256);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_addr(empty_addr0 + (uint32_t)slot * 8u);  // slot may be refilled
            if (++slot == p.n_slots) { slot = 0; phase ^= 1u; }
            if (!via_tma) {
#pragma unroll
                for (int k = 0; k < KW; ++k) {
                    v[k] = 0ull;
                    if (kp[k]) v[k] = ld_stream(col.in + wrow0 + k * 32 + lane);
                }
            }
            const int vsrc = p.col8_vsrc[c];
            if (vsrc >= 0) {  // placeholder 0 under a null (primitive.rs:175-178)
#pragma unroll
                for (int k = 0; k < KW; ++k) {
                    const uint32_t vm = staged32(vsrc, k, cur);
                    if (((vm >> lane) & 1u) == 0u) v[k] = 0ull;
                }
            }
            // one 64-bit base per item, 32-bit byte offsets per row (kept opaque so the adds are not re-associated)
            char* ob = reinterpret_cast<char*>(col.out + obase);
            asm volatile("" : "+l"(ob));
#pragma unroll
            for (int k = 0; k < KW; ++k)
                if (kp[k]) st_stream(reinterpret_cast<uint64_t*>(ob + off8[k]), v[k]);
        }

    }
}

static __global__ void __launch_bounds__(kBlock) gather_sparse_kernel(const __grid_constant__ CompactParams p) {
    __shared__ uint16_t s_rows[kWarps][kSparseCap];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_list = *p.list_count;
    const uint64_t base0 = p.base_in != nullptr ? (uint64_t)*p.base_in : 0ull;
    uint16_t* rows = s_rows[warp];
    // two-stage software prefetch (as in the dense kernel): the tile id two iterations ahead, its selection words and prefix words
    // one ahead — a sparse tile is a chain of dependent round trips (id -> words -> rows -> values), the first two now overlap
    // with the previous tile's gather
    const uint32_t stride = gridDim.x * kWarps;
    uint32_t i = blockIdx.x * kWarps + warp;
    uint32_t ntile = i < n_list ? p.list[i] : 0u;
    uint32_t nntile = i + stride < n_list ? p.list[i + stride] : 0u;
    uint32_t nw0 = 0, nw1 = 0;
    uint64_t ninfo = 0, ncbase = 0;
    if (i < n_list) {
        const uint32_t* sw = p.sel + (size_t)ntile * kTileWords;
        nw0 = __ldg(sw + lane); nw1 = __ldg(sw + 32 + lane);
        ninfo = p.tile_info[ntile]; ncbase = p.chunk_base[ntile / (uint32_t)p.tiles_per_chunk];
    }
#pragma unroll 1
    for (; i < n_list; i += stride) {
        const int64_t tile = (int64_t)ntile;
        const int64_t row0 = tile * kTileRows;
        const uint32_t w0 = nw0, w1 = nw1;
        const uint64_t prefix = ncbase + (ninfo >> kInfoShift);
        if (i + stride < n_list) {
            ntile = nntile;
            const uint32_t* sw = p.sel + (size_t)ntile * kTileWords;
            nw0 = __ldg(sw + lane); nw1 = __ldg(sw + 32 + lane);
            ninfo = p.tile_info[ntile]; ncbase = p.chunk_base[ntile / (uint32_t)p.tiles_per_chunk];
            if (i + 2 * stride < n_list) nntile = p.list[i + 2 * stride];
        }
        uint32_t e0, e1, total;
        tile_word_scan(w0, w1, lane, e0, e1, total);
        if (total > (uint32_t)kSparseCap) total = kSparseCap;  // cannot happen: pass 1 classifies with sparse_max <= kSparseCap
        uint32_t lim = total;
        if (p.limit >= 0) lim = prefix >= (uint64_t)p.limit ? 0u : (uint32_t)min((uint64_t)total, (uint64_t)p.limit - prefix);
        __syncwarp();  // previous tile's list is no longer read
        for (uint32_t w = w0, r = e0; w != 0u; w &= w - 1u, ++r)
            if (r < (uint32_t)kSparseCap) rows[r] = (uint16_t)(lane * 32 + (__ffs(w) - 1));
        for (uint32_t w = w1, r = e1; w != 0u; w &= w - 1u, ++r)
            if (r < (uint32_t)kSparseCap) rows[r] = (uint16_t)((lane + 32) * 32 + (__ffs(w) - 1));
        __syncwarp();
        const uint64_t obase = prefix - base0;

        // (column, survivor) pairs over the lanes, four rounds of loads in flight before the first store
        const uint32_t n_task = lim * (uint32_t)p.n_col8;
        for (uint32_t t0 = 0; t0 < n_task; t0 += 128) {
            uint64_t v[4];
            uint64_t* dst[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t t = t0 + u * 32 + lane;
                v[u] = 0ull; dst[u] = nullptr;
                if (t < n_task) {
                    const uint32_t c = t / lim, e = t - c * lim;
                    const Col8& col = p.col8[c];
                    const int64_t row = row0 + rows[e];
                    bool ok = true;
                    if (col.valid.words != nullptr) { const uint64_t bit = col.valid.bit0 + (uint64_t)row; ok = (__ldg(col.valid.words + (bit >> 5)) >> (bit & 31)) & 1u; }
                    if (ok) v[u] = ld_gather(col.in + row);  // 64-byte DRAM granule; placeholder 0 under a null (primitive.rs:175-178)
                    dst[u] = col.out + obase + e;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (dst[u] != nullptr) st_stream(dst[u], v[u]);
        }
    }
}

// ---- bit-packed columns ---------------------------------------------------------------------------------------------------
// Branch-free bit compress (parallel suffix, Hacker's Delight 7-4): mv[i] are the bits that move right by 2^i in round i.
struct CompressPlan {
    uint32_t m;
    uint32_t mv[5];
};
__device__ __forceinline__ CompressPlan compress_plan(uint32_t m) {
    CompressPlan c;
    c.m = m;
    uint32_t mk = ~m << 1;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        uint32_t mp = mk ^ (mk << 1);
        mp ^= mp << 2; mp ^= mp << 4; mp ^= mp << 8; mp ^= mp << 16;
        const uint32_t mv = mp & m;
        c.mv[i] = mv;
        m = (m ^ mv) | (mv >> (1 << i));
        mk &= ~mp;
    }
    return c;
}
// the bits of x selected by the plan's mask, packed towards bit 0 in order
__device__ __forceinline__ uint32_t compress_bits(uint32_t x, const CompressPlan& c) {
    x &= c.m;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const uint32_t t = x & c.mv[i];
        x = (x ^ t) | (t >> (1 << i));
    }
    return x;
}

// One tile of one bit-packed column: lane l holds the compressed survivors of words l (x0, n0 bits kept, tile rank e0) and l + 32
// (x1, n1, e1).  The 64 variable-length pieces are merged at their bit offsets in `buf` (kTileWords + 2 words of this warp) and
// stored at output bit `obase`: interior words plainly, the two boundary words OR-ed atomically (the neighbouring tiles own the rest).
__device__ __forceinline__ void emit_bit_tile(uint32_t* buf, int lane, uint32_t x0, uint32_t x1, uint32_t e0, uint32_t e1, uint32_t tot,
                                              uint64_t obase, uint32_t* out) {
    const uint32_t sh = (uint32_t)obase & 31u;
    const uint64_t first_word = obase >> 5;
    const uint32_t end = sh + tot;                    // bits [sh, end) of the tile's output words are ours
    const uint32_t n_words = (end + 31u) >> 5;        // <= 65
    buf[lane] = 0u; buf[lane + 32] = 0u;
    if (lane < 2) buf[64 + lane] = 0u;
    __syncwarp();
    if (x0 != 0u) {
        const uint32_t q = sh + e0, s = q & 31u;
        atomicOr(&buf[q >> 5], x0 << s);
        if (s != 0u && (x0 >> (32u - s)) != 0u) atomicOr(&buf[(q >> 5) + 1u], x0 >> (32u - s));
    }
    if (x1 != 0u) {
        const uint32_t q = sh + e1, s = q & 31u;
        atomicOr(&buf[q >> 5], x1 << s);
        if (s != 0u && (x1 >> (32u - s)) != 0u) atomicOr(&buf[(q >> 5) + 1u], x1 >> (32u - s));
    }
    __syncwarp();
    for (uint32_t w = (uint32_t)lane; w < n_words; w += 32u) {
        const uint32_t v = buf[w], lo = w * 32u;
        const bool owned = (w > 0u || sh == 0u) && (lo + 32u <= end);
        uint32_t* o = out + first_word + w;
        if (owned) *o = v;
        else if (v != 0u) atomicOr(o, v);
    }
    __syncwarp();
}

static __global__ void __launch_bounds__(kBlock) compact_bits_kernel(const __grid_constant__ CompactParams p) {
    __shared__ uint32_t s_out[kWarps][kTileWords + 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_tiles = (p.n_rows + kTileRows - 1) / kTileRows;
    const uint64_t base0 = p.base_in != nullptr ? (uint64_t)*p.base_in : 0ull;
    uint32_t* const buf = s_out[warp];
    // software prefetch: the next tile's info word, range base and selection words are requested one iteration ahead
    const int64_t stride = (int64_t)gridDim.x * kWarps;
    int64_t tile = (int64_t)blockIdx.x * kWarps + warp;
    uint64_t ninfo = 0, nbase = 0;
    uint32_t nw0 = 0, nw1 = 0;
    if (tile < n_tiles) {
        ninfo = p.tile_info[tile]; nbase = p.chunk_base[(uint32_t)tile / (uint32_t)p.tiles_per_chunk];
        const uint32_t* sw = p.sel + tile * kTileWords;
        nw0 = __ldg(sw + lane); nw1 = __ldg(sw + 32 + lane);
    }
#pragma unroll 1
    for (; tile < n_tiles; tile += stride) {
        const uint64_t info = ninfo;
        const uint64_t cbase = nbase;
        const uint32_t w0 = nw0, w1 = nw1;
        if (tile + stride < n_tiles) {
            const int64_t nt = tile + stride;
            ninfo = p.tile_info[nt]; nbase = p.chunk_base[(uint32_t)nt / (uint32_t)p.tiles_per_chunk];
            const uint32_t* sw = p.sel + nt * kTileWords;
            nw0 = __ldg(sw + lane); nw1 = __ldg(sw + 32 + lane);
        }
        if ((info & ((1ull << kInfoShift) - 1ull)) == 0ull) continue;  // no survivor in this tile
        const uint64_t prefix = cbase + (info >> kInfoShift);
        if (p.limit >= 0 && prefix >= (uint64_t)p.limit) continue;
        uint32_t e0, e1, total;
        tile_word_scan(w0, w1, lane, e0, e1, total);
        // survivors with a tile-local rank >= lim_rel lie beyond the LIMIT
        const uint32_t lim_rel = p.limit < 0 ? 0xFFFFFFFFu : (uint32_t)min((uint64_t)p.limit - prefix, (uint64_t)0xFFFFFFFFu);
        const uint32_t n0 = e0 >= lim_rel ? 0u : min((uint32_t)__popc(w0), lim_rel - e0);
        const uint32_t n1 = e1 >= lim_rel ? 0u : min((uint32_t)__popc(w1), lim_rel - e1);
        const uint32_t tot = min(total, lim_rel);
        const uint64_t obase = prefix - base0;            // output bit index of the tile's first survivor
        const CompressPlan c0 = compress_plan(w0), c1 = compress_plan(w1);
        const uint64_t r0 = (uint64_t)tile * kTileRows + (uint64_t)lane * 32u, r1 = r0 + 1024u;
        for (int b = 0; b < p.n_bits; ++b) {
            const BitCol& bc = p.bits[b];
            uint32_t x0 = compress_bits(load_bits32(bc.in, r0) & load_bits32(bc.mask, r0), c0);
            uint32_t x1 = compress_bits(load_bits32(bc.in, r1) & load_bits32(bc.mask, r1), c1);
            if (n0 < 32u) x0 &= (1u << n0) - 1u;
            if (n1 < 32u) x1 &= (1u << n1) - 1u;
            emit_bit_tile(buf, lane, x0, x1, e0, e1, tot, obase, bc.out);
        }
    }
}

}  // namespace rvl
