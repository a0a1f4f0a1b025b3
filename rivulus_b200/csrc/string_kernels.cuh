// string_kernels.cuh — K4: StringArray compaction (offset prefix-sum + byte copy) and the string
// comparison predicate.
//
// Replaces (reference, /root/reference/src):
//   execution/record_batch.rs:163-170  take_array(String): value(i).to_string() per survivor, then
//   execution/array/string.rs:19-58    StringArray::new: second copy + offsets rebuild + UTF-8 re-validation
//   datatypes/series.rs:111            str ordering for `col <op> "literal"` (byte-wise lexicographic)
//
// The predicate pass leaves a row-order selection bitmap and each tile's exclusive output row prefix.  Two kernels follow,
// the north star's "offset prefix-sum, then a byte copy staged through shared memory":
//   string_sizes_kernel   one warp per 2048-row tile, lane <-> row: survivor byte total of every tile; the last CTA to finish
//                         (ticket) turns the totals into exclusive prefixes.  No tile waits on another one.
//   string_gather_kernel  one CTA per tile: survivors ranked from the selection words, lengths -> byte offsets by one block scan,
//                         new int32 offsets written in rank order, bytes copied through a shared staging buffer.
// (A single-pass version with a decoupled look-back over the byte totals was measured first: with ~900 tiles in flight and
// ~10 us per tile, every tile waited on the slowest of its predecessors — the look-back cost 0.2 ms of a 0.65 ms launch.)
// UTF-8 validity is preserved by construction (whole strings are copied), so no re-validation pass.
#pragma once
#include "fused_filter.cuh"   // TMA 1-D bulk copy + mbarrier helpers

namespace rvl {

struct StrGatherParams {
    int64_t n_rows;
    int64_t limit;                      // < 0 none
    const uint32_t* sel;                // row-order selection words from the fused kernel; nullptr = every row (concat)
    const uint64_t* tile_prefix;        // exclusive output row index of each tile; nullptr = row_base + tile * 2048
    const uint64_t* chunk_base;         // two-pass plan: tile_prefix holds scan_kernels.cuh tile_info words and the row index is
    int64_t tiles_per_chunk;            //   chunk_base[tile / tiles_per_chunk] + (tile_prefix[tile] >> 12); nullptr = plain prefixes
    int64_t row_base;
    const unsigned long long* row_base_in;  // rows emitted by earlier batches (streaming): tile_prefix/limit are global, output index = rank - base
    const int32_t* offsets;             // row 0 of the view
    const uint8_t* data;
    BitSrc valid;
    int32_t* out_offsets;               // out_offsets[0] preset; this kernel writes [rank + 1]
    uint8_t* out_data;
    uint64_t* tile_bytes;               // [n_tiles + 1]: string_sizes_kernel leaves each tile's exclusive byte prefix here; the last
                                        //   word is the ticket counter of that kernel (zeroed before launch)
    const unsigned long long* byte_base_in;  // bytes already emitted before this launch (concat), or nullptr
    unsigned long long* bytes_total_out;     // byte base + bytes emitted by this launch
    // ---- round-2 kernels (string_sizes_ranges_kernel / string_gather_staged_kernel): 1024-row sub-tiles
    uint64_t* sub_bytes;                // [2 * n_tiles]: survivor bytes in front of each sub-tile inside its warp range
    uint64_t* range_bytes;              // [n_ranges + 1]: global byte base of each warp range (last CTA scans them); [n_ranges] = ticket
    int64_t tiles_per_range;            // sizes kernel: tiles per warp range
    int32_t n_ranges;
    uint32_t dense_min;                 // sub-tiles with at least this many survivors stream their whole source block through shared memory
    const uint8_t* data_lo;             // readable bounds of the data buffer (the TMA block is rounded to 16 bytes inside them)
    const uint8_t* data_hi;
};

// exclusive scan of one uint32 per thread across the 256-thread block; returns the thread's prefix, sets total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp /*[kWarps]*/, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const uint32_t c = s_warp[w];
        woff += (w < warp) ? c : 0u;
        tot += c;
    }
    __syncthreads();  // s_warp may be reused by the caller
    total = tot;
    return woff + incl - v;
}

// predicated byte moves with the predicate, the immediate offset and the 32-bit shared address spelled out: left to the compiler
// the address arithmetic is re-done under every predicate and the shared pointer goes through a generic-address conversion
template <int IMM>
__device__ __forceinline__ uint32_t ldg_u8_lt(const uint8_t* p, uint32_t k, uint32_t n) {
    uint32_t v;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %2, %3;\n\t@q ld.global.nc.u8 %0, [%1+%4];\n\t}" : "=r"(v) : "l"(p), "r"(k), "r"(n), "n"(IMM));
    return v;
}
template <int IMM>
__device__ __forceinline__ void sts_u8_lt(uint32_t addr, uint32_t v, uint32_t k, uint32_t n) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %2, %3;\n\t@q st.shared.u8 [%0+%4], %1;\n\t}" ::"r"(addr), "r"(v), "r"(k), "r"(n), "n"(IMM) : "memory");
}

// One warp copies the (clipped) strings described by its 32 lanes — lane i: my_n bytes from data + my_s to shared address
// stage_addr + my_d — with L lanes per string: 32 / L strings per step, lane `sub` of a string moves bytes sub, sub + L, ...;
// the first five rounds are unrolled with every load issued before the first store (5 x 32/L sectors in flight per warp),
// longer strings loop on.
template <int L>
__device__ __forceinline__ void copy_descriptors(const uint8_t* __restrict__ data, uint32_t stage_addr, uint32_t my_n, uint32_t my_d, uint32_t my_s, int lane) {
    constexpr int S = 32 / L;   // strings per step
    const uint32_t sub = (uint32_t)lane % L;
    const int which = lane / L;
    const uint32_t k1 = sub + L, k2 = sub + 2 * L, k3 = sub + 3 * L, k4 = sub + 4 * L;
#pragma unroll 1
    for (int j = 0; j < 32; j += S) {
        const uint32_t n = __shfl_sync(0xFFFFFFFFu, my_n, j + which);
        const uint32_t nmax = __reduce_max_sync(0xFFFFFFFFu, n);
        if (nmax == 0u) continue;
        const uint32_t d = stage_addr + __shfl_sync(0xFFFFFFFFu, my_d, j + which) + sub;
        const uint8_t* const src = data + __shfl_sync(0xFFFFFFFFu, my_s, j + which) + sub;
        const uint32_t v0 = ldg_u8_lt<0>(src, sub, n);
        const uint32_t v1 = ldg_u8_lt<L>(src, k1, n);
        const uint32_t v2 = ldg_u8_lt<2 * L>(src, k2, n);
        const uint32_t v3 = ldg_u8_lt<3 * L>(src, k3, n);
        const uint32_t v4 = ldg_u8_lt<4 * L>(src, k4, n);
        sts_u8_lt<0>(d, v0, sub, n);
        sts_u8_lt<L>(d, v1, k1, n);
        sts_u8_lt<2 * L>(d, v2, k2, n);
        sts_u8_lt<3 * L>(d, v3, k3, n);
        sts_u8_lt<4 * L>(d, v4, k4, n);
        if (nmax > (uint32_t)(5 * L)) {
            for (uint32_t k = sub + 5 * L; k < n; k += L) {
                const uint32_t v = ldg_u8_lt<0>(src + (k - sub), 0u, 1u);
                sts_u8_lt<0>(d + (k - sub), v, 0u, 1u);
            }
        }
    }
}

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// Short strings (the common case: tens of bytes): ONE LANE PER STRING.  The lane walks the 32-bit words of its destination
// range in the staging buffer; destination word i is the funnel shift of two consecutive aligned source words (the phase between
// source and destination alignment is constant along a string), so a whole word moves with one LDG.32 + SHF + STS.32 and the 32
// lanes of a warp move 32 strings at once — ~4 warp instructions per string instead of ~11 with lanes over bytes.  Only the first
// and the last word of a string (shared with its neighbours) are written byte-wise.  Source words are read only if they hold
// one of the string's bytes, so nothing outside the data buffer's words is touched.
__device__ __forceinline__ void copy_descriptor_per_lane(const uint8_t* __restrict__ data, uint32_t stage_addr, uint32_t n, uint32_t d_off, uint32_t s_off) {
    const uint32_t d = stage_addr + d_off;           // shared address of our first destination byte
    const uint32_t head = d & 3u;                    // bytes of destination word 0 in front of it
    const uint32_t dw = d - head;
    const uint32_t end = head + n;                   // one past our last byte, relative to dw
    const uint32_t nw = n != 0u ? (end + 3u) >> 2 : 0u;                       // destination words touched
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(data) + s_off - head;    // source address of destination word 0, byte 0
    const uint32_t ph = (uint32_t)a0 & 3u;
    const uint32_t sh = ph * 8u;
    const uint32_t* const src = reinterpret_cast<const uint32_t*>(a0 - ph);
    const uint32_t jl = (ph + head) >> 2;                                     // first aligned source word holding one of our bytes
    const uint32_t nl = n != 0u ? (ph + end + 3u) >> 2 : 0u;                  // one past the last
    auto L = [&](uint32_t j) -> uint32_t { return (j - jl) < (nl - jl) ? __ldg(src + j) : 0u; };
    auto put_full = [&](uint32_t i, uint32_t w) {
        const uint32_t lo = i * 4u;
        if (lo >= head && lo + 4u <= end) sts_u32(dw + lo, w);
    };
    auto put_partial = [&](uint32_t i, uint32_t w) {
        const uint32_t lo = i * 4u;
#pragma unroll
        for (uint32_t b = 0; b < 4u; ++b)
            if (lo + b >= head && lo + b < end) sts_u8(dw + lo + b, w >> (8u * b));
    };
    const uint32_t nmax = __reduce_max_sync(0xFFFFFFFFu, nw);
    uint32_t prev = nl > jl && jl == 0u ? __ldg(src) : 0u;
#pragma unroll 1
    for (uint32_t i0 = 0; i0 < nmax; i0 += 4u) {
        const uint32_t l1 = L(i0 + 1u), l2 = L(i0 + 2u), l3 = L(i0 + 3u), l4 = L(i0 + 4u);
        put_full(i0, __funnelshift_r(prev, l1, sh));
        put_full(i0 + 1u, __funnelshift_r(l1, l2, sh));
        put_full(i0 + 2u, __funnelshift_r(l2, l3, sh));
        put_full(i0 + 3u, __funnelshift_r(l3, l4, sh));
        prev = l4;
    }
    // the two boundary words
    if (nw != 0u && (head != 0u || end < 4u)) put_partial(0u, __funnelshift_r(L(0u), L(1u), sh));
    if (nw > 1u && (end & 3u) != 0u) put_partial(nw - 1u, __funnelshift_r(L(nw - 1u), L(nw), sh));
}

// Survivor byte total of every tile, then (last CTA, by ticket) their exclusive prefix scan in place.
//   tile_bytes[t]        <- sum over the tile's survivors (rank below the LIMIT, not null) of their byte lengths, then its prefix
//   *bytes_total_out     <- byte base + all survivor bytes
// One CTA per tile, thread t <-> rows [8t, 8t + 8): one selection byte, nine offsets, one validity byte per thread, so a tile is
// two dependent round trips deep and 8 CTAs per SM keep ~70 KB of loads in flight.
static __global__ void __launch_bounds__(kBlock, 6) string_sizes_kernel(const __grid_constant__ StrGatherParams p) {
    __shared__ uint32_t s_warp[kWarps];
    __shared__ uint64_t s_part[kWarps];
    __shared__ uint32_t s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_tiles = gridDim.x;
    const int64_t tile = blockIdx.x;
    const int64_t row0 = tile * kTileRows + (int64_t)tid * 8;
    uint32_t selbyte = 0;
    if (row0 < p.n_rows) {
        if (p.sel != nullptr) selbyte = reinterpret_cast<const uint8_t*>(p.sel)[tile * (kTileRows / 8) + tid];
        else selbyte = (p.n_rows - row0 >= 8) ? 0xFFu : ((1u << (p.n_rows - row0)) - 1u);
    }
    uint64_t rexcl = 0;
    if (p.limit >= 0) {
        rexcl = p.tile_prefix != nullptr ? p.tile_prefix[tile] : (uint64_t)(p.row_base + tile * kTileRows);
        if (p.chunk_base != nullptr) rexcl = p.chunk_base[tile / p.tiles_per_chunk] + (rexcl >> 12);
    }
    int32_t off[9];
    uint32_t vbits = 0xFFu, my_bytes = 0;
    if (selbyte != 0u) {
#pragma unroll
        for (int i = 0; i < 9; ++i) off[i] = (row0 + i <= p.n_rows) ? __ldg(p.offsets + row0 + i) : 0;
        if (p.valid.words != nullptr) vbits = load_bits32(p.valid, (uint64_t)row0) & 0xFFu;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if ((selbyte >> i) & (vbits >> i) & 1u) my_bytes += (uint32_t)(off[i + 1] - off[i]);
    }
    if (p.limit >= 0) {
        // survivors whose global rank reaches the LIMIT do not count (uniform branch: at most one tile per query is cut)
        uint32_t cnt_total;
        const uint32_t r0 = block_exclusive_scan(__popc(selbyte), s_warp, cnt_total);
        const uint32_t lim = rexcl >= (uint64_t)p.limit ? 0u : (uint32_t)min((uint64_t)cnt_total, (uint64_t)p.limit - rexcl);
        if (lim != cnt_total) {
            my_bytes = 0;
            uint32_t r = r0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if ((selbyte >> i) & 1u) {
                    if (r < lim && ((vbits >> i) & 1u)) my_bytes += (uint32_t)(off[i + 1] - off[i]);
                    ++r;
                }
            }
        }
    }
    const uint32_t wsum = __reduce_add_sync(0xFFFFFFFFu, my_bytes);
    if (lane == 0) s_part[warp] = wsum;
    __syncthreads();
    if (tid == 0) {
        uint64_t tb = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) tb += s_part[w];
        st_relaxed_gpu(p.tile_bytes + tile, tb);
        __threadfence();
        s_last = atomicAdd(reinterpret_cast<unsigned int*>(p.tile_bytes + n_tiles), 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (s_last == 0u) return;
    // ---- last CTA to finish: exclusive prefix over the tiles, in place, 2048 tiles a round (8 independent loads per thread)
    __threadfence();
    uint64_t carry = 0;
#pragma unroll 1
    for (int64_t base = 0; base < n_tiles; base += kBlock * 8) {
        const int64_t lo = base + (int64_t)tid * 8;
        uint64_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = lo + i < n_tiles ? ld_relaxed_gpu(p.tile_bytes + lo + i) : 0ull;
        uint64_t mine = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const uint64_t t = v[i]; v[i] = mine; mine += t; }
        uint64_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t nb = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += nb;
        }
        __syncthreads();   // s_part of the previous round has been read
        if (lane == 31) s_part[warp] = incl;
        __syncthreads();
        uint64_t woff = 0, all = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) { const uint64_t c = s_part[w]; woff += (w < warp) ? c : 0ull; all += c; }
        const uint64_t excl = carry + woff + incl - mine;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (lo + i < n_tiles) p.tile_bytes[lo + i] = excl + v[i];
        carry += all;
    }
    if (tid == 0) *p.bytes_total_out = (unsigned long long)((p.byte_base_in != nullptr ? (uint64_t)*p.byte_base_in : 0ull) + carry);
}

constexpr uint32_t kStrLong = 64;   // longer strings are copied cooperatively by the whole warp

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void atom_or_shared(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// bytes [blo, bhi) of a little-endian word, 0 <= blo < bhi <= 4
__device__ __forceinline__ uint32_t byte_mask(uint32_t blo, uint32_t bhi) {
    const uint32_t hi = bhi >= 4u ? 0xFFFFFFFFu : ((1u << (8u * bhi)) - 1u);
    return hi & ~((1u << (8u * blo)) - 1u);
}

// One lane, one string: n bytes from `src` (shared-memory byte address when SMEM, else a global pointer) to the ZEROED staging buffer
// at shared byte address dst.  Destination word i is the funnel shift of two consecutive aligned source words; interior words are
// stored, the first / last word (shared with neighbouring strings) are OR-ed in.  All 32 lanes run the loop to the warp's longest string.
template <bool SMEM>
__device__ __forceinline__ void copy_string_lane(uint32_t src_smem, const uint8_t* __restrict__ src_gmem, uint32_t dst, uint32_t n) {
    const uint32_t head = dst & 3u;
    const uint32_t dw = dst - head;
    const uint32_t end = head + n;
    const uint32_t nw = n != 0u ? (end + 3u) >> 2 : 0u;
    uint32_t sh, sa = 0;
    const uint32_t* gsrc = nullptr;
    uint32_t jl = 0, nl = 0;
    if (SMEM) {
        const uint32_t a0 = src_smem - head;      // shared address of the byte that lands in destination word 0, byte 0 (slack in front)
        sh = (a0 & 3u) * 8u;
        sa = a0 & ~3u;
    } else {
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(src_gmem) - head;
        const uint32_t ph = (uint32_t)a0 & 3u;
        sh = ph * 8u;
        gsrc = reinterpret_cast<const uint32_t*>(a0 - ph);
        jl = (ph + head) >> 2;                    // first aligned source word holding one of our bytes
        nl = n != 0u ? (ph + end + 3u) >> 2 : 0u; // one past the last: nothing outside [jl, nl) is dereferenced
    }
    auto load = [&](uint32_t j) -> uint32_t {
        if (SMEM) return lds_u32(sa + 4u * j);
        return (j - jl) < (nl - jl) ? __ldg(gsrc + j) : 0u;
    };
    const uint32_t nmax = __reduce_max_sync(0xFFFFFFFFu, nw);
    uint32_t prev = nw != 0u ? load(0u) : 0u;
#pragma unroll 1
    for (uint32_t i0 = 0; i0 < nmax; i0 += 2u) {
        // two words per round: both loads are issued before the first is used
        const uint32_t c1 = i0 < nw ? load(i0 + 1u) : 0u;
        const uint32_t c2 = i0 + 1u < nw ? load(i0 + 2u) : 0u;
        const uint32_t v0 = __funnelshift_r(prev, c1, sh), v1 = __funnelshift_r(c1, c2, sh);
        prev = c2;
        if (i0 < nw) {
            const uint32_t lo = i0 * 4u;
            if (lo >= head && lo + 4u <= end) sts_u32(dw + lo, v0);
            else atom_or_shared(dw + lo, v0 & byte_mask(lo >= head ? 0u : head - lo, min(end - lo, 4u)));
        }
        if (i0 + 1u < nw) {
            const uint32_t lo = i0 * 4u + 4u;
            if (lo + 4u <= end) sts_u32(dw + lo, v1);
            else atom_or_shared(dw + lo, v1 & byte_mask(0u, end - lo));
        }
    }
}

// The whole warp copies ONE long string: lane l moves destination words l, l + 32, ...
template <bool SMEM>
__device__ __forceinline__ void copy_string_warp(uint32_t src_smem, const uint8_t* __restrict__ src_gmem, uint32_t dst, uint32_t n, int lane) {
    const uint32_t head = dst & 3u;
    const uint32_t dw = dst - head;
    const uint32_t end = head + n;
    const uint32_t nw = (end + 3u) >> 2;
    uint32_t sh, sa = 0, jl = 0, nl = 0;
    const uint32_t* gsrc = nullptr;
    if (SMEM) {
        const uint32_t a0 = src_smem - head;
        sh = (a0 & 3u) * 8u; sa = a0 & ~3u;
    } else {
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(src_gmem) - head;
        const uint32_t ph = (uint32_t)a0 & 3u;
        sh = ph * 8u; gsrc = reinterpret_cast<const uint32_t*>(a0 - ph);
        jl = (ph + head) >> 2; nl = (ph + end + 3u) >> 2;
    }
    auto load = [&](uint32_t j) -> uint32_t {
        if (SMEM) return lds_u32(sa + 4u * j);
        return (j - jl) < (nl - jl) ? __ldg(gsrc + j) : 0u;
    };
    for (uint32_t i = (uint32_t)lane; i < nw; i += 32u) {
        const uint32_t v = __funnelshift_r(load(i), load(i + 1u), sh);
        const uint32_t lo = i * 4u;
        if (lo >= head && lo + 4u <= end) sts_u32(dw + lo, v);
        else atom_or_shared(dw + lo, v & byte_mask(lo >= head ? 0u : head - lo, min(end - lo, 4u)));
    }
}

// keep the lowest k set bits of w
__device__ __forceinline__ uint32_t keep_lowest_set(uint32_t w, uint32_t k) {
    while ((uint32_t)__popc(w) > k) w &= ~(0x80000000u >> __clz(w));
    return w;
}

constexpr uint32_t kStrChunk = 16 * 1024;   // staging bytes per pass (a 2048-row tile of ~24-byte strings at 50 % fits in one)

// COPY = 0: round-1 copy routines (byte-wise predicated stores at string boundaries).  COPY = 1: the round-2 lane / warp copy
// routines into a ZEROED staging buffer (boundary words OR-ed in with shared-memory atomics).
template <int COPY>
static __global__ void __launch_bounds__(kBlock, 6) string_gather_kernel(const __grid_constant__ StrGatherParams p) {
    __shared__ __align__(16) int32_t s_src[kTileRows];        // survivor r: first source byte
    __shared__ __align__(16) uint32_t s_dst[kTileRows + 8];   // survivor r: length, then first destination byte inside the tile's dense range; [count..] = total
    __shared__ __align__(16) uint8_t s_stage[kStrChunk + 16];
    __shared__ uint32_t s_warp[kWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t tile_row0 = tile * kTileRows;
    const uint32_t lt = lanemask_lt();

    // ---- A. the tile's 64 selection words, every warp redundantly (lane holds words lane and lane + 32): survivor counts in
    //         front of any warp are two warp reductions away, no barrier
    auto tail_mask = [&](int64_t r) -> uint32_t {   // rows [r, r + 32) that exist
        const int64_t rem = p.n_rows - r;
        return rem >= 32 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
    };
    uint32_t w0, w1;
    if (p.sel != nullptr) { const uint32_t* sw = p.sel + tile * kTileWords; w0 = __ldg(sw + lane); w1 = __ldg(sw + 32 + lane); }
    else { w0 = tail_mask(tile_row0 + 32 * lane); w1 = tail_mask(tile_row0 + 32 * (lane + 32)); }
    uint64_t rexcl = p.tile_prefix != nullptr ? p.tile_prefix[tile] : (uint64_t)(p.row_base + tile * kTileRows);
    if (p.chunk_base != nullptr) rexcl = p.chunk_base[tile / p.tiles_per_chunk] + (rexcl >> 12);
    const uint64_t rbase = p.row_base_in != nullptr ? (uint64_t)*p.row_base_in : 0ull;
    const uint32_t c0 = __popc(w0), c1 = __popc(w1);
    const uint32_t cnt_total = __reduce_add_sync(0xFFFFFFFFu, c0 + c1);
    const int first_word = warp * 8;               // this warp's rows are selection words [first_word, first_word + 8)
    uint32_t run = first_word < 32 ? __reduce_add_sync(0xFFFFFFFFu, lane < first_word ? c0 : 0u)
                                   : __reduce_add_sync(0xFFFFFFFFu, c0) + __reduce_add_sync(0xFFFFFFFFu, lane < first_word - 32 ? c1 : 0u);
    const uint32_t wsel = first_word < 32 ? w0 : w1;
    uint32_t cnt_lim = cnt_total;
    if (p.limit >= 0) cnt_lim = rexcl >= (uint64_t)p.limit ? 0u : (uint32_t)min((uint64_t)cnt_total, (uint64_t)p.limit - rexcl);

    // ---- B. lane <-> row: offsets read coalesced, (source offset, length) of every survivor stored at its rank
    //         (nulls are zero-length: string.rs:33-36; survivors beyond the LIMIT are simply not stored).
    //         A warp whose 256 rows hold no survivor skips its loads; otherwise it reads its 257 offsets in nine coalesced
    //         loads (lane <-> row of each 32-row group; the end of a group is lane 0 of the next one).
    const int64_t wrow0 = tile_row0 + (int64_t)warp * 256;
    const uint32_t wany = __ballot_sync(0xFFFFFFFFu, lane >= (first_word & 31) && lane < (first_word & 31) + 8 && wsel != 0u);
    if (wany != 0u) {
        uint32_t vwv = 0xFFFFFFFFu;
        if (p.valid.words != nullptr && lane < 8) vwv = load_bits32(p.valid, (uint64_t)(wrow0 + 32 * lane));
        int32_t o[9];
        const int32_t* const offw = p.offsets + wrow0 + lane;
        if (wrow0 + 256 + 32 <= p.n_rows) {
#pragma unroll
            for (int g = 0; g < 9; ++g) o[g] = __ldg(offw + 32 * g);
        } else {
#pragma unroll
            for (int g = 0; g < 9; ++g) o[g] = wrow0 + 32 * g + lane <= p.n_rows ? __ldg(offw + 32 * g) : 0;
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const uint32_t selw = __shfl_sync(0xFFFFFFFFu, wsel, (first_word + g) & 31);
            const uint32_t vw = __shfl_sync(0xFFFFFFFFu, vwv, g);
            int32_t o1 = __shfl_down_sync(0xFFFFFFFFu, o[g], 1);
            const int32_t onext = __shfl_sync(0xFFFFFFFFu, o[g + 1], 0);
            if (lane == 31) o1 = onext;
            const uint32_t r = run + __popc(selw & lt);
            run += __popc(selw);
            if (((selw >> lane) & 1u) != 0u && r < cnt_lim) {
                const uint32_t len = ((vw >> lane) & 1u) ? (uint32_t)(o1 - o[g]) : 0u;
                s_src[r] = o[g]; s_dst[r] = len;
                // pull the survivor's first line towards L2 now: the copy phase is a block scan away
                if (len != 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.data + o[g]));
            }
        }
    }
    __syncthreads();

    // ---- C. lengths -> exclusive byte offsets, in place: 8 ranks per thread, one block scan
    uint32_t b0, bytes_total;
    {
        const uint32_t base = (uint32_t)tid * 8u;
        uint32_t v[8];
        const uint4 x = *reinterpret_cast<const uint4*>(&s_dst[base]), y = *reinterpret_cast<const uint4*>(&s_dst[base + 4]);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const uint32_t t = base + i < cnt_lim ? v[i] : 0u; v[i] = sum; sum += t; }
        b0 = block_exclusive_scan(sum, s_warp, bytes_total);
        uint4 ox, oy;
        ox.x = b0 + v[0]; ox.y = b0 + v[1]; ox.z = b0 + v[2]; ox.w = b0 + v[3];
        oy.x = b0 + v[4]; oy.y = b0 + v[5]; oy.z = b0 + v[6]; oy.w = b0 + v[7];
        *reinterpret_cast<uint4*>(&s_dst[base]) = ox; *reinterpret_cast<uint4*>(&s_dst[base + 4]) = oy;
        if (tid == kBlock - 1) s_dst[kTileRows] = bytes_total;   // entries at and beyond cnt_lim hold the total
    }

    // global byte prefix of the tile: computed by string_sizes_kernel, so tiles do not wait on each other
    // (the ranges sizes pass folds the byte base into range_bytes and keeps one prefix per 1024-row half: the tile's is the first half's)
    const uint64_t bexcl = p.sub_bytes != nullptr ? p.range_bytes[tile / p.tiles_per_range] + p.sub_bytes[2 * tile]
                                                  : (p.byte_base_in != nullptr ? (uint64_t)*p.byte_base_in : 0ull) + p.tile_bytes[tile];
    __syncthreads();
    if (cnt_lim == 0u) return;

    // new offsets, in rank order: out_offsets[first survivor of the tile + r + 1] = end of survivor r (coalesced)
    {
        int32_t* oo = p.out_offsets + (rexcl - rbase) + 1;
        for (uint32_t r = tid; r < cnt_lim; r += kBlock) oo[r] = (int32_t)(bexcl + s_dst[r + 1]);
    }

    // Byte copy, staged through shared memory.  The tile's destination range is dense, so it is assembled in a staging buffer
    // laid out like the destination modulo 16 bytes and flushed with aligned 16-byte stores; only the first and last unit of
    // a tile (shared with the neighbouring tiles) are written byte-wise.  Gather side: every lane holds the descriptor of one
    // survivor of a 256-survivor group; short strings are moved one lane per string, word-wise (copy_descriptor_per_lane), long
    // ones cooperatively, the warp walking its 32 descriptors with all lanes over the bytes of one string (copy_descriptors<32>).
    // Ranges longer than the staging buffer take several chunks.
    const bool per_lane = bytes_total / cnt_lim <= 96u;   // mean survivor length: short strings go one lane per string
    const uint32_t pad = (uint32_t)(reinterpret_cast<uintptr_t>(p.out_data + bexcl) & 15u);
    const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(s_stage);
    uint32_t g_lo = 0;
#pragma unroll 1
    for (uint32_t c0 = 0; c0 < bytes_total; c0 += kStrChunk) {
        const uint32_t c1 = min(bytes_total, c0 + kStrChunk);
        if (COPY == 1) {
            for (uint32_t u = tid; u < (kStrChunk + 16) / 16; u += kBlock) *reinterpret_cast<uint4*>(s_stage + u * 16u) = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();
        }
        uint32_t g = g_lo;
#pragma unroll 1
        for (; g * kBlock < cnt_lim && s_dst[g * kBlock] < c1; ++g) {
            const uint32_t q = g * kBlock + (uint32_t)tid;
            uint32_t my_d = 0, my_n = 0, my_s = 0;
            if (q < cnt_lim) {
                const uint32_t d = s_dst[q], e = s_dst[q + 1];
                const uint32_t lo = max(d, c0), hi = min(e, c1);
                if (lo < hi) { my_n = hi - lo; my_d = lo - c0 + pad; my_s = (uint32_t)s_src[q] + (lo - d); }
            }
            if (__ballot_sync(0xFFFFFFFFu, my_n != 0u) == 0u) continue;
            if (COPY == 1) {
                uint32_t longs = __ballot_sync(0xFFFFFFFFu, my_n > kStrLong);
                copy_string_lane<false>(0u, p.data + my_s, stage_addr + my_d, my_n > kStrLong ? 0u : my_n);
                while (longs != 0u) {
                    const int l = __ffs(longs) - 1;
                    longs &= longs - 1u;
                    const uint32_t n = __shfl_sync(0xFFFFFFFFu, my_n, l), d = __shfl_sync(0xFFFFFFFFu, my_d, l), s0 = __shfl_sync(0xFFFFFFFFu, my_s, l);
                    copy_string_warp<false>(0u, p.data + s0, stage_addr + d, n, lane);
                }
            } else if (per_lane) copy_descriptor_per_lane(p.data, stage_addr, my_n, my_d, my_s);
            else copy_descriptors<32>(p.data, stage_addr, my_n, my_d, my_s, lane);
        }
        g_lo = g > g_lo ? g - 1u : g_lo;  // the last group may straddle the chunk boundary
        __syncthreads();
        uint8_t* const gbase = p.out_data + bexcl + c0 - pad;  // 16-byte aligned
        const uint32_t end = pad + (c1 - c0);
        for (uint32_t u = tid; u * 16u < end; u += kBlock) {
            const uint32_t b0 = u * 16u, b1 = b0 + 16u;
            if (b0 >= pad && b1 <= end) {
                const uint4 v = *reinterpret_cast<const uint4*>(s_stage + b0);
                asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(gbase + b0), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            } else {
                for (uint32_t k = max(b0, pad); k < min(b1, end); ++k) gbase[k] = s_stage[k];
            }
        }
        __syncthreads();
    }
}

// =====================================================================================================================
// Round-2 string kernels.  What round 1's ncu said about the pair above: the gather is instruction- and LSU-bound (~1000 warp
// instructions per warp and tile, every source word a divergent, bounds-checked LDG), the sizes pass is 24 K short CTAs plus a serial
// ticketed prefix tail.  Changes:
//   string_sizes_ranges_kernel   persistent; every warp owns a contiguous range of tiles (the predicate scan's shape), streams the
//                                offsets lane <-> row and leaves per-sub-tile byte prefixes inside its range; the last CTA scans
//                                the <= 2368 range totals once.  No tile waits on another, no serial tail.
//   string_gather_staged_kernel  persistent, three CTAs per SM, each walking 1024-row sub-tiles one iteration ahead of its global loads
//                                (selection words, prefixes and the 1025 offsets of the NEXT sub-tile sit in registers while the
//                                current one is processed; the first version, one CTA per sub-tile, spent 8 us per CTA on four
//                                dependent round trips: 670 us per 50 M rows whatever the selectivity).
//                                The sub-tile's source bytes are ONE contiguous block of the data
//                                buffer (consecutive rows are adjacent), so dense sub-tiles fetch it with a single TMA bulk copy
//                                (cp.async.bulk + mbarrier) into shared memory while the ranks / lengths / byte offsets are being
//                                computed from the offsets (also staged in shared memory); the word-wise funnel-shift copy then
//                                runs shared -> shared (unguarded LDS instead of guarded, divergent LDG), string boundaries are
//                                merged with shared-memory atomic ORs into a zeroed staging buffer instead of byte-wise
//                                predicated stores, and the dense destination range is flushed with 16-byte stores.
//                                Sparse sub-tiles (few survivors) and blocks larger than the buffer read the survivors' words from
//                                global memory with the same routine.  Strings longer than 64 bytes are copied by the whole warp.
constexpr int kStrRows = 1024;                 // rows per string sub-tile: half a selection tile
constexpr int kStrWords = kStrRows / 32;
constexpr uint32_t kStrSrcCap = 32 * 1024;     // source block buffer (a 1024-row sub-tile of ~24-byte strings is ~22 KB)
constexpr uint32_t kStrStage = 11 * 1024;      // destination staging chunk

struct __align__(128) StrSmem {
    uint8_t src[kStrSrcCap + 64];              // [16 bytes slack][block, 16-byte aligned][slack]
    uint8_t stage[kStrStage + 32];
    int32_t offs[2][kStrRows + 8];             // offsets of the current / the next sub-tile's rows (+ the end)
    int32_t s_src[kStrRows];                   // survivor r: first source byte (absolute offset into the data buffer)
    uint32_t s_dst[kStrRows + 8];              // survivor r: length, then first destination byte inside the sub-tile's dense range
    uint32_t s_warp[kWarps];
    uint64_t mbar;
};

// Survivor bytes of every 1024-row sub-tile as an exclusive prefix inside the owning warp's range of tiles, the range totals, and
// (last CTA, by ticket) the ranges' global byte bases + the grand total.
//   sub_bytes[2 t + h]   bytes of the range's survivors in front of half h of tile t
//   range_bytes[g]       byte base of range g (byte_base_in included)
static __global__ void __launch_bounds__(kBlock) string_sizes_ranges_kernel(const __grid_constant__ StrGatherParams p) {
    __shared__ uint64_t s_part[kWarps];
    __shared__ uint32_t s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_tiles = (p.n_rows + kTileRows - 1) / kTileRows;
    const int64_t g = (int64_t)blockIdx.x * kWarps + warp;
    const int64_t t0 = min(n_tiles, g * p.tiles_per_range), t1 = min(n_tiles, t0 + p.tiles_per_range);
    uint64_t run = 0;
#pragma unroll 1
    for (int64_t t = t0; t < t1; ++t) {
        const int64_t row0 = t * kTileRows;
        uint32_t w0, w1;
        if (p.sel != nullptr) { const uint32_t* sw = p.sel + t * kTileWords; w0 = __ldg(sw + lane); w1 = __ldg(sw + 32 + lane); }
        else {
            const int64_t r0 = p.n_rows - (row0 + 32 * lane), r1 = r0 - 1024;
            w0 = r0 >= 32 ? 0xFFFFFFFFu : (r0 <= 0 ? 0u : ((1u << r0) - 1u));
            w1 = r1 >= 32 ? 0xFFFFFFFFu : (r1 <= 0 ? 0u : ((1u << r1) - 1u));
        }
        if (p.limit >= 0) {
            // survivors whose global rank reaches the LIMIT do not count (at most one tile per query is cut)
            uint64_t rexcl = p.tile_prefix != nullptr ? p.tile_prefix[t] : (uint64_t)(p.row_base + row0);
            if (p.chunk_base != nullptr) rexcl = p.chunk_base[t / p.tiles_per_chunk] + (rexcl >> 12);
            const uint32_t c0 = __popc(w0), c1 = __popc(w1);
            uint32_t i0 = c0, i1 = c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, i0, o), b = __shfl_up_sync(0xFFFFFFFFu, i1, o);
                if (lane >= o) { i0 += a; i1 += b; }
            }
            const uint32_t h0 = __shfl_sync(0xFFFFFFFFu, i0, 31);
            const uint32_t total = h0 + __shfl_sync(0xFFFFFFFFu, i1, 31);
            if (rexcl >= (uint64_t)p.limit) { w0 = 0u; w1 = 0u; }
            else if (rexcl + total > (uint64_t)p.limit) {
                const uint32_t lim = (uint32_t)((uint64_t)p.limit - rexcl);
                const uint32_t e0 = i0 - c0, e1 = h0 + i1 - c1;
                w0 = e0 >= lim ? 0u : keep_lowest_set(w0, lim - e0);
                w1 = e1 >= lim ? 0u : keep_lowest_set(w1, lim - e1);
            }
        }
        // nulls are zero-length (string.rs:33-36)
        if (p.valid.words != nullptr) { w0 &= load_bits32(p.valid, (uint64_t)(row0 + 32 * lane)); w1 &= load_bits32(p.valid, (uint64_t)(row0 + 1024 + 32 * lane)); }
        uint32_t half_bytes[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t wh = h == 0 ? w0 : w1;
            uint32_t bytes = 0;
            if (__ballot_sync(0xFFFFFFFFu, wh != 0u) != 0u) {
                // lane <-> row of every 32-row group: all 33 loads of the half are issued before the first is used (one round trip);
                // a row's end is the next lane's start (lane 31: lane 0 of the next group)
                const int64_t hrow0 = row0 + h * 1024;
                const int32_t* const off = p.offsets + hrow0 + lane;
                int32_t a[33];
#pragma unroll
                for (int k = 0; k < 33; ++k) a[k] = hrow0 + k * 32 + lane <= p.n_rows ? __ldg(off + k * 32) : 0;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const uint32_t m = __shfl_sync(0xFFFFFFFFu, wh, k);
                    int32_t e = __shfl_down_sync(0xFFFFFFFFu, a[k], 1);
                    const int32_t e31 = __shfl_sync(0xFFFFFFFFu, a[k + 1], 0);
                    if (lane == 31) e = e31;
                    if ((m >> lane) & 1u) bytes += (uint32_t)(e - a[k]);
                }
            }
            half_bytes[h] = __reduce_add_sync(0xFFFFFFFFu, bytes);
        }
        if (lane == 0) { p.sub_bytes[2 * t] = run; p.sub_bytes[2 * t + 1] = run + half_bytes[0]; }
        run += (uint64_t)half_bytes[0] + half_bytes[1];
    }
    if (lane == 0 && g < p.n_ranges) p.range_bytes[g] = run;

    // ---- last CTA to finish: range totals -> global exclusive bases
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(reinterpret_cast<unsigned int*>(p.range_bytes + p.n_ranges), 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (s_last == 0u) return;
    __threadfence();
    const int G = p.n_ranges;
    const int per = (G + kBlock - 1) / kBlock;   // <= 10
    uint64_t vals[12];
    uint64_t mine = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const int idx = tid * per + i;
        vals[i] = (i < per && idx < G) ? ld_relaxed_gpu(p.range_bytes + idx) : 0ull;
        mine += vals[i];
    }
    uint64_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t nb = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += nb;
    }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    uint64_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { const uint64_t c = s_part[w]; woff += (w < warp) ? c : 0ull; total += c; }
    const uint64_t base0 = p.byte_base_in != nullptr ? (uint64_t)*p.byte_base_in : 0ull;
    uint64_t acc = base0 + woff + incl - mine;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const int idx = tid * per + i;
        if (i < per && idx < G) { p.range_bytes[idx] = acc; acc += vals[i]; }
    }
    if (tid == 0) *p.bytes_total_out = (unsigned long long)(base0 + total);
}

// Everything a sub-tile needs from global memory before its first instruction of real work, requested a whole iteration ahead
// and parked in registers: one selection word per lane (+ the first half's word for the second half's rank), the tile's row
// prefix words, the byte prefix words, and the sub-tile's 1025 offsets (thread t: entries t, t + 256, t + 512, t + 768; thread 0
// also entry 1024).
struct StrPrefetch {
    uint32_t w, wf;
    uint64_t tp, cb, rb, sb;
    int32_t o[5];
};
// what phase A derives from it (uniform across the CTA except `w`)
struct StrSub {
    int64_t row0;
    uint64_t rexcl, bexcl;
    uint32_t w, cnt_lim, src_bias;
    int rows_here;
    bool staged;
};

static __global__ void __launch_bounds__(kBlock, 3) string_gather_staged_kernel(const __grid_constant__ StrGatherParams p) {
    extern __shared__ __align__(128) unsigned char str_smem_raw[];
    StrSmem& sm = *reinterpret_cast<StrSmem*>(str_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_subs = (p.n_rows + kStrRows - 1) / kStrRows;
    const uint32_t lt = lanemask_lt();
    const uint64_t rbase = p.row_base_in != nullptr ? (uint64_t)*p.row_base_in : 0ull;
    uint8_t* const src_blk = sm.src + 16;   // 16-byte aligned (the struct is 128-byte aligned)
    const uint32_t stage_addr = smem_u32(sm.stage);

    auto prefetch = [&](int64_t sub) -> StrPrefetch {
        StrPrefetch f;
        const int64_t tile = sub >> 1, row0 = sub * kStrRows;
        const int half = (int)(sub & 1);
        const int rows_here = (int)min((int64_t)kStrRows, p.n_rows - row0);
        if (p.sel != nullptr) {
            const uint32_t* sw = p.sel + tile * kTileWords;
            f.w = __ldg(sw + half * 32 + lane);
            f.wf = half ? __ldg(sw + lane) : 0u;
        } else {
            const int64_t rem = p.n_rows - (row0 + 32 * lane);
            f.w = rem >= 32 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
            f.wf = half ? 0xFFFFFFFFu : 0u;
        }
        f.tp = p.tile_prefix != nullptr ? __ldg(p.tile_prefix + tile) : (uint64_t)(p.row_base + tile * kTileRows);
        f.cb = p.chunk_base != nullptr ? __ldg(p.chunk_base + (uint32_t)tile / (uint32_t)p.tiles_per_chunk) : 0ull;
        f.rb = __ldg(p.range_bytes + (uint32_t)tile / (uint32_t)p.tiles_per_range);
        f.sb = __ldg(p.sub_bytes + sub);
        const int32_t* const off = p.offsets + row0;
#pragma unroll
        for (int j = 0; j < 4; ++j) f.o[j] = j * kBlock + tid <= rows_here ? __ldg(off + j * kBlock + tid) : 0;
        f.o[4] = (tid == 0 && rows_here == kStrRows) ? __ldg(off + kStrRows) : 0;
        return f;
    };
    auto park_offsets = [&](const StrPrefetch& f, int buf) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sm.offs[buf][j * kBlock + tid] = f.o[j];
        if (tid == 0) sm.offs[buf][kStrRows] = f.o[4];
    };
    // phase A of a sub-tile whose offsets are already in sm.offs[buf] (visible): ranks in front, LIMIT cut, byte prefix, and the
    // TMA bulk copy of its source block when it is dense enough, fits the buffer and lies inside the readable bounds
    auto phase_a = [&](const StrPrefetch& f, int64_t sub, int buf) -> StrSub {
        StrSub s;
        s.row0 = sub * kStrRows;
        s.rows_here = (int)min((int64_t)kStrRows, p.n_rows - s.row0);
        s.w = f.w;
        s.rexcl = (p.chunk_base != nullptr ? f.cb + (f.tp >> 12) : f.tp) + __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(f.wf));
        const uint32_t cnt_total = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(f.w));
        s.cnt_lim = cnt_total;
        if (p.limit >= 0) s.cnt_lim = s.rexcl >= (uint64_t)p.limit ? 0u : (uint32_t)min((uint64_t)cnt_total, (uint64_t)p.limit - s.rexcl);
        s.bexcl = f.rb + f.sb;
        const int32_t first = sm.offs[buf][0], last = sm.offs[buf][s.rows_here];
        const uint8_t* const g0 = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(p.data + first) & ~uintptr_t(15));
        const uint8_t* const g1 = reinterpret_cast<const uint8_t*>((reinterpret_cast<uintptr_t>(p.data + last) + 15) & ~uintptr_t(15));
        const uint32_t blk = (uint32_t)(g1 - g0);
        s.staged = s.cnt_lim >= p.dense_min && s.cnt_lim != 0u && last > first && blk <= kStrSrcCap && g0 >= p.data_lo && g1 <= p.data_hi;
        // shared address of data byte `o` once staged: src_blk + (p.data + o - g0)
        s.src_bias = smem_u32(src_blk) - (uint32_t)(reinterpret_cast<uintptr_t>(g0) - reinterpret_cast<uintptr_t>(p.data));
        if (s.staged && tid == 0) {
            // the buffer was last read through the generic proxy (LDS of the previous sub-tile's copy): order those before the bulk write
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tma_load_1d(src_blk, g0, blk, &sm.mbar);
        }
        return s;
    };

    int64_t sub = blockIdx.x;
    if (sub >= n_subs) return;
    if (tid == 0) { mbar_init(&sm.mbar, 1); mbar_init_fence(); }
    int buf = 0;
    uint32_t tma_phase = 0;
    StrPrefetch pf = prefetch(sub);
    park_offsets(pf, 0);
    __syncthreads();
    StrSub cur = phase_a(pf, sub, 0);
    int64_t nsub = sub + gridDim.x;
    if (nsub < n_subs) pf = prefetch(nsub);

#pragma unroll 1
    while (true) {
        // the next sub-tile's inputs were requested an iteration ago: park its offsets now, request the one after it
        const bool has_next = nsub < n_subs;
        StrPrefetch nf = pf;
        if (has_next) {
            park_offsets(nf, buf ^ 1);
            if (nsub + gridDim.x < n_subs) pf = prefetch(nsub + gridDim.x);
        }
        StrSub nxt{};
        bool next_issued = false;
        const int32_t* const offs = sm.offs[buf];
        uint32_t bytes_total = 0;

        if (cur.cnt_lim != 0u) {
            // ---- B. lane <-> row: (source offset, length) of every survivor stored at its rank; nulls are zero-length (string.rs:33-36)
            {
                const int first_word = warp * 4;               // this warp's 128 rows = 4 selection words
                uint32_t run = __reduce_add_sync(0xFFFFFFFFu, lane < first_word ? (uint32_t)__popc(cur.w) : 0u);
                uint32_t vwv = 0xFFFFFFFFu;
                if (p.valid.words != nullptr && lane < 4) vwv = load_bits32(p.valid, (uint64_t)(cur.row0 + warp * 128 + 32 * lane));
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t selw = __shfl_sync(0xFFFFFFFFu, cur.w, first_word + g);
                    const uint32_t vw = __shfl_sync(0xFFFFFFFFu, vwv, g);
                    const int r_in = warp * 128 + g * 32 + lane;
                    const uint32_t r = run + __popc(selw & lt);
                    run += __popc(selw);
                    if (((selw >> lane) & 1u) != 0u && r < cur.cnt_lim) {
                        const int32_t o = offs[r_in], o1 = offs[r_in + 1];
                        sm.s_src[r] = o;
                        sm.s_dst[r] = ((vw >> lane) & 1u) ? (uint32_t)(o1 - o) : 0u;
                    }
                }
            }
            __syncthreads();
            // ---- C. lengths -> exclusive byte offsets, in place: 4 ranks per thread, one block scan
            {
                const uint32_t base = (uint32_t)tid * 4u;
                const uint4 x = *reinterpret_cast<const uint4*>(&sm.s_dst[base]);
                uint32_t v[4] = {x.x, x.y, x.z, x.w};
                uint32_t sum = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) { const uint32_t t = base + i < cur.cnt_lim ? v[i] : 0u; v[i] = sum; sum += t; }
                const uint32_t b0 = block_exclusive_scan(sum, sm.s_warp, bytes_total);
                uint4 ox; ox.x = b0 + v[0]; ox.y = b0 + v[1]; ox.z = b0 + v[2]; ox.w = b0 + v[3];
                *reinterpret_cast<uint4*>(&sm.s_dst[base]) = ox;
                if (tid == kBlock - 1) sm.s_dst[kStrRows] = bytes_total;   // entries at and beyond cnt_lim hold the total
            }
            __syncthreads();
            // new offsets, in rank order: out_offsets[first survivor of the sub-tile + r + 1] = end of survivor r (coalesced)
            {
                int32_t* oo = p.out_offsets + (cur.rexcl - rbase) + 1;
                for (uint32_t r = tid; r < cur.cnt_lim; r += kBlock) oo[r] = (int32_t)(cur.bexcl + sm.s_dst[r + 1]);
            }
        }
        if (cur.staged) { mbar_wait(&sm.mbar, tma_phase); tma_phase ^= 1u; }

        // ---- byte copy through the staging buffer, laid out like the destination modulo 16 bytes, in chunks
        if (bytes_total != 0u) {
            const uint32_t pad = (uint32_t)(reinterpret_cast<uintptr_t>(p.out_data + cur.bexcl) & 15u);
            uint32_t g_lo = 0;
#pragma unroll 1
            for (uint32_t c0 = 0; c0 < bytes_total; c0 += kStrStage) {
                const uint32_t c1 = min(bytes_total, c0 + kStrStage);
                // zero the staging buffer: string boundaries are OR-ed in
                for (uint32_t u = tid; u < (kStrStage + 32) / 16; u += kBlock) *reinterpret_cast<uint4*>(sm.stage + u * 16u) = make_uint4(0u, 0u, 0u, 0u);
                __syncthreads();
                uint32_t g = g_lo;
#pragma unroll 1
                for (; g * kBlock < cur.cnt_lim && sm.s_dst[g * kBlock] < c1; ++g) {
                    const uint32_t q = g * kBlock + (uint32_t)tid;
                    uint32_t my_d = 0, my_n = 0, my_s = 0;
                    if (q < cur.cnt_lim) {
                        const uint32_t d = sm.s_dst[q], e = sm.s_dst[q + 1];
                        const uint32_t lo = max(d, c0), hi = min(e, c1);
                        if (lo < hi) { my_n = hi - lo; my_d = lo - c0 + pad; my_s = (uint32_t)sm.s_src[q] + (lo - d); }
                    }
                    const uint32_t any = __ballot_sync(0xFFFFFFFFu, my_n != 0u);
                    if (any == 0u) continue;
                    uint32_t longs = __ballot_sync(0xFFFFFFFFu, my_n > kStrLong);
                    const uint32_t n_lane = my_n > kStrLong ? 0u : my_n;
                    if (cur.staged) copy_string_lane<true>(cur.src_bias + my_s, nullptr, stage_addr + my_d, n_lane);
                    else copy_string_lane<false>(0u, p.data + my_s, stage_addr + my_d, n_lane);
                    while (longs != 0u) {
                        const int l = __ffs(longs) - 1;
                        longs &= longs - 1u;
                        const uint32_t n = __shfl_sync(0xFFFFFFFFu, my_n, l), d = __shfl_sync(0xFFFFFFFFu, my_d, l), s0 = __shfl_sync(0xFFFFFFFFu, my_s, l);
                        if (cur.staged) copy_string_warp<true>(cur.src_bias + s0, nullptr, stage_addr + d, n, lane);
                        else copy_string_warp<false>(0u, p.data + s0, stage_addr + d, n, lane);
                    }
                }
                g_lo = g > g_lo ? g - 1u : g_lo;  // the last group may straddle the chunk boundary
                __syncthreads();
                if (c1 == bytes_total && has_next) {
                    // the source buffer is free: the next sub-tile's block starts flowing while this chunk is flushed
                    nxt = phase_a(nf, nsub, buf ^ 1);
                    next_issued = true;
                }
                uint8_t* const gbase = p.out_data + cur.bexcl + c0 - pad;  // 16-byte aligned
                const uint32_t end = pad + (c1 - c0);
                for (uint32_t u = tid; u * 16u < end; u += kBlock) {
                    const uint32_t b0 = u * 16u, b1 = b0 + 16u;
                    if (b0 >= pad && b1 <= end) {
                        const uint4 v = *reinterpret_cast<const uint4*>(sm.stage + b0);
                        asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(gbase + b0), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    } else {
                        for (uint32_t k = max(b0, pad); k < min(b1, end); ++k) gbase[k] = sm.stage[k];
                    }
                }
                __syncthreads();
            }
        }
        if (!has_next) break;
        if (!next_issued) {
            __syncthreads();   // everyone is done with this sub-tile's offsets, tables and source block
            nxt = phase_a(nf, nsub, buf ^ 1);
        }
        cur = nxt;
        sub = nsub;
        nsub += gridDim.x;
        buf ^= 1;
    }
}

// `string column <op> literal`: 3-way byte-wise compare -> truth-mask lookup -> one ballot word per warp.
// Validity / null semantics are applied afterwards by the fused kernel's bitmap-predicate mode.
static __global__ void __launch_bounds__(kBlock) string_predicate_kernel(int64_t n_rows, const int32_t* __restrict__ offsets,
                                                                 const uint8_t* __restrict__ data,
                                                                 const uint8_t* __restrict__ lit, int32_t lit_len,
                                                                 uint32_t truth, uint32_t* __restrict__ out_words) {
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    bool keep = false;
    if (row < n_rows) {
        const int32_t s = __ldg(offsets + row), e = __ldg(offsets + row + 1);
        const int32_t len = e - s;
        const int32_t m = len < lit_len ? len : lit_len;
        int c = 0;
        for (int32_t j = 0; j < m; ++j) {
            const int a = __ldg(data + s + j), b = __ldg(lit + j);
            if (a != b) { c = a < b ? -1 : 1; break; }
        }
        if (c == 0) c = len < lit_len ? -1 : (len > lit_len ? 1 : 0);
        const uint32_t code = c < 0 ? 1u : (c == 0 ? 2u : 4u);
        keep = (truth & code) != 0u;
    }
    const uint32_t w = __ballot_sync(0xFFFFFFFFu, keep);
    if ((threadIdx.x & 31) == 0 && (row - (row & 31)) < ((n_rows + 63) & ~63ll)) out_words[row >> 5] = w;
}

// Predicate over a Float64-dtype Series that also holds Int64 values (series.rs:210-212): per row the tag bit says which
// AnyValue variant `values` holds; the truth table of plan.rs:112-130 is applied with the row's own type.
// lit_kind: 0 = Null literal, 1 = Int64, 2 = Float64, 3 = any other type.  Output: final keep bits (null rule included).
static __global__ void __launch_bounds__(kBlock) mixed_predicate_kernel(int64_t n_rows, const uint64_t* __restrict__ values, BitSrc valid, BitSrc tag,
                                                                        int lit_kind, int64_t lit_bits, uint32_t truth,
                                                                        uint32_t* __restrict__ out_words) {
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    bool keep = false;
    if (row < n_rows) {
        bool ok = true;
        if (valid.words != nullptr) { const uint64_t bit = valid.bit0 + (uint64_t)row; ok = (__ldg(valid.words + (bit >> 5)) >> (bit & 31)) & 1u; }
        uint32_t code;
        if (!ok) code = lit_kind == 0 ? 2u : 1u;          // Null vs Null: Equal; Null vs value: Less (series.rs:105-106)
        else if (lit_kind == 0) code = 4u;                // value vs Null: Greater (series.rs:107)
        else {
            bool is_int = false;
            if (tag.words != nullptr) { const uint64_t bit = tag.bit0 + (uint64_t)row; is_int = (__ldg(tag.words + (bit >> 5)) >> (bit & 31)) & 1u; }
            const uint64_t v = values[row];
            if (is_int && lit_kind == 1) { const int64_t a = (int64_t)v; code = a < lit_bits ? 1u : (a == lit_bits ? 2u : 4u); }
            else if (!is_int && lit_kind == 2) {
                const double a = __longlong_as_double((long long)v), b = __longlong_as_double(lit_bits);
                code = a < b ? 1u : (a == b ? 2u : (a > b ? 4u : 8u));
            } else code = 8u;                             // different non-null types never compare (series.rs:95,114)
        }
        keep = (truth & code) != 0u;
    }
    const uint32_t w = __ballot_sync(0xFFFFFFFFu, keep);
    if ((threadIdx.x & 31) == 0 && (row - (row & 31)) < ((n_rows + 63) & ~63ll)) out_words[row >> 5] = w;
}

}  // namespace rvl
