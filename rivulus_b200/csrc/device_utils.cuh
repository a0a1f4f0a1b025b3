// device_utils.cuh — small sm_100a device helpers shared by the rivulus kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvl {

constexpr int kBlock = 256;                 // threads per CTA in every tile kernel
constexpr int kWarps = kBlock / 32;
constexpr int kTileRows = 2048;             // rows per tile: 8 warps x 4 groups x 64 rows
constexpr int kGroups = 4;                  // 64-row groups per warp
constexpr int kTileWords = kTileRows / 32;  // selection words per tile

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---- global memory accessors with explicit cache behaviour -------------------------------------
// 128-bit read-only streaming load (two rows of an 8-byte column)
__device__ __forceinline__ ulonglong2 ld_stream_v2(const uint64_t* p) {
    ulonglong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint64_t ld_stream(const uint64_t* p) {
    uint64_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
// Isolated gather load: B200 fetches the whole 128-byte line from DRAM for an ordinary load that misses L2, and 64 bytes when the
// load carries the .L2::64B prefetch-size qualifier (profiles/r02_microbench_sector.txt: 128 / 64 bytes of dram__bytes_read per
// touched sector; nothing fetches a lone 32-byte sector).  Sparse survivors therefore cost half the DRAM traffic through this one.
__device__ __forceinline__ uint64_t ld_gather(const uint64_t* p) {
    uint64_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
// streaming (evict-first) stores for outputs that are never re-read by this kernel
__device__ __forceinline__ void st_stream(uint64_t* p, uint64_t v) {
    asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_stream_v2(uint64_t* p, uint64_t a, uint64_t b) {
    asm volatile("st.global.cs.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
// tile-descriptor traffic for the decoupled look-back: relaxed, GPU scope (the 64-bit word carries
// both the status and the value, so no separate fence is needed)
__device__ __forceinline__ uint64_t ld_relaxed_gpu(const uint64_t* p) {
    uint64_t r;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_relaxed_gpu(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu_u32(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}

// ---- bitmaps -----------------------------------------------------------------------------------
// A bitmap operand: 32-bit-aligned word base, absolute bit index of row 0, readable word count.
struct BitSrc {
    const uint32_t* words;  // nullptr => "all ones"
    uint64_t bit0;
    uint64_t nwords;
};

// 32 bits for rows [row, row+32) (any bit alignment), zero beyond the buffer
__device__ __forceinline__ uint32_t load_bits32(const BitSrc& b, uint64_t row) {
    if (b.words == nullptr) return 0xFFFFFFFFu;
    const uint64_t bit = b.bit0 + row;
    const uint64_t w = bit >> 5;
    const uint32_t sh = (uint32_t)bit & 31u;
    const uint32_t lo = w < b.nwords ? __ldg(b.words + w) : 0u;
    const uint32_t hi = (sh != 0u && w + 1 < b.nwords) ? __ldg(b.words + w + 1) : 0u;
    return __funnelshift_r(lo, hi, sh);
}
// 64 bits for rows [row, row+64)
__device__ __forceinline__ uint64_t load_bits64(const BitSrc& b, uint64_t row) {
    if (b.words == nullptr) return ~0ull;
    const uint64_t bit = b.bit0 + row;
    const uint64_t w = bit >> 5;
    const uint32_t sh = (uint32_t)bit & 31u;
    const uint32_t w0 = w < b.nwords ? __ldg(b.words + w) : 0u;
    const uint32_t w1 = w + 1 < b.nwords ? __ldg(b.words + w + 1) : 0u;
    const uint32_t w2 = (sh != 0u && w + 2 < b.nwords) ? __ldg(b.words + w + 2) : 0u;
    const uint32_t lo = __funnelshift_r(w0, w1, sh);
    const uint32_t hi = __funnelshift_r(w1, w2, sh);
    return ((uint64_t)hi << 32) | lo;
}

// interleave two 32-bit masks: bit 2i <- even bit i, bit 2i+1 <- odd bit i (row order of a 64-row group)
__device__ __forceinline__ uint64_t spread_bits(uint32_t x) {
    uint64_t v = x;
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}
__device__ __forceinline__ uint64_t interleave_masks(uint32_t even, uint32_t odd) {
    return spread_bits(even) | (spread_bits(odd) << 1);
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// ---- decoupled look-back (Merrill & Garland single-pass scan) -----------------------------------
// descriptor = status (2 bits, top) | value (62 bits).  0 = not yet published.
constexpr uint64_t kStatusAggregate = 1ull << 62;
constexpr uint64_t kStatusPrefix = 2ull << 62;
constexpr uint64_t kStatusMask = 3ull << 62;
constexpr uint64_t kValueMask = ~kStatusMask;

// Decoupled look-back of one warp of tile `tile` (> 0): sums the aggregates of the predecessor tiles back to the nearest published
// inclusive prefix.  Tiles are processed in blockIdx order, so a predecessor is always resident or finished (same forward-progress
// argument as CUB's scan).
// 256-descriptor-wide look-back run by one warp (8 predecessors per lane per round)
__device__ __forceinline__ uint64_t lookback_exclusive_wide(const uint64_t* status, int64_t tile, int lane) {
    uint64_t exclusive = 0;
    int64_t base = tile - 1;
    while (true) {
        uint64_t s[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t idx = base - (j * 32 + lane);
            s[j] = idx >= 0 ? ld_relaxed_gpu(status + idx) : kStatusPrefix;  // virtual tile -1: prefix 0
        }
        bool found = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (!found) {
                const int64_t idx = base - (j * 32 + lane);
                while ((s[j] & kStatusMask) == 0) s[j] = ld_relaxed_gpu(status + idx);
                const uint32_t prefix_lanes = __ballot_sync(0xFFFFFFFFu, (s[j] & kStatusMask) == kStatusPrefix);
                const int first = __ffs(prefix_lanes) - 1;
                const uint64_t take = (first < 0 || lane <= first) ? (s[j] & kValueMask) : 0ull;
                exclusive += warp_sum_u64(take);
                found = first >= 0;
            }
        }
        if (found) break;
        base -= 256;
    }
    return exclusive;
}

}  // namespace rvl
