// string_kernels.cuh — K4: StringArray compaction (offset prefix-sum + byte copy) and the string
// comparison predicate.
//
// Replaces (reference, /root/reference/src):
//   execution/record_batch.rs:163-170  take_array(String): value(i).to_string() per survivor, then
//   execution/array/string.rs:19-58    StringArray::new: second copy + offsets rebuild + UTF-8 re-validation
//   datatypes/series.rs:111            str ordering for `col <op> "literal"` (byte-wise lexicographic)
//
// The fused kernel leaves a row-order selection bitmap and each tile's exclusive output prefix; this
// kernel walks the same 2048-row tiles: thread t owns rows [8t, 8t+8) (one selection byte), a block
// scan ranks the survivors, a second scan + decoupled look-back over survivor byte lengths yields the
// new int32 offsets in one pass, then the tile's strings are copied into its dense destination range.
// UTF-8 validity is preserved by construction (whole strings are copied), so no re-validation pass.
#pragma once
#include "device_utils.cuh"

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
    uint64_t* tile_status;              // descriptors for the byte prefix (zeroed before launch)
    const unsigned long long* byte_base_in;  // bytes already emitted before this launch (concat), or nullptr
    unsigned long long* bytes_total_out;     // byte base + bytes emitted by this launch
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

static __global__ void __launch_bounds__(kBlock, 4) string_gather_kernel(const __grid_constant__ StrGatherParams p) {
    __shared__ int32_t s_src[kTileRows];
    __shared__ int32_t s_dst[kTileRows];
    __shared__ int32_t s_len[kTileRows];
    __shared__ uint32_t s_warp[kWarps];
    __shared__ uint64_t s_bexcl;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t row0 = tile * kTileRows + (int64_t)tid * 8;

    uint32_t selbyte = 0;
    if (row0 < p.n_rows) {
        if (p.sel != nullptr) selbyte = reinterpret_cast<const uint8_t*>(p.sel)[tile * (kTileRows / 8) + tid];
        else selbyte = (p.n_rows - row0 >= 8) ? 0xFFu : ((1u << (p.n_rows - row0)) - 1u);
    }
    uint64_t rexcl = p.tile_prefix != nullptr ? p.tile_prefix[tile] : (uint64_t)(p.row_base + tile * kTileRows);
    if (p.chunk_base != nullptr) rexcl = p.chunk_base[tile / p.tiles_per_chunk] + (rexcl >> 12);
    const uint64_t rbase = p.row_base_in != nullptr ? (uint64_t)*p.row_base_in : 0ull;

    uint32_t cnt_total;
    const uint32_t r0 = block_exclusive_scan(__popc(selbyte), s_warp, cnt_total);
    uint32_t cnt_lim = cnt_total;
    if (p.limit >= 0) cnt_lim = rexcl >= (uint64_t)p.limit ? 0u : (uint32_t)min((uint64_t)cnt_total, (uint64_t)p.limit - rexcl);

    // survivor lengths of this thread's rows (nulls are zero-length: string.rs:33-36)
    int32_t off[9];
    uint32_t my_bytes = 0;
    uint32_t vbits = 0xFFu;
    if (selbyte != 0u) {
#pragma unroll
        for (int i = 0; i < 9; ++i) off[i] = (row0 + i <= p.n_rows) ? __ldg(p.offsets + row0 + i) : 0;
        if (p.valid.words != nullptr) vbits = load_bits32(p.valid, (uint64_t)row0) & 0xFFu;
        uint32_t r = r0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if ((selbyte >> i) & 1u) {
                if (r < cnt_lim && ((vbits >> i) & 1u)) my_bytes += (uint32_t)(off[i + 1] - off[i]);
                ++r;
            }
        }
    }
    uint32_t bytes_total;
    const uint32_t b0 = block_exclusive_scan(my_bytes, s_warp, bytes_total);

    // global byte prefix: second decoupled look-back
    if (warp == 0) {
        uint64_t bexcl;
        if (tile == 0) {
            bexcl = p.byte_base_in != nullptr ? (uint64_t)*p.byte_base_in : 0ull;
            if (lane == 0) st_relaxed_gpu(p.tile_status, kStatusPrefix | (bexcl + bytes_total));
        } else {
            if (lane == 0) st_relaxed_gpu(p.tile_status + tile, kStatusAggregate | (uint64_t)bytes_total);
            bexcl = lookback_exclusive(p.tile_status, tile, lane);
            if (lane == 0) st_relaxed_gpu(p.tile_status + tile, kStatusPrefix | (bexcl + bytes_total));
        }
        if (lane == 0) {
            s_bexcl = bexcl;
            if (tile == (int64_t)gridDim.x - 1) *p.bytes_total_out = (unsigned long long)(bexcl + bytes_total);
        }
    }
    __syncthreads();
    if (cnt_lim == 0u) return;
    const uint64_t bexcl = s_bexcl;

    // new offsets (rank order) + copy descriptors
    if (selbyte != 0u) {
        uint32_t r = r0, b = b0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if ((selbyte >> i) & 1u) {
                if (r < cnt_lim) {
                    const int32_t len = ((vbits >> i) & 1u) ? (off[i + 1] - off[i]) : 0;
                    s_src[r] = off[i]; s_dst[r] = (int32_t)b; s_len[r] = len;
                    b += (uint32_t)len;
                    p.out_offsets[rexcl - rbase + r + 1] = (int32_t)(bexcl + b);
                }
                ++r;
            }
        }
    }
    __syncthreads();

    // byte copy: one warp per string, lanes stride the bytes (destination range of the tile is dense)
    uint8_t* dst_base = p.out_data + bexcl;
    for (uint32_t q = warp; q < cnt_lim; q += kWarps) {
        const int32_t len = s_len[q];
        const uint8_t* src = p.data + s_src[q];
        uint8_t* dst = dst_base + s_dst[q];
        for (int32_t j = lane; j < len; j += 32) dst[j] = __ldg(src + j);
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
