// aux_kernels.cuh — support kernels around the hot path: synthetic table generator, order-sensitive
// checksums, bitmap popcount / bit-granular copy (concat, download rebasing), offset rebasing.
#pragma once
#include "../../include/rivulus_synth.h"
#include "device_utils.cuh"

namespace rvl {

// ---- synthetic columns (include/rivulus_synth.h) -----------------------------------------------
static __global__ void gen_col8_kernel(uint64_t* __restrict__ out, int kind, uint32_t col_id, uint64_t row0, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t u = rvl_synth_u(RVL_SYNTH_SEED, col_id, row0 + (uint64_t)i);
        uint64_t bits;
        if (kind == RVL_SYNTH_F64) bits = (uint64_t)__double_as_longlong(rvl_synth_f64(u));
        else bits = (uint64_t)rvl_synth_i64(u, kind);
        out[i] = bits;
    }
}

// one 32-bit word of a bit-packed column per thread: mode 0 = Boolean values (u & 1), 1 = validity
static __global__ void gen_bits_kernel(uint32_t* __restrict__ out, int mode, uint32_t col_id, uint32_t null_pct, uint64_t row0, int64_t n) {
    const int64_t nwords = (n + 31) / 32;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride) {
        uint32_t word = 0;
        for (int b = 0; b < 32; ++b) {
            const int64_t i = w * 32 + b;
            if (i >= n) break;
            const uint64_t row = row0 + (uint64_t)i;
            uint32_t bit;
            if (mode == 0) bit = (uint32_t)(rvl_synth_u(RVL_SYNTH_SEED, col_id, row) & 1ull);
            else bit = (uint32_t)rvl_synth_valid(RVL_SYNTH_SEED, col_id, row, null_pct);
            word |= bit << b;
        }
        out[w] = word;
    }
}

// string lengths (0 for nulls) -> lens[i]; offsets are produced by a host-driven scan over chunks
static __global__ void gen_strlen_kernel(int32_t* __restrict__ lens, uint32_t col_id, uint32_t null_pct, uint64_t row0, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t row = row0 + (uint64_t)i;
        const int valid = rvl_synth_valid(RVL_SYNTH_SEED, col_id, row, null_pct);
        lens[i] = valid ? (int32_t)rvl_synth_strlen(rvl_synth_u(RVL_SYNTH_SEED, col_id, row)) : 0;
    }
}
static __global__ void gen_strbytes_kernel(uint8_t* __restrict__ data, const int32_t* __restrict__ offsets, uint32_t col_id,
                                    uint64_t row0, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t u = rvl_synth_u(RVL_SYNTH_SEED, col_id, row0 + (uint64_t)i);
        const int32_t s = offsets[i], e = offsets[i + 1];
        for (int32_t j = 0; j < e - s; ++j) data[s + j] = rvl_synth_strbyte(u, (uint32_t)j);
    }
}

// single-block exclusive scan of int32 lengths into offsets[0..n] (offsets[0] = 0); n up to a few 10^8
// is handled by the chunk loop.  Only used by the generator (test/bench input), not by the hot path.
// *overflow (optional, zeroed by the caller) is set when a prefix exceeds INT32_MAX — the reference's `as i32` wraps silently there
// (string.rs:31); callers turn it into RVL_OFFSET_OVERFLOW.
static __global__ void __launch_bounds__(1024) scan_lengths_kernel(const int32_t* __restrict__ lens, int32_t* __restrict__ offsets, int64_t n,
                                                                   int32_t* __restrict__ overflow = nullptr) {
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_carry = 0; offsets[0] = 0; }
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + tid;
        long long v = i < n ? lens[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        long long woff = 0;
        for (int w = 0; w < warp; ++w) woff += s_warp[w];
        const long long carry = s_carry;
        if (i < n) {
            offsets[i + 1] = (int32_t)(carry + woff + incl);
            if (overflow != nullptr && carry + woff + incl > 2147483647ll) *overflow = 1;
        }
        __syncthreads();
        if (tid == 1023) s_carry = carry + woff + incl;
        __syncthreads();
    }
}

// ---- checksums (rivulus_synth.h: rvl_checksum_term) ---------------------------------------------
constexpr uint64_t kNullTag = 0x6E756C6C6E756C6Cull;  // "nullnull": what a null row contributes

__device__ __forceinline__ void block_sum_to(uint64_t v, unsigned long long* out) {
    __shared__ uint64_t s_part[32];
    v = warp_sum_u64(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_part[warp] = v;
    __syncthreads();
    if (warp == 0) {
        uint64_t t = lane < (int)(blockDim.x >> 5) ? s_part[lane] : 0ull;
        t = warp_sum_u64(t);
        if (lane == 0 && t != 0ull) atomicAdd(out, (unsigned long long)t);
    }
}

static __global__ void checksum_col8_kernel(const uint64_t* __restrict__ values, BitSrc valid, int64_t n, unsigned long long* out) {
    uint64_t acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        bool ok = true;
        if (valid.words != nullptr) { const uint64_t bit = valid.bit0 + (uint64_t)i; ok = (valid.words[bit >> 5] >> (bit & 31)) & 1u; }
        acc += rvl_checksum_term(ok ? values[i] : kNullTag, (uint64_t)i);
    }
    block_sum_to(acc, out);
}
static __global__ void checksum_bool_kernel(BitSrc vals, BitSrc valid, int64_t n, unsigned long long* out) {
    uint64_t acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        bool ok = true;
        if (valid.words != nullptr) { const uint64_t bit = valid.bit0 + (uint64_t)i; ok = (valid.words[bit >> 5] >> (bit & 31)) & 1u; }
        const uint64_t vb = vals.bit0 + (uint64_t)i;
        const uint64_t v = (vals.words[vb >> 5] >> (vb & 31)) & 1u;
        acc += rvl_checksum_term(ok ? v : kNullTag, (uint64_t)i);
    }
    block_sum_to(acc, out);
}
// strings: per-row hash h = fold(splitmix64(h ^ byte)) seeded with the length
static __global__ void checksum_str_kernel(const int32_t* __restrict__ offsets, const uint8_t* __restrict__ data, BitSrc valid, int64_t n,
                                    unsigned long long* out) {
    uint64_t acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        bool ok = true;
        if (valid.words != nullptr) { const uint64_t bit = valid.bit0 + (uint64_t)i; ok = (valid.words[bit >> 5] >> (bit & 31)) & 1u; }
        uint64_t h = kNullTag;
        if (ok) {
            const int32_t s = offsets[i], e = offsets[i + 1];
            h = (uint64_t)(e - s);
            for (int32_t j = s; j < e; ++j) h = rvl_splitmix64(h ^ (uint64_t)data[j]);
        }
        acc += rvl_checksum_term(h, (uint64_t)i);
    }
    block_sum_to(acc, out);
}

// ---- bitmap utilities ---------------------------------------------------------------------------
// number of 1 bits among rows [0, n) of `src`; n is read from *n_ptr when n_ptr != nullptr (device-side count)
static __global__ void count_ones_kernel(BitSrc src, int64_t n, const unsigned long long* n_ptr, const unsigned long long* sub_ptr,
                                  long long n_cap, unsigned long long* out) {
    // effective n = n_ptr ? min(*n_ptr, cap) - min(*sub_ptr, cap) : n
    int64_t nn = n;
    if (n_ptr != nullptr) {
        long long t = (long long)*n_ptr;
        long long sub = sub_ptr != nullptr ? (long long)*sub_ptr : 0;
        if (n_cap >= 0 && t > n_cap) t = n_cap;
        if (n_cap >= 0 && sub > n_cap) sub = n_cap;
        t -= sub;
        nn = t < 0 ? 0 : t;
    }
    uint64_t acc = 0;
    const int64_t nwords = (nn + 31) / 32;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride) {
        uint32_t bits = load_bits32(src, (uint64_t)w * 32);
        const int64_t rem = nn - w * 32;
        if (rem < 32) bits &= (1u << rem) - 1u;
        acc += __popc(bits);
    }
    block_sum_to(acc, out);
}

// dst bits [dst_bit0, dst_bit0 + n) |= src rows [0, n)   (dst zero-initialised; optional AND mask)
static __global__ void bitcopy_kernel(BitSrc src, BitSrc mask, uint32_t* __restrict__ dst, uint64_t dst_bit0, int64_t n) {
    // thread per destination word
    const uint64_t first_word = dst_bit0 >> 5;
    const uint32_t sh = (uint32_t)dst_bit0 & 31u;
    const int64_t n_words = (int64_t)((sh + (uint64_t)n + 31) >> 5);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_words; t += stride) {
        // destination word t holds source rows [32t - sh, 32t - sh + 32)
        const int64_t r0 = t * 32 - (int64_t)sh;
        uint32_t val = 0;
        if (r0 >= 0) {
            val = load_bits32(src, (uint64_t)r0) & load_bits32(mask, (uint64_t)r0);
            const int64_t rem = n - r0;
            if (rem < 32) val = rem <= 0 ? 0u : (val & ((1u << rem) - 1u));
        } else {
            // first (partial) word: rows [0, 32 - sh) shifted up by sh
            uint32_t lowbits = load_bits32(src, 0) & load_bits32(mask, 0);
            if (n < 32) lowbits &= (1u << n) - 1u;
            val = lowbits << sh;
        }
        const bool owned = (r0 >= 0) && (r0 + 32 <= n);
        if (owned) dst[first_word + t] = val;
        else if (val != 0u) atomicOr(dst + first_word + t, val);
    }
}

// out[i] = valid(i) ? in[i] : 0     (concat rebuilds through value(i): record_batch.rs:295-300)
static __global__ void copy_col8_zero_nulls_kernel(const uint64_t* __restrict__ in, BitSrc valid, uint64_t* __restrict__ out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        bool ok = true;
        if (valid.words != nullptr) { const uint64_t bit = valid.bit0 + (uint64_t)i; ok = (__ldg(valid.words + (bit >> 5)) >> (bit & 31)) & 1u; }
        out[i] = ok ? in[i] : 0ull;
    }
}

// ---- BooleanArray::{and, or, not} (array/boolean.rs:120-165): strict-null — the result is null where either input is null.
// One 32-row word per thread; inputs may be bit-offset views.  op: 0 = and, 1 = or, 2 = not (b ignored).
static __global__ void boolean_op_kernel(BitSrc a_vals, BitSrc a_valid, BitSrc b_vals, BitSrc b_valid, int op, int64_t n,
                                         uint32_t* __restrict__ out_vals, uint32_t* __restrict__ out_valid) {
    const int64_t nwords = (n + 31) / 32;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride) {
        const uint64_t row = (uint64_t)w * 32;
        uint32_t valid = load_bits32(a_valid, row);
        const uint32_t av = load_bits32(a_vals, row);
        uint32_t r;
        if (op == 2) r = ~av;
        else {
            valid &= load_bits32(b_valid, row);
            const uint32_t bv = load_bits32(b_vals, row);
            r = op == 0 ? (av & bv) : (av | bv);
        }
        const int64_t rem = n - w * 32;
        if (rem < 32) valid &= (1u << rem) - 1u;
        out_vals[w] = r & valid;  // append_null stores false under a null (boolean.rs:275-278)
        out_valid[w] = valid;
    }
}

// ---- take (record_batch.rs:108-178): gather rows by index, one output row per thread ------------------------------
// 8-byte columns: out[i] = valid(idx[i]) ? in[idx[i]] : 0, validity bit i = valid(idx[i]) (one ballot word per warp)
static __global__ void take_col8_kernel(const uint64_t* __restrict__ in, BitSrc valid, const int64_t* __restrict__ idx, int64_t n,
                                        uint64_t* __restrict__ out, uint32_t* __restrict__ out_valid) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (i < n) {
        const int64_t r = idx[i];
        ok = true;
        if (valid.words != nullptr) { const uint64_t bit = valid.bit0 + (uint64_t)r; ok = (__ldg(valid.words + (bit >> 5)) >> (bit & 31)) & 1u; }
        out[i] = ok ? in[r] : 0ull;
    }
    const uint32_t w = __ballot_sync(0xFFFFFFFFu, ok);
    if (out_valid != nullptr && (threadIdx.x & 31) == 0 && i < n) out_valid[i >> 5] = w;
}
// bit-packed values (Boolean) and/or validity
static __global__ void take_bits_kernel(BitSrc vals, BitSrc valid, const int64_t* __restrict__ idx, int64_t n, uint32_t* __restrict__ out_vals,
                                        uint32_t* __restrict__ out_valid) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false, v = false;
    if (i < n) {
        const uint64_t r = (uint64_t)idx[i];
        ok = true;
        if (valid.words != nullptr) { const uint64_t bit = valid.bit0 + r; ok = (__ldg(valid.words + (bit >> 5)) >> (bit & 31)) & 1u; }
        if (ok && vals.words != nullptr) { const uint64_t bit = vals.bit0 + r; v = (__ldg(vals.words + (bit >> 5)) >> (bit & 31)) & 1u; }
    }
    const uint32_t wv = __ballot_sync(0xFFFFFFFFu, v), wk = __ballot_sync(0xFFFFFFFFu, ok);
    if ((threadIdx.x & 31) == 0 && i < n) {
        if (out_vals != nullptr) out_vals[i >> 5] = wv;
        if (out_valid != nullptr) out_valid[i >> 5] = wk;
    }
}
// strings: lens[i] = byte length of the taken string (0 under a null)
static __global__ void take_strlen_kernel(const int32_t* __restrict__ offsets, BitSrc valid, const int64_t* __restrict__ idx, int64_t n,
                                          int32_t* __restrict__ lens) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t r = idx[i];
    bool ok = true;
    if (valid.words != nullptr) { const uint64_t bit = valid.bit0 + (uint64_t)r; ok = (__ldg(valid.words + (bit >> 5)) >> (bit & 31)) & 1u; }
    lens[i] = ok ? offsets[r + 1] - offsets[r] : 0;
}
// one warp per taken string: copy its bytes to the new offsets
static __global__ void take_strcopy_kernel(const int32_t* __restrict__ offsets, const uint8_t* __restrict__ data, const int64_t* __restrict__ idx,
                                           int64_t n, const int32_t* __restrict__ out_offsets, uint8_t* __restrict__ out_data) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int32_t len = out_offsets[i + 1] - out_offsets[i];
    const uint8_t* src = data + offsets[idx[i]];
    uint8_t* dst = out_data + out_offsets[i];
    for (int32_t j = lane; j < len; j += 32) dst[j] = __ldg(src + j);
}

// out[i + 1] = in[i + 1] - in[0] + byte_base   for i in [0, n)   (string concat / download rebasing)
static __global__ void rebase_offsets_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n, int32_t byte_base) {
    const int32_t first = in[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i + 1] = in[i + 1] - first + byte_base;
}

}  // namespace rvl
