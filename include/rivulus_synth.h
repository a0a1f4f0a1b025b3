/* rivulus_synth.h — the synthetic-table generator shared by tests, bench, oracle and device.
 *
 * Counter-based: every cell is a pure function of (seed, column id, global row index), so any
 * row range of any table is reproducible on the host and on any GPU without materialising the
 * table anywhere else.  Plain C99 integer code; also compiles
 * as CUDA device code.  This is test/bench input generation, not part of the hot path.
 */
#ifndef RIVULUS_SYNTH_H
#define RIVULUS_SYNTH_H
#include <stdint.h>

#if defined(__CUDACC__)
#define RVL_HD __host__ __device__ __forceinline__
#else
#define RVL_HD static inline
#endif

#define RVL_SYNTH_SEED 42ull
#define RVL_SYNTH_GOLDEN 0x9E3779B97F4A7C15ull
#define RVL_SYNTH_NULL_SALT 0x5851F42D4C957F2Dull

/* column "kinds" understood by rvl_gen_column / the oracle generator */
enum rvl_synth_kind {
    RVL_SYNTH_KEY1000 = 0, /* Int64   u % 1000   (k > 998/899/499/99  =>  0.1/10/50/90 %) */
    RVL_SYNTH_I64 = 1,     /* Int64   (int64)u */
    RVL_SYNTH_F64 = 2,     /* Float64 (u >> 11) * 2^-53 * 1000.0   in [0, 1000) */
    RVL_SYNTH_BOOL = 3,    /* Boolean u & 1 (bit-packed, LSB-first) */
    RVL_SYNTH_AGE100 = 4,  /* Int64   u % 100 */
    RVL_SYNTH_STR = 5      /* String  len = 8 + u % 33, bytes 'a' + splitmix64(u + j) % 26 */
};

RVL_HD uint64_t rvl_splitmix64(uint64_t x) {
    x += RVL_SYNTH_GOLDEN;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* the per-cell 64-bit draw */
RVL_HD uint64_t rvl_synth_u(uint64_t seed, uint32_t col_id, uint64_t row) {
    return rvl_splitmix64(seed + (uint64_t)col_id * RVL_SYNTH_GOLDEN + row);
}

/* 1 = valid, 0 = null; null_pct in [0,100] */
RVL_HD int rvl_synth_valid(uint64_t seed, uint32_t col_id, uint64_t row, uint32_t null_pct) {
    if (null_pct == 0) return 1;
    uint64_t v = rvl_splitmix64((seed ^ RVL_SYNTH_NULL_SALT) + (uint64_t)col_id * RVL_SYNTH_GOLDEN + row);
    return (v % 100ull) >= (uint64_t)null_pct;
}

RVL_HD int64_t rvl_synth_i64(uint64_t u, int kind) {
    if (kind == RVL_SYNTH_KEY1000) return (int64_t)(u % 1000ull);
    if (kind == RVL_SYNTH_AGE100) return (int64_t)(u % 100ull);
    return (int64_t)u;
}

RVL_HD double rvl_synth_f64(uint64_t u) {
    return (double)(u >> 11) * (1.0 / 9007199254740992.0) * 1000.0;
}

RVL_HD uint32_t rvl_synth_strlen(uint64_t u) { return 8u + (uint32_t)(u % 33ull); }
RVL_HD uint8_t rvl_synth_strbyte(uint64_t u, uint32_t j) {
    return (uint8_t)('a' + (uint32_t)(rvl_splitmix64(u + (uint64_t)j) % 26ull));
}

/* order-sensitive 64-bit column checksum:  sum_i mix(value_bits_i + (i+1)*GOLDEN)  (wrapping).
 * Position enters the mix, so any reordering changes the sum; being a sum it reduces in parallel. */
RVL_HD uint64_t rvl_checksum_term(uint64_t value_bits, uint64_t position) {
    return rvl_splitmix64(value_bits + (position + 1ull) * RVL_SYNTH_GOLDEN);
}

#endif /* RIVULUS_SYNTH_H */
