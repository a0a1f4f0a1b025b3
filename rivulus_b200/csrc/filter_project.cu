// filter_project.cu — host side of the fused Filter + Select + Limit operator: lowers an rvl_predicate to a
// kernel predicate plan (the eager truth table of physical_plan/plan.rs:112-130 over datatypes/series.rs:87-117,
// or the streaming mask rule of execution/record_batch.rs:235-240), sizes the outputs, launches the kernels of
// fused_filter.cuh / string_kernels.cuh on the context stream and builds the output RecordBatch.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "aux_kernels.cuh"
#include "chunk_kernels.cuh"
#include "compact_kernels.cuh"
#include "filter_project.cuh"
#include "fused_filter.cuh"
#include "string_kernels.cuh"

namespace rvl {

// truth masks over {bit0: value < literal, bit1: ==, bit2: >, bit3: unordered/incomparable}
static uint32_t truth_of(int op) {
    switch (op) {
        case RVL_OP_EQ: return 0x2u;     // ==  (PartialEq, series.rs:87-98)
        case RVL_OP_NOTEQ: return 0xDu;  // !=  true for <, > and incomparable (NaN, cross-type)
        case RVL_OP_LT: return 0x1u;     // partial_cmp == Some(Less)
        case RVL_OP_GT: return 0x4u;
        case RVL_OP_LTEQ: return 0x3u;
        case RVL_OP_GTEQ: return 0x6u;
        default: return 0u;
    }
}

struct PredPlan {
    int kind = kPredTrue;
    uint32_t truth = 0, keep_null = 0, pb_a = 0, pb_b = 0;
    int64_t lit_bits = 0;
    uint64_t range_lo = 0, range_span = 0;
    uint32_t range_neg = 0;
    const uint64_t* values = nullptr;
    int vec_ok = 0;
    BitSrc valid{nullptr, 0, 0};
    BitSrc pb_vals{nullptr, 0, 0};
    BufRef tmp_bits;  // string-compare result bitmap
    BufRef tmp_lit;
};

static int lower_predicate(const CoreRef& core, const rvl_batch* in, const rvl_predicate* pred, PredPlan* pp) {
    const int64_t n = in->num_rows;
    if (pred == nullptr || pred->mode == RVL_PRED_TRUE) { pp->kind = kPredTrue; return RVL_OK; }
    if (pred->column < 0 || pred->column >= (int32_t)in->cols.size())
        return fail(RVL_COLUMN_NOT_FOUND, "Column not found: index " + std::to_string(pred->column));
    const DevColumn& c = in->cols[pred->column];
    pp->valid = bitsrc_of(c.validity, c.offset, n);

    if (pred->mode == RVL_PRED_BOOL_COLUMN) {
        // FilterStream: keep only Some(true) (stream.rs:147-153, record_batch.rs:235-240)
        if (c.dtype != RVL_BOOLEAN) return fail(RVL_TYPE_MISMATCH, "Predicate column is not of boolean type");
        pp->kind = kPredBits; pp->pb_vals = bitsrc_of(c.values, c.offset, n); pp->pb_a = 1; pp->pb_b = 0; pp->keep_null = 0;
        return RVL_OK;
    }
    if (pred->mode != RVL_PRED_CMP_LITERAL) return fail(RVL_INVALID_ARGUMENT, "unknown predicate mode");
    const uint32_t truth = truth_of(pred->op);
    if (truth == 0u) return fail(RVL_INVALID_OPERATION, "Invalid operation: operator " + std::to_string(pred->op) + " is not a comparison");  // plan.rs:121-127

    if (pred->tag_column != 0) {
        // Float64-dtype Series holding Int64 values: evaluate row by row with the row's own type, into a bitmap
        const int32_t tc = pred->tag_column - 1;
        if (tc < 0 || tc >= (int32_t)in->cols.size()) return fail(RVL_COLUMN_NOT_FOUND, "Column not found: tag index " + std::to_string(tc));
        const DevColumn& t = in->cols[tc];
        if (t.dtype != RVL_BOOLEAN || c.dtype != RVL_FLOAT64) return fail(RVL_TYPE_MISMATCH, "tag_column needs a Float64 predicate column and a Boolean tag column");
        RVL_TRY(dev_alloc_zeroed(core, (size_t)((n + 63) / 64) * 8, &pp->tmp_bits));
        int lit_kind = 3;
        int64_t lit_bits = 0;
        if (pred->lit_dtype == RVL_NULL) lit_kind = 0;
        else if (pred->lit_dtype == RVL_INT64) { lit_kind = 1; lit_bits = pred->lit_i64; }
        else if (pred->lit_dtype == RVL_FLOAT64) { lit_kind = 2; std::memcpy(&lit_bits, &pred->lit_f64, 8); }
        if (n > 0) {
            mixed_predicate_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, core->stream>>>(
                n, (const uint64_t*)c.values->ptr + c.offset, pp->valid, bitsrc_of(t.values, t.offset, n), lit_kind, lit_bits, truth,
                (uint32_t*)pp->tmp_bits->ptr);
            core->launches++;
            RVL_CUDA_TRY(cudaGetLastError());
        }
        pp->kind = kPredBits; pp->pb_vals = bitsrc_of(pp->tmp_bits, 0, n); pp->pb_a = 1; pp->pb_b = 0; pp->keep_null = 0;
        pp->valid = BitSrc{nullptr, 0, 0};
        return RVL_OK;
    }

    // constant result on valid rows / on null rows, per the truth table
    auto constant = [&](uint32_t on_valid, uint32_t on_null) {
        pp->kind = kPredBits; pp->pb_vals = BitSrc{nullptr, 0, 0}; pp->pb_a = 0; pp->pb_b = on_valid; pp->keep_null = on_null;
        if (c.dtype == RVL_NULL) { pp->valid = BitSrc{nullptr, 0, 0}; pp->pb_b = on_null; }  // NullArray: every row is null
        return RVL_OK;
    };
    if (pred->lit_dtype == RVL_NULL) {
        // value vs Null: Greater (series.rs:107); Null vs Null: Equal (:105)
        return constant((truth >> 2) & 1u, (truth >> 1) & 1u);
    }
    const uint32_t null_row = truth & 1u;  // Null vs non-null literal: Less (series.rs:106)
    if (c.dtype == RVL_NULL || c.dtype != pred->lit_dtype) {
        // different non-null types never compare (series.rs:95,114): only != holds
        return constant((truth >> 3) & 1u, null_row);
    }
    pp->keep_null = null_row;
    pp->truth = truth;
    switch (c.dtype) {
        case RVL_INT64:
        case RVL_FLOAT64: {
            pp->kind = c.dtype == RVL_INT64 ? kPredI64 : kPredF64;
            pp->values = (const uint64_t*)c.values->ptr + c.offset;
            pp->vec_ok = (reinterpret_cast<uintptr_t>(pp->values) & 15) == 0;
            if (c.dtype == RVL_INT64) {
                // every i64 comparison is one unsigned range test in the kernel: keep = ((v - lo) <= span) != neg
                const int64_t L = pred->lit_i64;
                const uint64_t umin = (uint64_t)INT64_MIN, umax = (uint64_t)INT64_MAX;
                auto range = [&](uint64_t lo, uint64_t hi, uint32_t neg) { pp->range_lo = lo; pp->range_span = hi - lo; pp->range_neg = neg; };
                auto none = [&]() { pp->range_lo = 0; pp->range_span = UINT64_MAX; pp->range_neg = 1; };
                switch (pred->op) {
                    case RVL_OP_EQ: range((uint64_t)L, (uint64_t)L, 0); break;
                    case RVL_OP_NOTEQ: range((uint64_t)L, (uint64_t)L, 1); break;
                    case RVL_OP_LT: if (L == INT64_MIN) none(); else range(umin, (uint64_t)(L - 1), 0); break;
                    case RVL_OP_LTEQ: range(umin, (uint64_t)L, 0); break;
                    case RVL_OP_GT: if (L == INT64_MAX) none(); else range((uint64_t)(L + 1), umax, 0); break;
                    default: range((uint64_t)L, umax, 0); break;  // GTEQ
                }
                pp->lit_bits = L;
            } else {
                std::memcpy(&pp->lit_bits, &pred->lit_f64, 8);
            }
            return RVL_OK;
        }
        case RVL_BOOLEAN: {
            // false < true (series.rs:112): evaluate the op for value=false and value=true, fold into (vals & a) ^ b
            const int lit = pred->lit_bool ? 1 : 0;
            auto code = [&](int v) { return v < lit ? 1u : (v == lit ? 2u : 4u); };
            const uint32_t keep_f = (truth & code(0)) ? 1u : 0u, keep_t = (truth & code(1)) ? 1u : 0u;
            pp->kind = kPredBits; pp->pb_vals = bitsrc_of(c.values, c.offset, n);
            if (keep_t && !keep_f) { pp->pb_a = 1; pp->pb_b = 0; }
            else if (!keep_t && keep_f) { pp->pb_a = 1; pp->pb_b = 1; }
            else { pp->pb_a = 0; pp->pb_b = keep_t; }
            return RVL_OK;
        }
        case RVL_STRING: {
            if (pred->lit_str_len < 0 || pred->lit_str_len > INT32_MAX) return fail(RVL_INVALID_ARGUMENT, "bad string literal length");
            RVL_TRY(dev_alloc(core, (size_t)pred->lit_str_len + 1, &pp->tmp_lit));
            if (pred->lit_str_len > 0)
                RVL_CUDA_TRY(cudaMemcpyAsync(pp->tmp_lit->ptr, pred->lit_str, (size_t)pred->lit_str_len, cudaMemcpyHostToDevice, core->stream));
            RVL_TRY(dev_alloc_zeroed(core, (size_t)((n + 63) / 64) * 8, &pp->tmp_bits));
            if (n > 0) {
                string_predicate_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, core->stream>>>(
                    n, (const int32_t*)c.offsets->ptr + c.offset, (const uint8_t*)c.data->ptr, (const uint8_t*)pp->tmp_lit->ptr,
                    (int32_t)pred->lit_str_len, truth, (uint32_t*)pp->tmp_bits->ptr);
                core->launches++;
                RVL_CUDA_TRY(cudaGetLastError());
            }
            pp->kind = kPredBits; pp->pb_vals = bitsrc_of(pp->tmp_bits, 0, n); pp->pb_a = 1; pp->pb_b = 0;
            return RVL_OK;
        }
        default: return fail(RVL_INVALID_ARGUMENT, "bad dtype");
    }
}

static void launch_fused(const CoreRef& core, int kind, const FusedParams& fp) {
    // one CTA per super-tile of 8192 rows, taken in blockIdx order (the look-back relies on in-order dispatch)
    const dim3 grid((unsigned)fp.n_super), block(kBlock);
    const size_t smem = sizeof(FusedSmem);
    switch (kind) {
        case kPredI64: fused_filter_project_kernel<kPredI64><<<grid, block, smem, core->stream>>>(fp); break;
        case kPredF64: fused_filter_project_kernel<kPredF64><<<grid, block, smem, core->stream>>>(fp); break;
        case kPredBits: fused_filter_project_kernel<kPredBits><<<grid, block, smem, core->stream>>>(fp); break;
        default: fused_filter_project_kernel<kPredTrue><<<grid, block, smem, core->stream>>>(fp); break;
    }
    core->launches++;
}

// first pass of the two-pass plan (scan_kernels.cuh)
template <int PRED, int W, int R>
static int launch_scan_t(const CoreRef& core, const ScanParams& sp, int ctas) {
    const size_t smem = (size_t)W * sp.n_slots * (ScanShape<W, R>::kItemBytes + 8);
    predicate_scan_kernel<PRED, W, R><<<(unsigned)ctas, W * 32, smem, core->stream>>>(sp);
    core->launches++;
    RVL_CUDA_TRY(cudaGetLastError());
    return RVL_OK;
}
template <int W, int R>
static int launch_scan_w(const CoreRef& core, int kind, const ScanParams& sp, int ctas) {
    switch (kind) {
        case kPredI64: return launch_scan_t<kPredI64, W, R>(core, sp, ctas);
        case kPredF64: return launch_scan_t<kPredF64, W, R>(core, sp, ctas);
        case kPredBits: return launch_scan_t<kPredBits, W, R>(core, sp, ctas);
        default: return launch_scan_t<kPredTrue, W, R>(core, sp, ctas);
    }
}
// (warps, rows per ring slot): the default shapes hold 64 KB per ring depth; the small-slot shapes of 8 warps are the ones the
// round-2 sweep (profiles/r02_scan_sweep.txt) found fastest
static int scan_slot_cap(int warps, int item_rows) {   // deepest ring that fits in shared memory
    const int rows = item_rows > 0 ? item_rows : 8192 / warps;
    return std::max(1, std::min(16, (int)((220u * 1024u) / ((size_t)warps * ((size_t)rows * 8 + 8)))));
}
static int launch_scan(const CoreRef& core, int kind, const ScanParams& sp, int ctas, int warps, int item_rows) {
    if (warps == 32) return launch_scan_w<32, 0>(core, kind, sp, ctas);
    if (warps == 16) return item_rows == 256 ? launch_scan_w<16, 256>(core, kind, sp, ctas) : launch_scan_w<16, 0>(core, kind, sp, ctas);
    if (item_rows == 512) return launch_scan_w<8, 512>(core, kind, sp, ctas);
    if (item_rows == 256) return launch_scan_w<8, 256>(core, kind, sp, ctas);
    return launch_scan_w<8, 0>(core, kind, sp, ctas);
}

// second pass of the two-pass plan (compact_kernels.cuh): dense tiles through the TMA ring, sparse tiles gathered
constexpr int kDenseSmemMax = 227 * 1024;
// bitmap staging area of the dense kernel: [consumer warps][2 halves][n_bsrc][words per warp and tile + 1] 32-bit words
static size_t dense_stage_bytes(int warps, int n_bsrc) { return (size_t)warps * 2 * (size_t)n_bsrc * (size_t)(kTileWords / warps + 1) * 4; }
template <int CW>
static int launch_dense_t(const CoreRef& core, const CompactParams& cp, int per_sm) {
    const size_t smem = (size_t)cp.n_slots * (kSlotBytes + 16) + dense_stage_bytes(CW, cp.n_bsrc);
    compact_dense_kernel<CW><<<(unsigned)(core->sm_count * per_sm), (CW + 1) * 32, smem, core->stream>>>(cp);
    core->launches++;
    RVL_CUDA_TRY(cudaGetLastError());
    return RVL_OK;
}

// Every kernel instantiation that needs more than 48 KB of dynamic shared memory is opted in once per device, when the
// context is created (the attribute is per device and per function; doing it lazily from the launch path raced when
// two contexts were driven from different host threads).
template <int PRED, int W, int R>
static int scan_opt_in_shape() {
    RVL_CUDA_TRY(cudaFuncSetAttribute(predicate_scan_kernel<PRED, W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      W * scan_slot_cap(W, R) * (int)(ScanShape<W, R>::kItemBytes + 8)));
    return RVL_OK;
}
template <int PRED>
static int scan_opt_in() {
    RVL_TRY((scan_opt_in_shape<PRED, 8, 0>()));
    RVL_TRY((scan_opt_in_shape<PRED, 8, 512>()));
    RVL_TRY((scan_opt_in_shape<PRED, 8, 256>()));
    RVL_TRY((scan_opt_in_shape<PRED, 16, 0>()));
    RVL_TRY((scan_opt_in_shape<PRED, 16, 256>()));
    RVL_TRY((scan_opt_in_shape<PRED, 32, 0>()));
    RVL_CUDA_TRY(cudaFuncSetAttribute(fused_filter_project_kernel<PRED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem)));
    return RVL_OK;
}
int fp_init_device(int device) {
    RVL_CUDA_TRY(cudaSetDevice(device));
    RVL_TRY(scan_opt_in<kPredI64>());
    RVL_TRY(scan_opt_in<kPredF64>());
    RVL_TRY(scan_opt_in<kPredBits>());
    RVL_TRY(scan_opt_in<kPredTrue>());
    RVL_CUDA_TRY(cudaFuncSetAttribute(compact_dense_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDenseSmemMax));
    RVL_CUDA_TRY(cudaFuncSetAttribute(compact_dense_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDenseSmemMax));
    RVL_CUDA_TRY(cudaFuncSetAttribute(string_gather_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StrSmem)));
    RVL_CUDA_TRY(cudaFuncSetAttribute(chunk_filter_kernel<kPredI64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDenseSmemMax));
    RVL_CUDA_TRY(cudaFuncSetAttribute(chunk_filter_kernel<kPredF64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDenseSmemMax));
    return RVL_OK;
}

static int launch_compaction(const CoreRef& core, CompactParams cp, const uint32_t* dense_list, const uint32_t* sparse_list,
                             const uint32_t* list_counts, cudaStream_t bits_stream) {
    const int warps = core->dense_warps == 16 ? 16 : 8;
    const int per_sm = warps == 16 ? 1 : std::max(1, std::min(2, core->dense_ctas_per_sm));
    // every bitmap the dense kernel reads, listed once (staged ahead by its consumer warps)
    cp.n_bsrc = 0;
    auto add_src = [&](const BitSrc& b) -> int8_t {
        if (b.words == nullptr) return (int8_t)-1;
        cp.bsrc[cp.n_bsrc] = b;
        return (int8_t)cp.n_bsrc++;
    };
    for (int c = 0; c < kMaxCol8; ++c) cp.col8_vsrc[c] = c < cp.n_col8 ? add_src(cp.col8[c].valid) : (int8_t)-1;
    const int smem_budget = (per_sm == 1 ? kDenseSmemMax : 110 * 1024) - (int)dense_stage_bytes(warps, cp.n_bsrc);
    int max_slots = std::min(14, smem_budget / (int)(kSlotBytes + 16));
    // kernels forked onto the side stream (bit-packed compaction, string sizes) must find room on the SMs the persistent dense CTAs
    // occupy: with all 14 slots the dense kernel owns every byte of shared memory and the side kernels simply queue behind it
    // (measured: no overlap at all), so two slots (32 KB per SM) are left free whenever there is side work
    if (bits_stream != core->stream) max_slots = std::min(max_slots, 12);
    cp.n_slots = std::max(2, std::min(max_slots, core->dense_slots));
    if (cp.n_bits > 0) {
        // bit-packed columns (validity bitmaps, Boolean values) of every tile: one warp per tile; instruction-bound, barely touches
        // DRAM, so the caller forks it onto the side stream where it runs underneath the HBM-bound kernels below
        const int64_t n_tiles = (cp.n_rows + kTileRows - 1) / kTileRows;
        // on the side stream it shares the SMs with the persistent dense CTAs (544 threads each): two CTAs per SM leave them room
        const int per_sm_bits = bits_stream != core->stream ? 2 : 8;
        const int64_t ctas = std::max<int64_t>(1, std::min<int64_t>((n_tiles + kWarps - 1) / kWarps, (int64_t)core->sm_count * per_sm_bits));
        compact_bits_kernel<<<(unsigned)ctas, kBlock, 0, bits_stream>>>(cp);
        core->launches++;
        RVL_CUDA_TRY(cudaGetLastError());
    }
    if (cp.n_col8 > 0) {
        cp.list = dense_list; cp.list_count = list_counts;
        if (warps == 16) RVL_TRY(launch_dense_t<16>(core, cp, per_sm));
        else RVL_TRY(launch_dense_t<8>(core, cp, per_sm));
        cp.list = sparse_list; cp.list_count = list_counts + 1;
        gather_sparse_kernel<<<(unsigned)(core->sm_count * 8), kBlock, 0, core->stream>>>(cp);
        core->launches++;
        RVL_CUDA_TRY(cudaGetLastError());
    }
    return RVL_OK;
}

// order `to` behind everything enqueued on `from` so far
static int stream_after(const CoreRef& core, cudaStream_t from, cudaStream_t to) {
    cudaEvent_t ev = core->take_event();
    if (ev == nullptr) return fail(RVL_CUDA, "cudaEventCreate failed");
    cudaError_t e = cudaEventRecord(ev, from);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(to, ev, 0);
    core->give_event(ev);
    RVL_CUDA_TRY(e);
    return RVL_OK;
}

int str_prepare_sizes(const CoreRef& core, StrGatherParams& sp, const DevColumn& src, std::vector<BufRef>* keep) {
    const int64_t tiles = (sp.n_rows + kTileRows - 1) / kTileRows;
    if (tiles <= 0) return RVL_OK;
    if (core->string_kernel == 1) {
        BufRef tbytes;   // per-tile survivor bytes -> exclusive prefixes; last word = ticket counter of the sizes kernel
        RVL_TRY(dev_alloc(core, (size_t)(tiles + 1) * 8, &tbytes));
        RVL_CUDA_TRY(cudaMemsetAsync((uint64_t*)tbytes->ptr + tiles, 0, 8, core->stream));
        keep->push_back(tbytes);
        sp.tile_bytes = (uint64_t*)tbytes->ptr;
    } else {
        // every warp of a persistent grid owns a contiguous range of tiles
        const int ctas = (int)std::max<int64_t>(1, std::min<int64_t>((tiles + kWarps - 1) / kWarps, std::min(384, 2 * core->sm_count)));
        sp.n_ranges = ctas * kWarps;
        sp.tiles_per_range = std::max<int64_t>(1, (tiles + sp.n_ranges - 1) / sp.n_ranges);
        BufRef sub, ranges;
        RVL_TRY(dev_alloc(core, (size_t)tiles * 16, &sub));
        RVL_TRY(dev_alloc(core, (size_t)(sp.n_ranges + 1) * 8, &ranges));
        RVL_CUDA_TRY(cudaMemsetAsync((uint64_t*)ranges->ptr + sp.n_ranges, 0, 8, core->stream));
        keep->push_back(sub); keep->push_back(ranges);
        sp.sub_bytes = (uint64_t*)sub->ptr; sp.range_bytes = (uint64_t*)ranges->ptr;
        sp.dense_min = (uint32_t)std::max(0, core->string_dense_min);
        // bytes around the data buffer that may be read: the block a sub-tile stages is rounded to 16-byte boundaries
        sp.data_lo = src.data ? (const uint8_t*)src.data->ptr : nullptr;
        sp.data_hi = src.data ? (const uint8_t*)src.data->ptr + src.data->bytes : nullptr;
    }
    return RVL_OK;
}

int str_launch_sizes(const CoreRef& core, const StrGatherParams& sp, cudaStream_t stream) {
    const int64_t tiles = (sp.n_rows + kTileRows - 1) / kTileRows;
    if (tiles <= 0) return RVL_OK;
    if (core->string_kernel == 1) string_sizes_kernel<<<(unsigned)tiles, kBlock, 0, stream>>>(sp);
    else string_sizes_ranges_kernel<<<(unsigned)(sp.n_ranges / kWarps), kBlock, 0, stream>>>(sp);
    core->launches++;
    RVL_CUDA_TRY(cudaGetLastError());
    return RVL_OK;
}

int str_launch_gather(const CoreRef& core, const StrGatherParams& sp) {
    const int64_t tiles = (sp.n_rows + kTileRows - 1) / kTileRows;
    if (tiles <= 0) return RVL_OK;
    if (core->string_kernel == 1 || core->string_kernel == 3) string_gather_kernel<0><<<(unsigned)tiles, kBlock, 0, core->stream>>>(sp);
    else if (core->string_kernel == 4) string_gather_kernel<1><<<(unsigned)tiles, kBlock, 0, core->stream>>>(sp);
    else {
        // persistent: three CTAs per SM, each walks sub-tiles blockIdx.x, blockIdx.x + gridDim.x, ... one iteration ahead of its loads
        const int64_t subs = (sp.n_rows + kStrRows - 1) / kStrRows;
        const int64_t ctas = std::max<int64_t>(1, std::min<int64_t>(subs, (int64_t)core->sm_count * 3));
        string_gather_staged_kernel<<<(unsigned)ctas, kBlock, sizeof(StrSmem), core->stream>>>(sp);
    }
    core->launches++;
    RVL_CUDA_TRY(cudaGetLastError());
    return RVL_OK;
}

int fp_launch(const CoreRef& core, const rvl_batch* in, const rvl_predicate* pred, const int32_t* proj, int32_t nproj,
              int64_t limit, bool want_mask, const unsigned long long* base_in, unsigned long long* total_ext, FpPending** out,
              bool exact) {
    if (!in || !out || (nproj > 0 && !proj)) return fail(RVL_INVALID_ARGUMENT, "null argument");
    if (in->core->device != core->device) return fail(RVL_INVALID_ARGUMENT, "batch lives on another device than the context");
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    for (int i = 0; i < nproj; ++i)
        if (proj[i] < 0 || proj[i] >= (int32_t)in->cols.size())
            return fail(RVL_OUT_OF_BOUNDS, "Column index " + std::to_string(proj[i]) + " out of bounds for " + std::to_string(in->cols.size()) + " columns");
    const int64_t n = in->num_rows;
    if (limit < 0) limit = -1;
    const int64_t cap_worst = limit >= 0 ? std::min(n, limit) : n;
    const int64_t tiles = (n + kTileRows - 1) / kTileRows;

    auto pend = std::make_unique<FpPending>();
    pend->core = core; pend->n = n; pend->limit = limit; pend->n_launched = 0;
    // kernel-level timing (rvl_ctx_profile_*): one bracket around every kernel of this operator invocation
    struct ProfScope {
        const CoreRef& core; cudaEvent_t e0 = nullptr;
        explicit ProfScope(const CoreRef& c) : core(c) { if (core->profile) { cudaEventCreate(&e0); cudaEventRecord(e0, core->stream); } }
        void end() {
            if (!e0) return;
            cudaEvent_t e1; cudaEventCreate(&e1); cudaEventRecord(e1, core->stream);
            std::lock_guard<std::mutex> g(core->mu);
            core->prof_events.emplace_back(e0, e1); e0 = nullptr;
        }
        ~ProfScope() { if (e0) cudaEventDestroy(e0); }
    } prof(core);
    PredPlan pp;
    RVL_TRY(lower_predicate(core, in, pred, &pp));
    pend->temps.push_back(pp.tmp_bits); pend->temps.push_back(pp.tmp_lit);

    // ---- what each projected column needs; counters: [0] = base + survivors, then one per compacted validity bitmap / string column
    pend->outs.resize((size_t)nproj);
    pend->validity_counter.assign((size_t)nproj, -1);
    pend->bytes_counter.assign((size_t)nproj, -1);
    int n_counters = 1, n_col8 = 0, n_bitcols = 0, n_str = 0;
    for (int j = 0; j < nproj; ++j) {
        const DevColumn& s = in->cols[proj[j]];
        DevColumn& d = pend->outs[(size_t)j];
        d.dtype = s.dtype; d.offset = 0; d.length = 0;
        if (s.validity && s.dtype != RVL_NULL) { pend->validity_counter[(size_t)j] = n_counters++; ++n_bitcols; }
        if (s.dtype == RVL_INT64 || s.dtype == RVL_FLOAT64) ++n_col8;
        else if (s.dtype == RVL_BOOLEAN) ++n_bitcols;
        else if (s.dtype == RVL_STRING) { pend->bytes_counter[(size_t)j] = n_counters++; ++n_str; }
    }
    if ((size_t)n_counters + 1 > CtxCore::kSlotWords) return fail(RVL_INVALID_ARGUMENT, "too many projected columns");
    const int launches_needed = std::max<int>(1, std::max<int>((n_col8 + kMaxCol8 - 1) / kMaxCol8, (n_bitcols + kMaxBitCols - 1) / kMaxBitCols));
    // plan: one fused pass (small batches, one launch), or predicate scan + independent compaction pass (large batches)
    const bool two_pass = n > 0 && (n_col8 > 0 || n_bitcols > 0) && tiles < (1ll << 32) &&  // tile ids are 32-bit in the second pass
                          (core->plan_mode == 2 || (core->plan_mode == 0 && limit < 0 && n >= core->two_pass_min_rows));
    exact = exact && two_pass && base_in == nullptr;
    // per fused launch: one look-back descriptor per super-tile
    const int64_t n_super = (n + kSuperRows - 1) / kSuperRows;
    const size_t status_words = (size_t)n_super;
    const size_t counter_words = ((size_t)n_counters + 2 + 1) & ~(size_t)1;
    RVL_TRY(dev_alloc_zeroed(core, (counter_words + (two_pass || n == 0 ? 0 : status_words * (size_t)launches_needed)) * 8, &pend->counters));
    unsigned long long* dctr = (unsigned long long*)pend->counters->ptr;
    pend->n_counters = n_counters;
    // done flag lives right behind the counters
    uint32_t* done_flag = (uint32_t*)(dctr + n_counters);

    // ---- outputs + work lists.  Sized for `cap` rows; string bytes for the viewed window (or, in exact mode, the survivors' bytes).
    std::vector<Col8> col8s;
    std::vector<BitCol> bitcols;
    struct StrJob { int out_index; const DevColumn* src; StrGatherParams sp; };
    std::vector<StrJob> strjobs;
    for (int j = 0; j < nproj; ++j)
        if (in->cols[proj[j]].dtype == RVL_STRING) strjobs.push_back(StrJob{j, &in->cols[proj[j]], StrGatherParams{}});
    int64_t cap_alloc = cap_worst;
    auto alloc_outputs = [&](int64_t cap, const uint64_t* exact_counters) -> int {
        cap_alloc = cap;
        for (int j = 0; j < nproj; ++j) {
            const DevColumn& s = in->cols[proj[j]];
            DevColumn& d = pend->outs[(size_t)j];
            const BitSrc sv = bitsrc_of(s.validity, s.offset, n);
            if (s.validity && s.dtype != RVL_NULL) {
                RVL_TRY(dev_alloc_zeroed(core, (size_t)(cap + 7) / 8 + 8, &d.validity));
                bitcols.push_back(BitCol{sv, BitSrc{nullptr, 0, 0}, (uint32_t*)d.validity->ptr});
            }
            if (s.dtype == RVL_INT64 || s.dtype == RVL_FLOAT64) {
                RVL_TRY(dev_alloc(core, (size_t)cap * 8, &d.values));
                Col8 c8{};
                c8.in = (const uint64_t*)s.values->ptr + s.offset; c8.out = (uint64_t*)d.values->ptr; c8.valid = sv;
                c8.vec_ok = (reinterpret_cast<uintptr_t>(c8.in) & 15) == 0;
                col8s.push_back(c8);
            } else if (s.dtype == RVL_BOOLEAN) {
                RVL_TRY(dev_alloc_zeroed(core, (size_t)(cap + 7) / 8 + 8, &d.values));
                bitcols.push_back(BitCol{bitsrc_of(s.values, s.offset, n), sv, (uint32_t*)d.values->ptr});
            } else if (s.dtype == RVL_STRING) {
                RVL_TRY(dev_alloc(core, (size_t)(cap + 1) * 4, &d.offsets));
                RVL_CUDA_TRY(cudaMemsetAsync(d.offsets->ptr, 0, 4, core->stream));
                // survivors' bytes are a subset of the viewed window (the whole buffer when the window is not known on the host)
                size_t bytes = (size_t)(s.window_bytes >= 0 ? std::min(s.window_bytes, s.data_len) : s.data_len);
                if (exact_counters != nullptr) bytes = (size_t)exact_counters[pend->bytes_counter[(size_t)j]];
                RVL_TRY(dev_alloc(core, bytes, &d.data));
            }
        }
        return RVL_OK;
    };
    if (!exact) RVL_TRY(alloc_outputs(cap_worst, nullptr));

    if (n > 0) {
        const bool need_sel = want_mask || !strjobs.empty() || launches_needed > 1 || two_pass;
        BufRef sel, tile_prefix, lists;
        // look-back descriptors live behind the counters in the same zeroed allocation (one malloc + one memset per invocation)
        uint64_t* const status_base = two_pass ? nullptr : (uint64_t*)pend->counters->ptr + counter_words;
        if (need_sel) {
            // whole tiles, so the second-pass kernels can read 64 words per tile unconditionally
            // (the two-pass scan writes every word of every tile unless a LIMIT lets ranges stop early)
            if (two_pass && limit < 0) RVL_TRY(dev_alloc(core, (size_t)tiles * kTileWords * 4, &sel));
            else RVL_TRY(dev_alloc_zeroed(core, (size_t)tiles * kTileWords * 4, &sel));
            if (want_mask) pend->mask = sel; else pend->temps.push_back(sel);
        }
        if (!strjobs.empty() || two_pass) { RVL_TRY(dev_alloc(core, (size_t)tiles * 8, &tile_prefix)); pend->temps.push_back(tile_prefix); }

        // string columns: the per-tile byte prefixes (first kernel of the pair) only need pass 1's selection bitmap
        auto str_params = [&](StrJob& job, int64_t tiles_per_chunk, const BufRef& chunk_base) {
            const DevColumn& s = *job.src;
            DevColumn& d = pend->outs[(size_t)job.out_index];
            StrGatherParams& sp = job.sp;
            sp.n_rows = n; sp.limit = limit; sp.sel = (const uint32_t*)sel->ptr; sp.tile_prefix = (const uint64_t*)tile_prefix->ptr; sp.row_base = 0;
            if (two_pass) { sp.chunk_base = (const uint64_t*)chunk_base->ptr; sp.tiles_per_chunk = tiles_per_chunk; }
            sp.offsets = (const int32_t*)s.offsets->ptr + s.offset; sp.data = (const uint8_t*)s.data->ptr;
            sp.valid = bitsrc_of(s.validity, s.offset, n);
            sp.out_offsets = d.offsets ? (int32_t*)d.offsets->ptr : nullptr; sp.out_data = d.data ? (uint8_t*)d.data->ptr : nullptr;
            sp.byte_base_in = nullptr; sp.row_base_in = base_in;
            sp.bytes_total_out = dctr + pend->bytes_counter[(size_t)job.out_index];
        };
        // scratch of the sizes pass is allocated on the main stream; the kernel itself may run on the side stream
        auto prepare_str_sizes = [&](StrJob& job, int64_t tiles_per_chunk, const BufRef& chunk_base) -> int {
            str_params(job, tiles_per_chunk, chunk_base);
            return str_prepare_sizes(core, job.sp, *job.src, &pend->temps);
        };
        auto launch_str_sizes = [&](StrJob& job, int64_t tiles_per_chunk, const BufRef& chunk_base, cudaStream_t stream) -> int {
            str_params(job, tiles_per_chunk, chunk_base);
            return str_launch_sizes(core, job.sp, stream);
        };
        auto launch_str_gather = [&](StrJob& job, int64_t tiles_per_chunk, const BufRef& chunk_base) -> int {
            str_params(job, tiles_per_chunk, chunk_base);   // the output pointers may have been allocated since the sizes launch
            return str_launch_gather(core, job.sp);
        };

        int64_t tiles_per_chunk = 0;
        BufRef chunk_base;
        // Single-pass chunk plan (chunk_kernels.cuh): the predicate column is itself projected, so the two-pass plan would read it
        // twice.  Needs: a numeric predicate over a 16-byte-aligned view, no LIMIT (the look-back has no early exit), one column
        // group, worst-case output allocation (the count is only known at the end).
        int chunk_pred_col = -1;
        if (two_pass && core->chunk_plan != 0 && !exact && limit < 0 && launches_needed == 1 && (pp.kind == kPredI64 || pp.kind == kPredF64) && pp.vec_ok) {
            int k8 = 0;
            for (int j = 0; j < nproj && chunk_pred_col < 0; ++j) {
                const DevColumn& s = in->cols[proj[j]];
                if (s.dtype != RVL_INT64 && s.dtype != RVL_FLOAT64) continue;
                if (proj[j] == pred->column) chunk_pred_col = k8;
                ++k8;
            }
        }
        if (chunk_pred_col >= 0 && core->chunk_plan == 1) {
            // which side of the crossover is this batch on?
            double sel = core->sel_hint;
            if (core->sync_ok) {
                ChunkParams sp{};
                sp.n_rows = n; sp.pred_values = pp.values; sp.lit_bits = pp.lit_bits; sp.range_lo = pp.range_lo; sp.range_span = pp.range_span;
                sp.range_neg = pp.range_neg; sp.truth = pp.truth; sp.keep_null = pp.keep_null; sp.pred_valid = pp.valid;
                const int64_t groups = std::min<int64_t>(1024, std::max<int64_t>(1, n / 64));
                unsigned long long* cnt = dctr + n_counters + 1;   // scratch word of the (zeroed) counter block
                const unsigned grid = (unsigned)((groups * 32 + kBlock - 1) / kBlock);
                if (pp.kind == kPredI64) sample_selectivity_kernel<kPredI64><<<grid, kBlock, 0, core->stream>>>(sp, groups, cnt);
                else sample_selectivity_kernel<kPredF64><<<grid, kBlock, 0, core->stream>>>(sp, groups, cnt);
                core->launches++;
                RVL_CUDA_TRY(cudaGetLastError());
                RVL_CUDA_TRY(cudaMemcpyAsync(core->mailbox, cnt, 8, cudaMemcpyDeviceToHost, core->stream));
                RVL_CUDA_TRY(cudaMemsetAsync(cnt, 0, 8, core->stream));
                RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
                sel = (double)core->mailbox[0] / (double)(groups * 64);
            }
            if (!(sel >= core->chunk_min_sel)) chunk_pred_col = -1;
        }
        if (two_pass && chunk_pred_col >= 0) {
            const int64_t n_chunks = (n + kChunkRows - 1) / kChunkRows;
            BufRef status;   // look-back descriptors + the chunk ticket, zeroed; then one zero word standing in for the range bases
            RVL_TRY(dev_alloc_zeroed(core, (size_t)(n_chunks + 2) * 8, &status));
            pend->temps.push_back(status);
            chunk_base = status;                         // word [n_chunks + 1] stays 0: tile_info carries GLOBAL prefixes
            tiles_per_chunk = 0x7FFFFFFF;
            ChunkParams cp{};
            cp.n_rows = n; cp.n_chunks = n_chunks;
            cp.pred_values = pp.values; cp.lit_bits = pp.lit_bits; cp.range_lo = pp.range_lo; cp.range_span = pp.range_span;
            cp.range_neg = pp.range_neg; cp.truth = pp.truth; cp.keep_null = pp.keep_null; cp.pred_valid = pp.valid;
            cp.pred_col = chunk_pred_col; cp.n_col8 = (int)col8s.size();
            for (int k = 0; k < cp.n_col8; ++k) cp.col8[k] = col8s[(size_t)k];
            cp.sparse_max = (uint32_t)std::max(0, core->sparse_max);
            { const char* dbg = std::getenv("RVL_CHUNK_DEBUG"); cp.debug = dbg ? (uint32_t)std::atoi(dbg) : 0u; }
            { static const char* nap = std::getenv("RVL_CHUNK_NAP"); cp.producer_nap = nap ? (uint32_t)std::atoi(nap) : 0u; }
            if (cp.debug == 2u) { std::memset(core->mailbox + 128, 0, 128 * 8); cp.debug_words = (unsigned long long*)(core->mailbox + 128); }
            cp.base_in = base_in;
            cp.sel_out = (uint32_t*)sel->ptr; cp.tile_info = (uint64_t*)tile_prefix->ptr;
            cp.status = (uint64_t*)status->ptr; cp.ticket = (uint32_t*)((uint64_t*)status->ptr + n_chunks);
            cp.total_out = dctr;
            const size_t fixed = (sizeof(ChunkSmem) + 127) & ~size_t(127);
            cp.n_slots = std::max(1, std::min<int>(kChunkMaxSlots, (int)(((size_t)kDenseSmemMax - fixed) / kSlotBytes)));
            const size_t smem = fixed + (size_t)cp.n_slots * kSlotBytes;
            const unsigned grid = (unsigned)std::min<int64_t>(n_chunks, core->sm_count);
            if (pp.kind == kPredI64) chunk_filter_kernel<kPredI64><<<grid, kChunkThreads, smem, core->stream>>>(cp);
            else chunk_filter_kernel<kPredF64><<<grid, kChunkThreads, smem, core->stream>>>(cp);
            core->launches++;
            RVL_CUDA_TRY(cudaGetLastError());
            // the ticket / zero word sit behind the descriptors: point chunk_base at the zero word
            chunk_base = wrap_external(core, (uint64_t*)status->ptr + n_chunks + 1, 8);
            // bit-packed columns and string sizes read the selection bitmap + tile_info this kernel wrote
            for (StrJob& job : strjobs) {
                RVL_TRY(prepare_str_sizes(job, tiles_per_chunk, chunk_base));
                RVL_TRY(launch_str_sizes(job, tiles_per_chunk, chunk_base, core->stream));
            }
            if (!bitcols.empty()) {
                CompactParams bp{};
                bp.n_rows = n; bp.limit = limit; bp.sel = (const uint32_t*)sel->ptr; bp.tile_info = (const uint64_t*)tile_prefix->ptr;
                bp.chunk_base = (const uint64_t*)chunk_base->ptr; bp.tiles_per_chunk = tiles_per_chunk; bp.base_in = base_in;
                bp.n_col8 = 0; bp.n_bits = (int)bitcols.size();
                for (int k = 0; k < bp.n_bits; ++k) bp.bits[k] = bitcols[(size_t)k];
                RVL_TRY(launch_compaction(core, bp, nullptr, nullptr, nullptr, core->stream));
            }
        } else if (two_pass) {
            // pass 1: persistent predicate scan (scan_kernels.cuh), every warp owns a contiguous range of tiles
            const int scan_ctas = std::min(192, core->sm_count);
            const int scan_warps = core->scan_warps == 32 ? 32 : (core->scan_warps == 16 ? 16 : 8);
            const int64_t n_ranges = (int64_t)scan_ctas * scan_warps;
            tiles_per_chunk = std::max<int64_t>(1, (tiles + n_ranges - 1) / n_ranges);
            // [dense tile ids | sparse tile ids | n_dense, n_sparse, CTAs done, pad] and the per-range bases
            RVL_TRY(dev_alloc(core, (size_t)tiles * 8 + 16, &lists));
            pend->temps.push_back(lists);
            RVL_TRY(dev_alloc(core, (size_t)n_ranges * 8, &chunk_base));
            pend->temps.push_back(chunk_base);
            uint32_t* dense_list = (uint32_t*)lists->ptr;
            uint32_t* sparse_list = dense_list + tiles;
            uint32_t* list_counts = sparse_list + tiles;
            RVL_CUDA_TRY(cudaMemsetAsync(list_counts, 0, 16, core->stream));
            if (limit >= 0) RVL_CUDA_TRY(cudaMemsetAsync(tile_prefix->ptr, 0, (size_t)tiles * 8, core->stream));  // ranges may stop early
            ScanParams sp{};
            sp.n_rows = n; sp.n_tiles = tiles; sp.tiles_per_warp = tiles_per_chunk; sp.limit = limit;
            sp.pred_values = pp.values; sp.lit_bits = pp.lit_bits; sp.pred_valid = pp.valid; sp.truth = pp.truth;
            sp.range_lo = pp.range_lo; sp.range_span = pp.range_span; sp.range_neg = pp.range_neg;
            sp.keep_null = pp.keep_null; sp.pred_vec_ok = pp.vec_ok; sp.pb_a = pp.pb_a; sp.pb_b = pp.pb_b; sp.pb_vals = pp.pb_vals;
            const int scan_item_rows = (scan_warps == 8 && (core->scan_item_rows == 512 || core->scan_item_rows == 256)) ||
                                               (scan_warps == 16 && core->scan_item_rows == 256) ? core->scan_item_rows : 0;
            sp.n_slots = std::max(1, std::min(scan_slot_cap(scan_warps, scan_item_rows), core->scan_slots));
            sp.sparse_max = (uint32_t)std::max(0, std::min(kSparseCap, core->sparse_max));
            sp.base_in = base_in;
            sp.sel_out = (uint32_t*)sel->ptr;
            sp.tile_info = (uint64_t*)tile_prefix->ptr;
            sp.chunk_base = (uint64_t*)chunk_base->ptr;
            sp.dense_list = dense_list; sp.sparse_list = sparse_list; sp.list_counts = list_counts;
            sp.total_out = dctr;
            { static const char* dbg = std::getenv("RVL_SCAN_DEBUG"); sp.debug_skip = dbg ? (uint32_t)std::atoi(dbg) : 0u; }
            { static const char* h = std::getenv("RVL_SCAN_L2"); sp.l2_hints = h ? (uint32_t)std::atoi(h) : (uint32_t)core->scan_l2_hints; }
            RVL_TRY(launch_scan(core, pp.kind, sp, scan_ctas, scan_warps, scan_item_rows));
            if (exact) {
                // The survivor count (and every string column's survivor bytes) is known after pass 1: read it back and allocate the
                // outputs at their exact size.  One host round trip (~15 us) against a scan of megabytes to gigabytes; a 0.1 % query
                // over 10^9 rows then holds 32 MB of output instead of reserving 32 GB.
                for (StrJob& job : strjobs) {
                    RVL_TRY(prepare_str_sizes(job, tiles_per_chunk, chunk_base));
                    RVL_TRY(launch_str_sizes(job, tiles_per_chunk, chunk_base, core->stream));
                }
                RVL_CUDA_TRY(cudaMemcpyAsync(core->mailbox, dctr, (size_t)n_counters * 8, cudaMemcpyDeviceToHost, core->stream));
                RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
                const int64_t total = (int64_t)core->mailbox[0];
                RVL_TRY(alloc_outputs(std::min(cap_worst, total), core->mailbox));
            }
            // Everything below only reads what pass 1 left and writes its own buffers.  The HBM-bound kernels (dense / sparse compaction
            // of the 8-byte columns) stay on the main stream; the instruction- and latency-bound ones — bit-packed columns, the string
            // sizes pass — are forked onto the side stream and run underneath them; the main stream joins before the string gather.
            const bool fork = core->bits_overlap && core->side_stream != nullptr && !col8s.empty() && (!bitcols.empty() || (!strjobs.empty() && !exact));
            if (!exact) for (StrJob& job : strjobs) RVL_TRY(prepare_str_sizes(job, tiles_per_chunk, chunk_base));
            const cudaStream_t side = fork ? core->side_stream : core->stream;
            if (fork) RVL_TRY(stream_after(core, core->stream, side));
            if (!exact) for (StrJob& job : strjobs) RVL_TRY(launch_str_sizes(job, tiles_per_chunk, chunk_base, side));
            for (int L = 0; L < launches_needed; ++L) {
                CompactParams cp{};
                cp.n_rows = n; cp.limit = limit; cp.sel = (const uint32_t*)sel->ptr; cp.tile_info = (const uint64_t*)tile_prefix->ptr;
                cp.chunk_base = (const uint64_t*)chunk_base->ptr; cp.tiles_per_chunk = tiles_per_chunk;
                cp.base_in = base_in;
                const int c0 = L * kMaxCol8, c1 = std::min<int>((int)col8s.size(), c0 + kMaxCol8);
                cp.n_col8 = std::max(0, c1 - c0);
                for (int k = 0; k < cp.n_col8; ++k) cp.col8[k] = col8s[(size_t)(c0 + k)];
                const int b0 = L * kMaxBitCols, b1 = std::min<int>((int)bitcols.size(), b0 + kMaxBitCols);
                cp.n_bits = std::max(0, b1 - b0);
                for (int k = 0; k < cp.n_bits; ++k) cp.bits[k] = bitcols[(size_t)(b0 + k)];
                RVL_TRY(launch_compaction(core, cp, dense_list, sparse_list, list_counts, side));
            }
            if (fork) RVL_TRY(stream_after(core, side, core->stream));
        }

        for (int L = 0; L < (two_pass ? 0 : launches_needed); ++L) {
            FusedParams fp{};
            fp.n_rows = n; fp.n_super = n_super; fp.limit = limit;
            int kind = pp.kind;
            if (L == 0) {
                fp.pred_values = pp.values; fp.lit_bits = pp.lit_bits; fp.pred_valid = pp.valid; fp.truth = pp.truth;
                fp.range_lo = pp.range_lo; fp.range_span = pp.range_span; fp.range_neg = pp.range_neg;
                fp.keep_null = pp.keep_null; fp.pred_vec_ok = pp.vec_ok; fp.pb_a = pp.pb_a; fp.pb_b = pp.pb_b; fp.pb_vals = pp.pb_vals;
                fp.sel_out = need_sel ? (uint32_t*)sel->ptr : nullptr;
                fp.tile_prefix_out = tile_prefix ? (uint64_t*)tile_prefix->ptr : nullptr;
                fp.total_out = dctr;
            } else {
                // further column groups replay the selection bitmap written by the first launch
                kind = kPredBits;
                fp.pb_vals = bitsrc_of(sel, 0, n); fp.pb_a = 1; fp.pb_b = 0; fp.keep_null = 0; fp.pred_valid = BitSrc{nullptr, 0, 0};
                fp.total_out = dctr + n_counters + 1;  // scratch word
            }
            fp.base_in = base_in;
            fp.done_flag = done_flag;
            fp.tile_status = status_base + (size_t)L * status_words;
            const int c0 = L * kMaxCol8, c1 = std::min<int>((int)col8s.size(), c0 + kMaxCol8);
            fp.n_col8 = std::max(0, c1 - c0);
            for (int k = 0; k < fp.n_col8; ++k) fp.col8[k] = col8s[(size_t)(c0 + k)];
            const int b0 = L * kMaxBitCols, b1 = std::min<int>((int)bitcols.size(), b0 + kMaxBitCols);
            fp.n_bits = std::max(0, b1 - b0);
            for (int k = 0; k < fp.n_bits; ++k) fp.bits[k] = bitcols[(size_t)(b0 + k)];
            if (L > 0 && limit >= 0) RVL_CUDA_TRY(cudaMemsetAsync(done_flag, 0, 4, core->stream));
            launch_fused(core, kind, fp);
            RVL_CUDA_TRY(cudaGetLastError());
        }

        // strings: second kernel pair per column (offset prefix-sum + byte copy)
        for (StrJob& job : strjobs) {
            if (!two_pass) {
                RVL_TRY(prepare_str_sizes(job, tiles_per_chunk, chunk_base));
                RVL_TRY(launch_str_sizes(job, tiles_per_chunk, chunk_base, core->stream));
            }
            RVL_TRY(launch_str_gather(job, tiles_per_chunk, chunk_base));
        }

        // null counts of the compacted validity bitmaps (device-side row count, no host round trip)
        for (int j = 0; j < nproj; ++j) {
            if (pend->validity_counter[(size_t)j] < 0) continue;
            DevColumn& d = pend->outs[(size_t)j];
            count_ones_kernel<<<std::max<int>(1, (int)std::min<int64_t>((cap_alloc / 32 + 255) / 256, core->sm_count * 8)), 256, 0, core->stream>>>(
                bitsrc_of(d.validity, 0, cap_alloc), 0, dctr, base_in, limit, dctr + pend->validity_counter[(size_t)j]);
            core->launches++;
            RVL_CUDA_TRY(cudaGetLastError());
        }
        pend->n_launched = 1;
        prof.end();
    } else {
        if (exact) RVL_TRY(alloc_outputs(0, nullptr));
        if (base_in != nullptr)   // empty batch in a chained query: the running total passes through unchanged
            RVL_CUDA_TRY(cudaMemcpyAsync(dctr, base_in, 8, cudaMemcpyDeviceToDevice, core->stream));
    }
    pend->mailbox = core->take_slot();
    if (pend->mailbox == nullptr) return fail(RVL_OUT_OF_MEMORY, "cudaHostAlloc of a mailbox chunk failed");
    *reinterpret_cast<volatile uint64_t*>(pend->mailbox) = kMailboxPending;
    pend->chained = base_in != nullptr;
    RVL_CUDA_TRY(cudaMemcpyAsync(pend->mailbox, dctr, (size_t)n_counters * 8, cudaMemcpyDeviceToHost, core->stream));
    if (base_in != nullptr)
        RVL_CUDA_TRY(cudaMemcpyAsync(pend->mailbox + n_counters, base_in, 8, cudaMemcpyDeviceToHost, core->stream));
    if (total_ext != nullptr) RVL_CUDA_TRY(cudaMemcpyAsync(total_ext, dctr, 8, cudaMemcpyDeviceToDevice, core->stream));
    pend->done_event = core->take_event();
    if (pend->done_event == nullptr) return fail(RVL_CUDA, "cudaEventCreate failed");
    RVL_CUDA_TRY(cudaEventRecord(pend->done_event, core->stream));
    *out = pend.release();
    return RVL_OK;
}

int fp_finish(FpPending* pend, rvl_batch** out, rvl_batch** mask_out) {
    std::unique_ptr<FpPending> guard(pend);
    const CoreRef& core = pend->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    if (pend->done_event) {
        cudaError_t e = cudaEventSynchronize(pend->done_event);
        core->give_event(pend->done_event);
        pend->done_event = nullptr;
        pend->completed = true;
        if (e != cudaSuccess) return fail(RVL_CUDA, std::string("fused filter/project failed: ") + cudaGetErrorString(e));
    } else {
        RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
        pend->completed = true;
    }
    const uint64_t* mb = pend->mailbox;
    int64_t total = (int64_t)mb[0];
    int64_t base = pend->chained ? (int64_t)mb[pend->n_counters] : 0;
    if (pend->limit >= 0) { total = std::min(total, pend->limit); base = std::min(base, pend->limit); }
    const int64_t count = std::max<int64_t>(0, total - base);
    pend->base_rows = (uint64_t)base;
    if (out) {
        auto b = std::make_unique<rvl_batch>();
        b->core = core; b->num_rows = count;
        for (size_t j = 0; j < pend->outs.size(); ++j) {
            DevColumn d = pend->outs[j];
            d.length = count;
            if (d.dtype == RVL_NULL) d.null_count = count;
            else if (pend->validity_counter[j] >= 0) {
                d.null_count = count - (int64_t)mb[pend->validity_counter[j]];
                if (d.null_count == 0) d.validity.reset();  // bitmap only when a survivor is null (primitive.rs:180-185)
            } else d.null_count = 0;
            if (pend->bytes_counter[j] >= 0) { d.data_len = pend->n > 0 ? (int64_t)mb[pend->bytes_counter[j]] : 0; d.window_bytes = d.data_len; }
            b->cols.push_back(std::move(d));
        }
        *out = b.release();
    }
    if (mask_out) {
        auto m = std::make_unique<rvl_batch>();
        m->core = core; m->num_rows = pend->n;
        DevColumn d;
        d.dtype = RVL_BOOLEAN; d.length = pend->n; d.offset = 0; d.null_count = 0;
        if (pend->mask) d.values = pend->mask;
        else RVL_TRY(dev_alloc_zeroed(core, 8, &d.values));
        m->cols.push_back(std::move(d));
        *mask_out = m.release();
    }
    return RVL_OK;
}

}  // namespace rvl

using namespace rvl;

struct rvl_pending {
    rvl::FpPending* p;
};

extern "C" {

int32_t rvl_filter_project_launch(rvl_ctx* ctx, const rvl_batch* in, const rvl_predicate* pred, const int32_t* proj, int32_t nproj,
                                  int64_t limit, rvl_pending** pending) {
    if (!ctx || !pending) return fail(RVL_INVALID_ARGUMENT, "null argument");
    FpPending* p = nullptr;
    RVL_TRY(fp_launch(ctx->core, in, pred, proj, nproj, limit, false, nullptr, nullptr, &p));
    *pending = new rvl_pending{p};
    return RVL_OK;
}

int32_t rvl_filter_project_finish(rvl_ctx* ctx, rvl_pending* pending, rvl_batch** out) {
    if (!pending || !out) return fail(RVL_INVALID_ARGUMENT, "null argument");
    FpPending* p = pending->p;
    delete pending;
    return fp_finish(p, out, nullptr);
}

int32_t rvl_filter_project(rvl_ctx* ctx, const rvl_batch* in, const rvl_predicate* pred, const int32_t* proj, int32_t nproj,
                           int64_t limit, rvl_batch** out) {
    if (!ctx || !out) return fail(RVL_INVALID_ARGUMENT, "null argument");
    FpPending* p = nullptr;
    // Blocking call: the two-pass plan may wait for the survivor count and size the outputs exactly.  That costs one host round trip
    // in the middle of the operator (~50 us of idle device), so AUTO only pays it when the worst case — min(n, limit) rows of every
    // projected column — would pin more than a quarter of the device's memory.
    bool exact = ctx->core->exact_alloc == 1;
    if (ctx->core->exact_alloc == 2 && in != nullptr) {
        const int64_t cap = limit >= 0 ? std::min<int64_t>(in->num_rows, limit) : in->num_rows;
        size_t worst = 0;
        for (int j = 0; j < nproj && proj; ++j) {
            if (proj[j] < 0 || proj[j] >= (int32_t)in->cols.size()) continue;
            const DevColumn& s = in->cols[proj[j]];
            if (s.dtype == RVL_INT64 || s.dtype == RVL_FLOAT64) worst += (size_t)cap * 8;
            else if (s.dtype == RVL_STRING) worst += (size_t)cap * 4 + (size_t)(s.window_bytes >= 0 ? s.window_bytes : s.data_len);
            else worst += (size_t)cap / 8;
        }
        exact = worst > ctx->core->device_bytes / 4;
    }
    ctx->core->sync_ok = true;    // a blocking call may spend a host round trip on choosing its plan
    const int rc = fp_launch(ctx->core, in, pred, proj, nproj, limit, false, nullptr, nullptr, &p, exact);
    ctx->core->sync_ok = false;
    RVL_TRY(rc);
    return fp_finish(p, out, nullptr);
}

int32_t rvl_predicate_mask(rvl_ctx* ctx, const rvl_batch* in, const rvl_predicate* pred, rvl_batch** mask_out) {
    if (!ctx || !mask_out) return fail(RVL_INVALID_ARGUMENT, "null argument");
    FpPending* p = nullptr;
    RVL_TRY(fp_launch(ctx->core, in, pred, nullptr, 0, -1, true, nullptr, nullptr, &p));
    return fp_finish(p, nullptr, mask_out);
}

int32_t rvl_shard_range(int64_t n_rows, int32_t rank, int32_t world, int64_t* begin, int64_t* end) {
    if (world <= 0 || rank < 0 || rank >= world || n_rows < 0) return fail(RVL_INVALID_ARGUMENT, "bad shard arguments");
    // ceil(n / world) rounded up to a multiple of 64 rows so bitmap words never straddle two GPUs
    int64_t per = (n_rows + world - 1) / world;
    per = (per + 63) / 64 * 64;
    *begin = std::min<int64_t>(n_rows, per * rank);
    *end = std::min<int64_t>(n_rows, per * (rank + 1));
    return RVL_OK;
}

int32_t rvl_shard_limit_split(const int64_t* counts, int32_t world, int64_t limit, int64_t* take) {
    int64_t before = 0;
    for (int g = 0; g < world; ++g) {
        if (limit < 0) take[g] = counts[g];
        else take[g] = std::max<int64_t>(0, std::min<int64_t>(counts[g], limit - before));
        before += counts[g];
    }
    return RVL_OK;
}

int32_t rvl_filter_project_sharded(rvl_ctx* const* ctxs, int32_t n, const rvl_batch* const* shards, const rvl_predicate* pred,
                                   const int32_t* proj, int32_t nproj, int64_t limit, rvl_batch** outs, int64_t* counts) {
    if (!ctxs || !shards || !outs || n <= 0) return fail(RVL_INVALID_ARGUMENT, "null argument");
    // every GPU stops at `limit` local survivors; the ordered result takes clamp(limit - sum_{j<g} count_j, 0, count_g) from shard g
    std::vector<FpPending*> pend((size_t)n, nullptr);
    int rc = RVL_OK;
    for (int g = 0; g < n && rc == RVL_OK; ++g) rc = fp_launch(ctxs[g]->core, shards[g], pred, proj, nproj, limit, false, nullptr, nullptr, &pend[(size_t)g]);
    std::vector<int64_t> local((size_t)n, 0), take((size_t)n, 0);
    for (int g = 0; g < n; ++g) {
        outs[g] = nullptr;
        if (!pend[(size_t)g]) continue;
        const int r2 = fp_finish(pend[(size_t)g], &outs[g], nullptr);
        if (rc == RVL_OK) rc = r2;
        if (outs[g]) local[(size_t)g] = outs[g]->num_rows;
    }
    if (rc != RVL_OK) {
        for (int g = 0; g < n; ++g) { delete outs[g]; outs[g] = nullptr; }
        return rc;
    }
    rvl_shard_limit_split(local.data(), n, limit, take.data());
    for (int g = 0; g < n; ++g) {
        if (take[(size_t)g] < local[(size_t)g]) {
            rvl_batch* cut = nullptr;
            const int r3 = rvl_batch_slice(outs[g], 0, take[(size_t)g], &cut);
            if (r3 != RVL_OK) {   // hand nothing half-built back: the caller only sees the error
                for (int h = 0; h < n; ++h) { delete outs[h]; outs[h] = nullptr; }
                return r3;
            }
            delete outs[g];
            outs[g] = cut;
        }
        if (counts) counts[g] = take[(size_t)g];
    }
    return RVL_OK;
}

}  // extern "C"
