/* rivulus_gpu.h — C ABI of the B200-native filter / project / limit path.
 *
 * This is the drop-in boundary: a Rust `-sys` crate (bindgen over this header), or the C++ host
 * layer in rivulus_b200/host/, binds exactly these entry points.  The reference (CleConor/rivulus)
 * has no FFI of its own; each entry point below names the Rust-level seam it replaces
 * (file:line under /root/reference/src).  Conventions:
 *
 *   - every function returns an rvl_status (0 = OK); the message for the last failure on the calling
 *     thread is rvl_last_error().  No exception, abort or panic crosses this boundary: the reference's
 *     panics (slice/index out of bounds) become RVL_OUT_OF_BOUNDS.
 *   - plain pointers and sizes only; buffers passed in are borrowed for the duration of the call and
 *     never freed by the library; handles returned are owned by the caller and released explicitly.
 *   - layouts are the reference's Arrow layouts: 8-byte values, LSB-first bitmaps where 1 = valid/true
 *     (bitmap.rs:44-68), int32 string offsets (string.rs:8-16), (offset, length) views over whole
 *     buffers (primitive.rs:107-117).
 *   - there is NO CPU fallback: every data-path entry point runs hand-written sm_100a kernels and
 *     fails with RVL_CUDA when no device is usable.
 */
#ifndef RIVULUS_GPU_H
#define RIVULUS_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVL_ABI_VERSION 1

typedef enum rvl_status {
    RVL_OK = 0,
    RVL_COLUMN_NOT_FOUND = 1,  /* ExecutionError::ColumnNotFound   physical_plan/plan.rs:38 */
    RVL_TYPE_MISMATCH = 2,     /* ExecutionError::TypeMismatch     physical_plan/plan.rs:44; "Predicate must be a BooleanArray" record_batch.rs:233 */
    RVL_INVALID_OPERATION = 3, /* ExecutionError::InvalidOperation physical_plan/plan.rs:47 */
    RVL_LENGTH_MISMATCH = 4,   /* record_batch.rs:33-38, 223-227 */
    RVL_SCHEMA_MISMATCH = 5,   /* record_batch.rs:253; StreamError::SchemaMismatch stream.rs:12 */
    RVL_OUT_OF_BOUNDS = 6,     /* replaces the panics at record_batch.rs:93, primitive.rs:49,108 */
    RVL_OFFSET_OVERFLOW = 7,   /* new: int32 string offsets would wrap (string.rs:31,35 wraps silently) */
    RVL_CUDA = 8,              /* new: CUDA runtime / launch failure */
    RVL_INVALID_ARGUMENT = 9,
    RVL_OUT_OF_MEMORY = 10
} rvl_status;

/* execution/schema.rs:1-8 — declaration order */
typedef enum rvl_dtype { RVL_NULL = 0, RVL_BOOLEAN = 1, RVL_INT64 = 2, RVL_FLOAT64 = 3, RVL_STRING = 4 } rvl_dtype;

/* expressions/expr.rs:15-29 — declaration order */
typedef enum rvl_op {
    RVL_OP_PLUS = 0, RVL_OP_MINUS = 1, RVL_OP_MULTIPLY = 2, RVL_OP_DIVIDE = 3,
    RVL_OP_EQ = 4, RVL_OP_NOTEQ = 5, RVL_OP_LT = 6, RVL_OP_GT = 7, RVL_OP_LTEQ = 8, RVL_OP_GTEQ = 9,
    RVL_OP_AND = 10, RVL_OP_OR = 11
} rvl_op;

typedef enum rvl_location { RVL_HOST = 0, RVL_DEVICE = 1 } rvl_location;

/* One column = the raw buffers of a PrimitiveArray<i64|f64> (primitive.rs:20-28), BooleanArray
 * (boolean.rs:9-16), StringArray (string.rs:8-16) or NullArray (null.rs:5-9).  Row i of the view is
 * element (offset + i) of every buffer.  All buffers of one column live in the same `location`. */
typedef struct rvl_column {
    int32_t dtype;           /* rvl_dtype */
    int32_t location;        /* rvl_location of the buffers below */
    int64_t length;          /* rows in the view */
    int64_t offset;          /* first row of the view inside the buffers */
    const void* values;      /* Int64/Float64: 8-byte values; Boolean: LSB-first value bitmap; else NULL */
    const uint8_t* validity; /* LSB-first, 1 = valid; NULL = no nulls (the reference drops all-valid bitmaps) */
    const int32_t* offsets;  /* String: offset+length+1 entries reachable */
    const uint8_t* data;     /* String bytes */
    int64_t data_len;        /* String: bytes in `data` */
    int64_t null_count;      /* filled by the library on views it returns; ignored on input */
} rvl_column;

typedef enum rvl_pred_mode {
    RVL_PRED_CMP_LITERAL = 0,   /* eager engine: column <op> literal, truth table of plan.rs:112-130 over series.rs:87-117 */
    RVL_PRED_BOOL_COLUMN = 1,   /* streaming engine: keep rows where the Boolean column is Some(true), record_batch.rs:235-240 */
    RVL_PRED_TRUE = 2           /* no filter (Select / Limit only) */
} rvl_pred_mode;

/* PhysicalPlan::Filter { column, value: AnyValue, op } (physical_plan/plan.rs:17-22) or
 * FilterStream { predicate_column } (stream.rs:117-120), by column index. */
typedef struct rvl_predicate {
    int32_t mode;       /* rvl_pred_mode */
    int32_t column;     /* index into the input batch */
    int32_t op;         /* rvl_op, one of EQ..GTEQ (CMP_LITERAL only) */
    int32_t lit_dtype;  /* rvl_dtype of the literal AnyValue; RVL_NULL = AnyValue::Null */
    int64_t lit_i64;
    double lit_f64;
    const uint8_t* lit_str; /* host pointer, not NUL-terminated */
    int64_t lit_str_len;
    int32_t lit_bool;
    /* 0 = the predicate column is a plain array.  k + 1 = Boolean column k of the input batch tags the rows of a Float64-dtype
     * Series whose value is an AnyValue::Int64 (datatypes/series.rs:210-212 lets Int64 and Float64 share a Series): for tagged
     * rows `values` holds the int64 bit pattern and the row compares as an Int64 — i.e. against a Float64 literal it is a
     * different type and only != holds (series.rs:95,114); CMP_LITERAL mode, Float64-typed predicate column only. */
    int32_t tag_column;
} rvl_predicate;

typedef struct rvl_ctx rvl_ctx;       /* one per GPU: stream, memory pool, scratch, pinned mailbox */
typedef struct rvl_batch rvl_batch;   /* device-resident RecordBatch (record_batch.rs:8-13); immutable, ref-counted buffers */
typedef struct rvl_stream rvl_stream; /* device pipeline behind trait DataStream (stream.rs:25-54) */

/* ---- library ---------------------------------------------------------------------------- */
int32_t rvl_abi_version(void);
const char* rvl_last_error(void);
int32_t rvl_device_count(int32_t* count);

/* ---- context ---------------------------------------------------------------------------- */
int32_t rvl_ctx_create(int32_t device, rvl_ctx** ctx);
int32_t rvl_ctx_destroy(rvl_ctx* ctx);
int32_t rvl_ctx_synchronize(rvl_ctx* ctx);
/* Return the cached (freed, kept for reuse) blocks of the context's device memory pool to the driver.  The pool keeps everything it
 * has ever held so steady-state queries never call cudaMalloc; call this between workloads of very different footprints. */
int32_t rvl_ctx_trim(rvl_ctx* ctx);
/* Bytes the pool currently holds from the driver (reserved) and the part of them handed out to live buffers (used). */
int32_t rvl_ctx_pool_stats(rvl_ctx* ctx, uint64_t* reserved_bytes, uint64_t* used_bytes);
/* the cudaStream_t every kernel of this context is launched on (so callers can time with CUDA events on it) */
int32_t rvl_ctx_cuda_stream(rvl_ctx* ctx, void** cuda_stream);
int32_t rvl_ctx_device(rvl_ctx* ctx, int32_t* device);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
int32_t rvl_ctx_launch_count(rvl_ctx* ctx, int64_t* launches);
/* Kernel-level timing of the fused filter/project kernel: when enabled, every launch of that kernel is
 * bracketed by CUDA events on the context stream.  read() waits for the recorded launches, returns their
 * summed device time and count, and resets the accumulators (bench.py's roofline figure). */
int32_t rvl_ctx_profile_enable(rvl_ctx* ctx, int32_t enable);
int32_t rvl_ctx_profile_read(rvl_ctx* ctx, double* fused_kernel_ms, int64_t* fused_launches);
/* per-launch device times (ms, launch order) recorded since the last read; resets like profile_read */
int32_t rvl_ctx_profile_read_launches(rvl_ctx* ctx, double* ms_out, int64_t cap, int64_t* n);

/* Execution plan of rvl_filter_project.  Both plans produce identical bytes; they differ in how the work is laid out:
 *   FUSED     one kernel: predicate, look-back scan and compaction of a 8192-row super-tile per CTA (one launch; small batches,
 *             LIMIT early termination inside the kernel)
 *   TWO_PASS  persistent predicate scan (selection bitmap + per-tile output offsets, no inter-CTA dependency) followed by an
 *             independent compaction pass that streams dense tiles through a TMA shared-memory ring and gathers sparse tiles
 *   AUTO      TWO_PASS for unlimited queries over batches of at least RVL_OPT_TWO_PASS_MIN_ROWS rows, else FUSED */
typedef enum rvl_plan { RVL_PLAN_AUTO = 0, RVL_PLAN_FUSED = 1, RVL_PLAN_TWO_PASS = 2 } rvl_plan;
typedef enum rvl_option {
    RVL_OPT_PLAN = 0,               /* rvl_plan */
    RVL_OPT_TWO_PASS_MIN_ROWS = 1,  /* default 2 Mi rows */
    RVL_OPT_SPARSE_MAX = 2,         /* two-pass: 2048-row tiles with <= this many survivors are gathered (0..640) */
    RVL_OPT_DENSE_SLOTS = 3,        /* two-pass: 16 KB ring slots per CTA of the dense kernel (2..14) */
    RVL_OPT_DENSE_CTAS_PER_SM = 4,  /* 1 or 2 */
    RVL_OPT_SCAN_SLOTS = 5,         /* two-pass: ring slots per warp of the predicate scan (1..16, capped by what fits in shared memory) */
    RVL_OPT_SCAN_WARPS = 6,         /* two-pass: warps per CTA of the predicate scan (8 or 16) */
    RVL_OPT_DENSE_WARPS = 7,        /* two-pass: consumer warps per CTA of the dense kernel (8 or 16; 16 implies one CTA per SM) */
    RVL_OPT_BITS_OVERLAP = 8,       /* two-pass: run the bit-packed compaction kernel on a forked stream under the 8-byte kernels (default 1) */
    RVL_OPT_STRING_KERNEL = 10,     /* String compaction kernels: 1 = round-1 pair (CTA-per-tile sizes + gather from global memory);
                                       2 = persistent ranges sizes pass + persistent TMA-staged gather over 1024-row sub-tiles;
                                       3 = ranges sizes pass + round-1 gather (default); 4 = 3 with OR-merged string boundaries */
    RVL_OPT_STRING_DENSE_MIN = 11,  /* kernel 2: 1024-row sub-tiles with at least this many survivors fetch their whole source byte block
                                       with one TMA bulk copy (default 128); sparser ones read only the survivors' words */
    RVL_OPT_EXACT_ALLOC = 9,        /* blocking rvl_filter_project, two-pass plan: read the survivor count (and string bytes) back after the
                                       predicate scan and allocate the outputs at their exact size instead of min(n, limit) rows.
                                       0 = never, 1 = always, 2 = only when the worst case exceeds a quarter of device memory (default) */
    RVL_OPT_CHUNK_PLAN = 12,        /* two-pass plan, predicate column also projected (numeric, no LIMIT): run the single-pass chunk kernel
                                       that reads that column from HBM once instead of twice.  0 = never, 2 = always, 1 (default) = when
                                       at least ~30 % of the rows survive — estimated from a 64 K-row sample in blocking calls, taken from
                                       the previous batch in streams (below that the two-pass plan is faster) */
    RVL_OPT_SCAN_ITEM_ROWS = 13,    /* two-pass: rows per ring slot of the predicate scan: 0 = 8192 / warps (default), 512 or 256 with 8 warps,
                                       256 with 16 warps */
    RVL_OPT__COUNT = 14
} rvl_option;
int32_t rvl_ctx_set_option(rvl_ctx* ctx, int32_t option, int64_t value);

/* pinned host memory for the H2D/D2H staging of the streaming path (execution: pinned double-buffering) */
int32_t rvl_host_alloc(size_t bytes, void** ptr);
int32_t rvl_host_free(void* ptr);

/* ---- batches: RecordBatch (record_batch.rs) --------------------------------------------- */
/* RecordBatch::try_new (record_batch.rs:16-58) over borrowed host or device buffers: copies the viewed
 * window to the device (async on the context stream) and keeps the sub-64-row residual offset. */
int32_t rvl_batch_upload(rvl_ctx* ctx, const rvl_column* cols, int32_t ncols, rvl_batch** out);
/* wrap caller-owned DEVICE buffers without copying (caller keeps them alive and unmodified) */
int32_t rvl_batch_wrap_device(rvl_ctx* ctx, const rvl_column* cols, int32_t ncols, rvl_batch** out);
int32_t rvl_batch_release(rvl_batch* batch);
int32_t rvl_batch_num_rows(const rvl_batch* batch, int64_t* rows);       /* record_batch.rs:72 */
int32_t rvl_batch_num_columns(const rvl_batch* batch, int32_t* ncols);  /* record_batch.rs:76 */
/* device-side view of column i (pointers are DEVICE pointers, valid while the batch lives) */
int32_t rvl_batch_column(const rvl_batch* batch, int32_t i, rvl_column* view);
/* copy column i to caller-provided HOST buffers described by `dst` (same dtype; buffers sized from the view:
 * values length*8 bytes (or ceil(length/8) for Boolean), validity ceil(length/8) if the view has one, offsets
 * (length+1)*4, data data_len).  The copy is rebased to offset 0 like the reference's freshly built outputs. */
int32_t rvl_batch_download_column(rvl_ctx* ctx, const rvl_batch* batch, int32_t i, rvl_column* dst);

/* BooleanArray::count_true (array/boolean.rs: count_true): rows of Boolean column i that are Some(true) */
int32_t rvl_batch_count_true(rvl_ctx* ctx, const rvl_batch* batch, int32_t i, int64_t* count);

/* BooleanArray::{and, or, not} (array/boolean.rs:120-165), strict-null: the result is null wherever an input is null.  Inputs are
 * Boolean columns of device batches (any views); the result is a one-column batch built like BooleanArrayBuilder::finish leaves it.
 * RVL_LENGTH_MISMATCH "Array lengths must match for logical operations".  Masks combined this way feed RVL_PRED_BOOL_COLUMN. */
typedef enum rvl_bool_op { RVL_BOOL_AND = 0, RVL_BOOL_OR = 1, RVL_BOOL_NOT = 2 } rvl_bool_op;
int32_t rvl_boolean_op(rvl_ctx* ctx, int32_t op, const rvl_batch* a, int32_t a_col, const rvl_batch* b, int32_t b_col, rvl_batch** out);

/* RecordBatch::slice (record_batch.rs:92-106): zero-copy view; RVL_OUT_OF_BOUNDS instead of the panic */
int32_t rvl_batch_slice(const rvl_batch* batch, int64_t offset, int64_t length, rvl_batch** view);
/* RecordBatch::select_columns (record_batch.rs:180-206): zero-copy column pick */
int32_t rvl_batch_select(const rvl_batch* batch, const int32_t* indices, int32_t n, rvl_batch** view);
/* RecordBatch::take (record_batch.rs:108-129 -> take_array :131-178): gather rows by HOST indices (any order, repeats allowed);
 * RVL_OUT_OF_BOUNDS "Index {} out of bounds for {} rows".  Output freshly built like the reference's builders. */
int32_t rvl_batch_take(rvl_ctx* ctx, const rvl_batch* batch, const int64_t* indices, int64_t n, rvl_batch** out);
/* RecordBatch::concat (record_batch.rs:245-342): freshly built output, offset 0, bitmap iff nulls */
int32_t rvl_batch_concat(rvl_ctx* ctx, const rvl_batch* const* batches, int32_t n, rvl_batch** out);

/* ---- the hot path ----------------------------------------------------------------------- */
/* Fused Filter + Select + Limit:
 *   eager     PhysicalPlan::execute Filter/Select/Limit arms     physical_plan/plan.rs:68-173
 *   streaming FilterStream/SelectStream/LimitStream::next_batch  stream.rs:136-162,202-212; streaming.rs:268-287
 *             = RecordBatch::filter -> take -> take_array         record_batch.rs:221-243,108-178
 * One pass: predicate -> warp ballot/popc -> block scan -> decoupled look-back -> ordered compaction of every
 * projected column (values, validity and Boolean bitmaps; strings in a second kernel pair).
 * limit < 0 = no limit.  Output arrays are freshly built (offset 0, placeholder 0 under nulls, validity
 * present iff at least one surviving row is null), exactly as the reference's builders leave them. */
int32_t rvl_filter_project(rvl_ctx* ctx, const rvl_batch* in, const rvl_predicate* pred, const int32_t* proj,
                           int32_t nproj, int64_t limit, rvl_batch** out);

/* Standalone predicate -> BooleanArray-layout selection mask (the K1 kernel on its own): bit i = row i kept. */
int32_t rvl_predicate_mask(rvl_ctx* ctx, const rvl_batch* in, const rvl_predicate* pred, rvl_batch** mask_out);

/* Same, split in two so a caller can time / graph-capture the device work without the final readback:
 * launch() enqueues everything on the context stream; finish() waits and builds the output batch. */
typedef struct rvl_pending rvl_pending;
int32_t rvl_filter_project_launch(rvl_ctx* ctx, const rvl_batch* in, const rvl_predicate* pred, const int32_t* proj,
                                  int32_t nproj, int64_t limit, rvl_pending** pending);
int32_t rvl_filter_project_finish(rvl_ctx* ctx, rvl_pending* pending, rvl_batch** out);

/* ---- streaming executor: trait DataStream (stream.rs:25-54) ------------------------------ */
/* How the columns of a pushed batch reach the kernels.  STAGED: every needed column is copied to the device by the copy engine.
 * ZERO_COPY: only the predicate column is; fixed-width projected columns in pinned (page-locked) host memory are read in place by
 * the kernels, which fetch just the PCIe lines holding a survivor (dense tiles are streamed whole).  Pageable memory and String
 * columns are always staged.  AUTO = ZERO_COPY while the stream is selective; batches pushed after one that kept more than a
 * quarter of its rows are STAGED (the copy engine moves whole columns ~8 % faster than SM-issued reads).
 * In place means the caller's buffers must stay valid and unchanged until the batch's output has been returned by
 * rvl_stream_next / rvl_stream_collect. */
typedef enum rvl_transfer { RVL_TRANSFER_AUTO = 0, RVL_TRANSFER_STAGED = 1, RVL_TRANSFER_ZERO_COPY = 2 } rvl_transfer;
typedef struct rvl_stream_config {
    int64_t batch_rows; /* capacity of one staging slot, rows.  Pushed batches smaller than this are appended to the open slot and one
                           operator launch covers the whole group (a group needs 64-row-aligned joins, no String column, and — for
                           columns read in place — adjacent caller buffers; otherwise every batch is its own group).  LIMIT streams
                           grow their groups 1, 1, 2, 4, ... batches.  Set it to the batch size for one launch per pushed batch. */
    int32_t n_staging;  /* pinned/device staging slots (>= 2 for H2D / compute overlap) */
    int32_t transfer;   /* rvl_transfer */
} rvl_stream_config;

/* Opens Filter -> Select -> Limit over batches with the given input schema (dtypes of the pushed columns). */
int32_t rvl_stream_open(rvl_ctx* ctx, const int32_t* dtypes, int32_t ncols, const rvl_predicate* pred,
                        const int32_t* proj, int32_t nproj, int64_t limit, const rvl_stream_config* cfg,
                        rvl_stream** stream);
/* Feed one HOST batch (MemoryStream::next_batch upstream, stream.rs:105-113).  Asynchronous: H2D on the copy
 * stream of a free staging slot, fused kernel on the compute stream.  *accepted = 0 when the LIMIT has already
 * been reached and the batch was not transferred (LimitStream's early termination, streaming.rs:269-271).
 * Buffer lifetime: PAGEABLE sources are copied into the slot's pinned staging memory before push() returns and may be reused at
 * once.  PAGE-LOCKED sources (rvl_host_alloc / cudaHostAlloc / cudaHostRegister) are read asynchronously — by the copy engine, or
 * in place by the kernels (ZERO_COPY / AUTO) — and must stay valid and unchanged until the output covering the batch has been
 * returned by rvl_stream_next / rvl_stream_collect, or the stream is closed.
 * Any number of batches may be pushed before the first rvl_stream_next: beyond 32 outstanding launches push() finishes the oldest
 * ones itself and folds their outputs, 16 at a time, into exact-size batches that rvl_stream_next later hands out in order. */
int32_t rvl_stream_push(rvl_stream* stream, const rvl_column* host_cols, int32_t ncols, int32_t* accepted);
/* launch the operator over the batches appended to the open slot so far (next / collect do this implicitly) */
int32_t rvl_stream_flush(rvl_stream* stream);
/* next_batch() -> Option<RecordBatch>: *has_batch = 0 when no output is pending.  One output covers one group of pushed batches. */
int32_t rvl_stream_next(rvl_stream* stream, rvl_batch** out, int32_t* has_batch);
int32_t rvl_stream_limit_reached(rvl_stream* stream, int32_t* reached);
/* DataStream::concatenate / collect_stream_batches (stream.rs:41-53, streaming.rs:343-352) */
int32_t rvl_stream_collect(rvl_stream* stream, rvl_batch** out);
int32_t rvl_stream_stats(rvl_stream* stream, int64_t* batches_pushed, int64_t* batches_skipped, int64_t* h2d_bytes);
/* operator launches issued so far (= groups of pushed batches) */
int32_t rvl_stream_launches(rvl_stream* stream, int64_t* groups);
int32_t rvl_stream_close(rvl_stream* stream);

/* ---- inner equi-join ------------------------------------------------ */
/* PhysicalPlan::HashJoin, JoinType::Inner (physical_plan/plan.rs:174-284): for every probe row in order, every build row with an
 * equal key in ascending build order; *out holds the probe columns `probe_proj` followed by the build columns `build_proj`, each
 * taken by the matching row (materialize_join_result :208-254; bitmaps kept only where a taken row is null).  Key equality is
 * AnyValue's (series.rs:73-98): two nulls are equal, values are equal when they have the same type and the same value, an Int64
 * never equals a Float64, NaN equals nothing.  `*_tag_column`: 0, or k + 1 = Boolean column k marks the Int64 rows of a key column
 * that is a mixed Float64 / Int64 Series (as rvl_predicate::tag_column).  Both batches live on the context's device; each side is
 * limited to 2^32 - 1 rows. */
int32_t rvl_hash_join_inner(rvl_ctx* ctx, const rvl_batch* build, int32_t build_key, int32_t build_tag_column, const rvl_batch* probe,
                            int32_t probe_key, int32_t probe_tag_column, const int32_t* probe_proj, int32_t n_probe_proj,
                            const int32_t* build_proj, int32_t n_build_proj, rvl_batch** out, int64_t* n_pairs);

/* ---- multi-GPU: contiguous row ranges, no collective ---------------------- */
/* rows [begin, end) of shard `rank` of `world` for an n_rows table; boundaries are multiples of 64 rows */
int32_t rvl_shard_range(int64_t n_rows, int32_t rank, int32_t world, int64_t* begin, int64_t* end);
/* per-shard contributions to an ordered result under a global LIMIT: take[g] = clamp(limit - sum_{j<g} counts[j], 0, counts[g]) */
int32_t rvl_shard_limit_split(const int64_t* counts, int32_t world, int64_t limit, int64_t* take);
/* Ordered physical concatenation of per-GPU results (outs[] of rvl_filter_project_sharded, in shard order) on dst's device: the
 * concat kernels read the peers' buffers in place over NVLink (peer mappings are set up on first use); the order-preserving
 * concatenation of record_batch.rs:245-342 across devices.  Blocks until the result is complete. */
int32_t rvl_gather_to(rvl_ctx* dst, const rvl_batch* const* parts, int32_t n, rvl_batch** out);
/* The same ordered concatenation when every GPU is driven by its OWN process (torchrun, one rank per GPU).  The destination rank
 * creates the result (cudaMalloc-backed, zero-filled) and gets one CUDA IPC handle set per column; the handle array is plain bytes —
 * ship it to the other ranks by any means (torch.distributed broadcast in bench.py).  Every rank, the destination included, then
 * writes its part at its row offset (exclusive scan of the per-shard survivor counts; for String columns also its byte offset per
 * column) with rvl_gather_push — peer writes over NVLink, all ranks concurrently, bitmaps at any bit offset.  After a barrier the
 * destination calls rvl_gather_dest_finish (null counts; bitmaps kept only if something is null, primitive.rs:180-185).
 * Non-destination ranks open the handles with rvl_gather_dest_open and release the view with rvl_batch_release. */
#define RVL_IPC_HANDLE_BYTES 64
typedef struct rvl_gather_handle {
    int32_t dtype;
    int32_t has_validity;
    int64_t rows;
    int64_t data_bytes;
    uint8_t values[RVL_IPC_HANDLE_BYTES];
    uint8_t validity[RVL_IPC_HANDLE_BYTES];
    uint8_t offsets[RVL_IPC_HANDLE_BYTES];
    uint8_t data[RVL_IPC_HANDLE_BYTES];
} rvl_gather_handle;
int32_t rvl_gather_dest_create(rvl_ctx* ctx, const int32_t* dtypes, const int32_t* has_validity, const int64_t* data_bytes /* per column,
                               String only, may be NULL */, int32_t ncols, int64_t total_rows, rvl_batch** dest, rvl_gather_handle* handles);
int32_t rvl_gather_dest_open(rvl_ctx* ctx, const rvl_gather_handle* handles, int32_t ncols, rvl_batch** dest_view);
int32_t rvl_gather_push(rvl_ctx* ctx, const rvl_batch* part, rvl_batch* dest, int64_t row_offset, const int64_t* byte_offsets /* per column */);
int32_t rvl_gather_dest_finish(rvl_ctx* ctx, rvl_batch* dest);
/* one fused call per context (each on its own GPU/stream), all in flight together; outs[g] are in row order */
int32_t rvl_filter_project_sharded(rvl_ctx* const* ctxs, int32_t n, const rvl_batch* const* shards,
                                   const rvl_predicate* pred, const int32_t* proj, int32_t nproj, int64_t limit,
                                   rvl_batch** outs, int64_t* counts);

/* ---- synthetic tables and checksums (bench / test support; include/rivulus_synth.h) ------ */
/* generate rows [row0, row0+n) of synthetic column (kind, col_id) on the device */
int32_t rvl_gen_batch(rvl_ctx* ctx, const int32_t* kinds, const uint32_t* col_ids, const uint32_t* null_pct,
                      int32_t ncols, uint64_t row0, int64_t n, rvl_batch** out);
/* order-sensitive checksum of column i (values under validity; nulls hash as a fixed tag): rivulus_synth.h */
int32_t rvl_batch_checksum(rvl_ctx* ctx, const rvl_batch* batch, int32_t i, uint64_t* checksum);

#ifdef __cplusplus
}
#endif
#endif /* RIVULUS_GPU_H */
