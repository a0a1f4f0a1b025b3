// stream_exec.cu — the streaming executor: RecordBatches pushed from HOST memory flow through pinned,
// multi-buffered H2D staging on a copy stream while the fused Filter+Select+Limit kernel of the previous
// group runs on the compute stream; LIMIT stops further transfers as soon as the device-side running
// count (chained from launch to launch through a device word) reaches the limit.
//
// Groups.  A staging slot holds `batch_rows` rows.  Pushed batches smaller than that are APPENDED to the open slot (their H2D
// copies are enqueued immediately, back to back on the copy engine) and one operator launch covers the whole group — for 64 K-row
// batches that is 1 kernel, 1 count readback and 1 set of output buffers per 16 batches instead of per batch, which is what bounded
// small-batch streaming in round 1 (~25 us of API work + a 20 us kernel on 8 SMs per 1.5 MB batch).  LIMIT streams grow their
// groups 1, 1, 2, 4, ... batches, so a limit the first batch satisfies still costs exactly the transfers a LimitStream would pull
// (physical_plan/streaming.rs:269-271).  A group is launched when it is full, when the next batch cannot be appended (ragged row
// count, String columns, non-adjacent in-place buffers), or when the consumer asks for output.
//
// Transfer modes (rvl_stream_config::transfer).  STAGED: every needed column crosses PCIe on the copy engine.
// ZERO_COPY: only the predicate column is staged; the fixed-width projected columns stay in the caller's pinned
// memory and the kernels read them through the mapped host pointer — gather_sparse_kernel fetches just the 64-byte
// PCIe lines that hold a survivor, compact_dense_kernel streams dense tiles with TMA bulk copies at 93 % of the
// copy-engine rate (scripts/microbench_pcie.cu: 55.6 vs 51.5 GB/s; one survivor in 1000 / 100 / 10 rows costs
// 2 % / 17 % / 78 % of a full copy).  Columns that are neither the predicate nor projected never cross the bus.
//
// Replaces (reference, /root/reference/src): trait DataStream + MemoryStream/FilterStream/SelectStream
// (execution/stream.rs:25-213), LimitStream (physical_plan/streaming.rs:246-288) and the final
// collect_stream_batches concat (physical_plan/streaming.rs:343-352).
#include <cuda.h>

#include <algorithm>
#include <cstring>
#include <deque>

#include "filter_project.cuh"

using namespace rvl;

namespace {

struct StagingColumn {
    BufRef values, validity, offsets, data;  // device
    void *h_values = nullptr, *h_validity = nullptr, *h_offsets = nullptr, *h_data = nullptr;  // pinned host (lazy, CtxCore::take_pinned)
    size_t h_values_cap = 0, h_validity_cap = 0, h_offsets_cap = 0;
    size_t data_cap = 0, h_data_cap = 0;
};

struct Slot {
    std::vector<StagingColumn> cols;
    cudaEvent_t copied = nullptr;  // H2D of this slot finished
    cudaEvent_t free_ev = nullptr; // last kernel reading this slot finished
    bool used = false;
};

struct InFlight {
    FpPending* pend;
    int slot;
    int64_t rows;  // input rows of the group
};

// how one input column of the open group reaches the kernels
struct GroupCol {
    int mode = 0;                          // 0 = not needed (never transferred), 1 = staged into the slot, 2 = read in place
    bool has_validity = false;
    int64_t resid = 0;                     // sub-64-row residual of the first batch's view offset (bit offsets survive the copy)
    const uint8_t* ip_values = nullptr;    // in place: device view of the group's first row, rounded down to a 64-row boundary
    const uint8_t* ip_validity = nullptr;
    int64_t str_first = 0, str_last = 0;   // staged String: byte window of the batch
};

// What the driver says about a caller pointer, remembered per allocation: a streamed table is normally a handful of large pinned
// buffers sliced into many batches, so the (microsecond) attribute query is paid once per buffer instead of once per pushed column.
struct PtrRange {
    uintptr_t lo = 0, hi = 0;
    bool reachable = false;   // pinned host / device / managed: cudaMemcpyAsync from it is truly asynchronous
    bool has_dev = false;     // kernels can read it through host address + dev_delta
    intptr_t dev_delta = 0;
};

typedef CUresult (*PfnPointerGetAttribute)(void*, CUpointer_attribute, CUdeviceptr);

PfnPointerGetAttribute pointer_attr_fn() {
    // resolved through the runtime, so the library does not link libcuda
    static PfnPointerGetAttribute fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            f = nullptr;
        }
        return (PfnPointerGetAttribute)f;
    }();
    return fn;
}

}  // namespace

struct rvl_stream {
    CoreRef core;
    std::vector<int32_t> dtypes;
    rvl_predicate pred{};
    std::string lit_str;
    std::vector<int32_t> proj;
    int64_t limit = -1;
    int64_t batch_rows = 0;
    std::vector<Slot> slots;
    int next_slot = 0;
    std::deque<InFlight> inflight;
    std::deque<rvl_batch*> parts;      // outputs finished ahead of the consumer (a producer that pushes without draining)
    BufRef cursor;                     // two device words, ping-pong: rows emitted so far
    int64_t zero_copy_cols = 0;        // columns handed to the kernels in place so far
    bool adaptive = false;             // AUTO: groups go back to the copy engine while the stream is dense (recent_sel)
    double recent_sel = 0.0;           // survivors / rows of the most recent group whose count has arrived
    bool zero_copy = false;            // fixed-width projected columns are read in place from pinned host memory
    std::vector<uint8_t> needed;       // per input column: predicate (1) / projected (2) — anything else is not transferred
    bool any_string = false;           // a needed column is a String: batches are never coalesced (offsets would need rebasing)
    int64_t pushed = 0, skipped = 0, h2d_bytes = 0, groups = 0;
    bool limit_hit = false;
    int64_t rows_out = 0;              // rows of the finished outputs
    int64_t rows_seen = 0;             // newest running survivor total that has landed in a mailbox (LIMIT group sizing)
    // the open group: pushed batches appended to slot `next_slot`, operator not launched yet
    bool open = false;
    int64_t open_rows = 0, open_target = 0;
    std::vector<GroupCol> gcols;
    std::vector<PtrRange> ptr_cache;
};

namespace {

constexpr size_t kMaxInflight = 32;   // launches outstanding before push() starts finishing the oldest ones itself
constexpr size_t kMergeParts = 16;    // finished-ahead outputs are concatenated (at their exact size) this many at a time

PtrRange classify(rvl_stream* s, const void* p) {
    PtrRange r;
    if (p == nullptr) { r.reachable = true; return r; }
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    for (const PtrRange& c : s->ptr_cache)
        if (a >= c.lo && a < c.hi) return c;
    r.lo = a; r.hi = a + 1;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return r; }
    if (at.type != cudaMemoryTypeHost && at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) return r;  // pageable
    r.reachable = true;
    if (at.devicePointer != nullptr) { r.has_dev = true; r.dev_delta = (intptr_t)reinterpret_cast<uintptr_t>(at.devicePointer) - (intptr_t)a; }
    if (PfnPointerGetAttribute fn = pointer_attr_fn()) {
        CUdeviceptr start = 0; size_t size = 0;
        if (fn(&start, CU_POINTER_ATTRIBUTE_RANGE_START_ADDR, (CUdeviceptr)a) == CUDA_SUCCESS &&
            fn(&size, CU_POINTER_ATTRIBUTE_RANGE_SIZE, (CUdeviceptr)a) == CUDA_SUCCESS && size > 0 && (uintptr_t)start <= a && a < (uintptr_t)start + size) {
            r.lo = (uintptr_t)start; r.hi = (uintptr_t)start + size;
            if (s->ptr_cache.size() < 256) s->ptr_cache.push_back(r);
        }
    }
    return r;
}

// the slot's page-locked staging buffer for one column buffer, at least `need` bytes: from the context's cache of pinned blocks
int ensure_pinned(rvl_stream* s, void** p, size_t* cap, size_t need) {
    if (*p != nullptr && *cap >= need) return RVL_OK;
    if (*p != nullptr) {
        // variable-size staging (string bytes) grows by half again; copies still reading the old block must finish before it is recycled
        need = std::max(need, *cap + *cap / 2);
        cudaStreamSynchronize(s->core->copy_stream);
        s->core->give_pinned(*p);
        *p = nullptr; *cap = 0;
    }
    size_t got = 0;
    *p = s->core->take_pinned(need, &got);
    if (*p == nullptr) return fail(RVL_OUT_OF_MEMORY, "cudaHostAlloc(" + std::to_string(need) + ") failed");
    *cap = got;
    return RVL_OK;
}

// copy `bytes` from a caller buffer to the device slot on the copy stream.  Pageable sources are staged through the slot's pinned
// buffer first (at byte `pinned_off`), so the caller may reuse them as soon as push() returns; page-locked sources are read by the
// copy engine directly and must stay valid until the batch's output has been returned (include/rivulus_gpu.h).
int stage_copy(rvl_stream* s, void* dev_dst, const void* src, size_t bytes, void** pinned, size_t* pinned_cap, size_t pinned_need, size_t pinned_off) {
    if (bytes == 0) return RVL_OK;
    const void* from = src;
    if (!classify(s, src).reachable) {
        RVL_TRY(ensure_pinned(s, pinned, pinned_cap, std::max(pinned_need, pinned_off + bytes)));
        std::memcpy((uint8_t*)*pinned + pinned_off, src, bytes);
        from = (uint8_t*)*pinned + pinned_off;
    }
    RVL_CUDA_TRY(cudaMemcpyAsync(dev_dst, from, bytes, cudaMemcpyDefault, s->core->copy_stream));
    s->h2d_bytes += (int64_t)bytes;
    return RVL_OK;
}

void poll_limit(rvl_stream* s) {
    if (s->limit < 0 || s->limit_hit) return;
    for (const InFlight& f : s->inflight) {
        const uint64_t total = *reinterpret_cast<volatile uint64_t*>(f.pend->mailbox);
        if (total != kMailboxPending && (int64_t)total >= s->limit) { s->limit_hit = true; return; }
    }
}

// selectivity of the newest group whose counters have already landed in its mailbox (non-blocking)
void poll_selectivity(rvl_stream* s) {
    for (auto it = s->inflight.rbegin(); it != s->inflight.rend(); ++it) {
        const volatile uint64_t* mb = reinterpret_cast<volatile uint64_t*>(it->pend->mailbox);
        const uint64_t total = mb[0];
        if (total == kMailboxPending || it->rows <= 0) continue;
        const uint64_t base = it->pend->chained ? mb[it->pend->n_counters] : 0ull;
        if (base == kMailboxPending || base > total) continue;
        s->recent_sel = (double)(total - base) / (double)it->rows;
        s->rows_seen = std::max<int64_t>(s->rows_seen, (int64_t)total);
        return;
    }
}

// waits for the oldest launch and builds its output batch
int finish_oldest(rvl_stream* s, rvl_batch** out) {
    InFlight f = s->inflight.front();
    s->inflight.pop_front();
    rvl_batch* b = nullptr;
    RVL_TRY(fp_finish(f.pend, &b, nullptr));
    if (f.rows > 0) s->recent_sel = (double)b->num_rows / (double)f.rows;
    s->rows_out += b->num_rows;
    s->rows_seen = std::max(s->rows_seen, s->rows_out);
    if (s->limit >= 0 && s->rows_out >= s->limit) s->limit_hit = true;
    *out = b;
    return RVL_OK;
}

// A producer that never drains: finish the oldest launches ourselves and fold their (worst-case sized) outputs into exact-size
// batches, so neither the in-flight bookkeeping nor device memory grows with the number of pushed batches.
int retire_ahead(rvl_stream* s) {
    while (s->inflight.size() >= kMaxInflight) {
        rvl_batch* b = nullptr;
        RVL_TRY(finish_oldest(s, &b));
        s->parts.push_back(b);
    }
    if (s->parts.size() >= kMergeParts) {
        std::vector<const rvl_batch*> v(s->parts.begin(), s->parts.end());
        rvl_ctx tmp{s->core};
        rvl_batch* merged = nullptr;
        RVL_TRY(rvl_batch_concat(&tmp, v.data(), (int32_t)v.size(), &merged));
        for (rvl_batch* b : s->parts) delete b;
        s->parts.clear();
        s->parts.push_back(merged);
    }
    return RVL_OK;
}

// launch the operator over the open group
int flush_group(rvl_stream* s) {
    if (!s->open) return RVL_OK;
    const CoreRef& core = s->core;
    Slot& sl = s->slots[(size_t)s->next_slot];
    const int64_t n = s->open_rows;
    const int ncols = (int)s->dtypes.size();
    auto view = std::make_unique<rvl_batch>();
    view->core = core; view->num_rows = n;
    for (int c = 0; c < ncols; ++c) {
        const GroupCol& g = s->gcols[(size_t)c];
        StagingColumn& sc = sl.cols[(size_t)c];
        const int64_t span = g.resid + n;
        DevColumn d;
        d.dtype = s->dtypes[(size_t)c]; d.length = n; d.offset = g.resid;
        if (g.mode == 0) {
            // neither the predicate nor projected: the operator never looks at it, nothing crossed the bus
            d.dtype = RVL_NULL; d.null_count = n; d.offset = 0;
        } else if (g.mode == 2) {
            if (d.dtype == RVL_BOOLEAN) d.values = wrap_external(core, g.ip_values, (size_t)(span + 7) / 8);
            else d.values = wrap_external(core, g.ip_values, (size_t)span * 8);
            if (g.has_validity) d.validity = wrap_external(core, g.ip_validity, (size_t)(span + 7) / 8);
            else d.null_count = 0;
        } else {
            if (d.dtype == RVL_INT64 || d.dtype == RVL_FLOAT64 || d.dtype == RVL_BOOLEAN) d.values = sc.values;
            else if (d.dtype == RVL_STRING) {
                d.offsets = sc.offsets;
                // offsets stay absolute: present the data pointer shifted back by `first` (never dereferenced below the copy)
                const size_t nbytes = (size_t)(g.str_last - g.str_first);
                d.data = wrap_external(core, (const uint8_t*)sc.data->ptr - g.str_first, nbytes + (size_t)g.str_first);
                d.data_len = g.str_last;
                d.window_bytes = (int64_t)nbytes;   // sizes the output: the survivors' bytes are a subset of this window
            }
            if (g.has_validity) d.validity = sc.validity;
            else d.null_count = d.dtype == RVL_NULL ? n : 0;
        }
        view->cols.push_back(std::move(d));
    }
    RVL_CUDA_TRY(cudaEventRecord(sl.copied, core->copy_stream));
    RVL_CUDA_TRY(cudaStreamWaitEvent(core->stream, sl.copied, 0));

    unsigned long long* cur = (unsigned long long*)s->cursor->ptr;
    const int k = (int)(s->groups & 1);
    FpPending* pend = nullptr;
    core->sel_hint = s->groups > 0 ? s->recent_sel : -1.0;   // plan choice of the operator: what the previous group kept
    const int rc_launch = fp_launch(core, view.get(), &s->pred, s->proj.data(), (int32_t)s->proj.size(), s->limit, false, cur + k, cur + (k ^ 1), &pend);
    core->sel_hint = -1.0;
    RVL_TRY(rc_launch);
    RVL_CUDA_TRY(cudaEventRecord(sl.free_ev, core->stream));
    // the next H2D into this slot must not start before this kernel has read it
    sl.used = true;
    s->inflight.push_back(InFlight{pend, s->next_slot, n});
    s->next_slot = (s->next_slot + 1) % (int)s->slots.size();
    // make the copy stream wait for the slot it is going to overwrite next
    Slot& nx = s->slots[(size_t)s->next_slot];
    if (nx.used) RVL_CUDA_TRY(cudaStreamWaitEvent(core->copy_stream, nx.free_ev, 0));
    s->groups++;
    s->open = false; s->open_rows = 0;
    return RVL_OK;
}

}  // namespace

extern "C" {

int32_t rvl_stream_open(rvl_ctx* ctx, const int32_t* dtypes, int32_t ncols, const rvl_predicate* pred, const int32_t* proj, int32_t nproj,
                        int64_t limit, const rvl_stream_config* cfg, rvl_stream** stream) {
    if (!ctx || !stream || (ncols > 0 && !dtypes) || (nproj > 0 && !proj)) return fail(RVL_INVALID_ARGUMENT, "null argument");
    const CoreRef& core = ctx->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    auto s = std::make_unique<rvl_stream>();
    s->core = core;
    s->dtypes.assign(dtypes, dtypes + ncols);
    for (int i = 0; i < nproj; ++i) {
        if (proj[i] < 0 || proj[i] >= ncols)
            return fail(RVL_OUT_OF_BOUNDS, "Column index " + std::to_string(proj[i]) + " out of bounds for " + std::to_string(ncols) + " columns");
        s->proj.push_back(proj[i]);
    }
    if (pred) {
        s->pred = *pred;
        if (pred->mode != RVL_PRED_TRUE) {
            if (pred->column < 0 || pred->column >= ncols) return fail(RVL_COLUMN_NOT_FOUND, "Column not found: index " + std::to_string(pred->column));
            if (pred->mode == RVL_PRED_BOOL_COLUMN && dtypes[pred->column] != RVL_BOOLEAN)
                return fail(RVL_TYPE_MISMATCH, "Predicate column is not of boolean type");  // stream.rs:147-153
        }
        if (pred->lit_str && pred->lit_str_len > 0) { s->lit_str.assign((const char*)pred->lit_str, (size_t)pred->lit_str_len); s->pred.lit_str = (const uint8_t*)s->lit_str.data(); }
    } else {
        s->pred.mode = RVL_PRED_TRUE;
    }
    s->limit = limit < 0 ? -1 : limit;
    s->batch_rows = cfg && cfg->batch_rows > 0 ? cfg->batch_rows : (1 << 20);
    s->needed.assign((size_t)ncols, 0);
    for (int32_t pj : s->proj) s->needed[(size_t)pj] |= 2;
    if (s->pred.mode != RVL_PRED_TRUE) {
        s->needed[(size_t)s->pred.column] |= 1;
        if (s->pred.mode == RVL_PRED_CMP_LITERAL && s->pred.tag_column > 0 && s->pred.tag_column <= ncols) s->needed[(size_t)s->pred.tag_column - 1] |= 1;
    }
    for (int c = 0; c < ncols; ++c) s->any_string |= s->needed[(size_t)c] != 0 && dtypes[c] == RVL_STRING;
    const int transfer = cfg ? cfg->transfer : RVL_TRANSFER_AUTO;
    if (transfer != RVL_TRANSFER_AUTO && transfer != RVL_TRANSFER_STAGED && transfer != RVL_TRANSFER_ZERO_COPY)
        return fail(RVL_INVALID_ARGUMENT, "unknown transfer mode");
    s->zero_copy = transfer != RVL_TRANSFER_STAGED;
    s->adaptive = transfer == RVL_TRANSFER_AUTO;
    const int n_slots = cfg && cfg->n_staging >= 1 ? cfg->n_staging : 2;
    s->slots.resize((size_t)n_slots);
    s->gcols.resize((size_t)ncols);
    const size_t rows_cap = (size_t)s->batch_rows + 64;
    for (Slot& sl : s->slots) {
        sl.cols.resize((size_t)ncols);
        for (int c = 0; c < ncols; ++c) {
            StagingColumn& sc = sl.cols[(size_t)c];
            if (s->needed[(size_t)c] == 0) continue;  // never transferred
            if (dtypes[c] == RVL_INT64 || dtypes[c] == RVL_FLOAT64) RVL_TRY(dev_alloc(core, rows_cap * 8, &sc.values));
            else if (dtypes[c] == RVL_BOOLEAN) RVL_TRY(dev_alloc_zeroed(core, rows_cap / 8 + 8, &sc.values));
            else if (dtypes[c] == RVL_STRING) RVL_TRY(dev_alloc(core, (rows_cap + 1) * 4, &sc.offsets));
            if (dtypes[c] != RVL_NULL) RVL_TRY(dev_alloc_zeroed(core, rows_cap / 8 + 8, &sc.validity));
        }
        RVL_CUDA_TRY(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
        RVL_CUDA_TRY(cudaEventCreateWithFlags(&sl.free_ev, cudaEventDisableTiming));
    }
    RVL_TRY(dev_alloc_zeroed(core, 16, &s->cursor));
    RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
    *stream = s.release();
    return RVL_OK;
}

int32_t rvl_stream_push(rvl_stream* s, const rvl_column* cols, int32_t ncols, int32_t* accepted) {
    if (!s || !accepted || (ncols > 0 && !cols)) return fail(RVL_INVALID_ARGUMENT, "null argument");
    const CoreRef& core = s->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    *accepted = 0;
    if (ncols != (int32_t)s->dtypes.size()) return fail(RVL_SCHEMA_MISMATCH, "batch has " + std::to_string(ncols) + " columns but the stream schema has " + std::to_string(s->dtypes.size()));
    const int64_t n = ncols > 0 ? cols[0].length : 0;
    for (int c = 0; c < ncols; ++c) {
        if (cols[c].dtype != s->dtypes[(size_t)c]) return fail(RVL_SCHEMA_MISMATCH, "Column " + std::to_string(c) + " does not match the stream schema");
        if (cols[c].length != n) return fail(RVL_LENGTH_MISMATCH, "Column " + std::to_string(c) + " has length " + std::to_string(cols[c].length) + " but expected " + std::to_string(n));
        if (cols[c].location != RVL_HOST) return fail(RVL_INVALID_ARGUMENT, "rvl_stream_push takes host buffers");
    }
    if (n > s->batch_rows) return fail(RVL_INVALID_ARGUMENT, "batch of " + std::to_string(n) + " rows exceeds the stream's batch_rows=" + std::to_string(s->batch_rows));

    // LimitStream: once the limit is reached nothing more is pulled (streaming.rs:269-271)
    poll_limit(s);
    if (s->limit_hit || s->limit == 0) { s->limit_hit = true; s->skipped++; return RVL_OK; }

    // ---- can this batch join the open group?  Its rows land behind the group's rows in the same slot, so every needed column must
    // continue on a 64-row boundary (bitmap words), presence of validity must match the group's, and columns read in place must be
    // adjacent in the caller's memory.
    bool append = s->open && n > 0 && s->open_rows > 0 && !s->any_string && s->open_rows + n <= s->batch_rows && s->open_rows < s->open_target;
    for (int c = 0; c < ncols && append; ++c) {
        const GroupCol& g = s->gcols[(size_t)c];
        if (g.mode == 0) continue;
        const rvl_column& hc = cols[c];
        const int64_t pos = g.resid + s->open_rows;
        if (pos % 64 != 0 || hc.offset % 64 != 0 || (hc.validity != nullptr && hc.dtype != RVL_NULL) != g.has_validity) { append = false; break; }
        if (g.mode == 2) {
            const PtrRange r = classify(s, hc.values);
            const uint8_t* have = (const uint8_t*)hc.values + r.dev_delta + (hc.dtype == RVL_BOOLEAN ? hc.offset / 8 : hc.offset * 8);
            const uint8_t* want = g.ip_values + (hc.dtype == RVL_BOOLEAN ? pos / 8 : pos * 8);
            if (!r.reachable || !r.has_dev || want != have) append = false;
            if (append && g.has_validity) {
                const PtrRange rv = classify(s, hc.validity);
                if (!rv.reachable || !rv.has_dev || g.ip_validity + pos / 8 != hc.validity + rv.dev_delta + hc.offset / 8) append = false;
            }
        }
    }
    if (s->open && !append) RVL_TRY(flush_group(s));

    if (!s->open) {
        // ---- start a group in the next slot
        RVL_TRY(retire_ahead(s));
        Slot& sl0 = s->slots[(size_t)s->next_slot];
        if (sl0.used) {
            // back-pressure: this slot's previous group must have been consumed by its kernel; its mailbox is then
            // visible too, so a tripped limit is noticed before any further byte crosses PCIe
            RVL_CUDA_TRY(cudaEventSynchronize(sl0.free_ev));
            poll_limit(s);
        }
        if (!s->limit_hit && s->limit > 0 && s->groups >= 1 && s->recent_sel > 0.0) {
            // the groups still in flight are expected to satisfy the limit on their own: wait for their counts instead of
            // transferring another group that a LimitStream would never have pulled
            int64_t pending_rows = 0;
            for (const InFlight& f : s->inflight)
                if (*reinterpret_cast<volatile uint64_t*>(f.pend->mailbox) == kMailboxPending) pending_rows += f.rows;
            if (pending_rows > 0 && (double)s->rows_seen + 0.8 * s->recent_sel * (double)pending_rows >= (double)s->limit) {
                const int64_t n_slots = (int64_t)s->slots.size();
                Slot& w = s->slots[(size_t)((((int64_t)s->next_slot - 1) % n_slots + n_slots) % n_slots)];
                if (w.used) RVL_CUDA_TRY(cudaEventSynchronize(w.free_ev));
                poll_limit(s);
            }
        }
        if (!s->limit_hit && s->limit > 0 && s->groups >= 1) {
            // LIMIT streams start slowly: group 1 waits for group 0's count, groups 2 and 3 for the group two before them; only then
            // does the pipeline run at its full depth.  A limit that the first batches already satisfy (LIMIT 1000 over 64 K..1 M-row
            // batches) then costs exactly the batches a LimitStream would pull (streaming.rs:269-271), not pipeline-depth more.
            const int64_t i = s->groups, n_slots = (int64_t)s->slots.size();
            const int64_t window = i <= 1 ? 1 : (i <= 3 ? 2 : n_slots);
            if (window < n_slots) {
                Slot& w = s->slots[(size_t)((((int64_t)s->next_slot - window) % n_slots + n_slots) % n_slots)];
                if (w.used) RVL_CUDA_TRY(cudaEventSynchronize(w.free_ev));
                poll_limit(s);
            }
        }
        if (s->limit_hit) { s->skipped++; return RVL_OK; }
        // AUTO: a dense stream (more than one survivor in four rows: every PCIe line is needed anyway) goes through the copy engine,
        // which moves whole columns ~8 % faster than SM-issued reads; a selective one is read in place
        if (s->adaptive) poll_selectivity(s);
        const bool in_place = s->zero_copy && !(s->adaptive && s->recent_sel > 0.25);
        s->open = true; s->open_rows = 0;
        // group size: the whole slot, except that LIMIT streams grow 1, 1, 2, 4, ... batches
        s->open_target = s->batch_rows;
        if (s->limit > 0) {
            const int64_t g = s->groups;
            const int64_t nb = g <= 1 ? 1 : (g >= 40 ? ((int64_t)1 << 39) : ((int64_t)1 << (g - 1)));
            int64_t target = std::max<int64_t>(n, 1) * nb;
            // once a count has arrived the rows still needed are known to within the selectivity's noise: ask for 1.25x that,
            // so a 0.1 % stream over 64 K-row batches pulls ~17 batches for LIMIT 1000 instead of doubling past it
            poll_selectivity(s);
            if (s->recent_sel > 0.0) {
                const double need = (double)std::max<int64_t>(s->limit - s->rows_seen, 1) / s->recent_sel * 1.25;
                if (need < (double)target) target = std::max<int64_t>((int64_t)need, n);
                else if (need < 4.0e18) target = std::max<int64_t>(target, std::min<int64_t>((int64_t)need, s->batch_rows));
            }
            s->open_target = std::min<int64_t>(s->batch_rows, target);
        }
        for (int c = 0; c < ncols; ++c) {
            GroupCol& g = s->gcols[(size_t)c];
            g = GroupCol{};
            const rvl_column& hc = cols[c];
            if (s->needed[(size_t)c] == 0) continue;
            g.mode = 1;
            g.resid = hc.offset % 64;
            g.has_validity = hc.validity != nullptr && hc.dtype != RVL_NULL;
            if (in_place && (s->needed[(size_t)c] & 1) == 0 && (hc.dtype == RVL_INT64 || hc.dtype == RVL_FLOAT64 || hc.dtype == RVL_BOOLEAN) && hc.values != nullptr) {
                // projected fixed-width column: the kernels read the caller's pinned memory in place
                const PtrRange r = classify(s, hc.values);
                const PtrRange rv = classify(s, hc.validity);
                if (r.reachable && r.has_dev && (!g.has_validity || (rv.reachable && rv.has_dev))) {
                    const int64_t start = hc.offset - g.resid;
                    g.mode = 2;
                    g.ip_values = (const uint8_t*)hc.values + r.dev_delta + (hc.dtype == RVL_BOOLEAN ? start / 8 : start * 8);
                    if (g.has_validity) g.ip_validity = hc.validity + rv.dev_delta + start / 8;
                    s->zero_copy_cols++;
                }
            }
        }
    }

    // ---- append this batch's rows to the open group: the H2D copies are enqueued now, back to back on the copy engine
    Slot& sl = s->slots[(size_t)s->next_slot];
    const size_t rows_cap = (size_t)s->batch_rows + 64;
    const bool first = s->open_rows == 0;
    for (int c = 0; c < ncols; ++c) {
        GroupCol& g = s->gcols[(size_t)c];
        if (g.mode != 1) continue;
        const rvl_column& hc = cols[c];
        StagingColumn& sc = sl.cols[(size_t)c];
        // the first batch of a group keeps its sub-64-row residual; appended batches start on a 64-row boundary of the slot
        const int64_t resid = first ? g.resid : 0;
        const int64_t start = hc.offset - resid, span = resid + n;
        const int64_t at = first ? 0 : g.resid + s->open_rows;   // slot row the copy starts at
        if (hc.dtype == RVL_INT64 || hc.dtype == RVL_FLOAT64) {
            RVL_TRY(stage_copy(s, (uint8_t*)sc.values->ptr + at * 8, (const uint8_t*)hc.values + start * 8, (size_t)span * 8, &sc.h_values, &sc.h_values_cap, rows_cap * 8, (size_t)at * 8));
        } else if (hc.dtype == RVL_BOOLEAN) {
            RVL_TRY(stage_copy(s, (uint8_t*)sc.values->ptr + at / 8, (const uint8_t*)hc.values + start / 8, (size_t)(span + 7) / 8, &sc.h_values, &sc.h_values_cap, rows_cap / 8 + 8, (size_t)at / 8));
        } else if (hc.dtype == RVL_STRING) {
            RVL_TRY(stage_copy(s, sc.offsets->ptr, hc.offsets + start, (size_t)(span + 1) * 4, &sc.h_offsets, &sc.h_offsets_cap, (rows_cap + 1) * 4, 0));
            const int64_t b0 = hc.offsets[start], b1 = hc.offsets[start + span];
            const size_t nbytes = (size_t)(b1 - b0);
            if (sc.data_cap < nbytes || !sc.data) {
                // make sure no kernel still reads the old buffer: frees are ordered on the compute stream
                sc.data.reset();
                sc.data_cap = std::max(nbytes * 5 / 4 + 256, (size_t)1 << 20);
                RVL_TRY(dev_alloc(core, sc.data_cap, &sc.data));
                RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
            }
            RVL_TRY(stage_copy(s, sc.data->ptr, hc.data + b0, nbytes, &sc.h_data, &sc.h_data_cap, nbytes, 0));
            g.str_first = b0; g.str_last = b1;
        }
        if (g.has_validity)
            RVL_TRY(stage_copy(s, (uint8_t*)sc.validity->ptr + at / 8, hc.validity + start / 8, (size_t)(span + 7) / 8, &sc.h_validity, &sc.h_validity_cap, rows_cap / 8 + 8, (size_t)at / 8));
    }
    s->open_rows += n;
    s->pushed++;
    *accepted = 1;
    // launch as soon as the group is complete: full, or a shape nothing can be appended to
    if (n == 0 || s->any_string || s->open_rows >= s->open_target || s->open_rows + n > s->batch_rows) RVL_TRY(flush_group(s));
    return RVL_OK;
}

int32_t rvl_stream_flush(rvl_stream* s) {
    if (!s) return fail(RVL_INVALID_ARGUMENT, "null argument");
    RVL_CUDA_TRY(cudaSetDevice(s->core->device));
    return flush_group(s);
}

int32_t rvl_stream_next(rvl_stream* s, rvl_batch** out, int32_t* has_batch) {
    if (!s || !out || !has_batch) return fail(RVL_INVALID_ARGUMENT, "null argument");
    *has_batch = 0; *out = nullptr;
    RVL_CUDA_TRY(cudaSetDevice(s->core->device));
    if (!s->parts.empty()) {
        *out = s->parts.front(); s->parts.pop_front(); *has_batch = 1;
        return RVL_OK;
    }
    if (s->inflight.empty()) RVL_TRY(flush_group(s));   // the consumer is waiting on the open group
    if (s->inflight.empty()) return RVL_OK;
    rvl_batch* b = nullptr;
    RVL_TRY(finish_oldest(s, &b));
    *out = b; *has_batch = 1;
    return RVL_OK;
}

int32_t rvl_stream_limit_reached(rvl_stream* s, int32_t* reached) {
    if (!s || !reached) return fail(RVL_INVALID_ARGUMENT, "null argument");
    poll_limit(s);
    *reached = (s->limit_hit || s->limit == 0) ? 1 : 0;
    return RVL_OK;
}

int32_t rvl_stream_collect(rvl_stream* s, rvl_batch** out) {
    if (!s || !out) return fail(RVL_INVALID_ARGUMENT, "null argument");
    RVL_CUDA_TRY(cudaSetDevice(s->core->device));
    RVL_TRY(flush_group(s));
    std::vector<rvl_batch*> parts;
    int rc = RVL_OK;
    while (!s->parts.empty() || !s->inflight.empty()) {
        rvl_batch* b = nullptr; int32_t has = 0;
        rc = rvl_stream_next(s, &b, &has);
        if (rc != RVL_OK) break;
        // LimitStream stops yielding once the limit is met: later (empty) batches are never produced (streaming.rs:269-271)
        if (has) parts.push_back(b);
    }
    if (rc == RVL_OK) {
        // drop trailing batches produced after the limit was met (they are empty by construction)
        if (s->limit >= 0) {
            int64_t seen = 0; size_t keep = 0;
            for (; keep < parts.size(); ++keep) { if (seen >= s->limit) break; seen += parts[keep]->num_rows; }
            for (size_t i = keep; i < parts.size(); ++i) delete parts[i];
            parts.resize(keep);
        }
        rvl_ctx tmp{s->core};
        if (parts.empty()) {
            // RecordBatch::empty(schema) (streaming.rs:347-349, record_batch.rs:402-421)
            auto e = std::make_unique<rvl_batch>();
            e->core = s->core; e->num_rows = 0;
            for (int32_t p : s->proj) {
                DevColumn d;
                d.dtype = s->dtypes[(size_t)p]; d.null_count = 0;
                if (d.dtype == RVL_INT64 || d.dtype == RVL_FLOAT64 || d.dtype == RVL_BOOLEAN) rc = dev_alloc_zeroed(s->core, 8, &d.values);
                if (d.dtype == RVL_STRING) { rc = dev_alloc_zeroed(s->core, 8, &d.offsets); if (rc == RVL_OK) rc = dev_alloc_zeroed(s->core, 8, &d.data); }
                if (rc != RVL_OK) break;
                e->cols.push_back(d);
            }
            if (rc == RVL_OK) *out = e.release();
        } else {
            rc = rvl_batch_concat(&tmp, parts.data(), (int32_t)parts.size(), out);
        }
    }
    for (rvl_batch* b : parts) delete b;
    return rc;
}

int32_t rvl_stream_stats(rvl_stream* s, int64_t* pushed, int64_t* skipped, int64_t* h2d_bytes) {
    if (!s) return fail(RVL_INVALID_ARGUMENT, "null argument");
    if (pushed) *pushed = s->pushed;
    if (skipped) *skipped = s->skipped;
    if (h2d_bytes) *h2d_bytes = s->h2d_bytes;
    return RVL_OK;
}

int32_t rvl_stream_launches(rvl_stream* s, int64_t* groups) {
    if (!s || !groups) return fail(RVL_INVALID_ARGUMENT, "null argument");
    *groups = s->groups;
    return RVL_OK;
}

int32_t rvl_stream_close(rvl_stream* s) {
    if (!s) return RVL_OK;
    cudaSetDevice(s->core->device);
    s->open = false;   // rows copied into the open slot are simply dropped
    while (!s->inflight.empty()) {
        rvl_batch* b = nullptr;
        fp_finish(s->inflight.front().pend, &b, nullptr);
        delete b;
        s->inflight.pop_front();
    }
    for (rvl_batch* b : s->parts) delete b;
    s->parts.clear();
    cudaStreamSynchronize(s->core->copy_stream);
    cudaStreamSynchronize(s->core->stream);
    for (Slot& sl : s->slots) {
        if (sl.copied) cudaEventDestroy(sl.copied);
        if (sl.free_ev) cudaEventDestroy(sl.free_ev);
        for (StagingColumn& sc : sl.cols) {
            s->core->give_pinned(sc.h_values);
            s->core->give_pinned(sc.h_validity);
            s->core->give_pinned(sc.h_offsets);
            s->core->give_pinned(sc.h_data);
        }
    }
    delete s;
    return RVL_OK;
}

}  // extern "C"
