// runtime.cu — context, memory, device-resident RecordBatch (upload / views / download / slice /
// select / concat), synthetic generator and checksums behind include/rivulus_gpu.h.
//
// Reference seams (under /root/reference/src): execution/record_batch.rs:16-58 (try_new), :92-106 (slice),
// :180-206 (select_columns), :245-342 (concat); execution/array/*.rs for the buffer layouts.
#include "runtime.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>

#include "aux_kernels.cuh"
#include "filter_project.cuh"
#include "string_kernels.cuh"

namespace rvl {

thread_local std::string g_last_error;
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

CtxCore::~CtxCore() {
    cudaSetDevice(device);
    if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
    if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
    if (d2h_stream) { cudaStreamSynchronize(d2h_stream); cudaStreamDestroy(d2h_stream); }
    if (side_stream) { cudaStreamSynchronize(side_stream); cudaStreamDestroy(side_stream); }
    for (auto& e : prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (cudaEvent_t e : event_pool) cudaEventDestroy(e);
    if (mailbox) cudaFreeHost(mailbox);
    for (uint64_t* c : slot_chunks) cudaFreeHost(c);
    for (auto& kv : big_free) cudaFree(kv.second);   // streams are gone: synchronous free
    for (auto& kv : pinned_free) cudaFreeHost(kv.second);
}

void* CtxCore::take_pinned(size_t bytes, size_t* got) {
    bytes = bytes < ((size_t)1 << 20) ? ((bytes + 65535) & ~(size_t)65535) : ((bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1));
    if (bytes == 0) bytes = 65536;
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = pinned_free.lower_bound(bytes);
        if (it != pinned_free.end() && it->first <= 2 * bytes + ((size_t)1 << 20)) {
            void* p = it->second;
            *got = it->first;
            pinned_free.erase(it);
            return p;
        }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        drop_pinned();
        if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    std::lock_guard<std::mutex> g(mu);
    pinned_sizes[p] = bytes;
    *got = bytes;
    return p;
}
void CtxCore::give_pinned(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> g(mu);
    auto it = pinned_sizes.find(p);
    if (it == pinned_sizes.end()) { cudaFreeHost(p); return; }
    pinned_free.emplace(it->second, p);
}
void CtxCore::drop_pinned() {
    std::multimap<size_t, void*> blocks;
    { std::lock_guard<std::mutex> g(mu); blocks.swap(pinned_free); for (auto& kv : blocks) pinned_sizes.erase(kv.second); }
    for (auto& kv : blocks) cudaFreeHost(kv.second);
}
void* CtxCore::take_big(size_t bytes, size_t* got) {
    std::lock_guard<std::mutex> g(mu);
    auto it = big_free.lower_bound(bytes);
    if (it == big_free.end() || it->first > bytes + bytes / 8) return nullptr;
    void* p = it->second;
    *got = it->first;
    big_free_bytes -= it->first;
    big_free.erase(it);
    return p;
}
void CtxCore::give_big(void* p, size_t bytes) {
    std::lock_guard<std::mutex> g(mu);
    big_free.emplace(bytes, p);
    big_free_bytes += bytes;
}
void CtxCore::drop_big() {
    std::multimap<size_t, void*> blocks;
    { std::lock_guard<std::mutex> g(mu); blocks.swap(big_free); big_free_bytes = 0; }
    for (auto& kv : blocks) cudaFreeAsync(kv.second, stream);
}

uint64_t* CtxCore::take_slot() {
    std::lock_guard<std::mutex> g(mu);
    if (slot_free.empty()) {
        uint64_t* chunk = nullptr;
        if (cudaHostAlloc((void**)&chunk, kSlotsPerChunk * kSlotWords * sizeof(uint64_t), cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        slot_chunks.push_back(chunk);
        for (size_t i = kSlotsPerChunk; i-- > 0;) slot_free.push_back(chunk + i * kSlotWords);
    }
    uint64_t* s = slot_free.back();
    slot_free.pop_back();
    return s;
}

int CtxCore::prof_flush() {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs;
    { std::lock_guard<std::mutex> g(mu); evs.swap(prof_events); }
    for (auto& e : evs) {
        RVL_CUDA_TRY(cudaEventSynchronize(e.second));
        float ms = 0.f;
        RVL_CUDA_TRY(cudaEventElapsedTime(&ms, e.first, e.second));
        prof_ms += ms;
        prof_launches++;
        prof_times.push_back((double)ms);
        cudaEventDestroy(e.first); cudaEventDestroy(e.second);
    }
    return RVL_OK;
}

DevBuffer::~DevBuffer() {
    if (owned && ptr != nullptr && core) {
        cudaSetDevice(core->device);
        if (kind == 0) cudaFreeAsync(ptr, core->stream);
        else if (kind == 3) core->give_big(ptr, bytes);
        else {
            cudaStreamSynchronize(core->stream);
            if (kind == 1) cudaFree(ptr); else cudaIpcCloseMemHandle(ptr);
        }
    }
}

int dev_alloc(const CoreRef& core, size_t bytes, BufRef* out) {
    // pad so word / 16-byte vector reads at the tail of a buffer stay inside the allocation
    size_t padded = ((bytes + 64 + 255) / 256) * 256;
    void* p = nullptr;
    const bool big = padded >= CtxCore::kBigBlock;
    if (big) {
        padded = (padded + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);   // the driver maps 2 MiB pages anyway; fewer distinct sizes
        size_t got = 0;
        if ((p = core->take_big(padded, &got)) != nullptr) {
            auto b = std::make_shared<DevBuffer>();
            b->ptr = p; b->bytes = got; b->owned = true; b->kind = 3; b->core = core;
            *out = std::move(b);
            return RVL_OK;
        }
    }
    static const bool trace = std::getenv("RVL_TRACE_ALLOC") != nullptr;
    const auto t0 = trace ? std::chrono::steady_clock::now() : std::chrono::steady_clock::time_point();
    cudaError_t e = cudaMallocAsync(&p, padded, core->stream);
    if (e == cudaErrorMemoryAllocation && core->big_free_bytes > 0) {   // give the context's cached blocks back and try once more
        cudaGetLastError();
        core->drop_big();
        e = cudaMallocAsync(&p, padded, core->stream);
    }
    if (trace) {
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 0.2) std::fprintf(stderr, "[rvl] cudaMallocAsync(%zu) took %.3f ms\n", padded, ms);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? RVL_OUT_OF_MEMORY : RVL_CUDA,
                    std::string("cudaMallocAsync(") + std::to_string(padded) + "): " + cudaGetErrorString(e));
    }
    auto b = std::make_shared<DevBuffer>();
    b->ptr = p; b->bytes = padded; b->owned = true; b->kind = big ? 3 : 0; b->core = core;
    *out = std::move(b);
    return RVL_OK;
}

int dev_alloc_zeroed(const CoreRef& core, size_t bytes, BufRef* out) {
    RVL_TRY(dev_alloc(core, bytes, out));
    RVL_CUDA_TRY(cudaMemsetAsync((*out)->ptr, 0, (*out)->bytes, core->stream));
    return RVL_OK;
}

BufRef wrap_external(const CoreRef& core, const void* ptr, size_t bytes) {
    auto b = std::make_shared<DevBuffer>();
    b->ptr = const_cast<void*>(ptr); b->bytes = bytes; b->owned = false; b->core = core;
    return b;
}

BitSrc bitsrc_of(const BufRef& buf, int64_t offset, int64_t length) {
    BitSrc s{nullptr, 0, 0};
    if (!buf) return s;
    const uintptr_t p = reinterpret_cast<uintptr_t>(buf->ptr);
    const uintptr_t aligned = p & ~uintptr_t(3);
    s.words = reinterpret_cast<const uint32_t*>(aligned);
    s.bit0 = (uint64_t)(p - aligned) * 8ull + (uint64_t)offset;
    s.nwords = (s.bit0 + (uint64_t)length + 31ull) / 32ull;
    return s;
}

static inline int grid_for(int64_t n, int block, int sm_count) {
    int64_t g = (n + block - 1) / block;
    const int64_t cap = (int64_t)sm_count * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace rvl

using namespace rvl;

static const char* dtype_name(int d) {
    static const char* n[] = {"Null", "Boolean", "Int64", "Float64", "String"};
    return (d >= 0 && d <= 4) ? n[d] : "?";
}

extern "C" {

int32_t rvl_abi_version(void) { return RVL_ABI_VERSION; }
const char* rvl_last_error(void) { return g_last_error.c_str(); }

int32_t rvl_device_count(int32_t* count) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { cudaGetLastError(); *count = 0; return fail(RVL_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
    *count = c;
    return RVL_OK;
}

int32_t rvl_ctx_create(int32_t device, rvl_ctx** ctx) {
    if (!ctx) return fail(RVL_INVALID_ARGUMENT, "ctx is NULL");
    int n = 0;
    RVL_CUDA_TRY(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(RVL_INVALID_ARGUMENT, "device " + std::to_string(device) + " out of range (" + std::to_string(n) + " devices)");
    RVL_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    RVL_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(RVL_CUDA, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                                   "; this library carries sm_100a code only (no fallback path)");
    auto core = std::make_shared<CtxCore>();
    core->device = device;
    core->sm_count = prop.multiProcessorCount;
    core->device_bytes = prop.totalGlobalMem;
    RVL_CUDA_TRY(cudaStreamCreateWithFlags(&core->stream, cudaStreamNonBlocking));
    RVL_CUDA_TRY(cudaStreamCreateWithFlags(&core->copy_stream, cudaStreamNonBlocking));
    RVL_CUDA_TRY(cudaStreamCreateWithFlags(&core->d2h_stream, cudaStreamNonBlocking));
    RVL_CUDA_TRY(cudaStreamCreateWithFlags(&core->side_stream, cudaStreamNonBlocking));
    RVL_TRY(fp_init_device(device));
    core->mailbox_words = CtxCore::kSlotBase;
    RVL_CUDA_TRY(cudaHostAlloc((void**)&core->mailbox, core->mailbox_words * sizeof(uint64_t), cudaHostAllocDefault));
    // keep freed blocks cached in the pool: outputs are sized for the worst case and recycled call to call
    cudaMemPool_t pool;
    RVL_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t threshold = UINT64_MAX;
    RVL_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    // the gather of projected columns touches isolated 32-byte sectors at low selectivity: ask L2 not to widen
    // those misses to 128-byte fetches (measured: 128 B granularity moved 1.8x the algorithmic bytes at 10 %)
    if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32) != cudaSuccess) cudaGetLastError();
    *ctx = new rvl_ctx{core};
    return RVL_OK;
}

int32_t rvl_ctx_destroy(rvl_ctx* ctx) {
    if (!ctx) return RVL_OK;
    cudaSetDevice(ctx->core->device);
    cudaStreamSynchronize(ctx->core->stream);
    delete ctx;
    return RVL_OK;
}

int32_t rvl_ctx_synchronize(rvl_ctx* ctx) {
    RVL_CUDA_TRY(cudaSetDevice(ctx->core->device));
    RVL_CUDA_TRY(cudaStreamSynchronize(ctx->core->copy_stream));
    RVL_CUDA_TRY(cudaStreamSynchronize(ctx->core->stream));
    return RVL_OK;
}

int32_t rvl_ctx_trim(rvl_ctx* ctx) {
    if (!ctx) return fail(RVL_INVALID_ARGUMENT, "null context");
    RVL_CUDA_TRY(cudaSetDevice(ctx->core->device));
    ctx->core->drop_big();
    ctx->core->drop_pinned();
    RVL_CUDA_TRY(cudaStreamSynchronize(ctx->core->stream));
    cudaMemPool_t pool;
    RVL_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, ctx->core->device));
    RVL_CUDA_TRY(cudaMemPoolTrimTo(pool, 0));
    return RVL_OK;
}

int32_t rvl_ctx_pool_stats(rvl_ctx* ctx, uint64_t* reserved_bytes, uint64_t* used_bytes) {
    if (!ctx) return fail(RVL_INVALID_ARGUMENT, "null context");
    RVL_CUDA_TRY(cudaSetDevice(ctx->core->device));
    cudaMemPool_t pool;
    RVL_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, ctx->core->device));
    uint64_t r = 0, u = 0;
    RVL_CUDA_TRY(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &r));
    RVL_CUDA_TRY(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &u));
    if (reserved_bytes) *reserved_bytes = r;
    if (used_bytes) *used_bytes = u;
    return RVL_OK;
}

// debugging aid (not part of the ABI header): the pinned words kernels leave behind under RVL_CHUNK_DEBUG=2
int32_t rvl_debug_read(rvl_ctx* ctx, uint64_t* out, int32_t n) {
    for (int i = 0; i < n && i < 128; ++i) out[i] = ctx->core->mailbox[128 + i];
    return RVL_OK;
}

int32_t rvl_ctx_cuda_stream(rvl_ctx* ctx, void** s) { *s = (void*)ctx->core->stream; return RVL_OK; }
int32_t rvl_ctx_device(rvl_ctx* ctx, int32_t* d) { *d = ctx->core->device; return RVL_OK; }
int32_t rvl_ctx_launch_count(rvl_ctx* ctx, int64_t* n) { *n = ctx->core->launches.load(); return RVL_OK; }

int32_t rvl_ctx_set_option(rvl_ctx* ctx, int32_t option, int64_t value) {
    if (!ctx) return fail(RVL_INVALID_ARGUMENT, "null context");
    CtxCore& c = *ctx->core;
    switch (option) {
        case RVL_OPT_PLAN:
            if (value < RVL_PLAN_AUTO || value > RVL_PLAN_TWO_PASS) return fail(RVL_INVALID_ARGUMENT, "unknown plan");
            c.plan_mode = (int)value; return RVL_OK;
        case RVL_OPT_TWO_PASS_MIN_ROWS: c.two_pass_min_rows = value; return RVL_OK;
        case RVL_OPT_SPARSE_MAX:
            if (value < 0 || value > 640) return fail(RVL_INVALID_ARGUMENT, "sparse_max must be in [0, 640]");
            c.sparse_max = (int)value; return RVL_OK;
        case RVL_OPT_DENSE_SLOTS:
            if (value < 2 || value > 14) return fail(RVL_INVALID_ARGUMENT, "dense_slots must be in [2, 14]");
            c.dense_slots = (int)value; return RVL_OK;
        case RVL_OPT_DENSE_CTAS_PER_SM:
            if (value < 1 || value > 2) return fail(RVL_INVALID_ARGUMENT, "dense_ctas_per_sm must be 1 or 2");
            c.dense_ctas_per_sm = (int)value; return RVL_OK;
        case RVL_OPT_SCAN_SLOTS:
            if (value < 1 || value > 16) return fail(RVL_INVALID_ARGUMENT, "scan_slots must be in [1, 16]");
            c.scan_slots = (int)value; return RVL_OK;
        case RVL_OPT_SCAN_ITEM_ROWS:
            if (value != 0 && value != 256 && value != 512) return fail(RVL_INVALID_ARGUMENT, "scan_item_rows must be 0, 256 or 512");
            c.scan_item_rows = (int)value; return RVL_OK;
        case RVL_OPT_SCAN_WARPS:
            if (value != 8 && value != 16 && value != 32) return fail(RVL_INVALID_ARGUMENT, "scan_warps must be 8, 16 or 32");
            c.scan_warps = (int)value; return RVL_OK;
        case RVL_OPT_DENSE_WARPS:
            if (value != 8 && value != 16) return fail(RVL_INVALID_ARGUMENT, "dense_warps must be 8 or 16");
            c.dense_warps = (int)value; return RVL_OK;
        case RVL_OPT_BITS_OVERLAP: c.bits_overlap = value != 0; return RVL_OK;
        case RVL_OPT_CHUNK_PLAN:
            if (value < 0 || value > 2) return fail(RVL_INVALID_ARGUMENT, "chunk_plan must be 0 (never), 1 (by selectivity) or 2 (always)");
            c.chunk_plan = (int)value; return RVL_OK;
        case RVL_OPT_STRING_KERNEL:
            if (value < 1 || value > 4) return fail(RVL_INVALID_ARGUMENT, "string_kernel must be in [1, 4]");
            c.string_kernel = (int)value; return RVL_OK;
        case RVL_OPT_STRING_DENSE_MIN:
            if (value < 0 || value > 1025) return fail(RVL_INVALID_ARGUMENT, "string_dense_min must be in [0, 1025]");
            c.string_dense_min = (int)value; return RVL_OK;
        case RVL_OPT_EXACT_ALLOC:
            if (value < 0 || value > 2) return fail(RVL_INVALID_ARGUMENT, "exact_alloc must be 0 (never), 1 (always) or 2 (auto)");
            c.exact_alloc = (int)value; return RVL_OK;
        default: return fail(RVL_INVALID_ARGUMENT, "unknown option");
    }
}

int32_t rvl_ctx_profile_enable(rvl_ctx* ctx, int32_t enable) {
    if (!ctx) return fail(RVL_INVALID_ARGUMENT, "null argument");
    RVL_TRY(ctx->core->prof_flush());
    ctx->core->profile = enable != 0;
    ctx->core->prof_ms = 0.0; ctx->core->prof_launches = 0;
    return RVL_OK;
}
int32_t rvl_ctx_profile_read(rvl_ctx* ctx, double* ms, int64_t* launches) {
    if (!ctx) return fail(RVL_INVALID_ARGUMENT, "null argument");
    RVL_CUDA_TRY(cudaSetDevice(ctx->core->device));
    RVL_TRY(ctx->core->prof_flush());
    if (ms) *ms = ctx->core->prof_ms;
    if (launches) *launches = ctx->core->prof_launches;
    ctx->core->prof_ms = 0.0; ctx->core->prof_launches = 0; ctx->core->prof_times.clear();
    return RVL_OK;
}
int32_t rvl_ctx_profile_read_launches(rvl_ctx* ctx, double* ms_out, int64_t cap, int64_t* n) {
    if (!ctx || !n) return fail(RVL_INVALID_ARGUMENT, "null argument");
    RVL_CUDA_TRY(cudaSetDevice(ctx->core->device));
    RVL_TRY(ctx->core->prof_flush());
    const int64_t have = (int64_t)ctx->core->prof_times.size();
    *n = have;
    for (int64_t i = 0; i < have && i < cap && ms_out; ++i) ms_out[i] = ctx->core->prof_times[(size_t)i];
    ctx->core->prof_ms = 0.0; ctx->core->prof_launches = 0; ctx->core->prof_times.clear();
    return RVL_OK;
}

int32_t rvl_host_alloc(size_t bytes, void** ptr) {
    RVL_CUDA_TRY(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return RVL_OK;
}
int32_t rvl_host_free(void* ptr) {
    if (ptr) RVL_CUDA_TRY(cudaFreeHost(ptr));
    return RVL_OK;
}

// ------------------------------------------------------------------------------------------ batches
static int check_columns(const rvl_column* cols, int32_t ncols, int64_t* rows) {
    // record_batch.rs:25-40
    int64_t n = ncols > 0 ? cols[0].length : 0;
    for (int i = 0; i < ncols; ++i) {
        if (cols[i].dtype < RVL_NULL || cols[i].dtype > RVL_STRING) return fail(RVL_INVALID_ARGUMENT, "Column " + std::to_string(i) + " has unknown dtype");
        if (cols[i].length < 0 || cols[i].offset < 0) return fail(RVL_INVALID_ARGUMENT, "negative length/offset");
        if (cols[i].length != n)
            return fail(RVL_LENGTH_MISMATCH, "Column " + std::to_string(i) + " has length " + std::to_string(cols[i].length) + " but expected " + std::to_string(n));
        const bool need_values = cols[i].dtype == RVL_INT64 || cols[i].dtype == RVL_FLOAT64 || cols[i].dtype == RVL_BOOLEAN;
        if (need_values && cols[i].length > 0 && cols[i].values == nullptr) return fail(RVL_INVALID_ARGUMENT, "Column " + std::to_string(i) + " has no values buffer");
        if (cols[i].dtype == RVL_STRING && cols[i].offsets == nullptr) return fail(RVL_INVALID_ARGUMENT, "Column " + std::to_string(i) + " has no offsets buffer");
        if (cols[i].dtype == RVL_STRING && cols[i].data_len > (int64_t)INT32_MAX)
            return fail(RVL_OFFSET_OVERFLOW, "Column " + std::to_string(i) + ": string data exceeds the int32 offset range; split the batch");
    }
    *rows = n;
    return RVL_OK;
}

int32_t rvl_batch_upload(rvl_ctx* ctx, const rvl_column* cols, int32_t ncols, rvl_batch** out) {
    if (!ctx || !out || (ncols > 0 && !cols)) return fail(RVL_INVALID_ARGUMENT, "null argument");
    const CoreRef& core = ctx->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    int64_t n = 0;
    RVL_TRY(check_columns(cols, ncols, &n));
    auto b = std::make_unique<rvl_batch>();
    b->core = core; b->num_rows = n;
    for (int i = 0; i < ncols; ++i) {
        const rvl_column& c = cols[i];
        const cudaMemcpyKind kind = c.location == RVL_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        DevColumn d;
        d.dtype = c.dtype; d.length = n;
        const int64_t resid = c.offset % 64;  // keep the sub-64-row residual so bit offsets survive the copy
        const int64_t start = c.offset - resid;
        d.offset = resid;
        const int64_t span = resid + n;
        if (c.dtype == RVL_INT64 || c.dtype == RVL_FLOAT64) {
            RVL_TRY(dev_alloc(core, (size_t)span * 8, &d.values));
            if (span > 0) RVL_CUDA_TRY(cudaMemcpyAsync(d.values->ptr, (const uint8_t*)c.values + start * 8, (size_t)span * 8, kind, core->stream));
        } else if (c.dtype == RVL_BOOLEAN) {
            RVL_TRY(dev_alloc_zeroed(core, (size_t)(span + 7) / 8, &d.values));
            if (span > 0) RVL_CUDA_TRY(cudaMemcpyAsync(d.values->ptr, (const uint8_t*)c.values + start / 8, (size_t)(span + 7) / 8, kind, core->stream));
        } else if (c.dtype == RVL_STRING) {
            RVL_TRY(dev_alloc(core, (size_t)(span + 1) * 4, &d.offsets));
            RVL_CUDA_TRY(cudaMemcpyAsync(d.offsets->ptr, c.offsets + start, (size_t)(span + 1) * 4, kind, core->stream));
            RVL_TRY(dev_alloc(core, (size_t)c.data_len, &d.data));
            if (c.data_len > 0) RVL_CUDA_TRY(cudaMemcpyAsync(d.data->ptr, c.data, (size_t)c.data_len, kind, core->stream));
            d.data_len = c.data_len;
            if (c.location != RVL_DEVICE) d.window_bytes = (int64_t)c.offsets[c.offset + n] - (int64_t)c.offsets[c.offset];
        }
        if (c.validity != nullptr && c.dtype != RVL_NULL) {
            RVL_TRY(dev_alloc_zeroed(core, (size_t)(span + 7) / 8, &d.validity));
            if (span > 0) RVL_CUDA_TRY(cudaMemcpyAsync(d.validity->ptr, c.validity + start / 8, (size_t)(span + 7) / 8, kind, core->stream));
        } else {
            d.null_count = c.dtype == RVL_NULL ? n : 0;
        }
        b->cols.push_back(std::move(d));
    }
    // the source buffers are only borrowed for the duration of the call
    RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
    *out = b.release();
    return RVL_OK;
}

int32_t rvl_batch_wrap_device(rvl_ctx* ctx, const rvl_column* cols, int32_t ncols, rvl_batch** out) {
    if (!ctx || !out || (ncols > 0 && !cols)) return fail(RVL_INVALID_ARGUMENT, "null argument");
    int64_t n = 0;
    RVL_TRY(check_columns(cols, ncols, &n));
    auto b = std::make_unique<rvl_batch>();
    b->core = ctx->core; b->num_rows = n;
    for (int i = 0; i < ncols; ++i) {
        const rvl_column& c = cols[i];
        if (c.location != RVL_DEVICE) return fail(RVL_INVALID_ARGUMENT, "rvl_batch_wrap_device needs device buffers");
        DevColumn d;
        d.dtype = c.dtype; d.length = n; d.offset = c.offset;
        const int64_t span = c.offset + n;
        if (c.dtype == RVL_INT64 || c.dtype == RVL_FLOAT64) {
            if ((reinterpret_cast<uintptr_t>(c.values) & 7) != 0) return fail(RVL_INVALID_ARGUMENT, "values buffer must be 8-byte aligned");
            d.values = wrap_external(ctx->core, c.values, (size_t)span * 8);
        } else if (c.dtype == RVL_BOOLEAN) {
            d.values = wrap_external(ctx->core, c.values, (size_t)(span + 7) / 8);
        } else if (c.dtype == RVL_STRING) {
            if ((reinterpret_cast<uintptr_t>(c.offsets) & 3) != 0) return fail(RVL_INVALID_ARGUMENT, "offsets buffer must be 4-byte aligned");
            d.offsets = wrap_external(ctx->core, c.offsets, (size_t)(span + 1) * 4);
            d.data = wrap_external(ctx->core, c.data, (size_t)c.data_len);
            d.data_len = c.data_len;
        }
        if (c.validity != nullptr && c.dtype != RVL_NULL) d.validity = wrap_external(ctx->core, c.validity, (size_t)(span + 7) / 8);
        else d.null_count = c.dtype == RVL_NULL ? n : 0;
        b->cols.push_back(std::move(d));
    }
    *out = b.release();
    return RVL_OK;
}

int32_t rvl_batch_release(rvl_batch* batch) {
    delete batch;
    return RVL_OK;
}
int32_t rvl_batch_num_rows(const rvl_batch* batch, int64_t* rows) { *rows = batch->num_rows; return RVL_OK; }
int32_t rvl_batch_num_columns(const rvl_batch* batch, int32_t* n) { *n = (int32_t)batch->cols.size(); return RVL_OK; }

static int ensure_null_count(const rvl_batch* batch, int i) {
    DevColumn& c = const_cast<DevColumn&>(batch->cols[i]);
    if (c.null_count >= 0) return RVL_OK;
    if (!c.validity) { c.null_count = c.dtype == RVL_NULL ? c.length : 0; return RVL_OK; }
    const CoreRef& core = batch->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    BufRef counter;
    RVL_TRY(dev_alloc_zeroed(core, 8, &counter));
    count_ones_kernel<<<grid_for((c.length + 31) / 32, 256, core->sm_count), 256, 0, core->stream>>>(
        bitsrc_of(c.validity, c.offset, c.length), c.length, nullptr, nullptr, -1, (unsigned long long*)counter->ptr);
    core->launches++;
    RVL_CUDA_TRY(cudaGetLastError());
    RVL_CUDA_TRY(cudaMemcpyAsync(core->mailbox, counter->ptr, 8, cudaMemcpyDeviceToHost, core->stream));
    RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
    c.null_count = c.length - (int64_t)core->mailbox[0];
    return RVL_OK;
}

int32_t rvl_batch_column(const rvl_batch* batch, int32_t i, rvl_column* view) {
    if (!batch || !view) return fail(RVL_INVALID_ARGUMENT, "null argument");
    if (i < 0 || i >= (int32_t)batch->cols.size())
        return fail(RVL_OUT_OF_BOUNDS, "Column index " + std::to_string(i) + " out of bounds for " + std::to_string(batch->cols.size()) + " columns");
    RVL_TRY(ensure_null_count(batch, i));
    const DevColumn& c = batch->cols[i];
    std::memset(view, 0, sizeof *view);
    view->dtype = c.dtype; view->location = RVL_DEVICE; view->length = c.length; view->offset = c.offset;
    view->values = c.values ? c.values->ptr : nullptr;
    view->validity = c.validity ? (const uint8_t*)c.validity->ptr : nullptr;
    view->offsets = c.offsets ? (const int32_t*)c.offsets->ptr : nullptr;
    view->data = c.data ? (const uint8_t*)c.data->ptr : nullptr;
    view->data_len = c.data_len;
    view->null_count = c.null_count;
    return RVL_OK;
}

// copy `length` bits starting at bit `offset` of `src` to host `dst` rebased to bit 0, padding bits zero
static int download_bits(const CoreRef& core, const BufRef& src, int64_t offset, int64_t length, uint8_t* dst) {
    if (length == 0) return RVL_OK;
    const size_t nbytes = (size_t)(length + 7) / 8;
    if (offset % 8 == 0) {
        RVL_CUDA_TRY(cudaMemcpyAsync(dst, (const uint8_t*)src->ptr + offset / 8, nbytes, cudaMemcpyDeviceToHost, core->d2h_stream));
        RVL_CUDA_TRY(cudaStreamSynchronize(core->d2h_stream));
    } else {
        BufRef tmp;
        RVL_TRY(dev_alloc_zeroed(core, nbytes, &tmp));
        bitcopy_kernel<<<grid_for((length + 31) / 32 + 1, 256, core->sm_count), 256, 0, core->stream>>>(
            bitsrc_of(src, offset, length), BitSrc{nullptr, 0, 0}, (uint32_t*)tmp->ptr, 0, length);
        core->launches++;
        RVL_CUDA_TRY(cudaGetLastError());
        RVL_CUDA_TRY(cudaMemcpyAsync(dst, tmp->ptr, nbytes, cudaMemcpyDeviceToHost, core->stream));
        RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
    }
    if (length % 8 != 0) dst[nbytes - 1] &= (uint8_t)((1u << (length % 8)) - 1u);
    return RVL_OK;
}

int32_t rvl_batch_download_column(rvl_ctx* ctx, const rvl_batch* batch, int32_t i, rvl_column* dst) {
    if (!ctx || !batch || !dst) return fail(RVL_INVALID_ARGUMENT, "null argument");
    if (i < 0 || i >= (int32_t)batch->cols.size())
        return fail(RVL_OUT_OF_BOUNDS, "Column index " + std::to_string(i) + " out of bounds for " + std::to_string(batch->cols.size()) + " columns");
    const CoreRef& core = batch->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    const DevColumn& c = batch->cols[i];
    if (dst->dtype != c.dtype) return fail(RVL_TYPE_MISMATCH, std::string("Column ") + std::to_string(i) + " has type " + dtype_name(c.dtype) + " but destination expects " + dtype_name(dst->dtype));
    const int64_t n = c.length;
    if (c.dtype == RVL_INT64 || c.dtype == RVL_FLOAT64) {
        if (n > 0) {
            // finished batches are complete on every stream (their producers synchronise), so the copy-back
            // uses its own stream and overlaps later kernels / H2D staging
            RVL_CUDA_TRY(cudaMemcpyAsync(const_cast<void*>(dst->values), (const uint8_t*)c.values->ptr + c.offset * 8, (size_t)n * 8, cudaMemcpyDeviceToHost, core->d2h_stream));
            RVL_CUDA_TRY(cudaStreamSynchronize(core->d2h_stream));
        }
    } else if (c.dtype == RVL_BOOLEAN) {
        RVL_TRY(download_bits(core, c.values, c.offset, n, (uint8_t*)const_cast<void*>(dst->values)));
    } else if (c.dtype == RVL_STRING) {
        // rebase offsets to start at 0 and copy the referenced byte window
        int32_t first_last[2] = {0, 0};
        RVL_CUDA_TRY(cudaMemcpyAsync(&first_last[0], (const int32_t*)c.offsets->ptr + c.offset, 4, cudaMemcpyDeviceToHost, core->stream));
        RVL_CUDA_TRY(cudaMemcpyAsync(&first_last[1], (const int32_t*)c.offsets->ptr + c.offset + n, 4, cudaMemcpyDeviceToHost, core->stream));
        RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
        int32_t* hoff = const_cast<int32_t*>(dst->offsets);
        if (first_last[0] == 0) {
            RVL_CUDA_TRY(cudaMemcpyAsync(hoff, (const int32_t*)c.offsets->ptr + c.offset, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, core->stream));
        } else {
            BufRef tmp;
            RVL_TRY(dev_alloc(core, (size_t)(n + 1) * 4, &tmp));
            RVL_CUDA_TRY(cudaMemsetAsync(tmp->ptr, 0, 4, core->stream));
            if (n > 0) {
                rebase_offsets_kernel<<<grid_for(n, 256, core->sm_count), 256, 0, core->stream>>>((const int32_t*)c.offsets->ptr + c.offset, (int32_t*)tmp->ptr, n, 0);
                core->launches++;
                RVL_CUDA_TRY(cudaGetLastError());
            }
            RVL_CUDA_TRY(cudaMemcpyAsync(hoff, tmp->ptr, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, core->stream));
        }
        const int64_t nbytes = (int64_t)first_last[1] - first_last[0];
        if (nbytes > dst->data_len) return fail(RVL_LENGTH_MISMATCH, "destination data buffer too small: need " + std::to_string(nbytes) + " bytes");
        if (nbytes > 0) RVL_CUDA_TRY(cudaMemcpyAsync(const_cast<uint8_t*>(dst->data), (const uint8_t*)c.data->ptr + first_last[0], (size_t)nbytes, cudaMemcpyDeviceToHost, core->stream));
        RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
        dst->data_len = nbytes;
    }
    if (c.validity && dst->validity != nullptr) RVL_TRY(download_bits(core, c.validity, c.offset, n, const_cast<uint8_t*>(dst->validity)));
    dst->length = n; dst->offset = 0;
    RVL_TRY(ensure_null_count(batch, i));
    dst->null_count = c.null_count;
    return RVL_OK;
}

int32_t rvl_batch_count_true(rvl_ctx* ctx, const rvl_batch* batch, int32_t i, int64_t* count) {
    if (!ctx || !batch || !count) return fail(RVL_INVALID_ARGUMENT, "null argument");
    if (i < 0 || i >= (int32_t)batch->cols.size())
        return fail(RVL_OUT_OF_BOUNDS, "Column index " + std::to_string(i) + " out of bounds for " + std::to_string(batch->cols.size()) + " columns");
    const DevColumn& c = batch->cols[i];
    if (c.dtype != RVL_BOOLEAN) return fail(RVL_TYPE_MISMATCH, std::string("Column ") + std::to_string(i) + " has type " + dtype_name(c.dtype) + " but count_true expects Boolean");
    const CoreRef& core = batch->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    *count = 0;
    if (c.length == 0) return RVL_OK;
    BufRef counter, tmp;
    RVL_TRY(dev_alloc_zeroed(core, 8, &counter));
    BitSrc src = bitsrc_of(c.values, c.offset, c.length);
    if (c.validity) {
        // values AND validity, rebased to bit 0 (a null is never Some(true))
        RVL_TRY(dev_alloc_zeroed(core, (size_t)((c.length + 63) / 64) * 8, &tmp));
        bitcopy_kernel<<<grid_for((c.length + 31) / 32 + 1, 256, core->sm_count), 256, 0, core->stream>>>(
            src, bitsrc_of(c.validity, c.offset, c.length), (uint32_t*)tmp->ptr, 0, c.length);
        core->launches++;
        src = bitsrc_of(tmp, 0, c.length);
    }
    count_ones_kernel<<<grid_for((c.length + 31) / 32, 256, core->sm_count), 256, 0, core->stream>>>(src, c.length, nullptr, nullptr, -1,
                                                                                                      (unsigned long long*)counter->ptr);
    core->launches++;
    RVL_CUDA_TRY(cudaGetLastError());
    RVL_CUDA_TRY(cudaMemcpyAsync(core->mailbox, counter->ptr, 8, cudaMemcpyDeviceToHost, core->stream));
    RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
    *count = (int64_t)core->mailbox[0];
    return RVL_OK;
}

int32_t rvl_boolean_op(rvl_ctx* ctx, int32_t op, const rvl_batch* a, int32_t a_col, const rvl_batch* b, int32_t b_col, rvl_batch** out) {
    if (!ctx || !a || !out || (op != RVL_BOOL_NOT && !b)) return fail(RVL_INVALID_ARGUMENT, "null argument");
    if (op < RVL_BOOL_AND || op > RVL_BOOL_NOT) return fail(RVL_INVALID_ARGUMENT, "unknown boolean operation");
    if (a_col < 0 || a_col >= (int32_t)a->cols.size() || (op != RVL_BOOL_NOT && (b_col < 0 || b_col >= (int32_t)b->cols.size())))
        return fail(RVL_OUT_OF_BOUNDS, "Column index out of bounds");
    const DevColumn& ca = a->cols[a_col];
    const DevColumn* cb = op != RVL_BOOL_NOT ? &b->cols[b_col] : nullptr;
    if (ca.dtype != RVL_BOOLEAN || (cb && cb->dtype != RVL_BOOLEAN)) return fail(RVL_TYPE_MISMATCH, "logical operations need Boolean arrays");
    if (cb && cb->length != ca.length) return fail(RVL_LENGTH_MISMATCH, "Array lengths must match for logical operations");  // boolean.rs:121-123
    const CoreRef& core = ctx->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    const int64_t n = ca.length;
    auto res = std::make_unique<rvl_batch>();
    res->core = core; res->num_rows = n;
    DevColumn d;
    d.dtype = RVL_BOOLEAN; d.length = n; d.offset = 0; d.null_count = -1;
    const size_t bytes = (size_t)((n + 31) / 32) * 4 + 8;
    RVL_TRY(dev_alloc_zeroed(core, bytes, &d.values));
    RVL_TRY(dev_alloc_zeroed(core, bytes, &d.validity));
    if (n > 0) {
        boolean_op_kernel<<<grid_for((n + 31) / 32, 256, core->sm_count), 256, 0, core->stream>>>(
            bitsrc_of(ca.values, ca.offset, n), bitsrc_of(ca.validity, ca.offset, n),
            cb ? bitsrc_of(cb->values, cb->offset, n) : BitSrc{nullptr, 0, 0}, cb ? bitsrc_of(cb->validity, cb->offset, n) : BitSrc{nullptr, 0, 0},
            op, n, (uint32_t*)d.values->ptr, (uint32_t*)d.validity->ptr);
        core->launches++;
        RVL_CUDA_TRY(cudaGetLastError());
    }
    res->cols.push_back(std::move(d));
    RVL_TRY(ensure_null_count(res.get(), 0));
    if (res->cols[0].null_count == 0) res->cols[0].validity.reset();  // BooleanArrayBuilder::finish keeps a bitmap only with nulls (boolean.rs:280-286)
    *out = res.release();
    return RVL_OK;
}

int32_t rvl_batch_slice(const rvl_batch* batch, int64_t offset, int64_t length, rvl_batch** view) {
    if (!batch || !view) return fail(RVL_INVALID_ARGUMENT, "null argument");
    if (offset < 0 || length < 0 || offset + length > batch->num_rows) return fail(RVL_OUT_OF_BOUNDS, "Slice out of bounds");  // record_batch.rs:93
    auto b = std::make_unique<rvl_batch>();
    b->core = batch->core; b->num_rows = length;
    for (const DevColumn& c : batch->cols) {
        DevColumn d = c;  // shares the buffers (Arc clone in the reference)
        d.offset = c.offset + offset; d.length = length;
        if (length != c.length) d.window_bytes = -1;   // the narrower window's byte span is not known on the host
        d.null_count = c.dtype == RVL_NULL ? length : (c.validity ? -1 : 0);
        b->cols.push_back(std::move(d));
    }
    *view = b.release();
    return RVL_OK;
}

int32_t rvl_batch_select(const rvl_batch* batch, const int32_t* indices, int32_t n, rvl_batch** view) {
    if (!batch || !view || (n > 0 && !indices)) return fail(RVL_INVALID_ARGUMENT, "null argument");
    for (int i = 0; i < n; ++i)
        if (indices[i] < 0 || indices[i] >= (int32_t)batch->cols.size())
            return fail(RVL_OUT_OF_BOUNDS, "Column index " + std::to_string(indices[i]) + " out of bounds for " + std::to_string(batch->cols.size()) + " columns");  // record_batch.rs:183-187
    auto b = std::make_unique<rvl_batch>();
    b->core = batch->core; b->num_rows = batch->num_rows;
    for (int i = 0; i < n; ++i) b->cols.push_back(batch->cols[indices[i]]);
    *view = b.release();
    return RVL_OK;
}

// take_array (record_batch.rs:131-178) of the listed columns of `batch` by a DEVICE index list, appended to `res` (whose num_rows is n).
// Bitmaps are kept only when a taken row is null (primitive.rs:180-185).  Synchronises the stream.
int rvl_internal_take_rows(const CoreRef& core, const rvl_batch* batch, const int32_t* cols, int32_t ncols, const int64_t* idx, int64_t n, rvl_batch* res) {
    const size_t wbytes = (size_t)((n + 31) / 32) * 4 + 8;
    const unsigned grid = (unsigned)std::max<int64_t>(1, (n + 255) / 256);
    const size_t first = res->cols.size();
    for (int32_t ci = 0; ci < ncols; ++ci) {
        const DevColumn& s = batch->cols[(size_t)(cols ? cols[ci] : ci)];
        DevColumn d;
        d.dtype = s.dtype; d.length = n; d.offset = 0; d.null_count = -1;
        const BitSrc sv = bitsrc_of(s.validity, s.offset, s.length);
        if (s.validity && s.dtype != RVL_NULL) RVL_TRY(dev_alloc_zeroed(core, wbytes, &d.validity));
        uint32_t* ov = d.validity ? (uint32_t*)d.validity->ptr : nullptr;
        if (s.dtype == RVL_INT64 || s.dtype == RVL_FLOAT64) {
            RVL_TRY(dev_alloc(core, (size_t)std::max<int64_t>(n, 1) * 8, &d.values));
            if (n > 0) {
                take_col8_kernel<<<grid, 256, 0, core->stream>>>((const uint64_t*)s.values->ptr + s.offset, sv, idx, n, (uint64_t*)d.values->ptr, ov);
                core->launches++;
            }
        } else if (s.dtype == RVL_BOOLEAN) {
            RVL_TRY(dev_alloc_zeroed(core, wbytes, &d.values));
            if (n > 0) {
                take_bits_kernel<<<grid, 256, 0, core->stream>>>(bitsrc_of(s.values, s.offset, s.length), sv, idx, n, (uint32_t*)d.values->ptr, ov);
                core->launches++;
            }
        } else if (s.dtype == RVL_STRING) {
            BufRef lens;
            RVL_TRY(dev_alloc(core, (size_t)std::max<int64_t>(n, 1) * 4, &lens));
            RVL_TRY(dev_alloc(core, (size_t)(n + 1) * 4, &d.offsets));
            RVL_CUDA_TRY(cudaMemsetAsync(d.offsets->ptr, 0, 4, core->stream));
            int32_t total = 0;
            if (n > 0) {
                const int32_t* soff = (const int32_t*)s.offsets->ptr + s.offset;
                BufRef oflow;
                RVL_TRY(dev_alloc_zeroed(core, 4, &oflow));
                take_strlen_kernel<<<grid, 256, 0, core->stream>>>(soff, sv, idx, n, (int32_t*)lens->ptr);
                scan_lengths_kernel<<<1, 1024, 0, core->stream>>>((const int32_t*)lens->ptr, (int32_t*)d.offsets->ptr, n, (int32_t*)oflow->ptr);
                core->launches += 2;
                int32_t over = 0;
                RVL_CUDA_TRY(cudaMemcpyAsync(&total, (const int32_t*)d.offsets->ptr + n, 4, cudaMemcpyDeviceToHost, core->stream));
                RVL_CUDA_TRY(cudaMemcpyAsync(&over, oflow->ptr, 4, cudaMemcpyDeviceToHost, core->stream));
                RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
                // the reference wraps silently here (string.rs:31: `as i32`); a wrapped prefix can even come out positive, so the
                // scan itself reports the overflow
                if (over != 0 || total < 0) return fail(RVL_OFFSET_OVERFLOW, "taken string data exceeds the int32 offset range");
            }
            RVL_TRY(dev_alloc(core, (size_t)std::max<int32_t>(total, 1), &d.data));
            d.data_len = total;
            if (n > 0 && total > 0) {
                const int32_t* soff = (const int32_t*)s.offsets->ptr + s.offset;
                take_strcopy_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, core->stream>>>(soff, (const uint8_t*)s.data->ptr, idx, n,
                                                                                                    (const int32_t*)d.offsets->ptr, (uint8_t*)d.data->ptr);
                core->launches++;
            }
            if (ov != nullptr && n > 0) {
                take_bits_kernel<<<grid, 256, 0, core->stream>>>(BitSrc{nullptr, 0, 0}, sv, idx, n, nullptr, ov);
                core->launches++;
            }
        } else {
            d.null_count = n;
        }
        RVL_CUDA_TRY(cudaGetLastError());
        res->cols.push_back(std::move(d));
    }
    RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
    // the reference's builders keep a bitmap only when a taken row is null (primitive.rs:180-185)
    for (size_t c = first; c < res->cols.size(); ++c) {
        RVL_TRY(ensure_null_count(res, (int)c));
        if (res->cols[c].null_count == 0) res->cols[c].validity.reset();
    }
    return RVL_OK;
}

int32_t rvl_batch_take(rvl_ctx* ctx, const rvl_batch* batch, const int64_t* indices, int64_t n, rvl_batch** out) {
    if (!ctx || !batch || !out || n < 0 || (n > 0 && !indices)) return fail(RVL_INVALID_ARGUMENT, "null argument");
    const CoreRef& core = ctx->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    for (int64_t i = 0; i < n; ++i)
        if (indices[i] < 0 || indices[i] >= batch->num_rows)
            return fail(RVL_OUT_OF_BOUNDS, "Index " + std::to_string(indices[i]) + " out of bounds for " + std::to_string(batch->num_rows) + " rows");  // record_batch.rs:111-114
    auto res = std::make_unique<rvl_batch>();
    res->core = core; res->num_rows = n;
    BufRef didx;
    RVL_TRY(dev_alloc(core, (size_t)std::max<int64_t>(n, 1) * 8, &didx));
    if (n > 0) RVL_CUDA_TRY(cudaMemcpyAsync(didx->ptr, indices, (size_t)n * 8, cudaMemcpyHostToDevice, core->stream));
    RVL_TRY(rvl_internal_take_rows(core, batch, nullptr, (int32_t)batch->cols.size(), (const int64_t*)didx->ptr, n, res.get()));  // synchronises: `indices` may go
    *out = res.release();
    return RVL_OK;
}

int32_t rvl_batch_concat(rvl_ctx* ctx, const rvl_batch* const* batches, int32_t n, rvl_batch** out) {
    if (!ctx || !out) return fail(RVL_INVALID_ARGUMENT, "null argument");
    if (n <= 0 || !batches) return fail(RVL_INVALID_ARGUMENT, "Cannot concatenate empty batch list");  // record_batch.rs:247
    const CoreRef& core = ctx->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    const size_t ncols = batches[0]->cols.size();
    int64_t total = 0;
    for (int b = 0; b < n; ++b) {
        if (batches[b]->cols.size() != ncols) return fail(RVL_SCHEMA_MISMATCH, "All batches must have the same schema");  // :253
        for (size_t c = 0; c < ncols; ++c)
            if (batches[b]->cols[c].dtype != batches[0]->cols[c].dtype) return fail(RVL_SCHEMA_MISMATCH, "All batches must have the same schema");
        total += batches[b]->num_rows;
    }
    auto res = std::make_unique<rvl_batch>();
    res->core = core; res->num_rows = total;
    // device counters: [2c] ones in validity of column c, [2c+1] string bytes of column c
    BufRef counters;
    RVL_TRY(dev_alloc_zeroed(core, ncols * 16 + 16, &counters));
    unsigned long long* dctr = (unsigned long long*)counters->ptr;
    std::vector<BufRef> keep_alive;
    for (size_t c = 0; c < ncols; ++c) {
        DevColumn d;
        d.dtype = batches[0]->cols[c].dtype; d.length = total; d.offset = 0;
        bool any_validity = false;
        for (int b = 0; b < n; ++b) any_validity |= (bool)batches[b]->cols[c].validity;
        if (any_validity && d.dtype != RVL_NULL) RVL_TRY(dev_alloc_zeroed(core, (size_t)(total + 7) / 8, &d.validity));
        if (d.dtype == RVL_INT64 || d.dtype == RVL_FLOAT64) RVL_TRY(dev_alloc(core, (size_t)total * 8, &d.values));
        if (d.dtype == RVL_BOOLEAN) RVL_TRY(dev_alloc_zeroed(core, (size_t)(total + 7) / 8, &d.values));
        BufRef chain;
        int chain_k = 0;
        if (d.dtype == RVL_STRING) {
            int64_t cap = 0;
            for (int b = 0; b < n; ++b) cap += batches[b]->cols[c].data_len;
            if (cap > (int64_t)INT32_MAX) {
                // data_len is an upper bound; only fail when the bytes actually referenced overflow
                int64_t exact = 0;
                for (int b = 0; b < n; ++b) {
                    const DevColumn& s = batches[b]->cols[c];
                    int32_t fl[2];
                    RVL_CUDA_TRY(cudaMemcpyAsync(&fl[0], (const int32_t*)s.offsets->ptr + s.offset, 4, cudaMemcpyDeviceToHost, core->stream));
                    RVL_CUDA_TRY(cudaMemcpyAsync(&fl[1], (const int32_t*)s.offsets->ptr + s.offset + s.length, 4, cudaMemcpyDeviceToHost, core->stream));
                    RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
                    exact += (int64_t)fl[1] - fl[0];
                }
                if (exact > (int64_t)INT32_MAX) return fail(RVL_OFFSET_OVERFLOW, "concatenated string data (" + std::to_string(exact) + " bytes) exceeds the int32 offset range");
                cap = exact;
            }
            RVL_TRY(dev_alloc(core, (size_t)(total + 1) * 4, &d.offsets));
            RVL_CUDA_TRY(cudaMemsetAsync(d.offsets->ptr, 0, 4, core->stream));
            RVL_TRY(dev_alloc(core, (size_t)cap, &d.data));
            RVL_TRY(dev_alloc_zeroed(core, (size_t)(n + 1) * 8, &chain));      // bytes emitted before each part
            keep_alive.push_back(chain);
        }
        int64_t row_base = 0;
        for (int b = 0; b < n; ++b) {
            const DevColumn& s = batches[b]->cols[c];
            const int64_t m = s.length;
            if (m == 0) continue;
            const BitSrc sv = bitsrc_of(s.validity, s.offset, m);
            if (d.dtype == RVL_INT64 || d.dtype == RVL_FLOAT64) {
                if (!s.validity) {
                    // nothing to zero: a plain device copy (copy engine / peer copy over NVLink for rvl_gather_to parts)
                    RVL_CUDA_TRY(cudaMemcpyAsync((uint64_t*)d.values->ptr + row_base, (const uint64_t*)s.values->ptr + s.offset, (size_t)m * 8, cudaMemcpyDefault, core->stream));
                } else {
                    copy_col8_zero_nulls_kernel<<<grid_for(m, 256, core->sm_count), 256, 0, core->stream>>>(
                        (const uint64_t*)s.values->ptr + s.offset, sv, (uint64_t*)d.values->ptr + row_base, m);
                    core->launches++;
                }
            } else if (d.dtype == RVL_BOOLEAN) {
                bitcopy_kernel<<<grid_for((m + 31) / 32 + 1, 256, core->sm_count), 256, 0, core->stream>>>(
                    bitsrc_of(s.values, s.offset, m), sv, (uint32_t*)d.values->ptr, (uint64_t)row_base, m);
                core->launches++;
            } else if (d.dtype == RVL_STRING) {
                StrGatherParams sp{};
                sp.n_rows = m; sp.limit = -1; sp.sel = nullptr; sp.tile_prefix = nullptr; sp.row_base = row_base;
                sp.offsets = (const int32_t*)s.offsets->ptr + s.offset; sp.data = (const uint8_t*)s.data->ptr; sp.valid = sv;
                sp.out_offsets = (int32_t*)d.offsets->ptr; sp.out_data = (uint8_t*)d.data->ptr;
                unsigned long long* ch = (unsigned long long*)chain->ptr;
                sp.byte_base_in = ch + chain_k; sp.bytes_total_out = ch + chain_k + 1;
                ++chain_k;
                RVL_TRY(str_prepare_sizes(core, sp, s, &keep_alive));
                RVL_TRY(str_launch_sizes(core, sp, core->stream));
                RVL_TRY(str_launch_gather(core, sp));
            }
            if (d.validity) {
                bitcopy_kernel<<<grid_for((m + 31) / 32 + 1, 256, core->sm_count), 256, 0, core->stream>>>(
                    sv, BitSrc{nullptr, 0, 0}, (uint32_t*)d.validity->ptr, (uint64_t)row_base, m);
                core->launches++;
            }
            RVL_CUDA_TRY(cudaGetLastError());
            row_base += m;
        }
        if (d.dtype == RVL_STRING && chain)   // total string bytes = the last link of the per-part chain
            RVL_CUDA_TRY(cudaMemcpyAsync(dctr + 2 * c + 1, (unsigned long long*)chain->ptr + chain_k, 8, cudaMemcpyDeviceToDevice, core->stream));
        if (d.validity && total > 0) {
            count_ones_kernel<<<grid_for((total + 31) / 32, 256, core->sm_count), 256, 0, core->stream>>>(
                bitsrc_of(d.validity, 0, total), total, nullptr, nullptr, -1, dctr + 2 * c);
            core->launches++;
            RVL_CUDA_TRY(cudaGetLastError());
        }
        res->cols.push_back(std::move(d));
    }
    if (ncols * 2 > CtxCore::kSlotBase) return fail(RVL_INVALID_ARGUMENT, "too many columns");
    RVL_CUDA_TRY(cudaMemcpyAsync(core->mailbox, dctr, ncols * 16, cudaMemcpyDeviceToHost, core->stream));
    RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
    for (size_t c = 0; c < ncols; ++c) {
        DevColumn& d = res->cols[c];
        if (d.dtype == RVL_NULL) { d.null_count = total; continue; }
        if (d.validity) {
            d.null_count = total - (int64_t)core->mailbox[2 * c];
            if (d.null_count == 0) d.validity.reset();  // bitmap dropped when nothing is null (primitive.rs:180-185)
        } else d.null_count = 0;
        if (d.dtype == RVL_STRING) d.data_len = (int64_t)core->mailbox[2 * c + 1];
    }
    *out = res.release();
    return RVL_OK;
}

// ------------------------------------------------------------------------------------------ synthetic tables
// Ordered physical concatenation of per-GPU results on one device (SURVEY.md 8(e) form (2)): the concat kernels of the destination
// context read the other GPUs' buffers in place through NVLink peer mappings — no staging copy, bit offsets handled by the same
// bit-granular kernels as a local concat, string offsets rebased by each part's byte prefix.
int32_t rvl_gather_to(rvl_ctx* dst, const rvl_batch* const* parts, int32_t n, rvl_batch** out) {
    if (!dst || !out) return fail(RVL_INVALID_ARGUMENT, "null argument");
    if (n <= 0 || !parts) return fail(RVL_INVALID_ARGUMENT, "Cannot concatenate empty batch list");
    const CoreRef& core = dst->core;
    for (int i = 0; i < n; ++i) {
        if (!parts[i]) return fail(RVL_INVALID_ARGUMENT, "null batch");
        const CoreRef& src = parts[i]->core;
        if (src->device != core->device) {
            int can = 0;
            RVL_CUDA_TRY(cudaDeviceCanAccessPeer(&can, core->device, src->device));
            if (!can) return fail(RVL_CUDA, "device " + std::to_string(core->device) + " cannot map the memory of device " + std::to_string(src->device));
            RVL_CUDA_TRY(cudaSetDevice(core->device));
            cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(RVL_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
            // stream-ordered allocations come from the source device's default pool: grant the destination device access to it
            cudaMemPool_t pool = nullptr;
            RVL_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, src->device));
            cudaMemAccessDesc desc{};
            desc.location.type = cudaMemLocationTypeDevice; desc.location.id = core->device; desc.flags = cudaMemAccessFlagsProtReadWrite;
            RVL_CUDA_TRY(cudaMemPoolSetAccess(pool, &desc, 1));
        }
        if (src.get() != core.get()) {
            // the part's producing kernels run on its own context's stream: order the gather behind them
            RVL_CUDA_TRY(cudaSetDevice(src->device));
            cudaEvent_t ev;
            RVL_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            RVL_CUDA_TRY(cudaEventRecord(ev, src->stream));
            RVL_CUDA_TRY(cudaSetDevice(core->device));
            RVL_CUDA_TRY(cudaStreamWaitEvent(core->stream, ev, 0));
            RVL_CUDA_TRY(cudaEventDestroy(ev));
        }
    }
    RVL_TRY(rvl_batch_concat(dst, parts, n, out));
    // the parts may be released by the caller as soon as this returns
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
    return RVL_OK;
}


// ------------------------------------------------------------------------------------------ multi-process ordered gather
// One process per GPU (torchrun): the destination rank allocates the result with cudaMalloc and exports each buffer as a CUDA IPC
// handle; every rank maps them and WRITES its own rows at its row offset over NVLink / NVSwitch, all ranks at once — the
// destination's ingest link is the only shared resource.  The row offsets are the exclusive scan of the per-shard survivor
// counts; bitmaps continue at arbitrary bit offsets (boundary words are OR-ed with peer atomics), string
// offsets are rebased by the shard's byte prefix.  Same output as RecordBatch::concat of the parts (record_batch.rs:245-342).
static int alloc_ipc(const CoreRef& core, size_t bytes, BufRef* out) {
    const size_t padded = ((bytes + 64 + 255) / 256) * 256;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, padded);
    if (e == cudaErrorMemoryAllocation) {   // cached blocks of the context and of the pool go back to the driver, then once more
        cudaGetLastError();
        core->drop_big();
        cudaStreamSynchronize(core->stream);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, core->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        e = cudaMalloc(&p, padded);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? RVL_OUT_OF_MEMORY : RVL_CUDA, std::string("cudaMalloc(") + std::to_string(padded) + "): " + cudaGetErrorString(e));
    }
    RVL_CUDA_TRY(cudaMemsetAsync(p, 0, padded, core->stream));
    auto b = std::make_shared<DevBuffer>();
    b->ptr = p; b->bytes = padded; b->owned = true; b->kind = 1; b->core = core;
    *out = std::move(b);
    return RVL_OK;
}

static_assert(sizeof(cudaIpcMemHandle_t) == RVL_IPC_HANDLE_BYTES, "rvl_gather_handle layout");

int32_t rvl_gather_dest_create(rvl_ctx* ctx, const int32_t* dtypes, const int32_t* has_validity, const int64_t* data_bytes, int32_t ncols,
                               int64_t total_rows, rvl_batch** dest, rvl_gather_handle* handles) {
    if (!ctx || !dest || !handles || (ncols > 0 && (!dtypes || !has_validity)) || total_rows < 0) return fail(RVL_INVALID_ARGUMENT, "null argument");
    const CoreRef& core = ctx->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    auto b = std::make_unique<rvl_batch>();
    b->core = core; b->num_rows = total_rows;
    std::memset(handles, 0, sizeof(rvl_gather_handle) * (size_t)ncols);
    auto export_buf = [&](const BufRef& buf, uint8_t* h) -> int {
        cudaIpcMemHandle_t ih;
        RVL_CUDA_TRY(cudaIpcGetMemHandle(&ih, buf->ptr));
        std::memcpy(h, &ih, sizeof ih);
        return RVL_OK;
    };
    for (int c = 0; c < ncols; ++c) {
        DevColumn d;
        d.dtype = dtypes[c]; d.length = total_rows; d.offset = 0; d.null_count = -1;
        rvl_gather_handle& h = handles[c];
        if (d.dtype == RVL_INT64 || d.dtype == RVL_FLOAT64) { RVL_TRY(alloc_ipc(core, (size_t)total_rows * 8, &d.values)); RVL_TRY(export_buf(d.values, h.values)); }
        else if (d.dtype == RVL_BOOLEAN) { RVL_TRY(alloc_ipc(core, (size_t)(total_rows + 7) / 8, &d.values)); RVL_TRY(export_buf(d.values, h.values)); }
        else if (d.dtype == RVL_STRING) {
            const int64_t nb = data_bytes ? data_bytes[c] : 0;
            if (nb > (int64_t)INT32_MAX) return fail(RVL_OFFSET_OVERFLOW, "gathered string data (" + std::to_string(nb) + " bytes) exceeds the int32 offset range");
            RVL_TRY(alloc_ipc(core, (size_t)(total_rows + 1) * 4, &d.offsets)); RVL_TRY(export_buf(d.offsets, h.offsets));
            RVL_TRY(alloc_ipc(core, (size_t)nb, &d.data)); RVL_TRY(export_buf(d.data, h.data));
            d.data_len = nb; d.window_bytes = nb;
        } else if (d.dtype != RVL_NULL) return fail(RVL_INVALID_ARGUMENT, "unknown dtype");
        if (has_validity[c] && d.dtype != RVL_NULL) { RVL_TRY(alloc_ipc(core, (size_t)(total_rows + 7) / 8, &d.validity)); RVL_TRY(export_buf(d.validity, h.validity)); h.has_validity = 1; }
        else d.null_count = d.dtype == RVL_NULL ? total_rows : 0;
        h.dtype = d.dtype; h.rows = total_rows; h.data_bytes = d.data_len;
        b->cols.push_back(std::move(d));
    }
    RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));   // zero-filled before any peer writes
    *dest = b.release();
    return RVL_OK;
}

int32_t rvl_gather_dest_open(rvl_ctx* ctx, const rvl_gather_handle* handles, int32_t ncols, rvl_batch** dest_view) {
    if (!ctx || !handles || !dest_view) return fail(RVL_INVALID_ARGUMENT, "null argument");
    const CoreRef& core = ctx->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    auto b = std::make_unique<rvl_batch>();
    b->core = core; b->num_rows = ncols > 0 ? handles[0].rows : 0;
    auto open_buf = [&](const uint8_t* h, size_t bytes, BufRef* out) -> int {
        cudaIpcMemHandle_t ih;
        std::memcpy(&ih, h, sizeof ih);
        void* p = nullptr;
        RVL_CUDA_TRY(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
        auto buf = std::make_shared<DevBuffer>();
        buf->ptr = p; buf->bytes = bytes; buf->owned = true; buf->kind = 2; buf->core = core;
        *out = std::move(buf);
        return RVL_OK;
    };
    for (int c = 0; c < ncols; ++c) {
        const rvl_gather_handle& h = handles[c];
        DevColumn d;
        d.dtype = h.dtype; d.length = h.rows; d.offset = 0; d.null_count = -1;
        if (d.dtype == RVL_INT64 || d.dtype == RVL_FLOAT64) RVL_TRY(open_buf(h.values, (size_t)h.rows * 8, &d.values));
        else if (d.dtype == RVL_BOOLEAN) RVL_TRY(open_buf(h.values, (size_t)(h.rows + 7) / 8, &d.values));
        else if (d.dtype == RVL_STRING) {
            RVL_TRY(open_buf(h.offsets, (size_t)(h.rows + 1) * 4, &d.offsets));
            RVL_TRY(open_buf(h.data, (size_t)h.data_bytes, &d.data));
            d.data_len = h.data_bytes;
        }
        if (h.has_validity) RVL_TRY(open_buf(h.validity, (size_t)(h.rows + 7) / 8, &d.validity));
        b->cols.push_back(std::move(d));
    }
    *dest_view = b.release();
    return RVL_OK;
}

int32_t rvl_gather_push(rvl_ctx* ctx, const rvl_batch* part, rvl_batch* dest, int64_t row_offset, const int64_t* byte_offsets) {
    if (!ctx || !part || !dest) return fail(RVL_INVALID_ARGUMENT, "null argument");
    const CoreRef& core = ctx->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    if (part->cols.size() != dest->cols.size()) return fail(RVL_SCHEMA_MISMATCH, "All batches must have the same schema");
    const int64_t m = part->num_rows;
    if (row_offset < 0 || row_offset + m > dest->num_rows) return fail(RVL_OUT_OF_BOUNDS, "gather: rows [" + std::to_string(row_offset) + ", " + std::to_string(row_offset + m) + ") outside the destination's " + std::to_string(dest->num_rows) + " rows");
    if (part->core.get() != core.get()) {
        // the part's producing kernels ran on its own context's stream
        cudaEvent_t ev = core->take_event();
        if (!ev) return fail(RVL_CUDA, "cudaEventCreate failed");
        cudaSetDevice(part->core->device);
        cudaError_t e = cudaEventRecord(ev, part->core->stream);
        cudaSetDevice(core->device);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(core->stream, ev, 0);
        core->give_event(ev);
        RVL_CUDA_TRY(e);
    }
    for (size_t c = 0; c < part->cols.size(); ++c) {
        const DevColumn& s = part->cols[c];
        DevColumn& d = dest->cols[c];
        if (s.dtype != d.dtype) return fail(RVL_SCHEMA_MISMATCH, "All batches must have the same schema");
        if (m == 0) continue;
        const BitSrc sv = bitsrc_of(s.validity, s.offset, m);
        if (s.validity && !d.validity) return fail(RVL_INVALID_ARGUMENT, "gather: column " + std::to_string(c) + " carries nulls but the destination was created without a validity bitmap");
        if (d.dtype == RVL_INT64 || d.dtype == RVL_FLOAT64) {
            uint64_t* dst = (uint64_t*)d.values->ptr + row_offset;
            const uint64_t* src = (const uint64_t*)s.values->ptr + s.offset;
            if (!s.validity) RVL_CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)m * 8, cudaMemcpyDefault, core->stream));   // copy engine, peer write
            else { copy_col8_zero_nulls_kernel<<<grid_for(m, 256, core->sm_count), 256, 0, core->stream>>>(src, sv, dst, m); core->launches++; }
        } else if (d.dtype == RVL_BOOLEAN) {
            bitcopy_kernel<<<grid_for((m + 31) / 32 + 1, 256, core->sm_count), 256, 0, core->stream>>>(bitsrc_of(s.values, s.offset, m), sv, (uint32_t*)d.values->ptr, (uint64_t)row_offset, m);
            core->launches++;
        } else if (d.dtype == RVL_STRING) {
            if (!byte_offsets) return fail(RVL_INVALID_ARGUMENT, "gather: byte_offsets is required for String columns");
            int32_t fl[2];
            RVL_CUDA_TRY(cudaMemcpyAsync(&fl[0], (const int32_t*)s.offsets->ptr + s.offset, 4, cudaMemcpyDeviceToHost, core->stream));
            RVL_CUDA_TRY(cudaMemcpyAsync(&fl[1], (const int32_t*)s.offsets->ptr + s.offset + m, 4, cudaMemcpyDeviceToHost, core->stream));
            RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
            const int64_t nb = (int64_t)fl[1] - fl[0], bo = byte_offsets[c];
            if (bo < 0 || bo + nb > d.data_len) return fail(RVL_OUT_OF_BOUNDS, "gather: string bytes outside the destination's data buffer");
            rebase_offsets_kernel<<<grid_for(m, 256, core->sm_count), 256, 0, core->stream>>>((const int32_t*)s.offsets->ptr + s.offset, (int32_t*)d.offsets->ptr + row_offset, m, (int32_t)bo);
            core->launches++;
            if (nb > 0) RVL_CUDA_TRY(cudaMemcpyAsync((uint8_t*)d.data->ptr + bo, (const uint8_t*)s.data->ptr + fl[0], (size_t)nb, cudaMemcpyDefault, core->stream));
        }
        if (d.validity) {
            // a part without nulls still owns its bits of the destination bitmap: all ones
            bitcopy_kernel<<<grid_for((m + 31) / 32 + 1, 256, core->sm_count), 256, 0, core->stream>>>(sv, BitSrc{nullptr, 0, 0}, (uint32_t*)d.validity->ptr, (uint64_t)row_offset, m);
            core->launches++;
        }
        RVL_CUDA_TRY(cudaGetLastError());
    }
    RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));   // the rows are in the destination's memory when this returns
    return RVL_OK;
}

int32_t rvl_gather_dest_finish(rvl_ctx* ctx, rvl_batch* dest) {
    if (!ctx || !dest) return fail(RVL_INVALID_ARGUMENT, "null argument");
    RVL_CUDA_TRY(cudaSetDevice(ctx->core->device));
    for (size_t c = 0; c < dest->cols.size(); ++c) {
        DevColumn& d = dest->cols[c];
        d.null_count = d.dtype == RVL_NULL ? d.length : (d.validity ? -1 : 0);
        RVL_TRY(ensure_null_count(dest, (int)c));
        if (d.validity && d.null_count == 0) d.validity.reset();   // bitmap dropped when nothing is null (primitive.rs:180-185)
    }
    return RVL_OK;
}

int32_t rvl_gen_batch(rvl_ctx* ctx, const int32_t* kinds, const uint32_t* col_ids, const uint32_t* null_pct, int32_t ncols,
                      uint64_t row0, int64_t n, rvl_batch** out) {
    if (!ctx || !out || n < 0) return fail(RVL_INVALID_ARGUMENT, "bad argument");
    const CoreRef& core = ctx->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    auto b = std::make_unique<rvl_batch>();
    b->core = core; b->num_rows = n;
    const int g8 = grid_for(n, 256, core->sm_count), gw = grid_for((n + 31) / 32, 256, core->sm_count);
    for (int c = 0; c < ncols; ++c) {
        DevColumn d;
        d.length = n; d.offset = 0;
        const int kind = kinds[c];
        const uint32_t np = null_pct ? null_pct[c] : 0;
        if (kind == RVL_SYNTH_BOOL) {
            d.dtype = RVL_BOOLEAN;
            RVL_TRY(dev_alloc_zeroed(core, (size_t)(n + 7) / 8, &d.values));
            if (n > 0) { gen_bits_kernel<<<gw, 256, 0, core->stream>>>((uint32_t*)d.values->ptr, 0, col_ids[c], 0, row0, n); core->launches++; }
        } else if (kind == RVL_SYNTH_STR) {
            d.dtype = RVL_STRING;
            if (n * 40 > (int64_t)INT32_MAX) return fail(RVL_OFFSET_OVERFLOW, "synthetic string column of " + std::to_string(n) + " rows may exceed the int32 offset range; use <= 50M rows per batch");
            BufRef lens;
            RVL_TRY(dev_alloc(core, (size_t)n * 4, &lens));
            RVL_TRY(dev_alloc(core, (size_t)(n + 1) * 4, &d.offsets));
            RVL_CUDA_TRY(cudaMemsetAsync(d.offsets->ptr, 0, 4, core->stream));
            if (n > 0) {
                gen_strlen_kernel<<<g8, 256, 0, core->stream>>>((int32_t*)lens->ptr, col_ids[c], np, row0, n);
                scan_lengths_kernel<<<1, 1024, 0, core->stream>>>((const int32_t*)lens->ptr, (int32_t*)d.offsets->ptr, n);
                core->launches += 2;
            }
            int32_t total_bytes = 0;
            RVL_CUDA_TRY(cudaMemcpyAsync(&total_bytes, (const int32_t*)d.offsets->ptr + n, 4, cudaMemcpyDeviceToHost, core->stream));
            RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
            d.data_len = total_bytes; d.window_bytes = total_bytes;
            RVL_TRY(dev_alloc(core, (size_t)total_bytes, &d.data));
            if (n > 0) { gen_strbytes_kernel<<<g8, 256, 0, core->stream>>>((uint8_t*)d.data->ptr, (const int32_t*)d.offsets->ptr, col_ids[c], row0, n); core->launches++; }
        } else {
            d.dtype = kind == RVL_SYNTH_F64 ? RVL_FLOAT64 : RVL_INT64;
            RVL_TRY(dev_alloc(core, (size_t)n * 8, &d.values));
            if (n > 0) { gen_col8_kernel<<<g8, 256, 0, core->stream>>>((uint64_t*)d.values->ptr, kind, col_ids[c], row0, n); core->launches++; }
        }
        if (np > 0) {
            RVL_TRY(dev_alloc_zeroed(core, (size_t)(n + 7) / 8, &d.validity));
            if (n > 0) { gen_bits_kernel<<<gw, 256, 0, core->stream>>>((uint32_t*)d.validity->ptr, 1, col_ids[c], np, row0, n); core->launches++; }
        } else d.null_count = 0;
        RVL_CUDA_TRY(cudaGetLastError());
        b->cols.push_back(std::move(d));
    }
    RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
    *out = b.release();
    return RVL_OK;
}

int32_t rvl_batch_checksum(rvl_ctx* ctx, const rvl_batch* batch, int32_t i, uint64_t* checksum) {
    if (!ctx || !batch || !checksum) return fail(RVL_INVALID_ARGUMENT, "null argument");
    if (i < 0 || i >= (int32_t)batch->cols.size()) return fail(RVL_OUT_OF_BOUNDS, "Column index out of bounds");
    const CoreRef& core = batch->core;
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    const DevColumn& c = batch->cols[i];
    const int64_t n = c.length;
    BufRef counter;
    RVL_TRY(dev_alloc_zeroed(core, 8, &counter));
    unsigned long long* out = (unsigned long long*)counter->ptr;
    const BitSrc v = bitsrc_of(c.validity, c.offset, n);
    const int g = grid_for(n, 256, core->sm_count);
    if (n > 0) {
        if (c.dtype == RVL_INT64 || c.dtype == RVL_FLOAT64)
            checksum_col8_kernel<<<g, 256, 0, core->stream>>>((const uint64_t*)c.values->ptr + c.offset, v, n, out);
        else if (c.dtype == RVL_BOOLEAN)
            checksum_bool_kernel<<<g, 256, 0, core->stream>>>(bitsrc_of(c.values, c.offset, n), v, n, out);
        else if (c.dtype == RVL_STRING)
            checksum_str_kernel<<<g, 256, 0, core->stream>>>((const int32_t*)c.offsets->ptr + c.offset, (const uint8_t*)c.data->ptr, v, n, out);
        if (c.dtype != RVL_NULL) { core->launches++; RVL_CUDA_TRY(cudaGetLastError()); }
    }
    RVL_CUDA_TRY(cudaMemcpyAsync(core->mailbox, out, 8, cudaMemcpyDeviceToHost, core->stream));
    RVL_CUDA_TRY(cudaStreamSynchronize(core->stream));
    *checksum = core->mailbox[0];
    return RVL_OK;
}

}  // extern "C"
