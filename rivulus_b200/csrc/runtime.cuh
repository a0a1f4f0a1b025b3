// runtime.cuh — host-side runtime objects behind the C ABI (include/rivulus_gpu.h): context, stream-ordered
// device memory, device-resident RecordBatch columns.  Internal to librivulus_gpu.so.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <memory>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rivulus_gpu.h"
#include "device_utils.cuh"

namespace rvl {

extern thread_local std::string g_last_error;
int fail(int code, const std::string& msg);

#define RVL_CUDA_TRY(expr)                                                                                   \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) return ::rvl::fail(RVL_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define RVL_TRY(expr)                 \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != RVL_OK) return rc__; \
    } while (0)

// One per rvl_ctx; shared by every buffer allocated from it so frees stay stream-ordered and valid.
struct CtxCore {
    int device = 0;
    cudaStream_t stream = nullptr;   // compute stream: every kernel of the context
    cudaStream_t copy_stream = nullptr;  // H2D staging of the streaming executor
    int sm_count = 148;
    std::atomic<int64_t> launches{0};
    uint64_t* mailbox = nullptr;     // pinned host words the synchronous entry points (counts, checksums, concat) read back into
    size_t mailbox_words = 0;
    static constexpr size_t kSlotBase = 256;   // words of `mailbox` those synchronous readbacks may use
    // Every in-flight operator invocation (FpPending) owns one pinned kSlotWords-word slot from a free list until it is
    // finished; the list grows by whole pinned chunks, so any number of launches may be outstanding without two of them
    // ever sharing a slot (a fixed ring aliased once more than its size were in flight).
    static constexpr size_t kSlotWords = 64, kSlotsPerChunk = 64;
    std::mutex mu;                             // guards slot_free / slot_chunks / event_pool / prof_events / big_free
    // Blocks of at least kBigBlock bytes are recycled by the context itself instead of going back to the driver's pool: the pool
    // re-creates multi-GB blocks after a synchronisation under some free-list layouts (measured: one cudaMallocAsync(8 GB) of a
    // steady-state query taking 26-540 ms, ~5.5 ms per GB, profiles/r02_alloc_outlier.txt).  Every dev_alloc is ordered on
    // `stream`, and so is every release, so handing a released block to the next dev_alloc keeps the stream order the pool gave.
    static constexpr size_t kBigBlock = 16u << 20;
    std::multimap<size_t, void*> big_free;     // size -> block
    size_t big_free_bytes = 0;
    // page-locked staging buffers of the streaming executor outlive their stream the same way: cudaHostAlloc / cudaFreeHost take
    // milliseconds (and the process-wide mmap lock, against every other thread touching fresh memory), a query should not pay them
    std::multimap<size_t, void*> pinned_free;
    std::map<void*, size_t> pinned_sizes;      // every block handed out by take_pinned
    void* take_pinned(size_t bytes, size_t* got);              // cached block in [bytes, 2 * bytes + 1 MiB] or a new one; nullptr = out of memory
    void give_pinned(void* p);
    void drop_pinned();                                        // cudaFreeHost every cached block
    void* take_big(size_t bytes, size_t* got);                 // smallest cached block in [bytes, bytes * 9/8], or nullptr
    void give_big(void* p, size_t bytes);
    void drop_big();                                           // cudaFreeAsync every cached block (trim, OOM retry, teardown)
    std::vector<uint64_t*> slot_free;
    std::vector<uint64_t*> slot_chunks;
    uint64_t* take_slot();
    void give_slot(uint64_t* s) { if (s) { std::lock_guard<std::mutex> g(mu); slot_free.push_back(s); } }
    cudaStream_t d2h_stream = nullptr;   // downloads of finished batches (overlaps H2D staging and kernels)
    cudaStream_t side_stream = nullptr;  // forked from `stream` for the bit-packed compaction kernel (runs under the HBM-bound ones)
    bool bits_overlap = true;            // RVL_OPT_BITS_OVERLAP
    int string_kernel = 3;               // RVL_OPT_STRING_KERNEL: 1 = round-1 pair; 2 = ranges sizes pass + persistent TMA-staged gather;
                                         //   3 = ranges sizes pass + round-1 gather (default: fastest measured); 4 = 3 with the OR-merging copy
    int string_dense_min = 128;          // RVL_OPT_STRING_DENSE_MIN: survivors per 1024-row sub-tile from which the source block is TMA-staged
    int chunk_plan = 1;                  // RVL_OPT_CHUNK_PLAN: predicate column projected -> single-pass chunk kernel (chunk_kernels.cuh):
                                         //   0 never, 1 when the sampled / observed selectivity is at least chunk_min_sel, 2 always
    double chunk_min_sel = 0.30;         // measured crossover against the two-pass plan (profiles/r02_chunk_plan_sweep.txt)
    bool sync_ok = false;                // set by blocking entry points around fp_launch: a host round trip (sampling) is acceptable
    double sel_hint = -1.0;              // set by the streaming executor: selectivity of the most recent batch, or < 0
    int exact_alloc = 2;                 // RVL_OPT_EXACT_ALLOC: 0 never, 1 always, 2 when the worst case exceeds a quarter of device memory
    size_t device_bytes = 0;
    // optional kernel-level timing of the fused kernel (rvl_ctx_profile_*)
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    double prof_ms = 0.0;
    int64_t prof_launches = 0;
    std::vector<double> prof_times;
    // execution plan of the fused operator (rvl_ctx_set_option)
    int plan_mode = 0;                      // RVL_PLAN_AUTO / _FUSED / _TWO_PASS
    int64_t two_pass_min_rows = 2 << 20;    // AUTO: batches at least this large take the two-pass plan (measured crossover: equal at 2 Mi rows, 1.3x at 4 Mi)
    int sparse_max = 384;                   // two-pass: tiles with <= this many survivors (of 2048 rows) are gathered (64-byte-granule loads;
                                            //   measured crossover ~20 % with them, profiles/r02_sparse_max_sweep.txt; 11 % without)
    int dense_slots = 14;                   // two-pass: 16 KB ring slots per CTA of the dense compaction kernel
    int dense_ctas_per_sm = 1;
    int dense_warps = 16;                   // two-pass: consumer warps per CTA of the dense kernel (8 or 16)
    int scan_warps = 8;                     // two-pass: warps per CTA of the predicate scan (8, 16 or 32).  Round-2 sweep over warps x slot
                                            //   rows x ring depth (profiles/r02_scan_sweep.txt): 8 x 1024 x 2 = 6.70 TB/s, 16 x 512 x 2 (the
                                            //   round-1 default) 6.31; more than ~128 KB in flight per SM is slower in every shape
    int scan_slots = 2;                     // two-pass: ring slots per warp of the predicate scan (1..16, capped by shared memory)
    int scan_l2_hints = 0;                  // two-pass: bit 0 = predicate column evict_first, bit 1 = selection words evict_last
    int scan_item_rows = 0;                 // two-pass: rows per ring slot (0 = 8192 / warps; 512 or 256 with 8 warps, 256 with 16)
    // pooled timing-less events (one per in-flight operator invocation): create / destroy per batch costs microseconds
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t take_event() {
        {
            std::lock_guard<std::mutex> g(mu);
            if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
        }
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return e;
    }
    void give_event(cudaEvent_t e) { if (e) { std::lock_guard<std::mutex> g(mu); event_pool.push_back(e); } }
    int prof_flush();
    ~CtxCore();
};
using CoreRef = std::shared_ptr<CtxCore>;

struct DevBuffer {
    void* ptr = nullptr;
    size_t bytes = 0;
    bool owned = true;      // freed by the destructor (how: `kind`)
    int kind = 0;           // 0 = stream-ordered pool allocation (cudaFreeAsync), 1 = cudaMalloc (exportable over CUDA IPC: cudaFree),
                            // 2 = another process's allocation mapped through CUDA IPC (cudaIpcCloseMemHandle),
                            // 3 = pool allocation of >= CtxCore::kBigBlock bytes, recycled through CtxCore::big_free
    CoreRef core;
    ~DevBuffer();
};
using BufRef = std::shared_ptr<DevBuffer>;

// stream-ordered allocation from the device's default pool (release threshold = keep everything cached)
int dev_alloc(const CoreRef& core, size_t bytes, BufRef* out);
int dev_alloc_zeroed(const CoreRef& core, size_t bytes, BufRef* out);
BufRef wrap_external(const CoreRef& core, const void* ptr, size_t bytes);

struct DevColumn {
    int32_t dtype = RVL_NULL;
    int64_t length = 0;
    int64_t offset = 0;
    BufRef values;    // 8-byte values or Boolean value bitmap
    BufRef validity;  // optional
    BufRef offsets;   // String
    BufRef data;      // String
    int64_t data_len = 0;
    int64_t window_bytes = -1;  // String: bytes referenced by the viewed rows (offsets[offset+length] - offsets[offset]) when the host
                                //   knows it, else -1 (then data_len bounds it); sizes the output of a filter over this view
    int64_t null_count = -1;  // -1 = not computed yet
};

BitSrc bitsrc_of(const BufRef& buf, int64_t offset, int64_t length);

}  // namespace rvl

struct rvl_ctx {
    rvl::CoreRef core;
};
struct rvl_batch {
    rvl::CoreRef core;
    int64_t num_rows = 0;
    std::vector<rvl::DevColumn> cols;
};

// runtime.cu: take_array (record_batch.rs:131-178) of the listed columns (all when cols == nullptr) by a DEVICE index list, appended to
// `res`; synchronises the stream.  Internal (join.cu), not part of the ABI header.
extern "C" int rvl_internal_take_rows(const rvl::CoreRef& core, const rvl_batch* batch, const int32_t* cols, int32_t ncols, const int64_t* idx,
                                      int64_t n, rvl_batch* res);
