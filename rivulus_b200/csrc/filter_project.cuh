// filter_project.cuh — internal interface of the fused operator's host side (filter_project.cu),
// shared with the streaming executor (stream_exec.cu).
#pragma once
#include "runtime.cuh"

namespace rvl {

// An enqueued fused Filter+Select+Limit: output buffers sized for the worst case, device counters,
// and the pinned mailbox slot the counters are copied into when the stream reaches that point.
struct FpPending {
    CoreRef core;
    int64_t n = 0;       // input rows
    int64_t limit = -1;  // global limit (applies to base + local rank)
    int n_launched = 0;
    int n_counters = 0;
    std::vector<DevColumn> outs;
    std::vector<int> validity_counter;  // per projected column: counter index of "ones in compacted validity", or -1
    std::vector<int> bytes_counter;     // per projected column: counter index of "string bytes emitted", or -1
    BufRef counters;                    // device: [0] base + survivors, [1..] per-column counters, then done flag + scratch
    BufRef mask;                        // selection bitmap when requested
    std::vector<BufRef> temps;          // tile descriptors, selection bitmap, literals: released at finish
    uint64_t* mailbox = nullptr;        // pinned host slot; word 0 holds kMailboxPending until the copy lands
    cudaEvent_t done_event = nullptr;   // recorded after the mailbox copy
    uint64_t base_rows = 0;             // filled at finish for chained (streaming) launches
    bool chained = false;               // base_in was used: mailbox[n_counters] carries the base
    cudaStream_t side_stream = nullptr; // borrowed: bit-packed compaction runs here, joined before the mailbox copy
    // the slot and the event go back to the context's pools; a slot is never recycled while a copy into it may be in flight
    bool completed = false;             // fp_finish has waited for the launch
    ~FpPending() {
        if (!core) return;
        if (done_event) { cudaEventSynchronize(done_event); core->give_event(done_event); }
        else if (mailbox && !completed) { cudaSetDevice(core->device); cudaStreamSynchronize(core->stream); }
        core->give_slot(mailbox);
    }
};

constexpr uint64_t kMailboxPending = ~0ull;

// base_in:   device word holding the rows already emitted by earlier launches of the same query (or nullptr)
// total_ext: device word that receives base + survivors after this launch (or nullptr); must differ from base_in
// exact:     (blocking callers only) with the two-pass plan, wait for the predicate scan's survivor count (and the string
//            sizes pass) and allocate the outputs at their exact size instead of the worst case `min(n, limit)` rows
int fp_launch(const CoreRef& core, const rvl_batch* in, const rvl_predicate* pred, const int32_t* proj, int32_t nproj,
              int64_t limit, bool want_mask, const unsigned long long* base_in, unsigned long long* total_ext, FpPending** out,
              bool exact = false);
// one-time per device: opt every kernel instantiation in to its dynamic shared memory size (called by rvl_ctx_create)
int fp_init_device(int device);

// String compaction = a sizes pass (per-tile survivor byte prefixes) + a gather pass (new offsets + byte copy), string_kernels.cuh.
// `sp` carries the inputs; the sizes launch allocates the pass's scratch (kept alive through `keep`) and fills the matching fields
// of `sp`, which the gather launch then reads.  core->string_kernel picks the round-1 pair (1) or the round-2 pair (2).
struct StrGatherParams;
int str_prepare_sizes(const CoreRef& core, StrGatherParams& sp, const DevColumn& src, std::vector<BufRef>* keep);   // scratch, on core->stream
int str_launch_sizes(const CoreRef& core, const StrGatherParams& sp, cudaStream_t stream);
int str_launch_gather(const CoreRef& core, const StrGatherParams& sp);
// waits for the launch, builds the output batch (and/or the mask batch); deletes `pend`
int fp_finish(FpPending* pend, rvl_batch** out, rvl_batch** mask_out);
// device word holding base + survivors after this launch (valid until the pending object is finished)
inline const unsigned long long* fp_total_word(const FpPending* p) { return (const unsigned long long*)p->counters->ptr; }

}  // namespace rvl
