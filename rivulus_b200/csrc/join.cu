// join.cu — inner equi-join.
//
// Replaces (reference, /root/reference/src) PhysicalPlan::HashJoin, physical_plan/plan.rs:174-284:
//   build:  HashMap<AnyValue, Vec<usize>> over the build key column                      :186-193
//   probe:  for every probe row in order, every build row with an equal key, in build order  :198-205
//   result: probe columns taken by probe index, build columns taken by build index        :218-254
// Key equality is AnyValue's (series.rs:73-98): Null == Null, same variant and equal value otherwise; an Int64 never equals a
// Float64; NaN equals nothing.  (0.0 and -0.0 compare equal but hash differently there, so whether they meet depends on the
// HashMap's random state; here they are different keys, as they are in the reference with probability >= 127/128.)
//
// On the device the hash table is replaced by its sorted equivalent, which gives the reference's output order for free:
//   1. every row's key becomes (class, 64 bits): class = Null / Int64 / Float64 / String / Boolean / never-matches (NaN);
//      bits = the value's bit pattern, or a 64-bit hash of a string's bytes
//   2. the build rows are sorted by (class, bits), stably — equal keys stay in ascending row order (two LSD radix sorts, CUB)
//   3. every probe row binary-searches its key's run inside its class segment: run length = its number of matches; an exclusive scan turns the counts into
//      output offsets; a fill kernel writes the (probe row, build row) pairs in probe order, build order within a probe row
//   4. string keys: the pairs are candidates (equal hash); a compare kernel checks the bytes and the survivors are compacted, in order
//   5. the result columns are gathered by the two index lists (take kernels of runtime.cu)
// HBM roofline: the sort dominates (~10 passes over 12 B per build row); the gathers move 8-16 B per output row and column.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include "runtime.cuh"

#include <algorithm>
#include <cstring>

#include "device_utils.cuh"

namespace rvl {

constexpr uint8_t kKeyNull = 0, kKeyInt64 = 1, kKeyFloat64 = 2, kKeyString = 3, kKeyBoolean = 4, kKeyNever = 255;

struct KeyColumn {
    int32_t dtype;             // rvl_dtype of the key column
    const uint64_t* values;    // Int64 / Float64 (row 0 of the view)
    BitSrc bits;               // Boolean values
    BitSrc valid;              // validity (words == nullptr: no nulls)
    BitSrc int_tag;            // mixed Float64 / Int64 series: bit = 1 marks an Int64 row (words == nullptr: none)
    const int32_t* offsets;    // String (entry 0 of the view)
    const uint8_t* data;
};

__device__ __forceinline__ bool bit_at(const BitSrc& b, int64_t row) {
    const uint64_t pos = b.bit0 + (uint64_t)row;
    return (b.words[pos >> 5] >> (pos & 31u)) & 1u;
}

// FNV-1a over the bytes, finished with a 64-bit mixer (only has to spread; equality is verified on the bytes afterwards)
__device__ __forceinline__ uint64_t hash_bytes(const uint8_t* p, int32_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (int32_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h;
}

static __global__ void join_keys_kernel(KeyColumn k, int64_t n, uint8_t* __restrict__ cls, uint64_t* __restrict__ bits, uint32_t* __restrict__ row_ids) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint8_t c = kKeyNull;
    uint64_t b = 0;
    const bool valid = k.dtype != RVL_NULL && (k.valid.words == nullptr || bit_at(k.valid, r));
    if (valid) {
        switch (k.dtype) {
            case RVL_INT64: c = kKeyInt64; b = k.values[r]; break;
            case RVL_FLOAT64: {
                b = k.values[r];
                if (k.int_tag.words != nullptr && bit_at(k.int_tag, r)) c = kKeyInt64;
                else c = ((b & 0x7fffffffffffffffull) > 0x7ff0000000000000ull) ? kKeyNever : kKeyFloat64;   // NaN == nothing (series.rs:92)
                break;
            }
            case RVL_BOOLEAN: c = kKeyBoolean; b = bit_at(k.bits, r) ? 1u : 0u; break;
            case RVL_STRING: c = kKeyString; b = hash_bytes(k.data + k.offsets[r], k.offsets[r + 1] - k.offsets[r]); break;
            default: break;
        }
    }
    cls[r] = c; bits[r] = b;
    if (row_ids != nullptr) row_ids[r] = (uint32_t)r;
}

// cls_out[i] = class of the row that the first sort put at position i
static __global__ void join_gather_cls_kernel(const uint8_t* __restrict__ cls, const uint32_t* __restrict__ rows, int64_t n, uint8_t* __restrict__ cls_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cls_out[i] = cls[rows[i]];
}
static __global__ void join_iota_kernel(uint32_t* p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
// after the second (class) sort: bring the bits and row ids of the first sort into the final order
static __global__ void join_permute_kernel(const uint32_t* __restrict__ pos, const uint64_t* __restrict__ bits_in, const uint32_t* __restrict__ rows_in,
                                           int64_t n, uint64_t* __restrict__ bits_out, uint32_t* __restrict__ rows_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = pos[i];
    bits_out[i] = bits_in[p]; rows_out[i] = rows_in[p];
}

// [start, end) of every key class in the sorted build side (at most six classes occur); both arrays are zeroed before the launch
static __global__ void join_class_bounds_kernel(const uint8_t* __restrict__ scls, int64_t n, uint32_t* __restrict__ seg_start, uint32_t* __restrict__ seg_end) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t c = scls[i];
    if (i == 0 || scls[i - 1] != c) seg_start[c] = (uint32_t)i;
    if (i == n - 1 || scls[i + 1] != c) seg_end[c] = (uint32_t)(i + 1);
}

// run of build rows whose key equals the probe row's: [lo, lo + count) in the sorted order.  One lower-bound search over the 64-bit
// keys of the probe key's class segment (one dependent load per step), then the end of the run by galloping from its first element:
// two loads for a unique key.  (The first version ran a lower- and an upper-bound search over (class, bits) pairs — two loads per
// step, twice — and was 83 % of the whole join: 16.0 of 19.3 ms at 16 M x 64 M rows.)
static __global__ void join_probe_kernel(const uint8_t* __restrict__ pcls, const uint64_t* __restrict__ pbits, int64_t n_probe,
                                         const uint64_t* __restrict__ sbits, const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ seg_end,
                                         uint32_t* __restrict__ lo_out, unsigned long long* __restrict__ count_out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_probe) return;
    const uint8_t c = pcls[r];
    const uint64_t b = pbits[r];
    int64_t lo = seg_start[c];
    const int64_t end = seg_end[c];
    unsigned long long cnt = 0;
    if (c != kKeyNever && lo < end) {
        int64_t hi = end;
        while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (sbits[m] < b) lo = m + 1; else hi = m; }
        if (lo < end && sbits[lo] == b) {
            // gallop: the run ends inside (lo + step / 2, lo + step]
            int64_t step = 1;
            while (lo + step < end && sbits[lo + step] == b) step <<= 1;
            int64_t a = lo + (step >> 1) + 1, z = lo + step < end ? lo + step : end;   // first index that may differ .. first known to differ (or end)
            if (step == 1) a = lo + 1;
            while (a < z) { const int64_t m = (a + z) >> 1; if (sbits[m] == b) a = m + 1; else z = m; }
            cnt = (unsigned long long)(a - lo);
        }
    }
    lo_out[r] = (uint32_t)lo;
    count_out[r] = cnt;
}

static __global__ void join_fill_kernel(const uint32_t* __restrict__ lo, const unsigned long long* __restrict__ count, const unsigned long long* __restrict__ offset,
                                        const uint32_t* __restrict__ srows, int64_t n_probe, int64_t* __restrict__ probe_idx, int64_t* __restrict__ build_idx) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_probe) return;
    const unsigned long long cnt = count[r], o = offset[r];
    const uint32_t first = lo[r];
    for (unsigned long long k = 0; k < cnt; ++k) { probe_idx[o + k] = r; build_idx[o + k] = (int64_t)srows[first + k]; }
}

// string keys: a candidate pair survives when the two strings have the same bytes
static __global__ void join_verify_strings_kernel(KeyColumn probe, KeyColumn build, const int64_t* __restrict__ probe_idx, const int64_t* __restrict__ build_idx,
                                                  int64_t n_pairs, uint8_t* __restrict__ keep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const int64_t p = probe_idx[i], b = build_idx[i];
    bool same = true;
    const bool pv = probe.dtype == RVL_STRING && (probe.valid.words == nullptr || bit_at(probe.valid, p));
    const bool bv = build.dtype == RVL_STRING && (build.valid.words == nullptr || bit_at(build.valid, b));
    if (pv && bv) {   // both are strings (null pairs and non-string classes matched exactly already)
        const int32_t po = probe.offsets[p], pl = probe.offsets[p + 1] - po, bo = build.offsets[b], bl = build.offsets[b + 1] - bo;
        same = pl == bl;
        for (int32_t k = 0; same && k < pl; ++k) same = probe.data[po + k] == build.data[bo + k];
    }
    keep[i] = same ? 1 : 0;
}

static int key_column_of(const rvl_batch* batch, int32_t key, int32_t tag_column, KeyColumn* out) {
    if (key < 0 || key >= (int32_t)batch->cols.size()) return fail(RVL_OUT_OF_BOUNDS, "join key column index out of bounds");
    const DevColumn& c = batch->cols[(size_t)key];
    KeyColumn k{};
    k.dtype = c.dtype;
    k.valid = bitsrc_of(c.validity, c.offset, c.length);
    if (c.dtype == RVL_INT64 || c.dtype == RVL_FLOAT64) k.values = (const uint64_t*)c.values->ptr + c.offset;
    else if (c.dtype == RVL_BOOLEAN) k.bits = bitsrc_of(c.values, c.offset, c.length);
    else if (c.dtype == RVL_STRING) { k.offsets = (const int32_t*)c.offsets->ptr + c.offset; k.data = (const uint8_t*)c.data->ptr; }
    if (tag_column > 0) {
        if (tag_column - 1 >= (int32_t)batch->cols.size() || batch->cols[(size_t)tag_column - 1].dtype != RVL_BOOLEAN)
            return fail(RVL_TYPE_MISMATCH, "join key tag column must be a Boolean column");
        const DevColumn& t = batch->cols[(size_t)tag_column - 1];
        k.int_tag = bitsrc_of(t.values, t.offset, t.length);
    }
    *out = k;
    return RVL_OK;
}

}  // namespace rvl

using namespace rvl;

extern "C" int32_t rvl_hash_join_inner(rvl_ctx* ctx, const rvl_batch* build, int32_t build_key, int32_t build_tag_column, const rvl_batch* probe,
                                       int32_t probe_key, int32_t probe_tag_column, const int32_t* probe_proj, int32_t n_probe_proj,
                                       const int32_t* build_proj, int32_t n_build_proj, rvl_batch** out, int64_t* n_pairs) {
    if (!ctx || !build || !probe || !out || (n_probe_proj > 0 && !probe_proj) || (n_build_proj > 0 && !build_proj)) return fail(RVL_INVALID_ARGUMENT, "null argument");
    const CoreRef& core = ctx->core;
    if (build->core->device != core->device || probe->core->device != core->device) return fail(RVL_INVALID_ARGUMENT, "batch lives on another device than the context");
    RVL_CUDA_TRY(cudaSetDevice(core->device));
    for (int i = 0; i < n_probe_proj; ++i)
        if (probe_proj[i] < 0 || probe_proj[i] >= (int32_t)probe->cols.size()) return fail(RVL_OUT_OF_BOUNDS, "probe column index out of bounds");
    for (int i = 0; i < n_build_proj; ++i)
        if (build_proj[i] < 0 || build_proj[i] >= (int32_t)build->cols.size()) return fail(RVL_OUT_OF_BOUNDS, "build column index out of bounds");
    const int64_t nb = build->num_rows, np = probe->num_rows;
    if (nb >= (1ll << 32) || np >= (1ll << 32)) return fail(RVL_INVALID_ARGUMENT, "join sides are limited to 2^32 - 1 rows each");
    KeyColumn bk, pk;
    RVL_TRY(key_column_of(build, build_key, build_tag_column, &bk));
    RVL_TRY(key_column_of(probe, probe_key, probe_tag_column, &pk));
    cudaStream_t st = core->stream;
    auto grid = [](int64_t n) { return (unsigned)std::max<int64_t>(1, (n + 255) / 256); };

    auto res = std::make_unique<rvl_batch>();
    res->core = core;
    int64_t total = 0;
    BufRef probe_idx, build_idx;
    if (nb > 0 && np > 0) {
        // ---- 1. keys
        BufRef bcls, bbits, brows, pcls, pbits;
        RVL_TRY(dev_alloc(core, (size_t)nb, &bcls)); RVL_TRY(dev_alloc(core, (size_t)nb * 8, &bbits)); RVL_TRY(dev_alloc(core, (size_t)nb * 4, &brows));
        RVL_TRY(dev_alloc(core, (size_t)np, &pcls)); RVL_TRY(dev_alloc(core, (size_t)np * 8, &pbits));
        join_keys_kernel<<<grid(nb), 256, 0, st>>>(bk, nb, (uint8_t*)bcls->ptr, (uint64_t*)bbits->ptr, (uint32_t*)brows->ptr);
        join_keys_kernel<<<grid(np), 256, 0, st>>>(pk, np, (uint8_t*)pcls->ptr, (uint64_t*)pbits->ptr, nullptr);
        core->launches += 2;
        RVL_CUDA_TRY(cudaGetLastError());
        // ---- 2. build side sorted by (class, bits), stable: LSD = bits first, then class
        BufRef s1bits, s1rows, pos_in, pos_out, s2cls, sbits, srows, tmp;
        RVL_TRY(dev_alloc(core, (size_t)nb * 8, &s1bits)); RVL_TRY(dev_alloc(core, (size_t)nb * 4, &s1rows));
        RVL_TRY(dev_alloc(core, (size_t)nb * 4, &pos_in)); RVL_TRY(dev_alloc(core, (size_t)nb * 4, &pos_out));
        RVL_TRY(dev_alloc(core, (size_t)nb, &s2cls)); RVL_TRY(dev_alloc(core, (size_t)nb * 8, &sbits)); RVL_TRY(dev_alloc(core, (size_t)nb * 4, &srows));
        size_t t1 = 0, t2 = 0, t3 = 0;
        RVL_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t1, (const uint64_t*)bbits->ptr, (uint64_t*)s1bits->ptr, (const uint32_t*)brows->ptr, (uint32_t*)s1rows->ptr, nb, 0, 64, st));
        RVL_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t2, (const uint8_t*)bcls->ptr, (uint8_t*)s2cls->ptr, (const uint32_t*)pos_in->ptr, (uint32_t*)pos_out->ptr, nb, 0, 8, st));
        RVL_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, t3, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, np, st));
        RVL_TRY(dev_alloc(core, std::max({t1, t2, t3, (size_t)16}), &tmp));
        size_t tb = tmp->bytes;
        RVL_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp->ptr, tb, (const uint64_t*)bbits->ptr, (uint64_t*)s1bits->ptr, (const uint32_t*)brows->ptr, (uint32_t*)s1rows->ptr, nb, 0, 64, st));
        // classes in the order of the first sort, then sorted themselves (stable), carrying the position inside the first sort
        BufRef cls1;
        RVL_TRY(dev_alloc(core, (size_t)nb, &cls1));
        join_gather_cls_kernel<<<grid(nb), 256, 0, st>>>((const uint8_t*)bcls->ptr, (const uint32_t*)s1rows->ptr, nb, (uint8_t*)cls1->ptr);
        join_iota_kernel<<<grid(nb), 256, 0, st>>>((uint32_t*)pos_in->ptr, nb);
        tb = tmp->bytes;
        RVL_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp->ptr, tb, (const uint8_t*)cls1->ptr, (uint8_t*)s2cls->ptr, (const uint32_t*)pos_in->ptr, (uint32_t*)pos_out->ptr, nb, 0, 8, st));
        join_permute_kernel<<<grid(nb), 256, 0, st>>>((const uint32_t*)pos_out->ptr, (const uint64_t*)s1bits->ptr, (const uint32_t*)s1rows->ptr, nb,
                                                      (uint64_t*)sbits->ptr, (uint32_t*)srows->ptr);
        core->launches += 5;
        RVL_CUDA_TRY(cudaGetLastError());
        // ---- 3. probe: run per probe row, exclusive scan of the run lengths, pairs
        BufRef lo, cnt, off, seg;
        RVL_TRY(dev_alloc(core, (size_t)np * 4, &lo)); RVL_TRY(dev_alloc(core, (size_t)np * 8, &cnt)); RVL_TRY(dev_alloc(core, (size_t)np * 8, &off));
        RVL_TRY(dev_alloc_zeroed(core, 2 * 256 * 4, &seg));
        uint32_t* const seg_start = (uint32_t*)seg->ptr;
        uint32_t* const seg_end = seg_start + 256;
        join_class_bounds_kernel<<<grid(nb), 256, 0, st>>>((const uint8_t*)s2cls->ptr, nb, seg_start, seg_end);
        join_probe_kernel<<<grid(np), 256, 0, st>>>((const uint8_t*)pcls->ptr, (const uint64_t*)pbits->ptr, np, (const uint64_t*)sbits->ptr, seg_start, seg_end,
                                                    (uint32_t*)lo->ptr, (unsigned long long*)cnt->ptr);
        core->launches++;
        tb = tmp->bytes;
        RVL_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp->ptr, tb, (const unsigned long long*)cnt->ptr, (unsigned long long*)off->ptr, np, st));
        core->launches += 2;
        unsigned long long last_off = 0, last_cnt = 0;
        RVL_CUDA_TRY(cudaMemcpyAsync(&last_off, (const unsigned long long*)off->ptr + (np - 1), 8, cudaMemcpyDeviceToHost, st));
        RVL_CUDA_TRY(cudaMemcpyAsync(&last_cnt, (const unsigned long long*)cnt->ptr + (np - 1), 8, cudaMemcpyDeviceToHost, st));
        RVL_CUDA_TRY(cudaStreamSynchronize(st));
        total = (int64_t)(last_off + last_cnt);
        RVL_TRY(dev_alloc(core, (size_t)std::max<int64_t>(total, 1) * 8, &probe_idx));
        RVL_TRY(dev_alloc(core, (size_t)std::max<int64_t>(total, 1) * 8, &build_idx));
        if (total > 0) {
            join_fill_kernel<<<grid(np), 256, 0, st>>>((const uint32_t*)lo->ptr, (const unsigned long long*)cnt->ptr, (const unsigned long long*)off->ptr,
                                                       (const uint32_t*)srows->ptr, np, (int64_t*)probe_idx->ptr, (int64_t*)build_idx->ptr);
            core->launches++;
            RVL_CUDA_TRY(cudaGetLastError());
        }
        // ---- 4. string keys: equal hashes are only candidates
        if (total > 0 && bk.dtype == RVL_STRING && pk.dtype == RVL_STRING) {
            BufRef keep, p2, b2, nsel, tmp2;
            RVL_TRY(dev_alloc(core, (size_t)total, &keep)); RVL_TRY(dev_alloc(core, (size_t)total * 8, &p2)); RVL_TRY(dev_alloc(core, (size_t)total * 8, &b2));
            RVL_TRY(dev_alloc(core, 8, &nsel));
            join_verify_strings_kernel<<<grid(total), 256, 0, st>>>(pk, bk, (const int64_t*)probe_idx->ptr, (const int64_t*)build_idx->ptr, total, (uint8_t*)keep->ptr);
            core->launches++;
            size_t t4 = 0;
            RVL_CUDA_TRY(cub::DeviceSelect::Flagged(nullptr, t4, (const int64_t*)probe_idx->ptr, (const uint8_t*)keep->ptr, (int64_t*)p2->ptr, (long long*)nsel->ptr, total, st));
            RVL_TRY(dev_alloc(core, std::max<size_t>(t4, 16), &tmp2));
            size_t tb2 = tmp2->bytes;
            RVL_CUDA_TRY(cub::DeviceSelect::Flagged(tmp2->ptr, tb2, (const int64_t*)probe_idx->ptr, (const uint8_t*)keep->ptr, (int64_t*)p2->ptr, (long long*)nsel->ptr, total, st));
            tb2 = tmp2->bytes;
            RVL_CUDA_TRY(cub::DeviceSelect::Flagged(tmp2->ptr, tb2, (const int64_t*)build_idx->ptr, (const uint8_t*)keep->ptr, (int64_t*)b2->ptr, (long long*)nsel->ptr, total, st));
            core->launches += 2;
            long long kept = 0;
            RVL_CUDA_TRY(cudaMemcpyAsync(&kept, nsel->ptr, 8, cudaMemcpyDeviceToHost, st));
            RVL_CUDA_TRY(cudaStreamSynchronize(st));
            total = kept; probe_idx = p2; build_idx = b2;
        }
    }
    if (!probe_idx) { RVL_TRY(dev_alloc(core, 8, &probe_idx)); RVL_TRY(dev_alloc(core, 8, &build_idx)); }
    // ---- 5. materialize_join_result (:208-254): probe columns first, then the build columns
    res->num_rows = total;
    RVL_TRY(rvl_internal_take_rows(core, probe, probe_proj, n_probe_proj, (const int64_t*)probe_idx->ptr, total, res.get()));
    RVL_TRY(rvl_internal_take_rows(core, build, build_proj, n_build_proj, (const int64_t*)build_idx->ptr, total, res.get()));
    if (n_pairs) *n_pairs = total;
    *out = res.release();
    return RVL_OK;
}
