// Microbenchmark: reading a pinned HOST column from a B200 over PCIe — (a) copy engine (cudaMemcpyAsync), (b) TMA 1-D bulk copies
// issued by SMs straight from the host pointer (the predicate_scan_kernel ring structure), (c) plain 128-bit loads, (d) sparse
// 8-byte gathers (one row in every K) — to decide how the streaming executor should bring projected columns across.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_pcie microbench_pcie.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int W, int D>
__global__ void __launch_bounds__(W * 32, 1) persist_stream(const uint64_t* __restrict__ in, int64_t n_rows, unsigned long long* out) {
    constexpr int ITEM = 8192 / W;
    extern __shared__ __align__(128) unsigned char raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* ring = reinterpret_cast<uint64_t*>(raw) + (size_t)warp * D * ITEM;
    uint64_t* full = reinterpret_cast<uint64_t*>(raw + (size_t)W * D * ITEM * 8) + warp * D;
    const int64_t n_items = n_rows / ITEM;
    const int64_t G = (int64_t)gridDim.x * W, g = (int64_t)blockIdx.x * W + warp;
    const int64_t per = (n_items + G - 1) / G;
    auto item_of = [&](int64_t j) -> int64_t { return j * G + g; };
    auto issue = [&](int64_t j, int slot) {
        const int64_t it = item_of(j);
        if (j < per && it < n_items) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full + slot)), "r"(ITEM * 8) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + slot * ITEM)),
                         "l"(in + it * ITEM), "r"(ITEM * 8), "r"(smem_u32(full + slot)) : "memory");
        }
    };
    if (lane == 0) {
        for (int s = 0; s < D; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(full + s)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < D; ++s) issue(s, s);
    }
    __syncwarp();
    uint32_t acc = 0;
    int slot = 0; uint32_t phase = 0;
    for (int64_t j = 0; j < per; ++j) {
        if (item_of(j) >= n_items) break;
        uint32_t ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(full + slot)), "r"(phase) : "memory");
#pragma unroll
        for (int w = 0; w < ITEM / 32; ++w) acc += __popc(__ballot_sync(0xFFFFFFFFu, ring[slot * ITEM + w * 32 + lane] > 998));
        __syncwarp();
        if (lane == 0) issue(j + D, slot);
        if (++slot == D) { slot = 0; phase ^= 1u; }
    }
    if (acc == 0x7fffffffu) out[0] = acc;
}
__global__ void ldg_stream(const ulonglong2* __restrict__ in, int64_t n_vec, unsigned long long* out) {
    uint64_t acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride * 4) {
        ulonglong2 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i + k * stride < n_vec) ? __ldg(in + i + k * stride) : make_ulonglong2(0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += (v[k].x > 998) + (v[k].y > 998);
    }
    if (acc == 0x7fffffffffffull) out[0] = acc;
}
// one 8-byte read in every `every` rows (pseudo-random position inside each stride), 4 loads in flight per thread
__global__ void sparse_gather(const uint64_t* __restrict__ in, int64_t n_rows, int64_t every, uint64_t* __restrict__ dst) {
    const int64_t n_pick = n_rows / every;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pick; i += stride * 4) {
        uint64_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t q = i + k * stride;
            const int64_t row = q * every + (int64_t)((uint64_t)(q * 2654435761ull) % (uint64_t)every);
            v[k] = q < n_pick ? __ldg(in + row) : 0ull;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i + k * stride < n_pick) dst[i + k * stride] = v[k];
    }
}
template <typename F> float time_ms(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < reps; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}
int main() {
    const int64_t n = 1ll << 28;  // 2 GiB column in pinned host memory
    uint64_t *h, *d, *dst; unsigned long long* o;
    if (cudaHostAlloc(&h, n * 8, cudaHostAllocDefault) != cudaSuccess) { printf("cudaHostAlloc failed\n"); return 1; }
    for (int64_t i = 0; i < n; ++i) h[i] = (uint64_t)i * 2654435761ull % 1000;
    cudaMalloc(&d, n * 8); cudaMalloc(&o, 8); cudaMalloc(&dst, n * 8 / 4);
    auto report = [&](const char* name, float ms, double bytes) { printf("%-64s %9.3f ms  %8.2f GB/s\n", name, ms, bytes / ms / 1e6); fflush(stdout); };
    report("cudaMemcpyAsync H2D (copy engine), 2 GiB", time_ms([&] { cudaMemcpyAsync(d, h, n * 8, cudaMemcpyHostToDevice); }, 3), n * 8.0);
    report("cudaMemcpyAsync H2D, 128 MiB pieces", time_ms([&] { for (int64_t i = 0; i < 16; ++i) cudaMemcpyAsync(d + i * (n / 16), h + i * (n / 16), n / 2, cudaMemcpyHostToDevice); }, 3), n * 8.0);
#define PERSIST(W, D, CTAS)                                                                                              \
    {                                                                                                                    \
        auto k = persist_stream<W, D>;                                                                                   \
        int smem = W * D * (8192 / W) * 8 + W * D * 8;                                                                   \
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                                      \
        char nm[96]; snprintf(nm, 96, "TMA bulk from host pointer: %d CTAs x %d warps, ring depth %d", CTAS, W, D);      \
        report(nm, time_ms([&] { k<<<CTAS, W * 32, smem>>>(h, n, o); }, 3), n * 8.0);                                   \
    }
    PERSIST(16, 3, 148) PERSIST(8, 3, 148) PERSIST(16, 3, 64) PERSIST(16, 3, 16) PERSIST(16, 1, 148)
    for (int bps : {1, 4}) {
        char nm[96]; snprintf(nm, 96, "ldg.128 x4 from host pointer, %d CTA/SM x 256 thr", bps);
        report(nm, time_ms([&] { ldg_stream<<<148 * bps, 256>>>((const ulonglong2*)h, n / 2, o); }, 3), n * 8.0);
    }
    for (int64_t every : {1000, 100, 30, 10, 4}) {
        char nm[128];
        const float ms = time_ms([&] { sparse_gather<<<148 * 8, 256>>>(h, n, every, dst); }, 3);
        snprintf(nm, 128, "sparse 8-byte gather from host, 1 row in %lld (%.1f M reads, %.1f M reads/s)", (long long)every, n / every / 1e6, n / every / ms / 1e3);
        report(nm, ms, (double)(n / every) * 32.0);  // counted as 32-byte sectors
    }
    // duplex: TMA reads from host while the copy engine writes results back
    {
        uint64_t* h2; cudaHostAlloc(&h2, n * 8 / 2, cudaHostAllocDefault);
        cudaStream_t s2; cudaStreamCreate(&s2);
        auto k = persist_stream<16, 3>; int smem = 16 * 3 * 512 * 8 + 16 * 3 * 8;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const float ms = time_ms([&] { cudaMemcpyAsync(h2, d, n * 8 / 2, cudaMemcpyDeviceToHost, s2); k<<<148, 512, smem>>>(h, n, o); cudaStreamSynchronize(s2); }, 3);
        report("duplex: TMA from host (2 GiB) + D2H copy engine (1 GiB)", ms, n * 8.0);
        const float ms2 = time_ms([&] { cudaMemcpyAsync(h2, d, n * 8 / 2, cudaMemcpyDeviceToHost, s2); cudaMemcpyAsync(d, h, n * 8, cudaMemcpyHostToDevice); cudaStreamSynchronize(s2); }, 3);
        report("duplex: H2D copy engine (2 GiB) + D2H copy engine (1 GiB)", ms2, n * 8.0);
    }
    return 0;
}
