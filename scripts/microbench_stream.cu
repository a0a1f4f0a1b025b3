// Microbenchmark: how fast can a B200 stream one 8-byte column with (a) TMA 1-D bulk copies into shared memory in the
// CTA shape the fused kernel uses, (b) plain 128-bit loads.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int SUB, int THREADS>
__global__ void __launch_bounds__(THREADS) tma_stream(const uint64_t* __restrict__ in, int64_t n_rows, unsigned long long* out, int pad_smem) {
    extern __shared__ __align__(128) unsigned char raw[];
    uint64_t* buf = reinterpret_cast<uint64_t*>(raw);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(raw + SUB * 16384);
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * SUB * 2048;
    if (tid == 0) {
        for (int s = 0; s < SUB; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar + s)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < SUB; ++s) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar + s)), "r"(16384) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf + s * 2048)),
                         "l"(in + row0 + s * 2048), "r"(16384), "r"(smem_u32(mbar + s)) : "memory");
        }
    }
    __syncthreads();
    uint64_t acc = 0;
    for (int s = 0; s < SUB; ++s) {
        uint32_t ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(mbar + s)) : "memory");
        for (int i = tid; i < 2048; i += THREADS) acc += buf[s * 2048 + i] > 998 ? 1 : 0;
    }
    if (acc == 0x7fffffffffffull) out[0] = acc;
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
// persistent per-warp TMA rings (the predicate_scan_kernel structure): each warp owns ITEM-row slots, D deep.
// MODE 0: every warp streams its own contiguous range; MODE 1: items interleaved over all warps of the grid
// (the whole GPU reads one moving window); MODE 2: items interleaved over the CTAs, warps of a CTA take consecutive items
template <int W, int D, int MODE>
__global__ void __launch_bounds__(W * 32, 1) persist_stream(const uint64_t* __restrict__ in, int64_t n_rows, unsigned long long* out) {
    constexpr int ITEM = 8192 / W;
    extern __shared__ __align__(128) unsigned char raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* ring = reinterpret_cast<uint64_t*>(raw) + (size_t)warp * D * ITEM;
    uint64_t* full = reinterpret_cast<uint64_t*>(raw + (size_t)W * D * ITEM * 8) + warp * D;
    const int64_t n_items = n_rows / ITEM;
    const int64_t G = (int64_t)gridDim.x * W, g = (int64_t)blockIdx.x * W + warp;
    const int64_t per = (n_items + G - 1) / G;
    auto item_of = [&](int64_t j) -> int64_t {
        if (MODE == 0) return g * per + j;
        if (MODE == 1) return j * G + g;
        return (j * gridDim.x + blockIdx.x) * W + warp;
    };
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
template <typename F> float time_ms(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < reps; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}
int main() {
    const int64_t n = 1ll << 30;  // 8 GiB column
    uint64_t* d; unsigned long long* o;
    cudaMalloc(&d, n * 8); cudaMalloc(&o, 8); cudaMemset(d, 1, n * 8);
    auto report = [&](const char* name, float ms) { printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms, n * 8 / ms / 1e6); };
    {
        auto k = tma_stream<4, 256>; int smem = 4 * 16384 + 64;
        for (int extra : {0, 6000, 40000}) {  // extra smem lowers CTAs/SM: 3 -> 3 -> 2
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem + extra);
            int nb; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 256, smem + extra);
            char nm[96]; snprintf(nm, 96, "tma SUB=4 (64KB/CTA) %d CTA/SM", nb);
            report(nm, time_ms([&] { k<<<(unsigned)(n / 8192), 256, smem + extra>>>(d, n, o, 0); }, 5));
        }
    }
    {
        auto k = tma_stream<2, 256>; int smem = 2 * 16384 + 64;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int nb; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 256, smem);
        char nm[96]; snprintf(nm, 96, "tma SUB=2 (32KB/CTA) %d CTA/SM", nb);
        report(nm, time_ms([&] { k<<<(unsigned)(n / 4096), 256, smem>>>(d, n, o, 0); }, 5));
    }
    {
        auto k = tma_stream<1, 256>; int smem = 16384 + 64;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int nb; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 256, smem);
        char nm[96]; snprintf(nm, 96, "tma SUB=1 (16KB/CTA) %d CTA/SM", nb);
        report(nm, time_ms([&] { k<<<(unsigned)(n / 2048), 256, smem>>>(d, n, o, 0); }, 5));
    }
    {
        auto k = tma_stream<8, 256>; int smem = 8 * 16384 + 64;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int nb; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 256, smem);
        char nm[96]; snprintf(nm, 96, "tma SUB=8 (128KB/CTA) %d CTA/SM", nb);
        report(nm, time_ms([&] { k<<<(unsigned)(n / 16384), 256, smem>>>(d, n, o, 0); }, 5));
    }
#define PERSIST(W, D, MODE)                                                                                              \
    {                                                                                                                    \
        auto k = persist_stream<W, D, MODE>;                                                                             \
        int smem = W * D * (8192 / W) * 8 + W * D * 8;                                                                   \
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                                      \
        char nm[96]; snprintf(nm, 96, "persistent W=%d D=%d mode=%d (1 CTA/SM)", W, D, MODE);                          \
        report(nm, time_ms([&] { k<<<148, W * 32, smem>>>(d, n, o); }, 5));                                              \
    }
    PERSIST(8, 3, 0) PERSIST(8, 3, 1) PERSIST(8, 3, 2) PERSIST(16, 3, 0) PERSIST(16, 3, 1) PERSIST(16, 3, 2) PERSIST(8, 2, 1) PERSIST(16, 2, 1)
    for (int bps : {2, 4, 8}) {
        char nm[96]; snprintf(nm, 96, "ldg.128 x4 grid-stride, %d CTA/SM x 256 thr", bps);
        report(nm, time_ms([&] { ldg_stream<<<148 * bps, 256>>>((const ulonglong2*)d, n / 2, o); }, 5));
    }
    report("cudaMemsetAsync 8 GiB (write only)", time_ms([&] { cudaMemsetAsync(d, 2, n * 8); }, 3));
    return 0;
}
