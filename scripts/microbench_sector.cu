// microbench_sector.cu — what does B200 DRAM fetch when a kernel touches ONE 32-byte sector per N bytes?
//
// the 10 % query of configs[1] moves 1.71x its algorithmic bytes; DESIGN.md attributed that to 128-byte DRAM
// lines.  This isolates it.  An 8 GB column; every thread reads one 8-byte word from ONE 32-byte sector of each `stride`-byte window
// (window-local position random, or fixed at 0), stride 32 / 64 / 128 / 256 / 512 B, with four load flavours
//   0  ld.global.nc                      1  ld.global.nc.L1::no_allocate
//   2  ld.global.nc.L2::64B (prefetch)   3  ld.global.nc.L2::128B
// and with cudaLimitMaxL2FetchGranularity left at its default or set to 32 (argv[1]).  The program prints the time and the USEFUL
// sector bytes per launch; DRAM bytes per launch come from running it under
//   ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum --csv
// (launch order = print order).  If DRAM moved 32-byte sectors, dram__bytes_read = 8 GB * 32 / stride; with 64-byte (128-byte)
// granules it stays at 8 GB * 64 (128) / stride until the stride exceeds the granule.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/microbench_sector scripts/microbench_sector.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int LD> __device__ __forceinline__ uint64_t ld8(const uint64_t* p) {
    uint64_t v;
    if (LD == 0) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (LD == 1) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (LD == 2) asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else asm volatile("ld.global.nc.L2::128B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// window w covers bytes [w * stride, (w + 1) * stride); the thread reads word 0 of sector (random ? hash(w) % (stride / 32) : 0)
template <int LD>
__global__ void __launch_bounds__(256) sector_read_kernel(const uint64_t* __restrict__ col, uint64_t n_windows, uint32_t stride, int random, unsigned long long* sink) {
    const uint32_t spw = stride / 32u;   // sectors per window
    uint64_t acc = 0;
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w0 < n_windows; w0 += 4 * step) {
        uint64_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t w = w0 + (uint64_t)u * step;
            v[u] = 0;
            if (w < n_windows) {
                const uint64_t sec = w * spw + (random ? mix(w) % spw : 0ull);
                v[u] = ld8<LD>(col + sec * 4);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[u];
    }
    if (acc == 0x123456789ull) atomicAdd(sink, 1ull);
}

template <int LD>
static void run(const uint64_t* col, uint64_t bytes, uint32_t stride, int random, unsigned long long* sink, int sms) {
    const uint64_t n_windows = bytes / stride;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = sms * 16;
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        sector_read_kernel<LD><<<grid, 256>>>(col, n_windows, stride, random, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double useful = (double)n_windows * 32.0;
    printf("ld=%d stride=%4u %s  windows=%llu  useful_sector_bytes=%.3f GB  best_ms=%.3f  useful_GBs=%.1f  touched_sectors_per_us=%.1f\n", LD, stride,
           random ? "random" : "first ", (unsigned long long)n_windows, useful / 1e9, best, useful / 1e6 / best, (double)n_windows / (best * 1e3));
}

int main(int argc, char** argv) {
    const int gran = argc > 1 ? atoi(argv[1]) : 0;        // 0 = leave the default, else cudaLimitMaxL2FetchGranularity
    const uint64_t bytes = argc > 2 ? strtoull(argv[2], nullptr, 10) : (8ull << 30);
    CK(cudaSetDevice(0));
    if (gran > 0) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran));
    size_t g = 0; CK(cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("# %s, %d SMs, cudaLimitMaxL2FetchGranularity=%zu (requested %d), column %.2f GB\n", prop.name, prop.multiProcessorCount, g, gran, bytes / 1e9);
    uint64_t* col; unsigned long long* sink;
    CK(cudaMalloc(&col, bytes)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(col, 1, bytes)); CK(cudaMemset(sink, 0, 8));
    CK(cudaDeviceSynchronize());
    const uint32_t strides[] = {32, 64, 128, 256, 512};
    for (int random = 1; random >= 0; --random)
        for (uint32_t s : strides) {
            if (!random && s == 32) continue;
            run<0>(col, bytes, s, random, sink, prop.multiProcessorCount);
            run<1>(col, bytes, s, random, sink, prop.multiProcessorCount);
            run<2>(col, bytes, s, random, sink, prop.multiProcessorCount);
            run<3>(col, bytes, s, random, sink, prop.multiProcessorCount);
        }
    CK(cudaFree(col)); CK(cudaFree(sink));
    return 0;
}
