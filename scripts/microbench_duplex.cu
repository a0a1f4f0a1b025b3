// microbench_duplex.cu — what one GPU's PCIe link delivers while N - 1 other GPUs of the same host do the same thing.
//
// the end-to-end leg of bench.py scales 1.00 / 0.76 / 0.41 / 0.28 per GPU at 1 / 2 / 4 / 8 ranks and the builder
// blamed the host.  This measures it: every process drives ONE GPU through fixed wall-clock windows that all processes share
// (argv[2] = epoch second at which window 0 opens), so the N instances really overlap:
//   window 0  H2D, copy engine           (cudaMemcpyAsync from pinned memory, 256 MiB pieces)
//   window 1  D2H, copy engine
//   window 2  H2D + D2H, copy engines    (the e2e leg's STAGED mode)
//   window 3  H2D, SM-issued             (ld.global.nc.v2 from the mapped host pointer: the e2e leg's ZERO_COPY reads)
//   window 4  H2D SM-issued + D2H copy engine   (the e2e leg's AUTO mode)
// Each window lasts `win` seconds with a `gap` second pause; the program prints GB/s per direction per window.
// scripts/pcie_scaling.py launches N of these (optionally each bound to its GPU's CPU set) and tabulates the aggregate.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/microbench_duplex scripts/microbench_duplex.cu
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) host_read_kernel(const ulonglong2* __restrict__ src, ulonglong2* __restrict__ dst, size_t n16) {
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += 4 * step) {
        ulonglong2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i + u * step < n16) asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v[u].x), "=l"(v[u].y) : "l"(src + i + u * step));
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i + u * step < n16) dst[i + u * step] = v[u];
    }
}

static double now_s() { return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count(); }
static void sleep_until(double t) { const double d = t - now_s(); if (d > 0) std::this_thread::sleep_for(std::chrono::duration<double>(d)); }

int main(int argc, char** argv) {
    const int dev = argc > 1 ? atoi(argv[1]) : 0;
    const double t_start = argc > 2 ? atof(argv[2]) : now_s() + 1.0;
    const double win = argc > 3 ? atof(argv[3]) : 1.5, gap = 0.5;
    const size_t piece = 256ull << 20, n_piece = 8;   // 2 GiB of pinned memory each way
    CK(cudaSetDevice(dev));
    uint8_t *h_in, *h_out, *d_in, *d_out;
    CK(cudaHostAlloc(&h_in, piece * n_piece, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_out, piece * n_piece, cudaHostAllocDefault));
    CK(cudaMalloc(&d_in, piece * n_piece)); CK(cudaMalloc(&d_out, piece * n_piece));
    for (size_t i = 0; i < piece * n_piece; i += 4096) { h_in[i] = 1; h_out[i] = 1; }   // touch: pages exist before any window opens
    CK(cudaMemset(d_out, 1, piece * n_piece));
    void* d_view = nullptr;
    CK(cudaHostGetDevicePointer(&d_view, h_in, 0));
    cudaStream_t s_in, s_out;
    CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    const char* names[5] = {"H2D copy engine", "D2H copy engine", "H2D + D2H copy engines", "H2D SM-issued reads", "H2D SM-issued + D2H copy engine"};
    if (now_s() > t_start) fprintf(stderr, "warning: device %d was ready %.2f s after window 0 opened\n", dev, now_s() - t_start);
    for (int w = 0; w < 5; ++w) {
        const double t0 = t_start + w * (win + gap), t1 = t0 + win;
        sleep_until(t0);
        const bool in_ce = w == 0 || w == 2, in_sm = w == 3 || w == 4, out_ce = w == 1 || w == 2 || w == 4;
        size_t n_in = 0, n_out = 0;
        const double b0 = now_s();
        // keep two pieces in flight per direction until the window closes
        while (now_s() < t1) {
            for (int k = 0; k < 2; ++k) {
                const size_t off = ((n_in + k) % n_piece) * piece;
                if (in_ce) CK(cudaMemcpyAsync(d_in + off, h_in + off, piece, cudaMemcpyHostToDevice, s_in));
                if (in_sm) host_read_kernel<<<prop.multiProcessorCount, 256, 0, s_in>>>((const ulonglong2*)((uint8_t*)d_view + off), (ulonglong2*)(d_in + off), piece / 16);
                const size_t off2 = ((n_out + k) % n_piece) * piece;
                if (out_ce) CK(cudaMemcpyAsync(h_out + off2, d_out + off2, piece, cudaMemcpyDeviceToHost, s_out));
            }
            if (in_ce || in_sm) { CK(cudaStreamSynchronize(s_in)); n_in += 2; }
            if (out_ce) { CK(cudaStreamSynchronize(s_out)); n_out += 2; }
        }
        const double dt = now_s() - b0;
        printf("dev %d window %d %-34s  h2d %7.2f GB/s  d2h %7.2f GB/s\n", dev, w, names[w], n_in * piece / dt / 1e9, n_out * piece / dt / 1e9);
        fflush(stdout);
    }
    return 0;
}
