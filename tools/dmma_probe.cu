// dmma_probe.cu — latency / throughput of mma.sync.m8n8k4.f64 (DMMA) against plain DFMA on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu && tools/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CH> __global__ void k_dmma(double *out, int iters, long long *cyc) {
    double c[CH][2];
    for (int i = 0; i < CH; i++) c[i][0] = c[i][1] = 0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < CH; i++) dmma(c[i][0], c[i][1], a, b);
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH> __global__ void k_dfma(double *out, int iters, long long *cyc) {
    double c[CH];
    for (int i = 0; i < CH; i++) c[i] = 0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < CH; i++) c[i] = fma(a, b, c[i]);
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CH; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int warps : {1, 4, 8, 16, 32}) {
        k_dmma<1><<<1, warps * 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DMMA  chain=1 warps=%2d: %.1f cycles per DMMA per warp (dependent)\n", warps, (double)h / iters);
        k_dmma<8><<<1, warps * 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DMMA  chain=8 warps=%2d: %.2f cycles per DMMA per warp -> %.1f fp64 FMA/clk/SM\n", warps, (double)h / iters / 8, 256.0 * 8 * warps * iters / h);
        k_dfma<1><<<1, warps * 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA  chain=1 warps=%2d: %.1f cycles per DFMA (dependent)\n", warps, (double)h / iters);
        k_dfma<8><<<1, warps * 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA  chain=8 warps=%2d: %.2f cycles per DFMA per warp -> %.1f fp64 FMA/clk/SM\n", warps, (double)h / iters / 8, 32.0 * 8 * warps * iters / h);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
