// FP64 pipe probe: dependent-chain latency and per-SM throughput of DFMA on this part (tools only, not product code)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double* out, long long* cyc, double t2) {
    double a = 1.0, b = 0.3;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 4096; i++) { double n = fma(t2, b, -a); a = b; b = n; }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = b;
}
template <int ILP>
__global__ void k_thr(double* out, long long* cyc, double t2) {
    double a[ILP], b[ILP];
    for (int q = 0; q < ILP; q++) { a[q] = 1.0 + q; b[q] = 0.3 * threadIdx.x; }
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < 1024; i++) {
#pragma unroll
        for (int q = 0; q < ILP; q++) { double n = fma(t2, b[q], -a[q]); a[q] = b[q]; b[q] = n; }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0; for (int q = 0; q < ILP; q++) s += b[q];
    out[threadIdx.x] = s;
}
__global__ void k_shfl(double* out, long long* cyc) {
    double v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 1024; i++) v += __shfl_xor_sync(0xffffffffu, v, 1 + (i & 15));
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = v;
}
__global__ void k_sync(double* out, long long* cyc) {
    long long t0 = clock64();
    for (int i = 0; i < 1024; i++) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMallocManaged(&cyc, 64);
    k_lat<<<1, 32>>>(out, cyc, 1.7); cudaDeviceSynchronize(); printf("DFMA dependent latency (1 warp): %.1f cyc\n", cyc[0] / 4096.0);
    k_lat<<<1, 512>>>(out, cyc, 1.7); cudaDeviceSynchronize(); printf("DFMA dependent step (16 warps): %.1f cyc\n", cyc[0] / 4096.0);
    k_thr<4><<<1, 512>>>(out, cyc, 1.7); cudaDeviceSynchronize(); printf("DFMA 512 thr ILP4: %.2f lanes/clk\n", 512.0 * 4 * 1024 / cyc[0]);
    k_thr<8><<<1, 1024>>>(out, cyc, 1.7); cudaDeviceSynchronize(); printf("DFMA 1024 thr ILP8: %.2f lanes/clk\n", 1024.0 * 8 * 1024 / cyc[0]);
    k_shfl<<<1, 32>>>(out, cyc); cudaDeviceSynchronize(); printf("double shfl+add dependent: %.1f cyc\n", cyc[0] / 1024.0);
    k_sync<<<1, 512>>>(out, cyc); cudaDeviceSynchronize(); printf("__syncthreads 512 thr: %.1f cyc\n", cyc[0] / 1024.0);
    return 0;
}
