// latency probe for the barycentric group evaluation used by em_fast.cu (tools only)
#include <cstdio>
#include <cuda_runtime.h>
#include <math.h>
constexpr int CM = 64;
struct Sm { double xn[CM], wn[CM], f[CM]; };
__device__ __forceinline__ double ef_rcp(double y) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    r = r * fma(-y, r, 2.0);
    return fma(r, fma(-y, r, 1.0), r);
}
template <int G, int MODE>
__device__ __forceinline__ double bary(const Sm& s, double t) {
    const int sub = threadIdx.x & (G - 1);
    t = fmin(1.0, fmax(-1.0, t));
    double num = 0.0, den = 0.0, exact = 0.0;
    int hit = 0;
#pragma unroll
    for (int m = 0; m < CM / G; m++) {
        const int i = sub + G * m;
        const double d = t - s.xn[i], fi = s.f[i];
        if (d == 0.0) { hit = 1; exact = fi; }
        const double q = MODE == 1 ? s.wn[i] / d : s.wn[i] * ef_rcp(d);
        num = fma(q, fi, num);
        den += q;
    }
    if (MODE == 2) return num + den;   // no reduction: term phase only
#pragma unroll
    for (int d = G / 2; d; d >>= 1) {
        num += __shfl_xor_sync(0xffffffffu, num, d);
        den += __shfl_xor_sync(0xffffffffu, den, d);
        exact += __shfl_xor_sync(0xffffffffu, exact, d);
        hit |= __shfl_xor_sync(0xffffffffu, hit, d);
    }
    if (MODE == 3) return num + den;   // no final division
    return hit ? exact : num * ef_rcp(den);
}
template <int G, int MODE>
__global__ void k(double* out, long long* cyc) {
    __shared__ Sm s;
    if (threadIdx.x < CM) {
        double th = 3.14159265358979323846 * (threadIdx.x + 0.5) / CM;
        s.xn[threadIdx.x] = cos(th); s.wn[threadIdx.x] = (threadIdx.x & 1) ? -sin(th) : sin(th);
        s.f[threadIdx.x] = 0.9 * cos(th) + 0.01;
    }
    __syncthreads();
    double x = 0.77;
    long long t0 = clock64();
    for (int i = 0; i < 256; i++) x = bary<G, MODE>(s, x);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x;
}
template <int G, int MODE> void run(const char* name, int threads, double* out, long long* cyc) {
    k<G, MODE><<<1, threads>>>(out, cyc); cudaDeviceSynchronize();
    printf("%-44s %4d thr: %.0f cyc/eval\n", name, threads, cyc[0] / 256.0);
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMallocManaged(&cyc, 64);
    run<8, 0>("G=8 rcp", 32, out, cyc); run<8, 0>("G=8 rcp", 512, out, cyc);
    run<8, 1>("G=8 ieee div", 32, out, cyc);
    run<8, 2>("G=8 rcp, terms only", 32, out, cyc);
    run<8, 3>("G=8 rcp, terms+shuffles", 32, out, cyc);
    run<32, 0>("G=32 rcp", 32, out, cyc); run<32, 2>("G=32 terms only", 32, out, cyc);
    run<16, 0>("G=16 rcp", 32, out, cyc); run<16, 0>("G=16 rcp", 512, out, cyc);
    run<4, 0>("G=4 rcp", 32, out, cyc);
    return 0;
}
