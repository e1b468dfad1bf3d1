// What bounds the overlap pass (k_overlap, csrc/matrix.cu)?  250 M fragments, each: one returning atomicAdd on a per-barcode
// cursor (10 k hot barcodes take 90 % of the fragments, 200 k cold ones the rest) and one scattered 4-byte store behind it.
// Variants: returning atomic, non-returning reduction, no atomic at all (position known: 8-byte store at 2 * rank).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__global__ void k_fill(int32_t* bc, int64_t n, int hot, int cold) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t h = mix((uint32_t)i), g = mix(h ^ 0x9e3779b9U);
        bc[i] = (h % 10u) ? (int)(g % (uint32_t)hot) : hot + (int)(g % (uint32_t)cold);
    }
}
// exclusive scan of capacity 2 * count on the host is not needed: fixed-capacity segments
template <int MODE, int PER>
__global__ void __launch_bounds__(256) k_scatter(const int32_t* __restrict__ bc, const int32_t* __restrict__ rank, int64_t n, unsigned* cursor,
                                                 const int64_t* __restrict__ seg_off, int32_t* hits) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * PER;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x * PER + threadIdx.x; i0 < n; i0 += stride) {
        int b[PER]; unsigned p[PER]; int64_t so[PER];
#pragma unroll
        for (int u = 0; u < PER; u++) { const int64_t i = i0 + u * 256; b[u] = i < n ? bc[i] : -1; }
#pragma unroll
        for (int u = 0; u < PER; u++) {
            if (b[u] < 0) continue;
            if (MODE == 0) p[u] = atomicAdd(&cursor[b[u]], 1u);
            if (MODE == 1) { atomicAdd(&cursor[b[u]], 1u); p[u] = (unsigned)(i0 & 1023); }     // result unused: RED
            if (MODE == 2) p[u] = (unsigned)rank[i0 + u * 256];
            so[u] = seg_off[b[u]];
        }
#pragma unroll
        for (int u = 0; u < PER; u++) {
            if (b[u] < 0) continue;
            if (MODE == 2) *reinterpret_cast<int2*>(&hits[so[u] + 2 * (int64_t)p[u]]) = make_int2((int)i0, -1);
            else hits[so[u] + p[u]] = (int)i0;
        }
    }
}
__global__ void k_rank(const int32_t* bc, int64_t n, unsigned* cursor, int32_t* rank) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) rank[i] = (int)atomicAdd(&cursor[bc[i]], 1u);
}

int main() {
    const int64_t n = 250000000;
    const int hot = 10000, cold = 200000, nb = hot + cold;
    int32_t *bc, *rank, *hits; unsigned* cursor; int64_t* seg_off;
    cudaMalloc(&bc, 4 * n); cudaMalloc(&rank, 4 * n); cudaMalloc(&cursor, 4 * nb); cudaMalloc(&seg_off, 8 * (nb + 1));
    k_fill<<<148 * 8, 256>>>(bc, n, hot, cold);
    cudaMemset(cursor, 0, 4 * nb);
    k_rank<<<148 * 8, 256>>>(bc, n, cursor, rank);
    unsigned* h = new unsigned[nb]; int64_t* off = new int64_t[nb + 1];
    cudaMemcpy(h, cursor, 4 * nb, cudaMemcpyDeviceToHost);
    off[0] = 0; for (int i = 0; i < nb; i++) off[i + 1] = off[i] + 2 * (int64_t)h[i];
    cudaMemcpy(seg_off, off, 8 * (nb + 1), cudaMemcpyHostToDevice);
    cudaMalloc(&hits, 4 * off[nb]);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto kernel, int grid) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; rep++) {
            cudaMemset(cursor, 0, 4 * nb);
            cudaEventRecord(e0);
            kernel<<<grid, 256>>>(bc, rank, n, cursor, seg_off, hits);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
        }
        printf("%-40s %.3f ms  (%.1f G/s)  %s\n", name, best, n / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
    };
    run("returning atomic + 4 B store, 1/thread", k_scatter<0, 1>, 148 * 8);
    run("returning atomic + 4 B store, 2/thread", k_scatter<0, 2>, 148 * 8);
    run("returning atomic + 4 B store, 4/thread", k_scatter<0, 4>, 148 * 8);
    run("reduction (no return) + 4 B store, 2", k_scatter<1, 2>, 148 * 8);
    run("reduction (no return) + 4 B store, 4", k_scatter<1, 4>, 148 * 8);
    run("rank known: 8 B store at 2*rank, 2", k_scatter<2, 2>, 148 * 8);
    run("rank known: 8 B store at 2*rank, 4", k_scatter<2, 4>, 148 * 8);
    {
        float best = 1e9f;
        for (int rep = 0; rep < 4; rep++) {
            cudaMemset(cursor, 0, 4 * nb);
            cudaEventRecord(e0);
            k_rank<<<148 * 8, 256>>>(bc, n, cursor, rank);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
        }
        printf("%-40s %.3f ms\n", "rank = returning atomic, coalesced store", best);
    }
    return 0;
}
