// Exclusive prefix sums (int32/int64 -> int64): reduce / recurse on block sums / scan-with-offset.
// HBM-bound helper used for line offsets, per-barcode segment offsets and indptr.
#include "common.cuh"

namespace cratac {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int64_t block_exclusive_scan_256(int64_t v, int64_t* total, int64_t* warp_sums) {
    // inclusive warp scan
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < (SCAN_THREADS / 32) ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int64_t y = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += y;
        }
        if (lane < (SCAN_THREADS / 32)) warp_sums[lane] = w;   // inclusive warp totals
    }
    __syncthreads();
    int64_t warp_off = warp ? warp_sums[warp - 1] : 0;
    *total = warp_sums[SCAN_THREADS / 32 - 1];
    return warp_off + x - v;
}

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const TIn* __restrict__ in, int64_t n, int64_t* __restrict__ block_sums) {
    __shared__ int64_t warp_sums[SCAN_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) s += (int64_t)in[base + k];
    int64_t total;
    block_exclusive_scan_256(s, &total, warp_sums);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const TIn* in, int64_t n, const int64_t* __restrict__ block_offs,
                                                              int64_t* out, int64_t* total_out) {
    __shared__ int64_t warp_sums[SCAN_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? (int64_t)in[base + k] : 0;
        s += v[k];
    }
    int64_t total;
    int64_t off = block_exclusive_scan_256(s, &total, warp_sums) + (block_offs ? block_offs[blockIdx.x] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = off;
        off += v[k];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *total_out = off;
}

template <typename TIn>
static int scan_impl(Ctx* c, const TIn* d_in, int64_t* d_out, int64_t n, int64_t* d_total, int64_t* tmp, size_t tmp_elems) {
    if (n <= 0) {
        if (d_total) CRATAC_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int64_t), lane_stream(c)));
        return CRATAC_OK;
    }
    int64_t nblocks = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nblocks == 1) {
        k_scan_apply<TIn><<<1, SCAN_THREADS, 0, lane_stream(c)>>>(d_in, n, nullptr, d_out, d_total);
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
        return CRATAC_OK;
    }
    if ((size_t)nblocks > tmp_elems) { set_error("scan scratch too small"); return CRATAC_ERR_INTERNAL; }
    k_scan_reduce<TIn><<<(unsigned)nblocks, SCAN_THREADS, 0, lane_stream(c)>>>(d_in, n, tmp);
    c->launches++;
    CRATAC_CUDA(cudaGetLastError());
    CRATAC_TRY(scan_impl<int64_t>(c, tmp, tmp, nblocks, nullptr, tmp + nblocks, tmp_elems - nblocks));
    k_scan_apply<TIn><<<(unsigned)nblocks, SCAN_THREADS, 0, lane_stream(c)>>>(d_in, n, tmp, d_out, d_total);
    c->launches++;
    CRATAC_CUDA(cudaGetLastError());
    return CRATAC_OK;
}

static int scan_scratch(Ctx* c, int64_t n, int64_t** tmp, size_t* elems) {
    size_t need = 0;
    for (int64_t m = (n + SCAN_TILE - 1) / SCAN_TILE; m > 1; m = (m + SCAN_TILE - 1) / SCAN_TILE) need += (size_t)m;
    need += 8;
    CRATAC_TRY(lane_scan_tmp(c).reserve(need * sizeof(int64_t)));
    *tmp = lane_scan_tmp(c).as<int64_t>();
    *elems = need;
    return CRATAC_OK;
}

int exclusive_scan_i64(Ctx* c, const int64_t* d_in, int64_t* d_out, int64_t n, int64_t* d_total) {
    int64_t* tmp; size_t elems;
    CRATAC_TRY(scan_scratch(c, n, &tmp, &elems));
    return scan_impl<int64_t>(c, d_in, d_out, n, d_total, tmp, elems);
}

int exclusive_scan_i32_to_i64(Ctx* c, const int32_t* d_in, int64_t* d_out, int64_t n, int64_t* d_total) {
    int64_t* tmp; size_t elems;
    CRATAC_TRY(scan_scratch(c, n, &tmp, &elems));
    return scan_impl<int32_t>(c, d_in, d_out, n, d_total, tmp, elems);
}

}  // namespace cratac
