// K5b: bedgraph rows of one contig on the device.  Replaces write_chrom_bedgraph
// (mro/atac/stages/processing/count_cut_sites/__init__.py:27-48): a per-base Python loop that run-length encodes the window
// counts, in which a ZERO never closes the open run (:36-37) -- so equal counts either side of a zero gap are one row, and a
// row ends one past the last non-zero base before the next change of value.
//
// The window counts of the contig are written once to HBM by the pileup kernel (k_pileup<DUMP>, 4 bytes per base: they do
// not leave the device); then, with BG_TILE bases per CTA:
//   k_bg_last      per tile: position and value of its last non-zero base
//   k_bg_carry     one CTA: for every tile, the last non-zero base BEFORE it (a scan with "right operand wins if non-zero")
//   k_bg_heads     per tile: a non-zero base opens a row iff its value differs from the previous non-zero value; count them
//   scan           row offsets
//   k_bg_rows      per tile: start and count of the rows it opens, and the end of the row before each of them
// The rows come back as three int32 arrays (12 bytes per row over PCIe instead of 4 bytes per base).
#include <algorithm>

#include "common.cuh"

namespace cratac {

constexpr int BG_THREADS = 256;
constexpr int BG_PER = 16;                           // consecutive bases per thread
constexpr int BG_TILE = BG_THREADS * BG_PER;         // 4096 bases per CTA

int dump_window_counts(Ctx* c, int contig, int32_t* d_W);   // pileup.cu

struct LastNz { int32_t pos, val; };                 // pos < 0: none

__device__ __forceinline__ LastNz bg_pick(LastNz left, LastNz right) { return right.pos >= 0 ? right : left; }

// block-wide exclusive scan of "last non-zero so far" (thread order = base order); returns the value BEFORE this thread
__device__ __forceinline__ LastNz bg_block_exclusive(LastNz mine, LastNz carry_in, LastNz* warp_last) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    LastNz incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        LastNz o;
        o.pos = __shfl_up_sync(0xffffffffu, incl.pos, d);
        o.val = __shfl_up_sync(0xffffffffu, incl.val, d);
        if (lane >= d) incl = bg_pick(o, incl);
    }
    if (lane == 31) warp_last[warp] = incl;
    __syncthreads();
    LastNz before = carry_in;
    for (int w = 0; w < warp; w++) before = bg_pick(before, warp_last[w]);
    LastNz prev;
    prev.pos = __shfl_up_sync(0xffffffffu, incl.pos, 1);
    prev.val = __shfl_up_sync(0xffffffffu, incl.val, 1);
    if (lane > 0) before = bg_pick(before, prev);
    __syncthreads();
    return before;
}

__global__ void __launch_bounds__(BG_THREADS) k_bg_last(const int32_t* __restrict__ W, int64_t L, LastNz* __restrict__ tile_last) {
    __shared__ LastNz warp_last[BG_THREADS / 32];
    const int64_t p0 = (int64_t)blockIdx.x * BG_TILE + (int64_t)threadIdx.x * BG_PER;
    LastNz mine{-1, 0};
#pragma unroll
    for (int k = 0; k < BG_PER; k += 4) {
        const int64_t p = p0 + k;
        int4 v = make_int4(0, 0, 0, 0);
        if (p + 4 <= L) v = *reinterpret_cast<const int4*>(W + p);
        else { if (p < L) v.x = W[p]; if (p + 1 < L) v.y = W[p + 1]; if (p + 2 < L) v.z = W[p + 2]; }
        if (v.x) mine = LastNz{(int32_t)p, v.x};
        if (v.y) mine = LastNz{(int32_t)(p + 1), v.y};
        if (v.z) mine = LastNz{(int32_t)(p + 2), v.z};
        if (v.w) mine = LastNz{(int32_t)(p + 3), v.w};
    }
    const LastNz before = bg_block_exclusive(mine, LastNz{-1, 0}, warp_last);
    if (threadIdx.x == BG_THREADS - 1) tile_last[blockIdx.x] = bg_pick(before, mine);
}

// tile_carry[t] = last non-zero base before tile t; tile_carry[n_tiles] = last non-zero base of the contig
__global__ void __launch_bounds__(1024) k_bg_carry(const LastNz* __restrict__ tile_last, int64_t n_tiles, LastNz* __restrict__ tile_carry) {
    __shared__ LastNz warp_last[32];
    __shared__ LastNz s_carry;
    if (threadIdx.x == 0) s_carry = LastNz{-1, 0};
    __syncthreads();
    for (int64_t base = 0; base < n_tiles; base += 1024) {
        const int64_t t = base + threadIdx.x;
        const LastNz mine = t < n_tiles ? tile_last[t] : LastNz{-1, 0};
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        LastNz incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            LastNz o;
            o.pos = __shfl_up_sync(0xffffffffu, incl.pos, d);
            o.val = __shfl_up_sync(0xffffffffu, incl.val, d);
            if (lane >= d) incl = bg_pick(o, incl);
        }
        if (lane == 31) warp_last[warp] = incl;
        __syncthreads();
        LastNz before = s_carry;
        for (int w = 0; w < warp; w++) before = bg_pick(before, warp_last[w]);
        LastNz prev;
        prev.pos = __shfl_up_sync(0xffffffffu, incl.pos, 1);
        prev.val = __shfl_up_sync(0xffffffffu, incl.val, 1);
        if (lane > 0) before = bg_pick(before, prev);
        if (t < n_tiles) tile_carry[t] = before;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = bg_pick(before, mine);
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_carry[n_tiles] = s_carry;
}

// WRITE = false: heads per tile -> tile_heads.  WRITE = true: rows at row_off[tile] + rank.
template <bool WRITE>
__global__ void __launch_bounds__(BG_THREADS) k_bg_rows(const int32_t* __restrict__ W, int64_t L, const LastNz* __restrict__ tile_carry,
                                                         int32_t* __restrict__ tile_heads, const int64_t* __restrict__ row_off,
                                                         int32_t* __restrict__ r_start, int32_t* __restrict__ r_end, int32_t* __restrict__ r_count) {
    __shared__ LastNz warp_last[BG_THREADS / 32];
    __shared__ int warp_cnt[BG_THREADS / 32];
    const int64_t p0 = (int64_t)blockIdx.x * BG_TILE + (int64_t)threadIdx.x * BG_PER;
    int32_t v[BG_PER];
#pragma unroll
    for (int k = 0; k < BG_PER; k += 4) {
        const int64_t p = p0 + k;
        int4 q = make_int4(0, 0, 0, 0);
        if (p + 4 <= L) q = *reinterpret_cast<const int4*>(W + p);
        else { if (p < L) q.x = W[p]; if (p + 1 < L) q.y = W[p + 1]; if (p + 2 < L) q.z = W[p + 2]; }
        v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
    }
    LastNz mine{-1, 0};
#pragma unroll
    for (int k = 0; k < BG_PER; k++) if (v[k]) mine = LastNz{(int32_t)(p0 + k), v[k]};
    LastNz before = bg_block_exclusive(mine, tile_carry[blockIdx.x], warp_last);
    // heads among my bases
    unsigned heads = 0;
    LastNz run = before;
#pragma unroll
    for (int k = 0; k < BG_PER; k++) {
        if (v[k]) {
            if (run.pos < 0 || run.val != v[k]) heads |= 1u << k;
            run = LastNz{(int32_t)(p0 + k), v[k]};
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int my = __popc(heads);
    int incl = my;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
    if (lane == 31) warp_cnt[warp] = incl;
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < BG_THREADS / 32; w++) { const int x = warp_cnt[w]; if (w < warp) off += x; total += x; }
    if (!WRITE) {
        if (threadIdx.x == 0) tile_heads[blockIdx.x] = total;
        return;
    }
    int64_t row = row_off[blockIdx.x] + off + incl - my;
    run = before;
#pragma unroll
    for (int k = 0; k < BG_PER; k++) {
        if (v[k]) {
            if (heads & (1u << k)) {
                r_start[row] = (int32_t)(p0 + k);
                r_count[row] = v[k];
                if (row > 0) r_end[row - 1] = run.pos + 1;      // the row before ends one past the last non-zero base before this one
                row++;
            }
            run = LastNz{(int32_t)(p0 + k), v[k]};
        }
    }
}

__global__ void k_bg_close(const LastNz* __restrict__ tile_carry, int64_t n_tiles, const int64_t* __restrict__ n_rows, int32_t* __restrict__ r_end) {
    const int64_t n = *n_rows;
    if (n > 0) r_end[n - 1] = tile_carry[n_tiles].pos + 1;
}

// rows of `contig` -> c->d_bg_* (int32 start / end / count), c->bg_rows; cached until the fragments change
int bedgraph_rows_device(Ctx* c, int contig) {
    if (c->bg_contig == contig && c->bg_valid) return CRATAC_OK;
    const int64_t L = c->contig_lens[contig];
    if (L >= 0x7fffffffLL) { set_error("contig too long for 32-bit bedgraph rows"); return CRATAC_ERR_ARG; }
    cudaStream_t s = c->stream;
    const int64_t n_tiles = std::max<int64_t>((L + BG_TILE - 1) / BG_TILE, 1);
    CRATAC_TRY(c->d_bg_W.reserve(4 * (size_t)std::max<int64_t>(L, 1) + 16));
    CRATAC_TRY(c->d_bg_tiles.reserve((size_t)n_tiles * (8 + 8 + 4 + 8) + 64 + 16));
    LastNz* tile_last = c->d_bg_tiles.as<LastNz>();
    LastNz* tile_carry = tile_last + n_tiles;                                   // n_tiles + 1 entries
    int64_t* row_off = reinterpret_cast<int64_t*>(tile_carry + n_tiles + 1);     // n_tiles + 1 entries (last = total)
    int32_t* tile_heads = reinterpret_cast<int32_t*>(row_off + n_tiles + 1);
    int32_t* W = c->d_bg_W.as<int32_t>();
    CRATAC_TRY(dump_window_counts(c, contig, W));
    k_bg_last<<<(unsigned)n_tiles, BG_THREADS, 0, s>>>(W, L, tile_last);
    k_bg_carry<<<1, 1024, 0, s>>>(tile_last, n_tiles, tile_carry);
    k_bg_rows<false><<<(unsigned)n_tiles, BG_THREADS, 0, s>>>(W, L, tile_carry, tile_heads, nullptr, nullptr, nullptr, nullptr);
    c->launches += 3;
    CRATAC_CUDA(cudaGetLastError());
    CRATAC_TRY(exclusive_scan_i32_to_i64(c, tile_heads, row_off, n_tiles, row_off + n_tiles));
    int64_t n_rows = 0;
    CRATAC_CUDA(cudaMemcpyAsync(&n_rows, row_off + n_tiles, 8, cudaMemcpyDeviceToHost, s));
    CRATAC_CUDA(cudaStreamSynchronize(s));
    const size_t nn = (size_t)std::max<int64_t>(n_rows, 1);
    CRATAC_TRY(c->d_bg_start.reserve(4 * nn)); CRATAC_TRY(c->d_bg_end.reserve(4 * nn)); CRATAC_TRY(c->d_bg_count.reserve(4 * nn));
    if (n_rows > 0) {
        k_bg_rows<true><<<(unsigned)n_tiles, BG_THREADS, 0, s>>>(W, L, tile_carry, nullptr, row_off, c->d_bg_start.as<int32_t>(), c->d_bg_end.as<int32_t>(),
                                                                c->d_bg_count.as<int32_t>());
        k_bg_close<<<1, 1, 0, s>>>(tile_carry, n_tiles, row_off + n_tiles, c->d_bg_end.as<int32_t>());
        c->launches += 2;
        CRATAC_CUDA(cudaGetLastError());
    }
    c->bg_rows = n_rows;
    c->bg_contig = contig;
    c->bg_valid = true;
    return CRATAC_OK;
}

}  // namespace cratac
