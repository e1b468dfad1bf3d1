// The other scans of the fragments file that a decoded SoA in HBM serves: the fragment x
// region siblings of the peak matrix.
//
//   cratac_low_targeting      REMOVE_LOW_TARGETING_BARCODES.main (stages/processing/cell_calling/remove_low_targeting_barcodes/
//                             __init__.py:176-245): per barcode, fragments, fragments whose INTERVAL overlaps a peak, padded
//                             fragment lengths and the bases covered by the union of the padded fragments
//   cratac_insert_sizes       REPORT_INSERT_SIZES.main (stages/metrics/report_insert_sizes/__init__.py:76-113): per barcode
//                             histogram of fragment sizes 1..1000 and > 1000
//   cratac_counts_by_barcode  get_counts_by_barcode (lib/python/tools/io.py:429-555): per barcode counts against the TSS /
//                             CTCF / DNase / enhancer / promoter / blacklist / peak tracks and the TSS / CTCF profiles
//
// The reference reads fragments.tsv.gz once more for each of them, a Python loop per line.  Here the fragments are already
// decoded; the only structure missing is "the fragments of one barcode in file order" (the covered-bases union walks them in
// order), which a stable LSD radix sort of the fragment indices by barcode provides (4 bits per pass, block-local ranks
// from packed per-thread digit counts).  Region lookups follow lib/python/tools/regions.py to the letter, including the
// bisect over region ENDS that are not sorted when regions nest (:83-93).
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace cratac {

int build_barcodes(Ctx* c);   // decode.cu

// ------------------------------------------------------------------ stable radix sort of fragment indices by barcode
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void k_rs_keys(const int32_t* __restrict__ bc, const int32_t* __restrict__ slot_to_dense, int64_t n, int32_t* __restrict__ key,
                          int32_t* __restrict__ idx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = bc[i];
        key[i] = slot_to_dense ? slot_to_dense[b] : b;
        idx[i] = (int32_t)i;
    }
}

// digit counts of one tile: hist[d * n_tiles + tile]
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const int32_t* __restrict__ key, int64_t n, int shift, int64_t n_tiles, int32_t* __restrict__ hist) {
    __shared__ int cnt[16];
    if (threadIdx.x < 16) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)threadIdx.x * RS_ITEMS;
    unsigned long long packed = 0;                     // 16 digits x 4 bits: at most 8 items of one digit per thread
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++)
        if (base + k < n) packed += 1ULL << (4 * ((key[base + k] >> shift) & 15));
    // 8 items of the same digit overflow a 4-bit field into the next digit: count them apart
    // (all 8 equal  <=>  the field would read 8 = 0b1000, which fits; 4 bits hold 0..15, so no overflow at all)
#pragma unroll
    for (int d = 0; d < 16; d++) {
        const int c = (int)((packed >> (4 * d)) & 15ULL);
        if (c) atomicAdd(&cnt[d], c);
    }
    __syncthreads();
    if (threadIdx.x < 16) hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = cnt[threadIdx.x];
}

// scatter with stable ranks: position = global offset of (digit, tile) + items of that digit earlier in the tile
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const int32_t* __restrict__ key_in, const int32_t* __restrict__ idx_in, int64_t n, int shift,
                                                            int64_t n_tiles, const int64_t* __restrict__ offs, int32_t* __restrict__ key_out,
                                                            int32_t* __restrict__ idx_out) {
    __shared__ unsigned warp_tot[RS_THREADS / 32][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)threadIdx.x * RS_ITEMS;
    int32_t k[RS_ITEMS], v[RS_ITEMS];
    unsigned cntw[8] = {0, 0, 0, 0, 0, 0, 0, 0};       // my digit counts, two 16-bit fields per word (digit d: word d / 2, half d % 2)
#pragma unroll
    for (int q = 0; q < RS_ITEMS; q++) {
        if (base + q < n) {
            k[q] = key_in[base + q]; v[q] = idx_in[base + q];
            const int d = (k[q] >> shift) & 15;
            cntw[d >> 1] += 1u << (16 * (d & 1));
        } else { k[q] = 0x7fffffff; v[q] = -1; }
    }
    // exclusive prefix of the packed counts over the threads of the tile (fields stay below 2048 < 65536)
    unsigned incl[8];
#pragma unroll
    for (int w = 0; w < 8; w++) {
        unsigned x = cntw[w];
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, dd); if (lane >= dd) x += y; }
        incl[w] = x;
        if (lane == 31) warp_tot[warp][w] = x;
    }
    __syncthreads();
    unsigned run[8];
#pragma unroll
    for (int w = 0; w < 8; w++) {
        unsigned before = incl[w] - cntw[w];
        for (int q = 0; q < warp; q++) before += warp_tot[q][w];
        run[w] = before;
    }
#pragma unroll
    for (int q = 0; q < RS_ITEMS; q++) {
        if (base + q < n) {
            const int d = (k[q] >> shift) & 15;
            const unsigned r = (run[d >> 1] >> (16 * (d & 1))) & 0xFFFFu;
            run[d >> 1] += 1u << (16 * (d & 1));
            const int64_t pos = offs[(int64_t)d * n_tiles + blockIdx.x] + r;
            key_out[pos] = k[q]; idx_out[pos] = v[q];
        }
    }
}

// fragment indices grouped by dense barcode id, file order inside a barcode -> c->d_sib_perm; segment offsets -> c->d_sib_seg
static int fragments_by_barcode(Ctx* c) {
    if (c->sib_valid) return CRATAC_OK;
    const int64_t n = c->n_fragments, nb = c->n_barcodes;
    if (n >= 0x7fffffffLL) { set_error("too many fragments for 32-bit fragment indices"); return CRATAC_ERR_ARG; }
    cudaStream_t s = c->stream;
    const size_t nn = (size_t)std::max<int64_t>(n, 1);
    CRATAC_TRY(c->d_sib_perm.reserve(4 * nn)); CRATAC_TRY(c->d_sib_tmp.reserve(12 * nn));
    CRATAC_TRY(c->d_sib_seg.reserve(8 * (size_t)(nb + 2)));
    int32_t* key_a = c->d_sib_tmp.as<int32_t>();
    int32_t* key_b = key_a + nn;
    int32_t* idx_b = key_b + nn;
    int32_t* idx_a = c->d_sib_perm.as<int32_t>();
    if (n > 0) {
        k_rs_keys<<<NUM_SMS * 8, 256, 0, s>>>(c->d_bc.as<int32_t>(), c->bc_is_slot ? c->d_slot_to_dense.as<int32_t>() : nullptr, n, key_a, idx_a);
        c->launches++;
    }
    const int64_t n_tiles = std::max<int64_t>((n + RS_TILE - 1) / RS_TILE, 1);
    CRATAC_TRY(c->d_sib_hist.reserve((4 + 8) * 16 * (size_t)n_tiles + 64));
    int32_t* hist = c->d_sib_hist.as<int32_t>();
    int64_t* offs = reinterpret_cast<int64_t*>(hist + 16 * n_tiles + (16 * n_tiles & 1));
    int bits = 1;
    while (((int64_t)1 << bits) < std::max<int64_t>(nb, 2)) bits++;
    int passes = (bits + 3) / 4;
    if (passes & 1) passes++;                          // an even number of passes ends in d_sib_perm
    for (int p = 0; p < passes && n > 0; p++) {
        const int32_t* kin = (p & 1) ? key_b : key_a; const int32_t* iin = (p & 1) ? idx_b : idx_a;
        int32_t* kout = (p & 1) ? key_a : key_b; int32_t* iout = (p & 1) ? idx_a : idx_b;
        k_rs_hist<<<(unsigned)n_tiles, RS_THREADS, 0, s>>>(kin, n, 4 * p, n_tiles, hist);
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
        CRATAC_TRY(exclusive_scan_i32_to_i64(c, hist, offs, 16 * n_tiles, nullptr));
        k_rs_scatter<<<(unsigned)n_tiles, RS_THREADS, 0, s>>>(kin, iin, n, 4 * p, n_tiles, offs, kout, iout);
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
    }
    // segment offsets from the fragments per barcode the decoder counted
    if (!c->dense_nfrag_valid) { set_error("fragments per barcode are not available"); return CRATAC_ERR_INTERNAL; }
    CRATAC_TRY(exclusive_scan_i32_to_i64(c, c->d_dense_nfrag.as<int32_t>(), c->d_sib_seg.as<int64_t>(), nb, c->d_sib_seg.as<int64_t>() + nb));
    c->sib_valid = true;
    return CRATAC_OK;
}

// ------------------------------------------------------------------ region lookups (lib/python/tools/regions.py)
// overlaps_region (:83-93): k = bisect_left(ends, start); k < n and end > starts[k]
__device__ __forceinline__ bool region_overlaps(const int32_t* __restrict__ starts, const int32_t* __restrict__ ends, int64_t lo0, int64_t hi0, int start,
                                                int end) {
    int64_t lo = lo0, hi = hi0;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (ends[mid] < start) lo = mid + 1; else hi = mid; }
    return lo < hi0 && end > starts[lo];
}
// get_region_containing_point (:66-81): k = bisect(starts, point) - 1; starts[k] <= point < ends[k]  -> index or -1
__device__ __forceinline__ int64_t region_with_point(const int32_t* __restrict__ starts, const int32_t* __restrict__ ends, int64_t lo0, int64_t hi0, int point) {
    int64_t lo = lo0, hi = hi0;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (point < starts[mid]) hi = mid; else lo = mid + 1; }
    const int64_t k = lo - 1;
    return (k >= lo0 && starts[k] <= point && point < ends[k]) ? k : -1;
}

// ------------------------------------------------------------------ REMOVE_LOW_TARGETING_BARCODES
constexpr int LT_MAX_PAD = 4;
struct LowTargetArgs {
    const int32_t* perm; const int64_t* seg; int64_t nb;
    const int32_t *chrom, *start, *end;
    const int64_t* contig_lens;
    const int64_t* reg_off;                 // per contig (+ total): regions of the contig, tools/regions.py order
    const int32_t *reg_start, *reg_end;
    int n_pad; long long pad[LT_MAX_PAD];
    long long* out;                         // [2 + 2 * n_pad][nb]: fragments, targeted, lengths[p], covered[p]  (dense barcode order)
};

// one warp per barcode, 32 of its fragments (file order) per round
__global__ void __launch_bounds__(256) k_low_targeting(LowTargetArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < a.nb; b += n_warps) {
        const int64_t s0 = a.seg[b], s1 = a.seg[b + 1];
        long long targeted = 0, len[LT_MAX_PAD] = {0, 0, 0, 0}, cov[LT_MAX_PAD] = {0, 0, 0, 0};
        int carry_c = -1;
        long long carry_stop = 0;           // largest stop of the earlier fragments of this barcode on contig carry_c
        for (int64_t base = s0; base < s1; base += 32) {
            const bool live = base + lane < s1;
            int c = -2, st = 0, en = 0;
            if (live) { const int i = a.perm[base + lane]; c = a.chrom[i]; st = a.start[i]; en = a.end[i]; }
            // exclusive prefix maximum of the stops over the earlier lanes of MY contig (the fragments of a barcode are in
            // file order, so a contig's lanes are consecutive), seeded with the carry of the previous rounds
            const int c_prev = __shfl_up_sync(0xffffffffu, c, 1);
            const int c_first = __shfl_sync(0xffffffffu, c, 0);
            long long sm = en;
            bool flag = lane == 0 || c_prev != c;          // head of a run of equal contigs
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long om = __shfl_up_sync(0xffffffffu, sm, d);
                const bool of = __shfl_up_sync(0xffffffffu, flag, d);
                if (lane >= d && !flag) { sm = om > sm ? om : sm; flag = of; }
            }
            // sm: maximum over the lanes of my run up to and including me
            long long pm = __shfl_up_sync(0xffffffffu, sm, 1);
            bool have = lane > 0 && c_prev == c;
            if (!have) pm = 0;
            if (c == c_first && carry_c == c) { pm = have ? (pm > carry_stop ? pm : carry_stop) : carry_stop; have = true; }
            if (live) {
                const long long L = a.contig_lens[c];
                const int64_t r0 = a.reg_off[c], r1 = a.reg_off[c + 1];
                if (r1 > r0 && region_overlaps(a.reg_start, a.reg_end, r0, r1, st, en)) targeted++;
                for (int p = 0; p < a.n_pad; p++) {
                    const long long lo = st - a.pad[p] > 0 ? st - a.pad[p] : 0;
                    const long long hi = en + a.pad[p] < L ? en + a.pad[p] : L;
                    len[p] += hi - lo;
                    if (have) {
                        const long long pmb = pm + a.pad[p] < L ? pm + a.pad[p] : L;     // running stop of the open run
                        if (lo < pmb) cov[p] += hi > pmb ? hi - pmb : 0;                 // joins it
                        else cov[p] += hi - lo;                                          // opens a new one
                    } else {
                        cov[p] += hi - lo;
                    }
                }
            }
            // carry: contig and running maximum at the last live lane
            const int last = (int)((s1 - base < 32 ? s1 - base : 32) - 1);
            const int lc = __shfl_sync(0xffffffffu, c, last);
            long long lm = __shfl_sync(0xffffffffu, sm, last);
            if (lc == c_first && carry_c == lc) lm = lm > carry_stop ? lm : carry_stop;
            carry_c = lc; carry_stop = lm;
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            targeted += __shfl_xor_sync(0xffffffffu, targeted, d);
            for (int p = 0; p < LT_MAX_PAD; p++) { len[p] += __shfl_xor_sync(0xffffffffu, len[p], d); cov[p] += __shfl_xor_sync(0xffffffffu, cov[p], d); }
        }
        if (lane == 0) {
            a.out[0 * a.nb + b] = s1 - s0;
            a.out[1 * a.nb + b] = targeted;
            for (int p = 0; p < a.n_pad; p++) { a.out[(2 + p) * a.nb + b] = len[p]; a.out[(2 + a.n_pad + p) * a.nb + b] = cov[p]; }
        }
    }
}

// regions (contig index, start, end), sorted by contig and, inside a contig, in tools/regions.py order -> device + per-contig offsets
static int upload_regions(Ctx* c, int64_t n_regions, const int32_t* chrom, const int32_t* start, const int32_t* end, DevBuf& d_off, DevBuf& d_start, DevBuf& d_end) {
    std::vector<int64_t> off((size_t)c->n_contigs + 1, 0);
    for (int64_t i = 0; i < n_regions; i++) {
        if (chrom[i] < 0 || chrom[i] >= c->n_contigs || (i > 0 && chrom[i] < chrom[i - 1])) { set_error("regions must be grouped by contig in reference order"); return CRATAC_ERR_ARG; }
        off[(size_t)chrom[i] + 1]++;
    }
    for (int k = 0; k < c->n_contigs; k++) off[(size_t)k + 1] += off[(size_t)k];
    const size_t nr = (size_t)std::max<int64_t>(n_regions, 1);
    CRATAC_TRY(d_off.reserve(8 * off.size())); CRATAC_TRY(d_start.reserve(4 * nr)); CRATAC_TRY(d_end.reserve(4 * nr));
    CRATAC_CUDA(cudaMemcpyAsync(d_off.p, off.data(), 8 * off.size(), cudaMemcpyHostToDevice, c->stream));
    if (n_regions > 0) {
        CRATAC_CUDA(cudaMemcpyAsync(d_start.p, start, 4 * (size_t)n_regions, cudaMemcpyHostToDevice, c->stream));
        CRATAC_CUDA(cudaMemcpyAsync(d_end.p, end, 4 * (size_t)n_regions, cudaMemcpyHostToDevice, c->stream));
    }
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));     // `off` is a local
    return CRATAC_OK;
}

// dense-barcode-order rows -> matrix column order on the host
static void to_columns(const Ctx* c, const std::vector<long long>& dense, int rows, int64_t nb, int64_t* out) {
    for (int r = 0; r < rows; r++)
        for (int64_t j = 0; j < nb; j++) out[(int64_t)r * nb + j] = dense[(size_t)r * nb + (size_t)c->h_c2d[(size_t)j]];
}

int low_targeting(Ctx* c, int64_t n_regions, const int32_t* chrom, const int32_t* start, const int32_t* end, int n_pad, const int64_t* paddings,
                  int64_t* out) {
    if (n_pad < 0 || n_pad > LT_MAX_PAD) { set_error("at most %d paddings", LT_MAX_PAD); return CRATAC_ERR_ARG; }
    CRATAC_TRY(build_barcodes(c));
    CRATAC_TRY(fragments_by_barcode(c));
    DevBuf d_off, d_rs, d_re, d_out;
    int rc = upload_regions(c, n_regions, chrom, start, end, d_off, d_rs, d_re);
    const int64_t nb = c->n_barcodes;
    const int rows = 2 + 2 * n_pad;
    if (rc == CRATAC_OK) rc = d_out.reserve(8 * (size_t)rows * (size_t)std::max<int64_t>(nb, 1));
    if (rc == CRATAC_OK && nb > 0) {
        LowTargetArgs a;
        a.perm = c->d_sib_perm.as<int32_t>(); a.seg = c->d_sib_seg.as<int64_t>(); a.nb = nb;
        a.chrom = c->d_chrom.as<int32_t>(); a.start = c->d_start.as<int32_t>(); a.end = c->d_end.as<int32_t>();
        a.contig_lens = c->d_contig_lens.as<int64_t>();
        a.reg_off = d_off.as<int64_t>(); a.reg_start = d_rs.as<int32_t>(); a.reg_end = d_re.as<int32_t>();
        a.n_pad = n_pad;
        for (int p = 0; p < LT_MAX_PAD; p++) a.pad[p] = p < n_pad ? paddings[p] : 0;
        a.out = d_out.as<long long>();
        k_low_targeting<<<NUM_SMS * 8, 256, 0, c->stream>>>(a);
        c->launches++;
        std::vector<long long> dense((size_t)rows * (size_t)nb);
        if (cudaGetLastError() != cudaSuccess || cudaMemcpyAsync(dense.data(), d_out.p, 8 * dense.size(), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = CRATAC_ERR_CUDA;
        if (rc == CRATAC_OK) rc = sync_status(c);
        if (rc == CRATAC_OK) to_columns(c, dense, rows, nb, out);
    }
    d_off.release(); d_rs.release(); d_re.release(); d_out.release();
    return rc;
}

// ------------------------------------------------------------------ REPORT_INSERT_SIZES
constexpr int IS_MAX = 1000;                 // stages/metrics/report_insert_sizes/__init__.py:32

__global__ void k_insert_sizes(const int32_t* __restrict__ chrom, const int32_t* __restrict__ start, const int32_t* __restrict__ end,
                               const int32_t* __restrict__ bc, const int32_t* __restrict__ slot_to_dense, const int32_t* __restrict__ row_of_dense,
                               const int32_t* __restrict__ primary, int exclude_non_primary, int64_t n, unsigned long long* __restrict__ table) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int b = bc[i];
        if (slot_to_dense) b = slot_to_dense[b];
        const int row = row_of_dense[b];
        if (row < 0) continue;
        if (exclude_non_primary && !primary[chrom[i]]) continue;
        const int size = end[i] - start[i];
        if (size > IS_MAX) atomicAdd(&table[(int64_t)row * (IS_MAX + 1) + IS_MAX], 1ULL);
        else if (size >= 1) atomicAdd(&table[(int64_t)row * (IS_MAX + 1) + size - 1], 1ULL);
    }
}

// cols[n_sel]: matrix columns (barcodes) wanted, table_out[n_sel][1001]
int insert_sizes(Ctx* c, int64_t n_sel, const int64_t* cols, const int32_t* h_primary, int exclude_non_primary, int64_t* table_out) {
    CRATAC_TRY(build_barcodes(c));
    const int64_t nb = c->n_barcodes;
    std::vector<int32_t> row_of_dense((size_t)std::max<int64_t>(nb, 1), -1);
    for (int64_t r = 0; r < n_sel; r++) {
        if (cols[r] < 0 || cols[r] >= nb) { set_error("column %lld out of range", (long long)cols[r]); return CRATAC_ERR_ARG; }
        row_of_dense[(size_t)c->h_c2d[(size_t)cols[r]]] = (int32_t)r;
    }
    DevBuf d_row, d_prim, d_tab;
    const size_t cells = (size_t)std::max<int64_t>(n_sel, 1) * (IS_MAX + 1);
    int rc = d_row.reserve(4 * row_of_dense.size());
    if (rc == CRATAC_OK) rc = d_prim.reserve(4 * (size_t)c->n_contigs);
    if (rc == CRATAC_OK) rc = d_tab.reserve(8 * cells);
    if (rc == CRATAC_OK) {
        cudaStream_t s = c->stream;
        if (cudaMemcpyAsync(d_row.p, row_of_dense.data(), 4 * row_of_dense.size(), cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaMemcpyAsync(d_prim.p, h_primary, 4 * (size_t)c->n_contigs, cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaMemsetAsync(d_tab.p, 0, 8 * cells, s) != cudaSuccess) rc = CRATAC_ERR_CUDA;
        if (rc == CRATAC_OK && c->n_fragments > 0 && n_sel > 0) {
            k_insert_sizes<<<NUM_SMS * 8, 256, 0, s>>>(c->d_chrom.as<int32_t>(), c->d_start.as<int32_t>(), c->d_end.as<int32_t>(), c->d_bc.as<int32_t>(),
                                                      c->bc_is_slot ? c->d_slot_to_dense.as<int32_t>() : nullptr, d_row.as<int32_t>(), d_prim.as<int32_t>(),
                                                      exclude_non_primary, c->n_fragments, d_tab.as<unsigned long long>());
            c->launches++;
            if (cudaGetLastError() != cudaSuccess) rc = CRATAC_ERR_CUDA;
        }
        if (rc == CRATAC_OK && n_sel > 0 && cudaMemcpyAsync(table_out, d_tab.p, 8 * (size_t)n_sel * (IS_MAX + 1), cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = CRATAC_ERR_CUDA;
        if (rc == CRATAC_OK) rc = sync_status(c);
    }
    d_row.release(); d_prim.release(); d_tab.release();
    return rc;
}

// ------------------------------------------------------------------ get_counts_by_barcode
// tracks, in this order: tss (padded 2000), ctcf (padded 250), dnase, enhancer, promoter, blacklist, peaks
enum { TR_TSS = 0, TR_CTCF, TR_DNASE, TR_ENHANCER, TR_PROMOTER, TR_BLACKLIST, TR_PEAKS, TR_COUNT };
// rows of the per-barcode result
enum { CB_PASSED = 0, CB_TOTAL, CB_DUPLICATE, CB_CUTSITES, CB_TSS, CB_DNASE, CB_ENHANCER, CB_PROMOTER, CB_ON_TARGET, CB_BLACKLIST, CB_PEAK_FRAGS, CB_FIXED };

struct CountsArgs {
    int64_t n;
    const int32_t *chrom, *start, *end, *bc, *dup;
    const int32_t* slot_to_dense;
    int n_contigs;
    int only_contig;                        // -1: all
    const int64_t* tr_off;                  // [TR_COUNT][n_contigs + 1]: regions of (track, contig); offsets into the arrays below
    const int32_t *r_start, *r_end;
    const int8_t* r_minus;                  // 1: strand '-'
    const int32_t* species_of_contig;       // null: single species
    int n_species;
    int64_t nb;
    unsigned long long* out;                // [CB_FIXED + 2 * n_species][nb], dense barcode order
    unsigned long long* relpos;             // [2][2 * rel_half + 1]: TSS, CTCF profiles; index = relative position + rel_half
    int rel_half;
    int64_t* status;
};

__device__ __forceinline__ int rel_position(int rs, int re, bool minus, int point) {
    const int width = re - rs - 1;
    const int half = width >> 1;            // Python's floor division, also for width = -1
    return minus ? (re - point) - half : (point - rs) - half;
}

__global__ void __launch_bounds__(256) k_counts_by_barcode(CountsArgs a) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = a.chrom[i];
        if (a.only_contig >= 0 && c != a.only_contig) continue;
        int b = a.bc[i];
        if (a.slot_to_dense) b = a.slot_to_dense[b];
        const int st = a.start[i], en = a.end[i], dups = a.dup[i];
        unsigned long long* o = a.out + b;
        atomicAdd(o + CB_PASSED * a.nb, 1ULL);
        atomicAdd(o + CB_TOTAL * a.nb, (unsigned long long)(long long)dups);
        atomicAdd(o + CB_DUPLICATE * a.nb, (unsigned long long)(long long)(dups - 1));
        bool ov[TR_COUNT];
#pragma unroll
        for (int t = 0; t < TR_COUNT; t++) {
            const int64_t r0 = a.tr_off[(int64_t)t * (a.n_contigs + 1) + c], r1 = a.tr_off[(int64_t)t * (a.n_contigs + 1) + c + 1];
            ov[t] = r1 > r0 && region_overlaps(a.r_start, a.r_end, r0, r1, st, en);
        }
        int cuts = 0;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int point = q ? en : st;
#pragma unroll
            for (int t = 0; t < 2; t++) {                       // TSS, CTCF profiles
                const int64_t r0 = a.tr_off[(int64_t)t * (a.n_contigs + 1) + c], r1 = a.tr_off[(int64_t)t * (a.n_contigs + 1) + c + 1];
                if (r1 <= r0) continue;
                const int64_t k = region_with_point(a.r_start, a.r_end, r0, r1, point);
                if (k < 0) continue;
                const int rel = rel_position(a.r_start[k], a.r_end[k], a.r_minus[k] != 0, point);
                if (rel < -a.rel_half || rel > a.rel_half) { raise_error(a.status, CRATAC_ERR_CAPACITY, -500 - t); continue; }
                atomicAdd(&a.relpos[(int64_t)t * (2 * a.rel_half + 1) + rel + a.rel_half], 1ULL);
            }
            const int64_t p0 = a.tr_off[(int64_t)TR_PEAKS * (a.n_contigs + 1) + c], p1 = a.tr_off[(int64_t)TR_PEAKS * (a.n_contigs + 1) + c + 1];
            if (p1 > p0 && region_with_point(a.r_start, a.r_end, p0, p1, point) >= 0) cuts++;
        }
        if (cuts) atomicAdd(o + CB_CUTSITES * a.nb, (unsigned long long)cuts);
        if (ov[TR_TSS]) atomicAdd(o + CB_TSS * a.nb, 1ULL);
        if (ov[TR_DNASE]) atomicAdd(o + CB_DNASE * a.nb, 1ULL);
        if (ov[TR_ENHANCER]) atomicAdd(o + CB_ENHANCER * a.nb, 1ULL);
        if (ov[TR_PROMOTER]) atomicAdd(o + CB_PROMOTER * a.nb, 1ULL);
        if (ov[TR_TSS] || ov[TR_DNASE] || ov[TR_ENHANCER] || ov[TR_PROMOTER]) atomicAdd(o + CB_ON_TARGET * a.nb, 1ULL);
        if (ov[TR_BLACKLIST]) atomicAdd(o + CB_BLACKLIST * a.nb, 1ULL);
        if (ov[TR_PEAKS]) atomicAdd(o + CB_PEAK_FRAGS * a.nb, 1ULL);
        if (a.species_of_contig) {
            const int sp = a.species_of_contig[c];
            if (sp >= 0 && sp < a.n_species) {
                atomicAdd(o + (int64_t)(CB_FIXED + sp) * a.nb, 1ULL);
                if (ov[TR_PEAKS]) atomicAdd(o + (int64_t)(CB_FIXED + a.n_species + sp) * a.nb, 1ULL);
            }
        }
    }
}

// track_off[TR_COUNT + 1]: regions of every track in r_* (inside a track: grouped by contig in reference order, tools/regions.py order
// inside a contig); out[CB_FIXED + 2 * n_species][n_barcodes] in matrix column order; relpos[2][2 * rel_half + 1]
int counts_by_barcode(Ctx* c, const int64_t* track_off, const int32_t* r_chrom, const int32_t* r_start, const int32_t* r_end, const int8_t* r_minus,
                      int only_contig, int n_species, const int32_t* species_of_contig, int rel_half, int64_t* out, int64_t* relpos) {
    CRATAC_TRY(build_barcodes(c));
    if (n_species < 0 || n_species > 8 || rel_half < 0) { set_error("cratac_counts_by_barcode: bad argument"); return CRATAC_ERR_ARG; }
    const int64_t n_reg = track_off[TR_COUNT];
    const int nc = c->n_contigs;
    std::vector<int64_t> off((size_t)TR_COUNT * (nc + 1), 0);
    for (int t = 0; t < TR_COUNT; t++) {
        int64_t* o = &off[(size_t)t * (nc + 1)];
        for (int64_t i = track_off[t]; i < track_off[t + 1]; i++) {
            if (r_chrom[i] < 0 || r_chrom[i] >= nc || (i > track_off[t] && r_chrom[i] < r_chrom[i - 1])) { set_error("track %d: regions must be grouped by contig in reference order", t); return CRATAC_ERR_ARG; }
            o[r_chrom[i] + 1]++;
        }
        o[0] = track_off[t];
        for (int k = 0; k < nc; k++) o[k + 1] += o[k];
    }
    const int64_t nb = c->n_barcodes;
    const int rows = CB_FIXED + 2 * n_species;
    const size_t nr = (size_t)std::max<int64_t>(n_reg, 1), nrel = 2 * (size_t)(2 * rel_half + 1);
    DevBuf d_off, d_rs, d_re, d_rm, d_sp, d_out, d_rel;
    int rc = d_off.reserve(8 * off.size());
    if (rc == CRATAC_OK) rc = d_rs.reserve(4 * nr);
    if (rc == CRATAC_OK) rc = d_re.reserve(4 * nr);
    if (rc == CRATAC_OK) rc = d_rm.reserve(nr);
    if (rc == CRATAC_OK) rc = d_sp.reserve(4 * (size_t)nc);
    if (rc == CRATAC_OK) rc = d_out.reserve(8 * (size_t)rows * (size_t)std::max<int64_t>(nb, 1));
    if (rc == CRATAC_OK) rc = d_rel.reserve(8 * nrel);
    cudaStream_t s = c->stream;
    if (rc == CRATAC_OK) {
        bool ok = cudaMemcpyAsync(d_off.p, off.data(), 8 * off.size(), cudaMemcpyHostToDevice, s) == cudaSuccess;
        if (n_reg > 0) {
            ok = ok && cudaMemcpyAsync(d_rs.p, r_start, 4 * (size_t)n_reg, cudaMemcpyHostToDevice, s) == cudaSuccess;
            ok = ok && cudaMemcpyAsync(d_re.p, r_end, 4 * (size_t)n_reg, cudaMemcpyHostToDevice, s) == cudaSuccess;
            ok = ok && cudaMemcpyAsync(d_rm.p, r_minus, (size_t)n_reg, cudaMemcpyHostToDevice, s) == cudaSuccess;
        }
        if (species_of_contig) ok = ok && cudaMemcpyAsync(d_sp.p, species_of_contig, 4 * (size_t)nc, cudaMemcpyHostToDevice, s) == cudaSuccess;
        ok = ok && cudaMemsetAsync(d_out.p, 0, 8 * (size_t)rows * (size_t)std::max<int64_t>(nb, 1), s) == cudaSuccess;
        ok = ok && cudaMemsetAsync(d_rel.p, 0, 8 * nrel, s) == cudaSuccess;
        if (!ok) rc = CRATAC_ERR_CUDA;
    }
    if (rc == CRATAC_OK && c->n_fragments > 0) {
        CountsArgs a;
        a.n = c->n_fragments; a.chrom = c->d_chrom.as<int32_t>(); a.start = c->d_start.as<int32_t>(); a.end = c->d_end.as<int32_t>();
        a.bc = c->d_bc.as<int32_t>(); a.dup = c->d_dup.as<int32_t>(); a.slot_to_dense = c->bc_is_slot ? c->d_slot_to_dense.as<int32_t>() : nullptr;
        a.n_contigs = nc; a.only_contig = only_contig;
        a.tr_off = d_off.as<int64_t>(); a.r_start = d_rs.as<int32_t>(); a.r_end = d_re.as<int32_t>(); a.r_minus = d_rm.as<int8_t>();
        a.species_of_contig = species_of_contig ? d_sp.as<int32_t>() : nullptr; a.n_species = n_species;
        a.nb = nb; a.out = d_out.as<unsigned long long>(); a.relpos = d_rel.as<unsigned long long>(); a.rel_half = rel_half;
        a.status = c->d_status.as<int64_t>();
        k_counts_by_barcode<<<NUM_SMS * 8, 256, 0, s>>>(a);
        c->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = CRATAC_ERR_CUDA;
    }
    if (rc == CRATAC_OK) {
        std::vector<long long> dense((size_t)rows * (size_t)std::max<int64_t>(nb, 1));
        if (cudaMemcpyAsync(dense.data(), d_out.p, 8 * dense.size(), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaMemcpyAsync(relpos, d_rel.p, 8 * nrel, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = CRATAC_ERR_CUDA;
        if (rc == CRATAC_OK) rc = sync_status(c);
        if (rc == CRATAC_OK && nb > 0) to_columns(c, dense, rows, nb, out);
    }
    d_off.release(); d_rs.release(); d_re.release(); d_rm.release(); d_sp.release(); d_out.release(); d_rel.release();
    return rc;
}

}  // namespace cratac
