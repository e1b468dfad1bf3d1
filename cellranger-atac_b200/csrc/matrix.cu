// K6-K8: fragment ends x peaks, counted per barcode, into canonical CSC.  Replaces
// generate_peak_matrix/__init__.py:107-119 (per fragment: two bisects in tenkit Regions
// (tenkit/lib/python/tenkit/regions.py:136-145), Counter[(barcode, peak)] += 1, COO -> CSC) for ALL
// barcodes in one pass instead of 30 barcode chunks that each re-read the file.
//
//   sizes        per-barcode segment capacity = 2 x the barcode's fragments (counted by the decoder, or by
//                k_count_frags for fragments handed over as arrays); scan -> segment offsets.  No count pass.
//   K6   overlap per fragment: upper_bound(peak starts, pos) - 1, half-open containment test, for pos in
//                (start, stop) -- stop used as is, dup_count ignored; ONE record per (fragment, peak)
//                (row << 1 | both-ends: a fragment with both ends in the same peak is one record worth 2),
//                appended to the barcode's segment through a per-barcode cursor.  Tiles of 2048 start-sorted
//                fragments search a peak window staged in shared memory (found by a warp-wide 33-ary search).
//   K7   reduce  per barcode segment [0, cursor): sort the records (one warp for <= 1024, a CTA for <= 4096: bitonic in
//                shared memory; range-partitioned counting in shared memory for longer ones) and run-length
//                reduce to (row, count).
//   K8   build   indptr = scan of distinct counts in column order; copy (row, count) segments to
//                indices int64 / data int32, chunked over the output entries.
//   select_columns (FILTER_PEAK_MATRIX) and merge_rank_csc (multi-GPU) reuse the chunked copy.
#include <algorithm>

#include "common.cuh"

namespace cratac {

constexpr int MX_THREADS = 256;
constexpr int SORT_SMALL = 4096;        // bitonic path capacity
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARP = 1024;         // segments up to this length are sorted by a single warp
constexpr int COUNT_RANGE = 16384;      // counter words per pass on the dense path (64 KiB; 8-, 16- or 32-bit counters in them)
constexpr int COUNT_THREADS = 512;

struct OverlapArgs {
    int64_t n;
    const int32_t *chrom, *start, *end, *bc;
    const int32_t* slot_to_dense;          // null: bc already dense
    const int64_t* peak_off;               // per contig
    const int32_t *pk_start, *pk_end, *pk_row;
    const int64_t* seg_off;                // per dense barcode: start of its segment (capacity 2 x its fragments)
    unsigned int* cursor;
    int32_t* hits;
    long long maxpos, maxneg;              // bounds of end - start over all fragments
    int64_t* n_hits;                       // device counter of fragment ends inside peaks
};

__device__ __forceinline__ int find_peak(const OverlapArgs& a, int c, int pos) {
    int64_t lo = a.peak_off[c], hi = a.peak_off[c + 1];
    const int64_t base = lo;
    while (lo < hi) {                       // bisect.bisect(starts, pos): first start > pos
        int64_t mid = (lo + hi) >> 1;
        if (pos < __ldg(&a.pk_start[mid])) hi = mid; else lo = mid + 1;
    }
    int64_t idx = lo - 1;
    if (idx < base) return -1;
    if (pos < __ldg(&a.pk_end[idx])) return __ldg(&a.pk_row[idx]);   // pos >= start holds by construction
    return -1;
}

// first index in [lo, hi) whose value exceeds key (hi if none), searched by a whole warp: 32 split points per round
__device__ __forceinline__ int64_t warp_upper_bound(const int32_t* __restrict__ arr, int64_t lo, int64_t hi, long long key) {
    const int lane = threadIdx.x & 31;
    while (hi - lo > 32) {
        const int64_t step = (hi - lo) / 33 + 1;
        const int64_t p = lo + step * (lane + 1) - 1;                     // probes ascend with the lane
        const bool gt = p < hi ? key < (long long)__ldg(&arr[p]) : true;
        const unsigned m = __ballot_sync(0xffffffffu, gt);
        if (m == 0u) { lo = lo + step * 32; continue; }                  // every probe <= key: the answer lies beyond the last one
        const int j = __ffs(m) - 1;                                      // first probe with a larger value
        const int64_t pj = lo + step * (j + 1) - 1;
        const int64_t nlo = j > 0 ? lo + step * j : lo;                  // one past the previous probe
        hi = pj < hi ? pj : hi;                                          // arr[pj] > key (or pj is past the end): answer <= pj
        lo = nlo;
    }
    const int64_t p = lo + lane;
    const bool gt = p < hi ? key < (long long)__ldg(&arr[p]) : true;
    const unsigned m = __ballot_sync(0xffffffffu, gt);
    return m ? lo + (__ffs(m) - 1) : hi;
}

// Fragments are start-sorted, so the OV_TILE fragments a CTA takes at a time touch a handful of consecutive
// peaks: the CTA finds that window with two bisects, stages it in shared memory and every fragment end is
// then located there.  Tiles that straddle a contig boundary or a very dense peak region fall back to the
// global bisect.
constexpr int OV_TILE = 2048;
constexpr int OV_WIN = 1024;

__global__ void __launch_bounds__(MX_THREADS) k_overlap(OverlapArgs a) {
    __shared__ int s_start[OV_WIN], s_end[OV_WIN], s_row[OV_WIN];
    __shared__ long long s_plo;
    __shared__ int s_w;
    const int64_t n_tiles = (a.n + OV_TILE - 1) / OV_TILE;
    long long my_hits = 0;                                  // fragment ends inside peaks (K of the size accounting)
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i0 = tile * OV_TILE;
        const int64_t i1 = i0 + OV_TILE < a.n ? i0 + OV_TILE : a.n;
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp 0 finds the tile's peak window with two 33-ary searches (every lane probes one split point per
            // round: 3 rounds of dependent loads for a contig's few thousand peaks instead of 13 for a bisect)
            int w = -1;
            const int c0 = a.chrom[i0];
            if (c0 == a.chrom[i1 - 1]) {
                const long long lo_pos = (long long)a.start[i0] - a.maxneg;
                const long long hi_pos = (long long)a.start[i1 - 1] + a.maxpos;
                const int64_t base = a.peak_off[c0], top = a.peak_off[c0 + 1];
                const int64_t lo = warp_upper_bound(a.pk_start, base, top, lo_pos);      // first peak start > lo_pos
                const int64_t p_lo = lo > base ? lo - 1 : base;
                const int64_t hi = warp_upper_bound(a.pk_start, p_lo, top, hi_pos);      // first peak start > hi_pos
                if (hi - p_lo <= OV_WIN) { w = (int)(hi - p_lo); if (threadIdx.x == 0) s_plo = p_lo; }
            }
            if (threadIdx.x == 0) s_w = w;
        }
        __syncthreads();
        const int w = s_w;
        if (w > 0) {
            const int64_t p_lo = s_plo;
            for (int k = threadIdx.x; k < w; k += MX_THREADS) {
                s_start[k] = a.pk_start[p_lo + k]; s_end[k] = a.pk_end[p_lo + k]; s_row[k] = a.pk_row[p_lo + k];
            }
        }
        __syncthreads();
        // two fragments per thread and round: their bisections interleave and both cursor atomics are in flight together
        // (the kernel waits on returning L2 atomics: 24 % issue at full occupancy, profiles/r02e_kernels_full_250m.csv)
        for (int64_t ia = i0 + threadIdx.x; ia < i1; ia += 2 * MX_THREADS) {
            int64_t idx[2] = {ia, ia + MX_THREADS};
            int b[2], r0[2], r1[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                b[u] = -1; r0[u] = -1; r1[u] = -1;
                const int64_t i = idx[u];
                if (i >= i1) continue;
                b[u] = a.bc[i];                               // table slot or dense id: the cursors are indexed by whichever it is
                if (w >= 0) {
                    int pos[2] = {a.start[i], a.end[i]};
                    int r[2];
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        int lo = 0, hi = w;                   // first staged start > pos
                        while (lo < hi) { int mid = (lo + hi) >> 1; if (pos[q] < s_start[mid]) hi = mid; else lo = mid + 1; }
                        r[q] = (lo > 0 && pos[q] < s_end[lo - 1]) ? s_row[lo - 1] : -1;
                    }
                    r0[u] = r[0]; r1[u] = r[1];
                } else {
                    const int c = a.chrom[i];
                    r0[u] = find_peak(a, c, a.start[i]);
                    r1[u] = find_peak(a, c, a.end[i]);
                }
            }
            // one record per (fragment, peak): value = row << 1 | (both ends in that peak).  Most fragments that hit
            // a peak lie inside it, so this nearly halves the scattered 4-byte stores that bound the scatter pass.
            bool same[2];
            int nrec[2];
            unsigned int p[2] = {0, 0};
            int64_t so[2] = {0, 0};
#pragma unroll
            for (int u = 0; u < 2; u++) {
                same[u] = r0[u] >= 0 && r1[u] == r0[u];
                nrec[u] = (r0[u] >= 0) + (r1[u] >= 0 && !same[u]);
                if (b[u] < 0) nrec[u] = 0;
                if (nrec[u]) {
                    my_hits += (r0[u] >= 0) + (r1[u] >= 0);
                    p[u] = atomicAdd(&a.cursor[b[u]], (unsigned)nrec[u]);
                    so[u] = a.seg_off[b[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (!nrec[u]) continue;
                int64_t o = so[u] + p[u];
                if (r0[u] >= 0) a.hits[o++] = (r0[u] << 1) | (int)same[u];
                if (r1[u] >= 0 && !same[u]) a.hits[o] = r1[u] << 1;
            }
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) my_hits += __shfl_xor_sync(0xffffffffu, my_hits, d);
    if ((threadIdx.x & 31) == 0 && my_hits) atomicAdd(reinterpret_cast<unsigned long long*>(a.n_hits), (unsigned long long)my_hits);
}

// the overlap kernel indexes cursors and segment offsets by what bc[] holds.  After a decode that is the barcode's hash-table
// slot: mapping it to the dense id first would put one more dependent L2 access in front of every cursor atomic.
__global__ void k_slot_segments(const int32_t* __restrict__ slot_to_dense, int cap, const int64_t* __restrict__ seg_off, int64_t* __restrict__ slot_seg_off,
                                int32_t* __restrict__ dense_slot) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap) return;
    const int d = slot_to_dense[s];
    if (d < 0) return;
    slot_seg_off[s] = seg_off[d];
    dense_slot[d] = s;
}
__global__ void k_slot_cursors(const unsigned int* __restrict__ slot_cursor, const int32_t* __restrict__ dense_slot, int64_t nb, unsigned int* __restrict__ cursor) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d < nb) cursor[d] = slot_cursor[dense_slot[d]];
}

// segment capacities: a barcode with f fragments has at most 2 f hit records
__global__ void k_seg_caps(const int32_t* __restrict__ nfrag, int64_t nb, int32_t* __restrict__ caps) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) caps[b] = 2 * nfrag[b];
}
// fragments per dense barcode for fragments that did not come through the decoder (cratac_set_fragments)
__global__ void k_count_frags(const int32_t* __restrict__ bc, int64_t n, int64_t nb, int32_t* __restrict__ nfrag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = bc[i];
        if (b >= 0 && b < nb) atomicAdd(&nfrag[b], 1);
    }
}

// ---- K7: per-barcode sort + run-length reduce -----------------------------------------------------
// medium segments (SORT_WARP < records <= SORT_SMALL): bitonic sort in shared memory; one CTA per barcode
__global__ void __launch_bounds__(SORT_THREADS) k_reduce_small(const int32_t* __restrict__ hits, const int64_t* __restrict__ seg_off, const unsigned int* __restrict__ seg_len,
                                                                int64_t n_bc, int32_t* __restrict__ out_row, int32_t* __restrict__ out_cnt,
                                                                int32_t* __restrict__ nnz_col) {
    __shared__ int s[SORT_SMALL];
    __shared__ int warp_tot[SORT_THREADS / 32];
    __shared__ int s_base;
    for (int64_t b = blockIdx.x; b < n_bc; b += gridDim.x) {
        const int64_t o0 = seg_off[b];
        const int64_t len64 = seg_len[b];
        if (len64 > SORT_SMALL || len64 <= SORT_WARP) continue;   // long path / warp path
        const int len = (int)len64;
        int m = 32;
        while (m < len) m <<= 1;
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += SORT_THREADS) s[i] = i < len ? hits[o0 + i] : 0x7fffffff;
        __syncthreads();
        for (int k = 2; k <= m; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < m; i += SORT_THREADS) {
                    int ixj = i ^ j;
                    if (ixj > i) {
                        int x = s[i], y = s[ixj];
                        bool up = (i & k) == 0;
                        if ((x > y) == up) { s[i] = y; s[ixj] = x; }
                    }
                }
                __syncthreads();
            }
        }
        // run-length reduce, in order
        if (threadIdx.x == 0) s_base = 0;
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int i0 = 0; i0 < len; i0 += SORT_THREADS) {
            int i = i0 + threadIdx.x;
            int head = (i < len) && (i == 0 || (s[i] >> 1) != (s[i - 1] >> 1));
            int x = head;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
            if (lane == 31) warp_tot[warp] = x;
            __syncthreads();
            int off = s_base;
            for (int w = 0; w < warp; w++) off += warp_tot[w];
            if (head) {
                const int row = s[i] >> 1;
                int e = i, cnt = 0;
                while (e < len && (s[e] >> 1) == row) { cnt += 1 + (s[e] & 1); e++; }
                out_row[o0 + off + x - 1] = row;
                out_cnt[o0 + off + x - 1] = cnt;
            }
            __syncthreads();
            if (threadIdx.x == SORT_THREADS - 1) s_base = off + x;
            __syncthreads();
        }
        if (threadIdx.x == 0) nnz_col[b] = s_base;
    }
}

// very short segments (most barcodes of a library are background with a few dozen hits): one warp per barcode,
// bitonic sort in the warp's own slice of shared memory, no block barriers
__global__ void __launch_bounds__(SORT_THREADS) k_reduce_warp(const int32_t* __restrict__ hits, const int64_t* __restrict__ seg_off, const unsigned int* __restrict__ seg_len,
                                                               int64_t n_bc, int32_t* __restrict__ out_row, int32_t* __restrict__ out_cnt,
                                                               int32_t* __restrict__ nnz_col) {
    __shared__ int s_all[(SORT_THREADS / 32) * SORT_WARP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* s = s_all + warp * SORT_WARP;
    const int64_t n_warps = (int64_t)gridDim.x * (SORT_THREADS / 32);
    for (int64_t b = (int64_t)blockIdx.x * (SORT_THREADS / 32) + warp; b < n_bc; b += n_warps) {
        const int64_t o0 = seg_off[b];
        const int64_t len64 = seg_len[b];
        if (len64 > SORT_WARP) continue;
        const int len = (int)len64;
        if (len == 0) { if (lane == 0) nnz_col[b] = 0; continue; }
        int m = 32;
        while (m < len) m <<= 1;
        __syncwarp();
        for (int i = lane; i < m; i += 32) s[i] = i < len ? hits[o0 + i] : 0x7fffffff;
        __syncwarp();
        for (int k = 2; k <= m; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                // m / 2 compare-exchanges: pair index q -> element i with bit j clear
                for (int q = lane; q < (m >> 1); q += 32) {
                    const int i = ((q & ~(j - 1)) << 1) | (q & (j - 1));
                    const int x = s[i], y = s[i | j];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { s[i] = y; s[i | j] = x; }
                }
                __syncwarp();
            }
        }
        int base = 0;
        for (int i0 = 0; i0 < len; i0 += 32) {
            const int i = i0 + lane;
            const int head = (i < len) && (i == 0 || (s[i] >> 1) != (s[i - 1] >> 1));
            const unsigned hm = __ballot_sync(0xffffffffu, head);
            if (head) {
                const int row = s[i] >> 1;
                int e = i, cnt = 0;
                while (e < len && (s[e] >> 1) == row) { cnt += 1 + (s[e] & 1); e++; }
                const int pos = base + __popc(hm & ((1u << lane) - 1u));
                out_row[o0 + pos] = row;
                out_cnt[o0 + pos] = cnt;
            }
            base += __popc(hm);
        }
        if (lane == 0) nnz_col[b] = base;
    }
}

// long segments: counting over ranges of peak rows whose counters are held in shared memory.  The counters are packed
// PACK to a 32-bit word and a pass covers PACK * COUNT_RANGE rows: 8-bit counters first (a (barcode, peak) count above
// 255 is rare: the whole GRCh38 matrix is one pass instead of three with 16-bit counters), and a segment in which one
// overflows is counted again with 16-bit (fewer than 32768 records, each worth 1 or 2: cannot overflow) or 32-bit ones.
__global__ void __launch_bounds__(COUNT_THREADS, 3) k_reduce_large(const int32_t* __restrict__ hits, const int64_t* __restrict__ seg_off, const unsigned int* __restrict__ seg_len,
                                                                 int64_t n_bc, const int64_t* __restrict__ status, int64_t n_peaks_host,
                                                                 int32_t* __restrict__ out_row, int32_t* __restrict__ out_cnt,
                                                                 int32_t* __restrict__ nnz_col) {
    extern __shared__ int cnt[];           // COUNT_RANGE counter words, then 4 * COUNT_RANGE / 32 bitmap words
    unsigned* bm = reinterpret_cast<unsigned*>(cnt + COUNT_RANGE);
    __shared__ int warp_tot[COUNT_THREADS / 32];
    __shared__ int s_base, s_overflow;
    const int64_t n_peaks = n_peaks_host >= 0 ? n_peaks_host : status[ST_N_PEAKS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // counters and bitmap start clean and are cleaned again as they are emitted: no clearing pass per range
    for (int i = threadIdx.x; i < COUNT_RANGE + 4 * COUNT_RANGE / 32; i += COUNT_THREADS) cnt[i] = 0;
    if (threadIdx.x == 0) s_overflow = 0;
    for (int64_t b = blockIdx.x; b < n_bc; b += gridDim.x) {
        const int64_t o0 = seg_off[b];
        const int64_t len = seg_len[b];
        if (len <= SORT_SMALL) continue;
        if (len > 0x7fffffffLL) { if (threadIdx.x == 0) raise_error(const_cast<int64_t*>(status), CRATAC_ERR_CAPACITY, b); continue; }
        int lg = 2;                            // log2(PACK): 8-bit counters first
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) s_base = 0;
            const int bits = 32 >> lg;
            const unsigned fmask = lg ? (1u << bits) - 1u : 0xFFFFFFFFu;
            const int range = COUNT_RANGE << lg;
            bool overflow = false;
            for (int64_t r0 = 0; r0 < n_peaks && !overflow; r0 += range) {
                // count; the first hit of a row also sets its bit in the occupancy bitmap.  Eight loads in flight per
                // thread: the loop is bound by the latency of the global loads otherwise.
                constexpr int UN = 8;
                const int32_t* __restrict__ seg = hits + o0;
                const int len32 = (int)len, r0_32 = (int)r0;
                bool over = false;
                for (int i0 = threadIdx.x; i0 < len32; i0 += UN * COUNT_THREADS) {
                    unsigned r[UN];
                    int inc[UN];
#pragma unroll
                    for (int u = 0; u < UN; u++) {
                        const int i = i0 + u * COUNT_THREADS;
                        const int v = i < len32 ? seg[i] : -1;
                        r[u] = v >= 0 ? (unsigned)((v >> 1) - r0_32) : 0xFFFFFFFFu;
                        inc[u] = 1 + (v & 1);
                    }
                    unsigned old[UN];
#pragma unroll
                    for (int u = 0; u < UN; u++) {
                        old[u] = 1u;
                        if (r[u] < (unsigned)range) {
                            const int sh = bits * (int)(r[u] & ((1u << lg) - 1u));
                            old[u] = ((unsigned)atomicAdd(&cnt[r[u] >> lg], inc[u] << sh) >> sh) & fmask;
                            // the first overflow of a word sees an intact field: enough to notice that one happened
                            over |= lg == 2 && old[u] + (unsigned)inc[u] > 255u;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < UN; u++)
                        if (old[u] == 0u) atomicOr(&bm[r[u] >> 5], 1u << (r[u] & 31u));
                }
                if (over) s_overflow = 1;
                __syncthreads();
                overflow = s_overflow != 0;
                if (overflow) break;
                // emit the set bits in row order
                for (int w0 = 0; w0 < range / 32; w0 += COUNT_THREADS) {
                    const int wi = w0 + threadIdx.x;
                    unsigned word = wi < range / 32 ? bm[wi] : 0u;
                    const unsigned word0 = word;
                    int c = __popc(word);
                    int x = c;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
                    if (lane == 31) warp_tot[warp] = x;
                    __syncthreads();
                    int off = s_base;
                    for (int w = 0; w < warp; w++) off += warp_tot[w];
                    int64_t pos = o0 + off + x - c;
                    while (word) {
                        const int bit = __ffs(word) - 1;
                        word &= word - 1;
                        const int row = wi * 32 + bit;
                        out_row[pos] = (int32_t)(r0 + row);
                        out_cnt[pos] = (int)(((unsigned)cnt[row >> lg] >> (bits * (row & ((1 << lg) - 1)))) & fmask);
                        pos++;
                    }
                    if (word0) {   // this thread owns the 32 rows of its bitmap word (whole counter words): clean up behind it
                        bm[wi] = 0u;
                        for (int k = 0; k < (32 >> lg); k++) cnt[((wi * 32) >> lg) + k] = 0;
                    }
                    __syncthreads();
                    if (threadIdx.x == COUNT_THREADS - 1) s_base = off + x;
                    __syncthreads();
                }
            }
            if (!overflow) break;
            // an 8-bit counter overflowed: wipe everything and count this segment again with wider counters
            __syncthreads();
            for (int i = threadIdx.x; i < COUNT_RANGE + 4 * COUNT_RANGE / 32; i += COUNT_THREADS) cnt[i] = 0;
            if (threadIdx.x == 0) s_overflow = 0;
            lg = len < 32768 ? 1 : 0;
        }
        if (threadIdx.x == 0) nnz_col[b] = s_base;
    }
}

// ---- K8 -------------------------------------------------------------------------------------------
__global__ void k_gather_nnz(const int32_t* __restrict__ nnz_col, const int32_t* __restrict__ col_to_dense, int64_t n_cols, int32_t* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cols) return;
    int64_t d = col_to_dense ? col_to_dense[j] : j;
    out[j] = d >= 0 ? nnz_col[d] : 0;       // d < 0: the barcode of this (global) column does not occur in this shard
}

// One CTA per CSC_CHUNK consecutive output entries (not per column: a cell's column is 100x a background
// barcode's).  The columns a chunk touches are found by two binary searches in indptr; a chunk inside few
// columns is copied by the whole CTA column after column, a chunk over many short columns one warp per column.
constexpr int CSC_CHUNK = 4096;
__global__ void __launch_bounds__(256) k_csc_copy(const int32_t* __restrict__ out_row, const int32_t* __restrict__ out_cnt, const int64_t* __restrict__ seg_off,
                                                   const int32_t* __restrict__ col_to_dense, const int64_t* __restrict__ indptr, int64_t n_cols, int64_t nnz,
                                                   int64_t* __restrict__ indices, int32_t* __restrict__ data) {
    __shared__ int64_t s_j[2];
    const int64_t p0 = (int64_t)blockIdx.x * CSC_CHUNK;
    const int64_t p1 = p0 + CSC_CHUNK < nnz ? p0 + CSC_CHUNK : nnz;
    if (threadIdx.x < 2) {       // largest j < n_cols with indptr[j] <= target
        const int64_t target = threadIdx.x == 0 ? p0 : p1 - 1;
        int64_t lo = 0, hi = n_cols - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (indptr[mid] <= target) lo = mid; else hi = mid - 1;
        }
        s_j[threadIdx.x] = lo;
    }
    __syncthreads();
    const int64_t j0 = s_j[0], j1 = s_j[1];
    const bool by_warp = j1 - j0 >= 8;
    const int step = by_warp ? 8 : 1, width = by_warp ? 32 : 256;
    const int me = by_warp ? (threadIdx.x & 31) : threadIdx.x;
    for (int64_t j = j0 + (by_warp ? (threadIdx.x >> 5) : 0); j <= j1; j += step) {
        const int64_t c0 = indptr[j], c1 = indptr[j + 1];
        const int64_t a = c0 > p0 ? c0 : p0, b = c1 < p1 ? c1 : p1;
        if (b <= a) continue;
        const int64_t d = col_to_dense ? col_to_dense[j] : j;
        const int64_t src = seg_off[d] + (a - c0);
        const int len = (int)(b - a);
        // destination-aligned groups of four entries: two 16-byte stores of row indices (widened to int64) and one of counts;
        // the source segment starts wherever the reduce left it, so its loads stay scalar
        const int head = (int)(((a + 3) & ~(int64_t)3) - a) < len ? (int)(((a + 3) & ~(int64_t)3) - a) : len;
        const int groups = (len - head) >> 2;
        for (int k = me; k < head; k += width) {
            indices[a + k] = (int64_t)out_row[src + k];
            data[a + k] = out_cnt[src + k];
        }
        for (int g = me; g < groups; g += width) {
            const int k = head + 4 * g;
            const int r0 = out_row[src + k], r1 = out_row[src + k + 1], r2 = out_row[src + k + 2], r3 = out_row[src + k + 3];
            const int4 cn = make_int4(out_cnt[src + k], out_cnt[src + k + 1], out_cnt[src + k + 2], out_cnt[src + k + 3]);
            longlong2* ip = reinterpret_cast<longlong2*>(indices + a + k);
            ip[0] = make_longlong2((long long)r0, (long long)r1);
            ip[1] = make_longlong2((long long)r2, (long long)r3);
            *reinterpret_cast<int4*>(data + a + k) = cn;
        }
        for (int k = head + 4 * groups + me; k < len; k += width) {
            indices[a + k] = (int64_t)out_row[src + k];
            data[a + k] = out_cnt[src + k];
        }
    }
}

int build_barcodes(Ctx* c);   // decode.cu

// ---- multi-GPU: column blocks of the shards (same global columns, disjoint ascending row ranges) -> one CSC
__global__ void k_merge_counts(const int64_t* __restrict__ indptr_all, int world, int64_t n_cols, int32_t* __restrict__ cnt) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cols) return;
    int64_t t = 0;
    for (int r = 0; r < world; r++) { const int64_t* ip = indptr_all + (int64_t)r * (n_cols + 1); t += ip[j + 1] - ip[j]; }
    cnt[j] = (int32_t)t;
}
// chunked over the merged output like k_csc_copy; inside a column the rank segments follow each other in rank
// order (= genome order = ascending rows)
__global__ void __launch_bounds__(256) k_merge_copy(const int64_t* __restrict__ indptr_all, const int64_t* __restrict__ rank_off, int world, int64_t n_cols,
                                                    const int64_t* __restrict__ indices_all, const int32_t* __restrict__ data_all,
                                                    const int64_t* __restrict__ indptr, int64_t nnz, int64_t* __restrict__ indices, int32_t* __restrict__ data) {
    __shared__ int64_t s_j[2];
    const int64_t p0 = (int64_t)blockIdx.x * CSC_CHUNK;
    const int64_t p1 = p0 + CSC_CHUNK < nnz ? p0 + CSC_CHUNK : nnz;
    if (threadIdx.x < 2) {       // largest j < n_cols with indptr[j] <= target
        const int64_t target = threadIdx.x == 0 ? p0 : p1 - 1;
        int64_t lo = 0, hi = n_cols - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (indptr[mid] <= target) lo = mid; else hi = mid - 1;
        }
        s_j[threadIdx.x] = lo;
    }
    __syncthreads();
    const int64_t j0 = s_j[0], j1 = s_j[1];
    const bool by_warp = j1 - j0 >= 8;
    const int step = by_warp ? 8 : 1, width = by_warp ? 32 : 256;
    const int me = by_warp ? (threadIdx.x & 31) : threadIdx.x;
    for (int64_t j = j0 + (by_warp ? (threadIdx.x >> 5) : 0); j <= j1; j += step) {
        int64_t dst = indptr[j];
        if (indptr[j + 1] <= p0 || dst >= p1) continue;
        for (int r = 0; r < world; r++) {
            const int64_t* ip = indptr_all + (int64_t)r * (n_cols + 1);
            const int64_t s0 = ip[j], len = ip[j + 1] - s0;
            const int64_t a = dst > p0 ? dst : p0, b = dst + len < p1 ? dst + len : p1;
            if (b > a) {
                const int64_t src = rank_off[r] + s0 + (a - dst);
                const int n = (int)(b - a);
                for (int k = me; k < n; k += width) { indices[a + k] = indices_all[src + k]; data[a + k] = data_all[src + k]; }
            }
            dst += len;
        }
    }
}

int merge_rank_csc(Ctx* c, int world, int64_t n_cols, const int64_t* d_indptr_all, const int64_t* d_rank_off, const int64_t* d_indices_all,
                   const int32_t* d_data_all, int64_t total_nnz) {
    cudaStream_t s = c->stream;
    size_t ncc = (size_t)std::max<int64_t>(n_cols, 1);
    CRATAC_TRY(c->d_indptr.reserve(8 * (ncc + 1)));
    CRATAC_TRY(c->d_nnz_ordered.reserve(4 * ncc));
    size_t zz = (size_t)std::max<int64_t>(total_nnz, 1);
    CRATAC_TRY(c->d_indices.reserve(8 * zz)); CRATAC_TRY(c->d_data.reserve(4 * zz));
    if (n_cols > 0) {
        k_merge_counts<<<(unsigned)((n_cols + 255) / 256), 256, 0, s>>>(d_indptr_all, world, n_cols, c->d_nnz_ordered.as<int32_t>());
        c->launches++;
    }
    CRATAC_TRY(exclusive_scan_i32_to_i64(c, c->d_nnz_ordered.as<int32_t>(), c->d_indptr.as<int64_t>(), n_cols, c->d_indptr.as<int64_t>() + n_cols));
    if (n_cols > 0) {
        if (total_nnz > 0)
            k_merge_copy<<<(unsigned)((total_nnz + CSC_CHUNK - 1) / CSC_CHUNK), 256, 0, s>>>(d_indptr_all, d_rank_off, world, n_cols, d_indices_all, d_data_all,
                                                                                            c->d_indptr.as<int64_t>(), total_nnz, c->d_indices.as<int64_t>(),
                                                                                            c->d_data.as<int32_t>());
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
    }
    CRATAC_CUDA(cudaStreamSynchronize(s));
    c->nnz = total_nnz;
    c->n_cols = n_cols;
    return CRATAC_OK;
}

// K6 + K7: hit records and their per-barcode reduction.  Needs the peaks (and their row offset) but not the matrix columns.
int peak_matrix_hits(Ctx* c) {
    cudaStream_t s = c->stream;
    int64_t* st = c->d_status.as<int64_t>();
    const int64_t n = c->n_fragments, nb = c->n_barcodes;
    size_t nbb = (size_t)std::max<int64_t>(nb, 1);
    CRATAC_TRY(c->d_bc_hits.reserve(4 * nbb));                  // segment capacities
    CRATAC_TRY(c->d_seg_off.reserve(8 * (nbb + 1)));
    CRATAC_TRY(c->d_cursor.reserve(4 * nbb));
    CRATAC_TRY(c->d_nnz_col.reserve(4 * nbb * 2));
    CRATAC_TRY(c->d_indptr.reserve(8 * (nbb + 1)));
    CRATAC_CUDA(cudaMemsetAsync(c->d_cursor.p, 0, 4 * nbb, s));
    // Per-barcode segments are sized from the barcode's fragment count (at most 2 records per fragment; the decoder
    // counts fragments per barcode as it inserts them), so the hits are placed in ONE pass over the fragments: no count pass
    // and no host round trip for the number of hits.
    if (!c->dense_nfrag_valid) {
        CRATAC_TRY(c->d_dense_nfrag.reserve(4 * nbb));
        CRATAC_CUDA(cudaMemsetAsync(c->d_dense_nfrag.p, 0, 4 * nbb, s));
        if (n > 0) {
            k_count_frags<<<NUM_SMS * 8, 256, 0, s>>>(c->d_bc.as<int32_t>(), n, nb, c->d_dense_nfrag.as<int32_t>());
            c->launches++;
        }
        c->dense_nfrag_valid = true;
    }
    int32_t* caps = c->d_bc_hits.as<int32_t>();
    if (nb > 0) {
        k_seg_caps<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(c->d_dense_nfrag.as<int32_t>(), nb, caps);
        c->launches++;
    }
    CRATAC_TRY(exclusive_scan_i32_to_i64(c, caps, c->d_seg_off.as<int64_t>(), nb, c->d_seg_off.as<int64_t>() + nb));
    const size_t kk = (size_t)std::max<int64_t>(2 * n, 1);      // total capacity = 2 N
    CRATAC_TRY(c->d_hits.reserve(4 * kk)); CRATAC_TRY(c->d_out_row.reserve(4 * kk)); CRATAC_TRY(c->d_out_cnt.reserve(4 * kk));
    OverlapArgs a;
    a.n = n; a.chrom = c->d_chrom.as<int32_t>(); a.start = c->d_start.as<int32_t>(); a.end = c->d_end.as<int32_t>(); a.bc = c->d_bc.as<int32_t>();
    a.slot_to_dense = nullptr;
    a.peak_off = c->d_peak_off.as<int64_t>();
    a.pk_start = c->d_peak_start.as<int32_t>(); a.pk_end = c->d_peak_end.as<int32_t>(); a.pk_row = c->d_peak_row.as<int32_t>();
    a.maxpos = c->h_status[ST_MAX_POS_LEN]; a.maxneg = c->h_status[ST_MAX_NEG_LEN];
    a.seg_off = c->d_seg_off.as<int64_t>(); a.cursor = c->d_cursor.as<unsigned int>(); a.hits = c->d_hits.as<int32_t>();
    a.n_hits = &st[ST_N_HITS];
    const bool by_slot = c->bc_is_slot && nb > 0;
    if (by_slot) {
        const size_t cap = (size_t)c->bc_table_cap;
        CRATAC_TRY(c->d_slot_seg.reserve(8 * cap + 4 * cap + 4 * nbb));
        int64_t* slot_seg_off = c->d_slot_seg.as<int64_t>();
        unsigned int* slot_cursor = reinterpret_cast<unsigned int*>(slot_seg_off + cap);
        int32_t* dense_slot = reinterpret_cast<int32_t*>(slot_cursor + cap);
        CRATAC_CUDA(cudaMemsetAsync(slot_cursor, 0, 4 * cap, s));
        k_slot_segments<<<(unsigned)((cap + 255) / 256), 256, 0, s>>>(c->d_slot_to_dense.as<int32_t>(), (int)cap, a.seg_off, slot_seg_off, dense_slot);
        c->launches++;
        a.seg_off = slot_seg_off; a.cursor = slot_cursor;
    }
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_N_HITS], 0, sizeof(int64_t), s));
    int grid = (int)std::min<int64_t>((n + OV_TILE - 1) / OV_TILE, (int64_t)NUM_SMS * 8);
    if (grid < 1) grid = 1;
    c->stage_begin(SG_OVERLAP_SCATTER);
    k_overlap<<<grid, MX_THREADS, 0, s>>>(a);
    if (by_slot) {
        const size_t cap = (size_t)c->bc_table_cap;
        unsigned int* slot_cursor = reinterpret_cast<unsigned int*>(c->d_slot_seg.as<int64_t>() + cap);
        k_slot_cursors<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(slot_cursor, reinterpret_cast<int32_t*>(slot_cursor + cap), nb, c->d_cursor.as<unsigned int>());
        c->launches++;
        a.seg_off = c->d_seg_off.as<int64_t>(); a.cursor = c->d_cursor.as<unsigned int>();
    }
    c->stage_end(SG_OVERLAP_SCATTER);
    c->launches++;
    CRATAC_CUDA(cudaGetLastError());
    c->host_mark("matrix_launch");
    int32_t* nnz_col = c->d_nnz_col.as<int32_t>();
    if (nb > 0) {
        static bool configured = false;
        if (!configured) {
            CRATAC_CUDA(cudaFuncSetAttribute(k_reduce_large, cudaFuncAttributeMaxDynamicSharedMemorySize, COUNT_RANGE * 4 + COUNT_RANGE / 2));
            configured = true;
        }
        int g1 = (int)std::min<int64_t>(nb, (int64_t)NUM_SMS * 32);
        c->stage_begin(SG_REDUCE_SMALL);
        k_reduce_warp<<<NUM_SMS * 8, SORT_THREADS, 0, s>>>(a.hits, a.seg_off, a.cursor, nb, c->d_out_row.as<int32_t>(), c->d_out_cnt.as<int32_t>(), nnz_col);
        k_reduce_small<<<g1, SORT_THREADS, 0, s>>>(a.hits, a.seg_off, a.cursor, nb, c->d_out_row.as<int32_t>(), c->d_out_cnt.as<int32_t>(), nnz_col);
        c->stage_end(SG_REDUCE_SMALL);
        c->stage_begin(SG_REDUCE_LARGE);
        int g2 = (int)std::min<int64_t>(nb, (int64_t)NUM_SMS * 3);          // 72 KiB of shared memory: three CTAs per SM
        k_reduce_large<<<g2, COUNT_THREADS, COUNT_RANGE * 4 + COUNT_RANGE / 2, s>>>(a.hits, a.seg_off, a.cursor, nb, st,
                                                                 c->row_range > 0 ? c->row_range : (c->peaks_on_device_count ? -1 : c->n_peaks),
                                                                 c->d_out_row.as<int32_t>(), c->d_out_cnt.as<int32_t>(), nnz_col);
        c->stage_end(SG_REDUCE_LARGE);
        c->launches += 3;
        CRATAC_CUDA(cudaGetLastError());
    }
    return CRATAC_OK;
}

// K8: CSC in column order.  The columns are the shard's own barcodes in matrix order, or the global column map of a
// multi-GPU run (cratac_union_columns / cratac_set_column_map).
int peak_matrix_csc(Ctx* c) {
    cudaStream_t s = c->stream;
    int64_t* st = c->d_status.as<int64_t>();
    const int64_t nb = c->n_barcodes;
    int32_t* nnz_col = c->d_nnz_col.as<int32_t>();
    // matrix column order: host work (CPython-2.7 set emulation over the barcode hashes) overlaps the kernels
    // queued so far (pileup, fit, peaks, overlap, reduce)
    if (!c->global_cols) CRATAC_TRY(build_barcodes(c));
    c->host_mark("build_barcodes");
    const int32_t* c2d = c->global_cols ? c->d_global_c2d.as<int32_t>() : c->d_col_to_dense.as<int32_t>();
    const int64_t n_cols = c->global_cols ? c->n_global_cols : nb;
    c->n_cols = n_cols;
    size_t ncc = (size_t)std::max<int64_t>(n_cols, 1);
    CRATAC_TRY(c->d_indptr.reserve(8 * (ncc + 1)));
    CRATAC_TRY(c->d_nnz_ordered.reserve(4 * ncc));
    int32_t* nnz_ordered = c->d_nnz_ordered.as<int32_t>();
    if (n_cols > 0) {
        k_gather_nnz<<<(unsigned)((n_cols + 255) / 256), 256, 0, s>>>(nnz_col, c2d, n_cols, nnz_ordered);
        c->launches++;
    }
    CRATAC_TRY(exclusive_scan_i32_to_i64(c, nnz_ordered, c->d_indptr.as<int64_t>(), n_cols, c->d_indptr.as<int64_t>() + n_cols));
    CRATAC_CUDA(cudaMemcpyAsync(&c->h_status[ST_NNZ], c->d_indptr.as<int64_t>() + n_cols, 8, cudaMemcpyDeviceToHost, s));
    CRATAC_CUDA(cudaMemcpyAsync(&c->h_status[ST_N_HITS], &st[ST_N_HITS], 8, cudaMemcpyDeviceToHost, s));
    CRATAC_CUDA(cudaStreamSynchronize(s));
    c->nnz = c->h_status[ST_NNZ];
    c->n_hits = c->h_status[ST_N_HITS];                     // fragment ends inside peaks
    size_t zz = (size_t)std::max<int64_t>(c->nnz, 1);
    CRATAC_TRY(c->d_indices.reserve(8 * zz)); CRATAC_TRY(c->d_data.reserve(4 * zz));
    if (n_cols > 0) {
        c->stage_begin(SG_CSC_BUILD);
        if (c->nnz > 0)
            k_csc_copy<<<(unsigned)((c->nnz + CSC_CHUNK - 1) / CSC_CHUNK), 256, 0, s>>>(c->d_out_row.as<int32_t>(), c->d_out_cnt.as<int32_t>(), c->d_seg_off.as<int64_t>(), c2d,
                                                                                       c->d_indptr.as<int64_t>(), n_cols, c->nnz, c->d_indices.as<int64_t>(),
                                                                                       c->d_data.as<int32_t>());
        c->stage_end(SG_CSC_BUILD);
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
    }
    return CRATAC_OK;
}

int peak_matrix_device(Ctx* c) {
    CRATAC_TRY(peak_matrix_hits(c));
    return peak_matrix_csc(c);
}

// ---- FILTER_PEAK_MATRIX: m[:, cols] (+ drop the columns left without a positive entry) -----------------------------
// filter_peak_matrix/__init__.py:46-70 (prune) and :93 (filter_barcodes -> cellranger/matrix.py:658-662 m[:, indices]).
// one warp per selected column: entries and whether any is > 0
__global__ void __launch_bounds__(256) k_sel_count(const int64_t* __restrict__ cols, int64_t n_sel, int64_t n_cols_old, const int64_t* __restrict__ indptr,
                                                    const int32_t* __restrict__ data, int drop_empty, int32_t* __restrict__ cnt,
                                                    int32_t* __restrict__ keep, int64_t* status) {
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n_sel; j += n_warps) {
        const int64_t c = cols[j];
        if (c < 0 || c >= n_cols_old) { if (lane == 0) { raise_error(status, CRATAC_ERR_ARG, j); cnt[j] = 0; keep[j] = 0; } continue; }
        const int64_t s = indptr[c], e = indptr[c + 1];
        bool pos = false;
        for (int64_t k = s + lane; k < e && !pos; k += 32) pos = data[k] > 0;
        pos = __any_sync(0xffffffffu, pos);
        const bool kp = !drop_empty || pos;
        if (lane == 0) { cnt[j] = kp ? (int32_t)(e - s) : 0; keep[j] = kp; }
    }
}
__global__ void k_sel_compact(const int64_t* __restrict__ cols, int64_t n_sel, const int64_t* __restrict__ indptr, const int32_t* __restrict__ cnt,
                              const int32_t* __restrict__ keep, const int64_t* __restrict__ kpos, int64_t* __restrict__ kept, int32_t* __restrict__ out_cnt,
                              int64_t* __restrict__ src_start) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_sel || !keep[j]) return;
    const int64_t o = kpos[j];
    kept[o] = j;
    out_cnt[o] = cnt[j];
    src_start[o] = indptr[cols[j]];
}
// chunked over the output entries like k_csc_copy
__global__ void __launch_bounds__(256) k_sel_copy(const int64_t* __restrict__ src_start, const int64_t* __restrict__ indptr_out, int64_t n_out, int64_t nnz,
                                                   const int64_t* __restrict__ indices_in, const int32_t* __restrict__ data_in,
                                                   int64_t* __restrict__ indices, int32_t* __restrict__ data) {
    __shared__ int64_t s_j[2];
    const int64_t p0 = (int64_t)blockIdx.x * CSC_CHUNK;
    const int64_t p1 = p0 + CSC_CHUNK < nnz ? p0 + CSC_CHUNK : nnz;
    if (threadIdx.x < 2) {
        const int64_t target = threadIdx.x == 0 ? p0 : p1 - 1;
        int64_t lo = 0, hi = n_out - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (indptr_out[mid] <= target) lo = mid; else hi = mid - 1;
        }
        s_j[threadIdx.x] = lo;
    }
    __syncthreads();
    const int64_t j0 = s_j[0], j1 = s_j[1];
    const bool by_warp = j1 - j0 >= 8;
    const int step = by_warp ? 8 : 1, width = by_warp ? 32 : 256;
    const int me = by_warp ? (threadIdx.x & 31) : threadIdx.x;
    for (int64_t j = j0 + (by_warp ? (threadIdx.x >> 5) : 0); j <= j1; j += step) {
        const int64_t c0 = indptr_out[j], c1 = indptr_out[j + 1];
        const int64_t a = c0 > p0 ? c0 : p0, b = c1 < p1 ? c1 : p1;
        if (b <= a) continue;
        const int64_t src = src_start[j] + (a - c0);
        const int len = (int)(b - a);
        for (int k = me; k < len; k += width) { indices[a + k] = indices_in[src + k]; data[a + k] = data_in[src + k]; }
    }
}

int select_columns(Ctx* c, int64_t n_sel, const int64_t* h_cols, int drop_empty) {
    cudaStream_t s = c->stream;
    int64_t* st = c->d_status.as<int64_t>();
    const size_t ns = (size_t)std::max<int64_t>(n_sel, 1);
    // scratch: cols i64 | kpos i64 | src_start i64 | cnt i32 | keep i32 | out_cnt i32
    CRATAC_TRY(c->d_sel_scratch.reserve(ns * (8 * 3 + 4 * 3) + 64));
    int64_t* cols = c->d_sel_scratch.as<int64_t>();
    int64_t* kpos = cols + ns;
    int64_t* src_start = kpos + ns;
    int32_t* cnt = reinterpret_cast<int32_t*>(src_start + ns);
    int32_t* keep = cnt + ns;
    int32_t* out_cnt = keep + ns;
    CRATAC_TRY(c->d_sel_kept.reserve(8 * ns));
    CRATAC_TRY(c->d_sel_indptr.reserve(8 * (ns + 1)));
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_ERROR], 0, sizeof(int64_t) * 2, s));
    c->n_sel_cols = 0; c->sel_nnz = 0;
    if (n_sel > 0) {
        CRATAC_CUDA(cudaMemcpyAsync(cols, h_cols, 8 * (size_t)n_sel, cudaMemcpyHostToDevice, s));
        const int g = (int)std::min<int64_t>((n_sel * 32 + 255) / 256, (int64_t)NUM_SMS * 16);
        k_sel_count<<<g, 256, 0, s>>>(cols, n_sel, c->n_cols, c->d_indptr.as<int64_t>(), c->d_data.as<int32_t>(), drop_empty, cnt, keep, st);
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
    }
    CRATAC_TRY(exclusive_scan_i32_to_i64(c, keep, kpos, n_sel, &st[ST_TICKET2]));
    CRATAC_CUDA(cudaMemcpyAsync(&c->h_status[ST_TICKET2], &st[ST_TICKET2], 8, cudaMemcpyDeviceToHost, s));
    CRATAC_TRY(sync_status(c));
    const int64_t n_out = c->h_status[ST_TICKET2];
    if (n_sel > 0) {
        k_sel_compact<<<(unsigned)((n_sel + 255) / 256), 256, 0, s>>>(cols, n_sel, c->d_indptr.as<int64_t>(), cnt, keep, kpos, c->d_sel_kept.as<int64_t>(), out_cnt,
                                                                     src_start);
        c->launches++;
    }
    CRATAC_TRY(exclusive_scan_i32_to_i64(c, out_cnt, c->d_sel_indptr.as<int64_t>(), n_out, c->d_sel_indptr.as<int64_t>() + n_out));
    CRATAC_CUDA(cudaMemcpyAsync(&c->h_status[ST_TICKET2], c->d_sel_indptr.as<int64_t>() + n_out, 8, cudaMemcpyDeviceToHost, s));
    CRATAC_CUDA(cudaStreamSynchronize(s));
    const int64_t nnz = c->h_status[ST_TICKET2];
    const size_t zz = (size_t)std::max<int64_t>(nnz, 1);
    CRATAC_TRY(c->d_sel_indices.reserve(8 * zz)); CRATAC_TRY(c->d_sel_data.reserve(4 * zz));
    if (nnz > 0) {
        k_sel_copy<<<(unsigned)((nnz + CSC_CHUNK - 1) / CSC_CHUNK), 256, 0, s>>>(src_start, c->d_sel_indptr.as<int64_t>(), n_out, nnz, c->d_indices.as<int64_t>(),
                                                                                c->d_data.as<int32_t>(), c->d_sel_indices.as<int64_t>(), c->d_sel_data.as<int32_t>());
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
    }
    CRATAC_CUDA(cudaStreamSynchronize(s));
    c->n_sel_cols = n_out; c->sel_nnz = nnz;
    return CRATAC_OK;
}

}  // namespace cratac
