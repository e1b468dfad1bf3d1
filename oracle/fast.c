/*
 * Plain-C restatement of the reference stage loops -- TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  Per-base loops are kept literal
 * (one iteration per genome base, as the Python reference does) so that this
 * file is an independent check on the closed forms used by the CUDA kernels.
 *
 * Reference lines (cellranger-atac 1.2.0):
 *   orc_window_counts   mro/atac/stages/processing/count_cut_sites/__init__.py:94-101
 *   orc_count_hist      count_cut_sites/__init__.py:104
 *   orc_bedgraph_rows   count_cut_sites/__init__.py:27-48  (zero-gap quirk at :36-37)
 *   orc_identify_peaks  lib/python/utils.py:105-115 + detect_peaks/__init__.py:40-62
 *   orc_peak_matrix     generate_peak_matrix/__init__.py:107-119 +
 *                       tenkit/lib/python/tenkit/regions.py:136-145
 *   orc_parse           lib/python/tools/io.py:75-79
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC fast.c -o liboracle.so)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HALF_WINDOW 200
#define MERGE_DIST 500

/* Cuts[max(0,s-200) : min(s+201,L)] += 1 for s in (start, stop), via a
 * difference array (identical result to the slice adds). */
int orc_window_counts(int64_t L, const int32_t *starts, const int32_t *stops, int64_t n, int32_t *W)
{
    int64_t *diff = (int64_t *)calloc((size_t)L + 1, sizeof(int64_t));
    if (!diff) return -1;
    for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < 2; k++) {
            int64_t s = k ? stops[i] : starts[i];
            int64_t lo = s - HALF_WINDOW; if (lo < 0) lo = 0;
            int64_t hi = s + HALF_WINDOW + 1; if (hi > L) hi = L;
            if (lo < hi) { diff[lo] += 1; diff[hi] -= 1; }
        }
    }
    int64_t run = 0;
    for (int64_t p = 0; p < L; p++) { run += diff[p]; W[p] = (int32_t)run; }
    free(diff);
    return 0;
}

/* Counter(v for v in Cuts if v > 0): hist[v] += 1.  Returns max v seen, or -1 if v >= cap. */
int64_t orc_count_hist(const int32_t *W, int64_t L, int64_t *hist, int64_t cap)
{
    int64_t mx = 0;
    for (int64_t p = 0; p < L; p++) {
        int32_t v = W[p];
        if (v > 0) {
            if (v >= cap) return -1;
            hist[v] += 1;
            if (v > mx) mx = v;
        }
    }
    return mx;
}

/* write_chrom_bedgraph: literal loop.  Returns number of rows (may exceed cap; only cap are stored). */
int64_t orc_bedgraph_rows(const int32_t *W, int64_t L, int64_t *rs, int64_t *re, int64_t *rc, int64_t cap)
{
    int64_t n = 0;
    int have = 0;
    int64_t last_count = 0, last_start = 0, last_pos = 0;
    for (int64_t pos = 0; pos < L; pos++) {
        int64_t count = W[pos];
        if (count == 0) continue;
        if (have && last_count == count) {
            last_pos = pos;
        } else {
            if (have && last_count >= 1) {
                if (n < cap) { rs[n] = last_start; re[n] = last_pos + 1; rc[n] = last_count; }
                n++;
            }
            last_start = pos; last_pos = pos; last_count = count; have = 1;
        }
    }
    if (have && last_count >= 1) {
        if (n < cap) { rs[n] = last_start; re[n] = last_pos + 1; rc[n] = last_count; }
        n++;
    }
    return n;
}

/* identify_peaks over the per-base stream the bedgraph rows expand to (one contig). */
int64_t orc_identify_peaks(const int64_t *rs, const int64_t *re, const int64_t *rc, int64_t nrows,
                           int64_t threshold, int64_t *ps, int64_t *pe, int64_t cap)
{
    int64_t n = 0;
    int open = 0;
    int64_t w0 = 0, w1 = 0;
    for (int64_t r = 0; r < nrows; r++) {
        for (int64_t pos = rs[r]; pos < re[r]; pos++) {
            if (rc[r] < threshold) continue;
            if (!open) { w0 = pos; w1 = pos + 1; open = 1; }
            else if (pos - MERGE_DIST <= w1) { w1 = pos + 1; }
            else {
                if (n < cap) { ps[n] = w0; pe[n] = w1; }
                n++;
                w0 = pos; w1 = pos + 1;
            }
        }
    }
    if (open) { if (n < cap) { ps[n] = w0; pe[n] = w1; } n++; }
    return n;
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

/* bisect.bisect(starts, pt) on [lo, hi) */
static int64_t bisect_right(const int32_t *a, int64_t lo, int64_t hi, int32_t x)
{
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (x < a[mid]) hi = mid; else lo = mid + 1; }
    return lo;
}

/* Peaks are given per contig: peak_off[c]..peak_off[c+1] index sorted, non-overlapping
 * (start,end) rows whose matrix row id is peak_row[i].  Fragments carry contig id, start, stop, and
 * column id (bc < 0 = barcode not in this chunk).  Output: canonical CSC.  indptr has n_bc+1
 * entries; indices/data have capacity cap.  Returns nnz (or -1 on allocation failure). */
int64_t orc_peak_matrix(const int64_t *peak_off, const int32_t *peak_start, const int32_t *peak_end,
                        const int32_t *peak_row, int32_t n_contigs,
                        const int32_t *f_chrom, const int32_t *f_start, const int32_t *f_stop,
                        const int32_t *f_bc, int64_t n, int64_t n_bc,
                        int64_t *indptr, int64_t *indices, int64_t *data, int64_t cap)
{
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(2 * n + 1));
    if (!keys) return -1;
    int64_t k = 0;
    for (int64_t i = 0; i < n; i++) {
        int32_t c = f_chrom[i];
        if (f_bc[i] < 0 || c < 0 || c >= n_contigs) continue;
        int64_t lo = peak_off[c], hi = peak_off[c + 1];
        for (int e = 0; e < 2; e++) {
            int32_t pos = e ? f_stop[i] : f_start[i];
            int64_t idx = bisect_right(peak_start, lo, hi, pos) - 1;
            if (idx < lo) continue;
            if (pos >= peak_start[idx] && pos < peak_end[idx])
                keys[k++] = ((uint64_t)(uint32_t)f_bc[i] << 32) | (uint32_t)peak_row[idx];
        }
    }
    qsort(keys, (size_t)k, sizeof(uint64_t), cmp_u64);
    for (int64_t b = 0; b <= n_bc; b++) indptr[b] = 0;
    int64_t nnz = 0;
    for (int64_t i = 0; i < k;) {
        int64_t j = i;
        while (j < k && keys[j] == keys[i]) j++;
        if (nnz < cap) { indices[nnz] = (int64_t)(keys[i] & 0xffffffffu); data[nnz] = j - i; }
        indptr[(keys[i] >> 32) + 1] += 1;
        nnz++;
        i = j;
    }
    for (int64_t b = 0; b < n_bc; b++) indptr[b + 1] += indptr[b];
    free(keys);
    return nnz;
}

/* line.strip().split("\t") -> contig, int(start), int(stop), barcode, int(dup).
 * Contig and barcode are returned as (offset, length) into the text.  Returns the number of
 * lines parsed, or -(line_no+1) on a malformed line. */
int64_t orc_parse(const char *text, int64_t nbytes, int64_t cap,
                  int64_t *contig_off, int32_t *contig_len, int32_t *start, int32_t *stop,
                  int64_t *bc_off, int32_t *bc_len, int32_t *dup)
{
    int64_t n = 0, p = 0;
    while (p < nbytes) {
        int64_t eol = p;
        while (eol < nbytes && text[eol] != '\n') eol++;
        int64_t a = p, b = eol;                      /* [a,b) = line without '\n' */
        while (a < b && (text[a] == ' ' || text[a] == '\t' || text[a] == '\r')) a++;
        while (b > a && (text[b - 1] == ' ' || text[b - 1] == '\t' || text[b - 1] == '\r')) b--;
        if (a == b) { p = eol + 1; continue; }
        if (n >= cap) return n + 1;
        int64_t f[6]; int nf = 0; f[nf++] = a;
        for (int64_t q = a; q < b && nf < 6; q++) if (text[q] == '\t') f[nf++] = q + 1;
        if (nf != 5) return -(n + 1);
        f[5] = b + 1;
        int64_t vals[3]; int which[3] = {1, 2, 4};
        for (int t = 0; t < 3; t++) {
            int64_t v = 0; int64_t q = f[which[t]], e = f[which[t] + 1] - 1;
            if (q >= e) return -(n + 1);
            for (; q < e; q++) { if (text[q] < '0' || text[q] > '9') return -(n + 1); v = v * 10 + (text[q] - '0'); }
            vals[t] = v;
        }
        contig_off[n] = f[0]; contig_len[n] = (int32_t)(f[1] - 1 - f[0]);
        start[n] = (int32_t)vals[0]; stop[n] = (int32_t)vals[1];
        bc_off[n] = f[3]; bc_len[n] = (int32_t)(f[4] - 1 - f[3]);
        dup[n] = (int32_t)vals[2];
        n++;
        p = eol + 1;
    }
    return n;
}
