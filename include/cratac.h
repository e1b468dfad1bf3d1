/*
 * libcratac -- C ABI of the B200-native fragment -> peaks -> peak x barcode matrix path.
 *
 * The reference (cellranger-atac 1.2.0) has no FFI for this path: the three Martian stages
 * COUNT_CUT_SITES, DETECT_PEAKS and GENERATE_PEAK_MATRIX are pure Python.  The entry points
 * below are what a ctypes binding inside those stage modules calls instead of the Python
 * loops; each one names the reference lines it replaces (paths relative to the reference
 * checkout).  Style follows the reference's only ctypes precedent,
 * tenkit/lib/python/striped_smith_waterman/ssw_wrap.py:54-72: explicit argtypes/restype, an
 * opaque handle, an explicit destroy.
 *
 * Conventions
 *   - every function returns CRATAC_OK (0) or a negative CRATAC_ERR_* code;
 *     cratac_last_error() returns a human-readable message for the calling thread.
 *   - host arrays are caller-allocated, plain pointers + element counts; no torch types.
 *   - pointers named d_* are CUDA device pointers on the context's device.
 *   - coordinates are 0-based, fragments are (start, stop) exactly as in fragments.tsv.gz.
 *   - there is NO CPU fallback: without a CUDA device cratac_create fails.
 */
#ifndef CRATAC_H
#define CRATAC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRATAC_OK 0
#define CRATAC_ERR_CUDA (-1)          /* CUDA runtime failure (message has file:line) */
#define CRATAC_ERR_ARG (-2)           /* bad argument / call order */
#define CRATAC_ERR_PARSE (-3)         /* malformed fragments line (message has the line number) */
#define CRATAC_ERR_UNSORTED (-4)      /* fragments not sorted by (contig order, start) */
#define CRATAC_ERR_CONTIG (-5)        /* contig name not in the reference */
#define CRATAC_ERR_CAPACITY (-6)      /* an internal table overflowed after retries */
#define CRATAC_ERR_EM_ZERODIV (-7)    /* the reference fit raises ZeroDivisionError on this histogram */
#define CRATAC_ERR_EM_INDEX (-8)      /* the reference fit raises IndexError (fewer than 2 counts > 15) */
#define CRATAC_ERR_PEAKS_OVERLAP (-9) /* user peaks overlap (reference: KeyError at generate_peak_matrix:114) */
#define CRATAC_ERR_INTERNAL (-10)
#define CRATAC_ERR_IO (-11)

/* matrix column (barcode) order */
#define CRATAC_BC_ORDER_PY2SET 0      /* CPython-2.7 list(set(...)) order: generate_peak_matrix/__init__.py:44 */
#define CRATAC_BC_ORDER_FIRST_SEEN 1  /* order of first appearance in the file */
#define CRATAC_BC_ORDER_SORTED 2      /* lexicographic */

typedef struct cratac_ctx cratac_ctx;

const char* cratac_last_error(void);
int cratac_abi_version(void);

/* One context per GPU.  contig_names/contig_lens: every contig of fasta/genome.fa.fai in file
 * order (lib/python/tools/ref_manager.py:45-52); is_primary[i] != 0 for the contigs that
 * ReferenceManager.primary_contigs(allow_sex_chromosomes=True) returns (ref_manager.py:140-146,
 * the chunking of count_cut_sites/__init__.py:55-62 and detect_peaks/__init__.py:67-81):
 * peaks are called on primary contigs only; fragments on other contigs still count in the matrix. */
int cratac_create(int device, int n_contigs, const char* const* contig_names, const int64_t* contig_lens,
                  const int32_t* is_primary, cratac_ctx** out);
void cratac_destroy(cratac_ctx* ctx);
/* options: "barcode_order" (CRATAC_BC_ORDER_*), "barcode_table_slots", "run_capacity", "em_fast" (0 plain fixed
 * point, 1 accelerated, 2 on-chip kernel first), "profile" (per-kernel CUDA events), "defer_sync" (see
 * cratac_fit_threshold_result), "stream_decode" (0 off / 1 host texts of 64 MiB and more / 2 always: the host text is
 * copied in chunks and each chunk is decoded while the next one is on the PCIe bus), "stream_chunk_tiles" (16 KiB
 * text tiles per chunk, 0 = 128 MiB worth), "py2_device_min" (from this many distinct barcodes on the PY2SET column
 * order is worked out on the device instead of by the host simulation; default 300000), "decode_kernel" (0, default:
 * newline-count pass + general parse kernel; 1: the text is read ONCE by k_parse_stream (decoupled look-back for the line
 * numbers), tiles it declines go through the general kernel; 2: same with 32 KiB tiles -- both measured slower than 0 on
 * B200), "decode_ctas_per_sm" (cap on resident CTAs of k_parse_stream) */
int cratac_set_option(cratac_ctx* ctx, const char* key, int64_t value);
/* kernels launched by this context since creation / last reset */
int64_t cratac_launch_count(cratac_ctx* ctx, int reset);
/* text ranges (of at most 16 KiB) the single-pass decoder handed to the general parse kernel since creation / last reset:
 * 0 on a clean fragments.tsv (tools/io.py:75-92 shaped lines of at most 64 bytes) */
int64_t cratac_decode_declined(cratac_ctx* ctx, int reset);
/* option "decode_profile" = 1: k_parse_stream adds up SM cycles per phase of its tile loop (one thread's clock per CTA);
 * this reads the first n (<= 10) counters of the last decode: wait for the tile, separator masks, separator list, structure
 * check, look-back, field parse, commit, tiles, look-back rounds, look-back poll retries.  Diagnostics only (tools/decode_bench.py). */
int cratac_decode_profile(cratac_ctx* ctx, int64_t* cycles_out, int n);
int cratac_synchronize(cratac_ctx* ctx);
/* the CUDA stream every kernel of this context is launched on (cudaStream_t as an integer) */
int cratac_stream(cratac_ctx* ctx, uint64_t* stream_out);
/* per-kernel device times of the last run, measured with CUDA events on that stream when the
 * option "profile" is 1: ms_out has cratac_num_stages() entries, -1 for stages that did not run. */
int cratac_num_stages(void);
const char* cratac_stage_name(int id);
int cratac_stage_times(cratac_ctx* ctx, float* ms_out);
/* host wall-clock phases of the last cratac_run as "name=ms;..." (profiling aid) */
int cratac_host_times(cratac_ctx* ctx, char* out, int64_t cap);
/* SM cycles per phase of the last threshold fit, 9 doubles (profiling aid) */
int cratac_em_debug(cratac_ctx* ctx, double* out);

/* ---- (1) fragment decode: replaces lib/python/tools/io.py:75-79 (open_fragment_file) and
 *      :82-92 (parsed_fragments_from_contig).  `text` is the DECOMPRESSED fragments.tsv
 *      (contig \t start \t stop \t barcode \t dup_count \n), position sorted.  Produces SoA int32
 *      (chrom, start, end, barcode index, dup count) in HBM plus the barcode dictionary.
 *      Limits the reference's line.split() does not have: a line of more than 256 bytes is reported as malformed
 *      (CRATAC_ERR_PARSE with its line number; a text tile sees 256 bytes of what precedes it), a barcode field of more
 *      than 63 bytes likewise; positions are int32.  Lines of cellranger-atac's own fragments.tsv.gz are 40-60 bytes. */
int cratac_decode_host(cratac_ctx* ctx, const char* text, int64_t nbytes);
int cratac_decode_device(cratac_ctx* ctx, const char* d_text, int64_t nbytes);
/* already-parsed fragments (host SoA, n each); bc = index into `barcodes` (n_barcodes C strings,
 * which become the matrix columns in that order). */
int cratac_set_fragments(cratac_ctx* ctx, int64_t n, const int32_t* chrom, const int32_t* start,
                         const int32_t* end, const int32_t* bc, const int32_t* dup,
                         int64_t n_barcodes, const char* const* barcodes);
int cratac_num_fragments(cratac_ctx* ctx, int64_t* n_fragments, int64_t* n_barcodes);
/* D2H of the SoA; bc[i] is the matrix column of fragment i.  Any pointer may be NULL. */
int cratac_get_fragments(cratac_ctx* ctx, int32_t* chrom, int32_t* start, int32_t* end, int32_t* bc, int32_t* dup);
/* barcode strings in matrix column order, NUL-terminated, packed at `stride` bytes each
 * (replaces the set comprehension of generate_peak_matrix/__init__.py:44). */
int cratac_get_barcodes(cratac_ctx* ctx, char* out, int64_t stride);
int cratac_max_barcode_len(cratac_ctx* ctx, int64_t* len);

/* ---- (2) COUNT_CUT_SITES: replaces count_cut_sites/__init__.py:94-104 (pileup of +-200 bp
 *      windows around both fragment ends, then Counter(v for v in Cuts if v > 0)) for all
 *      primary contigs, and the Counter merge of :77-81.  The histogram stays on the device for
 *      cratac_fit_threshold_device; this call also copies it out (values ascending). */
int cratac_count_cut_sites(cratac_ctx* ctx, int64_t* values, int64_t* weights, int64_t cap, int64_t* n_bins);
/* windowed counts of one contig, D2H (Cuts of count_cut_sites/__init__.py:96-101); W has contig_len entries */
int cratac_window_counts(cratac_ctx* ctx, int contig, int32_t* W);
/* bedgraph rows of one contig (write_chrom_bedgraph, count_cut_sites/__init__.py:27-48, incl. the
 * zero-gap behaviour of :36-37).  Call with cap = 0 to get the row count. */
int cratac_bedgraph_rows(cratac_ctx* ctx, int contig, int64_t* start, int64_t* end, int64_t* count,
                         int64_t cap, int64_t* n_rows);

/* ---- (3) threshold fit: replaces lib/python/analysis/peaks.py:173-193 (estimate_final_threshold)
 *      with everything it calls (peaks.py:28-170, lib/python/analysis/em_fitting.py:10-87).
 *      has_params = 0 and threshold = 15 when the gate of peaks.py:176-177 fires. */
int cratac_fit_threshold(cratac_ctx* ctx, const int64_t* values, const int64_t* weights, int64_t n_bins,
                         double odds_ratio, double params_out[6], int* has_params, int64_t* threshold_out,
                         int64_t* em_iterations, int64_t* inner_iterations);
/* same, on the histogram left on the device by cratac_count_cut_sites; result stays on the
 * device for cratac_call_peaks(threshold = -1) and is also returned. */
int cratac_fit_threshold_device(cratac_ctx* ctx, double odds_ratio, double params_out[6], int* has_params,
                                int64_t* threshold_out);
/* With option "defer_sync" = 1, cratac_window_histogram, cratac_compact_histogram and
 * cratac_fit_threshold_device only queue their kernels on the context's stream (outputs untouched), so
 * that a multi-GPU host can exchange the barcode dictionary while the fit runs; this call waits for the
 * fit and returns its numbers (the Counter merge + estimate_final_threshold of detect_peaks/__init__.py:36-58). */
int cratac_fit_threshold_result(cratac_ctx* ctx, double params_out[6], int* has_params, int64_t* threshold_out);

/* ---- (4) DETECT_PEAKS.main/join: replaces lib/python/utils.py:105-115 + detect_peaks/__init__.py:40-62
 *      (threshold, 500 bp merge) over all primary contigs and the ordering of
 *      lib/python/tools/regions.py:392-417 (.fai contig order, then start).  threshold < 0 uses
 *      the value fitted on the device.  The peaks become the current peak set of the context. */
int cratac_call_peaks(cratac_ctx* ctx, int64_t threshold, int64_t* n_peaks);
/* the same merge for the DETECT_PEAKS stage proper, whose input is the bedgraph (cut_sites) and not the fragments:
 * `n` signal intervals = bedgraph rows with count >= threshold (lib/python/utils.py:105-115), sorted by (contig, start) */
int cratac_peaks_from_rows(cratac_ctx* ctx, int64_t n, const int32_t* chrom, const int64_t* start, const int64_t* end, int64_t* n_peaks);
int cratac_get_peaks(cratac_ctx* ctx, int32_t* chrom, int32_t* start, int32_t* end);
/* user-supplied peaks in peaks.bed row order (reanalyze / GENERATE_PEAK_MATRIX stage input);
 * row i of the matrix is peak i.  Touching peaks are fine, overlapping ones are an error. */
int cratac_set_peaks(cratac_ctx* ctx, int64_t n, const int32_t* chrom, const int32_t* start, const int32_t* end);

/* ---- (5) GENERATE_PEAK_MATRIX: replaces generate_peak_matrix/__init__.py:107-119 (both fragment ends
 *      tested with tenkit/regions.py:136-145, Counter[(barcode, peak)], COO -> canonical CSC) for all
 *      barcodes at once and the hstack of :62-72.  Output layout = cellranger/h5_constants.py:25:
 *      indptr int64[B+1], indices int64[nnz] (row = peak), data int32[nnz]. */
int cratac_peak_matrix(cratac_ctx* ctx, int64_t* nnz);
int cratac_get_matrix(cratac_ctx* ctx, int64_t* indptr, int64_t* indices, int32_t* data);
/* device views of the current results (valid until the next call that recomputes them) */
int cratac_matrix_device(cratac_ctx* ctx, const int64_t** d_indptr, const int64_t** d_indices, const int32_t** d_data);

/* ---- (6) FILTER_PEAK_MATRIX (the stage right after the path): replaces
 *      filter_peak_matrix/__init__.py:84-96 -> cellranger/matrix.py:874-883 (filter_barcodes: the columns of
 *      the sorted cell barcodes, m[:, indices] of matrix.py:658-662) and prune (:46-70: optional random
 *      subset of barcodes, then the columns with at least one positive entry).  The caller passes the
 *      column indices in output order (the random subset is drawn on the host with numpy's legacy
 *      generator, as the reference does); with drop_empty the columns without a positive entry are
 *      removed.  cratac_set_matrix loads a raw matrix from host CSC arrays (raw_peak_bc_matrix.h5);
 *      after cratac_peak_matrix / cratac_run the matrix in HBM is used as it is.  kept[j] = position in
 *      `cols` of output column j. */
int cratac_set_matrix(cratac_ctx* ctx, int64_t n_cols, const int64_t* indptr, const int64_t* indices, const int32_t* data);
int cratac_select_columns(cratac_ctx* ctx, int64_t n_sel, const int64_t* cols, int drop_empty, int64_t* n_cols_out,
                          int64_t* nnz_out);
int cratac_get_selected(cratac_ctx* ctx, int64_t* indptr, int64_t* indices, int32_t* data, int64_t* kept);
int cratac_hist_device(cratac_ctx* ctx, const uint64_t** d_hist, int64_t* cap);

/* ---- whole path in one call, no host round trip between kernels:
 *      decode -> pileup/window/histogram -> fit -> peaks -> matrix.  `text_is_device` selects
 *      whether `text` is a host or a device pointer.  Results are fetched with the getters. */
typedef struct cratac_run_result {
    int64_t n_fragments, n_barcodes, n_bins, threshold, has_params, n_peaks, nnz, em_iterations, em_inner_iterations;
    double params[6];
    int64_t n_hits;   /* fragment ends that fell inside a peak (sum of the matrix) */
} cratac_run_result;
int cratac_run(cratac_ctx* ctx, const char* text, int64_t nbytes, int text_is_device, double odds_ratio,
               cratac_run_result* result);
/* same, starting from fragments already decoded in this context (SoA resident in HBM) */
int cratac_run_from_fragments(cratac_ctx* ctx, double odds_ratio, cratac_run_result* result);
/* The cell caller's mixture fit: two negative binomials (background / cell-containing barcodes) fitted to
 * the histogram {fragments per barcode: barcodes} (ascending values), detect_cell_barcodes/__init__.py:84-201, with the same
 * dispersion solver as the peak fit.  out = mu_noise, alpha_noise, mu_signal, alpha_signal, frac_noise (:50-51), the initial
 * threshold (:150-164), the threshold of estimate_threshold (:125-147; -1 when it is None), EM iterations. */
int cratac_fit_cell_mixture(cratac_ctx* ctx, const int64_t* values, const int64_t* weights, int64_t n_bins, double odds_ratio, double out[8]);
/* ---- the fragment x region siblings of the peak matrix (csrc/siblings.cu).  They read the fragments the
 *      context holds (cratac_decode_*), so one decode serves all of them; per-barcode results come in matrix column order
 *      (cratac_get_barcodes).  Regions are given as (contig index, start, end) arrays grouped by contig in reference order
 *      and, inside a contig, in the order lib/python/tools/regions.py keeps them (sorted (start, end) pairs). */
/* REMOVE_LOW_TARGETING_BARCODES.main, cell_calling/remove_low_targeting_barcodes/__init__.py:176-245, summed over the contigs:
 * out[2 + 2 * n_paddings][n_barcodes] = fragments; fragments whose interval overlaps a region (tools/regions.py:83-93);
 * sum of padded, clipped fragment lengths per padding; bases covered by the union of the padded fragments per padding
 * (per contig, in file order: __init__.py:212-241).  At most 4 paddings. */
int cratac_low_targeting(cratac_ctx* ctx, int64_t n_regions, const int32_t* chrom, const int32_t* start, const int32_t* end, int n_paddings,
                         const int64_t* paddings, int64_t* out);
/* REPORT_INSERT_SIZES.main, stages/metrics/report_insert_sizes/__init__.py:76-113: table_out[n_sel][1001] for the barcodes of
 * the matrix columns cols[]: column s - 1 = fragments of size s (1..1000), column 1000 = longer ones; size 0 is dropped;
 * exclude_non_primary = 1 skips fragments on contigs that are not primary */
int cratac_insert_sizes(cratac_ctx* ctx, int64_t n_sel, const int64_t* cols, int exclude_non_primary, int64_t* table_out);
/* get_counts_by_barcode, lib/python/tools/io.py:429-555.  Seven tracks in this order: TSS (already padded by 2000), CTCF
 * (padded by 250), DNase, enhancer, promoter, blacklist, peaks; track t owns regions [track_off[t], track_off[t + 1]) of the
 * r_* arrays (r_minus: 1 for strand '-').  only_contig >= 0 restricts the scan to one contig.  out[11 + 2 * n_species]
 * [n_barcodes]: passed_filters, total, duplicate, peak_region_cutsites, TSS_fragments, DNase_sensitive_region_fragments,
 * enhancer_region_fragments, promoter_region_fragments, on_target_fragments, blacklist_region_fragments,
 * peak_region_fragments, then passed_filters_<species> and peak_region_fragments_<species> per species
 * (species_of_contig[n_contigs], null with one species).  relpos[2][2 * rel_half + 1]: cut sites by position relative to the
 * TSS / CTCF region they fall in (index = position + rel_half; tools/regions.py:24-36). */
int cratac_counts_by_barcode(cratac_ctx* ctx, const int64_t* track_off, const int32_t* r_chrom, const int32_t* r_start, const int32_t* r_end,
                             const int8_t* r_minus, int only_contig, int n_species, const int32_t* species_of_contig, int rel_half, int64_t* out,
                             int64_t* relpos);
/* cratac_run* start peak calling and the matrix from the threshold the mixture fit shows after option "spec_iter" (32) EM
 * iterations, on a second stream, while the fit -- a single CTA, peaks.py:144-193 is sequential -- runs to convergence;
 * the speculative result is kept only when the fit ends on the same threshold, otherwise both are done again (option
 * "speculate" = 0 switches this off).  An integer threshold that is still on its way at iteration 32 (the real-valued
 * crossing of peaks.py:43-63 drifts past an integer later) is caught by the fit itself: every eighth iteration it
 * extrapolates the crossing's limit.  While that limit is more than an integer away the first guess is held back (at most
 * 64 iterations); a limit that can be trusted -- no integer between here and there -- is the guess; and after a first
 * guess the first trusted limit either confirms it or is published as the second guess, for which peaks + matrix are
 * queued once more, still beside the fit.
 * Counters since creation: runs whose speculative result was kept / redone; runs with a second guess. */
int cratac_speculation_stats(cratac_ctx* ctx, int64_t* kept, int64_t* redone);
/* Genome bases per pileup tile (one CTA pass of K2/K3/K5; count_cut_sites/__init__.py:94-104 has no tiles): for tests that put
 * run ends and zero gaps on tile borders. */
int64_t cratac_pileup_tile_bases(void);
int64_t cratac_speculation_second_guesses(cratac_ctx* ctx);
/* The same as building blocks for the multi-GPU step (every rank fits the same all-reduced histogram, so all ranks see the
 * same tentative threshold): launch the fit so that it publishes one; wait for it (published = 0 when the fit ended before
 * its spec_iter-th iteration -- then there is nothing to speculate on, on any rank); route the library's calls to the
 * speculation stream with the tentative threshold (on = 1; on = 2: with the fit's second guess) and back (on = 0).
 * cratac_fit_revised waits for the second guess or the end of the fit (published = 0: the fit did not correct itself). */
int cratac_fit_speculative_launch(cratac_ctx* ctx, double odds_ratio);
int cratac_fit_tentative(cratac_ctx* ctx, int* published, int64_t* threshold);
int cratac_fit_revised(cratac_ctx* ctx, int* published, int64_t* threshold);
int cratac_speculation_scope(cratac_ctx* ctx, int on, uint64_t* stream_out);
/* fragments.tsv.gz (BGZF; plain gzip and plain text are read too) -> fragments in HBM.  Replaces, for the whole file, the
 * `gzip -c -d` pipe of lib/python/tools/io.py:197-217 and the per-line split of :75-92: `threads` host threads (0: all
 * cores) inflate the BGZF blocks with the library's own DEFLATE decoder (CRC-32 checked) into a ring of pinned chunks
 * (option "file_chunk_bytes", default 32 MiB each, 4 slots); each chunk is copied to the device and parsed as soon as its
 * blocks are there, so inflate, PCIe copy and parse overlap.  The inflated text never exists in host memory as a whole. */
int cratac_decode_file(cratac_ctx* ctx, const char* fragments_path, int threads);
/* text -> BGZF file (blocks of <= 0xff00 input bytes + the EOF block; the container of fragments.tsv.gz,
 * mark_duplicates/__init__.py:399), deflated at `level` on `threads` host threads (0: all cores) */
int cratac_write_bgzf(const char* path, const char* data, int64_t nbytes, int level, int threads);
/* HDF5 chunk filters (shuffle + deflate) of a 1-D array on `threads` host threads (0: all cores): chunk k of chunk_elems
 * elements (the last one zero-padded) -> out + k * stride (stride >= zlib's compressBound of a chunk), sizes[k] its filtered
 * length.  The storage of the matrix arrays in cellranger/matrix.py:441-447 (chunks of 80000, gzip, shuffle). */
int cratac_h5_filter_chunks(const void* data, int64_t n_elems, int item_size, int64_t chunk_elems, int level, int threads, unsigned char* out,
                            int64_t stride, int64_t* sizes);
/* cratac_decode_file + the rest of the path (as cratac_run) */
int cratac_run_file(cratac_ctx* ctx, const char* fragments_path, int threads, double odds_ratio, cratac_run_result* result);

/* ---- multi-GPU building blocks (one process per GPU, genome sharded by contiguous contig ranges; the
 *      exchanges are NCCL calls made by the host on these device buffers).  They replace the file-based
 *      merges of the reference: `Counter +=` (count_cut_sites/__init__.py:77-81) and `hstack`
 *      (generate_peak_matrix/__init__.py:62-72). */
int cratac_status_device(cratac_ctx* ctx, int64_t** d_status, int64_t* n_slots);   /* slot 6 = max window count, 10 = n_peaks */
int cratac_window_histogram(cratac_ctx* ctx);
int cratac_compact_histogram(cratac_ctx* ctx);
int cratac_barcode_table_device(cratac_ctx* ctx, const uint64_t** d_keys, const uint64_t** d_first, const int64_t** d_py2hash, int64_t* n);
int cratac_set_column_map(cratac_ctx* ctx, int64_t n_cols, const int32_t* d_col_to_dense);   /* n_cols < 0 resets */
int cratac_set_row_offset(cratac_ctx* ctx, int64_t row_base, int64_t total_rows);
/* The barcode universe of all shards -> global matrix columns, on the device (csrc/multi.cu): replaces the `set` of
 * generate_peak_matrix/__init__.py:44, which the reference fills by reading the whole fragments file.  d_gathered is the
 * all-gather of every rank's cratac_barcode_table_device triple, laid out [world][3][max_barcodes] (keys, shard-local first
 * lines, CPython-2.7 hashes; rows padded to max_barcodes); n_fragments / n_barcodes are host arrays [world].  Every rank
 * computes the same columns (hash-table union with atomicMin on the first global line, insertion rank by bitmap +
 * popcount, CPython-2.7 set order or first-seen order per option "barcode_order"); the context's column map is set as by
 * cratac_set_column_map.  `stream` (0: the context's own) is where the work is queued: a side stream lets it run under the
 * mixture fit; the context's stream is ordered behind it. */
int cratac_union_columns(cratac_ctx* ctx, int world, int64_t max_barcodes, const int64_t* d_gathered, const int64_t* n_fragments,
                         const int64_t* n_barcodes, uint64_t stream, int64_t* n_cols);
/* 64-bit barcode keys of the global columns, in column order (device pointer, valid until the next union) */
int cratac_column_keys_device(cratac_ctx* ctx, const int64_t** d_keys, int64_t* n_cols);
/* cratac_peak_matrix in two halves: `begin` queues overlap + per-barcode reduction (generate_peak_matrix/__init__.py:107-119:
 * needs the peaks, not the columns), `finish` builds the CSC in column order (:62-72) -- so a column order that is still
 * being worked out (cratac_union_columns on a side stream, the host's set simulation) costs no GPU idle time */
int cratac_peak_matrix_begin(cratac_ctx* ctx);
int cratac_peak_matrix_finish(cratac_ctx* ctx, int64_t* nnz);
int cratac_merge_rank_csc(cratac_ctx* ctx, int world, int64_t n_cols, const int64_t* d_indptr_all, const int64_t* d_rank_off,
                          const int64_t* d_indices_all, const int32_t* d_data_all, int64_t total_nnz);
int cratac_py2_set_order_hashes(int64_t n, const int64_t* hashes, int32_t* order_out);
/* the same on the device: d_hashes int64[n] in insertion order, d_order_out int32[n]; the tables of
 * every growth step are rebuilt in parallel (fixed point of "a slot keeps the earliest key that asks for
 * it"), which gives exactly the layout of the one-by-one insertion.  Used by the multi-GPU step, where
 * the host simulation of every rank would sit on the critical path.  stream = 0: the context's stream,
 * otherwise the cudaStream_t to run on (the fit may be occupying the context's own). */
int cratac_py2_set_order_hashes_device(cratac_ctx* ctx, int64_t n, const int64_t* d_hashes, int32_t* d_order_out,
                                       uint64_t stream);

/* ---- host helpers */
/* CPython-2.7 list(set(keys)) order for `n` NUL-terminated strings given in insertion order:
 * order_out[j] = index of the key at column j (generate_peak_matrix/__init__.py:44). */
int cratac_py2_set_order(int64_t n, const char* const* keys, int64_t* order_out);
/* multi-threaded BGZF/gzip inflate of fragments.tsv.gz into a caller buffer (replaces the
 * `gzip -c -d` pipe of lib/python/tools/io.py:197-217).  cap = 0 returns the inflated size. */
int cratac_inflate_file(const char* path, int threads, char* out, int64_t cap, int64_t* nbytes);
/* the same for one BGZF virtual-offset range [voff_begin, voff_end) of the file (virtual offset = compressed offset of
 * a block << 16 | offset in the inflated block, the unit of a tabix index): what a rank that owns a range of contigs
 * reads instead of the whole file (replaces pysam.TabixFile(...).fetch(contig), lib/python/tools/io.py:86-99).
 * cap = 0 returns the inflated size. */
int cratac_inflate_range(const char* path, int threads, uint64_t voff_begin, uint64_t voff_end, char* out, int64_t cap, int64_t* nbytes);
/* BGZF blocks are decoded by the library's own DEFLATE decoder (csrc/inflate.cpp); a block it declines goes to zlib.
 * Process-wide counters of both, and a switch for measurements (mode 0: zlib only, 1: own decoder first, the default). */
/* Matrix Market coordinate body ("row col value" 1-based, column by column) after the caller's header lines: the per-entry
 * loop of CountMatrix.save_mex (lib/cellranger/lib/python/cellranger/matrix.py:812-814), formatted on `threads` threads.
 * Host arrays: indptr[n_cols + 1], indices / data [indptr[n_cols]]. */
int cratac_write_mtx(const char* path, const char* header, int64_t n_cols, const int64_t* indptr, const int64_t* indices,
                     const int32_t* data, int threads);
/* The cut_sites bedgraph is the FILE between COUNT_CUT_SITES and DETECT_PEAKS (the stage interface keeps it).  Rows
 * "contig\tstart\tend\tcount" of one contig appended to `path` -- the per-row format() of count_cut_sites/__init__.py:38-48 --
 * and the rows with count >= threshold read back -- the `for track in f` loop of lib/python/utils.py:105-115 with the
 * integer test of detect_peaks/__init__.py:49 -- both on `threads` host threads.  cratac_read_bedgraph allocates its three
 * result arrays (chrom = index into contig_names); release them with cratac_free_host. */
int cratac_write_bedgraph(const char* path, int append, const char* contig, int64_t n, const int64_t* start, const int64_t* end,
                          const int64_t* count, int threads);
int cratac_read_bedgraph(const char* path, int n_contigs, const char* const* contig_names, int has_threshold, double threshold,
                         int threads, int32_t** chrom_out, int64_t** start_out, int64_t** end_out, int64_t* n_out);
void cratac_free_host(void* p);
int cratac_inflate_stats(int64_t* fast_blocks, int64_t* zlib_blocks, int reset);
int cratac_inflate_mode(int mode);
/* synthetic text encoder (SoA on the device -> fragments.tsv text on the device), used by bench/tests
 * to build large inputs without a host formatting loop.  bc_seq = 2-bit packed 16-mers. */
int cratac_encode_text_device(cratac_ctx* ctx, int64_t n, const int32_t* d_chrom, const int32_t* d_start,
                              const int32_t* d_end, const uint32_t* d_bc_seq, const int32_t* d_dup,
                              char* d_out, int64_t cap, int64_t* nbytes);

#ifdef __cplusplus
}
#endif
#endif /* CRATAC_H */
