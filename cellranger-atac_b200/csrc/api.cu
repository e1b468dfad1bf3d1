// C ABI of libcratac (include/cratac.h): context management, host<->device plumbing, whole-path driver.
#include <stdarg.h>
#include <string.h>
#include <zlib.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <unordered_map>
#include <string>
#include <memory>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <chrono>

#include "common.cuh"

namespace cratac {

static thread_local char g_err[1024] = "";
thread_local Lane* tl_lane = nullptr;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap && p != nullptr) return CRATAC_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = std::max<size_t>(bytes, 256);
    want = (want + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e)); p = nullptr; return CRATAC_ERR_CUDA; }
    cap = want;
    return CRATAC_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
}

int sync_status(Ctx* c) {
    CRATAC_CUDA(cudaMemcpyAsync(c->h_status, c->d_status.p, sizeof(int64_t) * ST_SLOTS, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    int64_t err = c->h_status[ST_ERROR];
    if (err == 0) return CRATAC_OK;
    int64_t arg = c->h_status[ST_ERROR_ARG];
    switch ((int)err) {
        case CRATAC_ERR_UNSORTED: set_error("fragment %lld is out of order: fragments must be sorted by (reference contig order, start)", (long long)arg); break;
        case CRATAC_ERR_PARSE: set_error("malformed fragments line %lld", (long long)(arg + 1)); break;
        case CRATAC_ERR_CONTIG: set_error("fragments line %lld: contig is not in the reference", (long long)(arg + 1)); break;
        case CRATAC_ERR_CAPACITY: set_error("device table overflow (detail %lld)", (long long)arg); break;
        case CRATAC_ERR_EM_ZERODIV: set_error("threshold fit: a mixture component received no weight (the reference raises ZeroDivisionError here)"); break;
        case CRATAC_ERR_EM_INDEX: set_error("threshold fit: fewer than two distinct window counts above 15 (the reference raises IndexError here)"); break;
        default: set_error("device error %lld (detail %lld)", (long long)err, (long long)arg);
    }
    cudaMemsetAsync(c->d_status.p, 0, sizeof(int64_t) * 2, c->stream);
    return (int)err;
}

int pileup_tile_bases();                                                                 // pileup.cu
int compact_histogram(Ctx* c);                                                           // pileup.cu
int offset_peak_rows(Ctx* c, int64_t base);                                              // pileup.cu
int merge_runs_from_host(Ctx* c, int64_t n, const int32_t* chrom, const int64_t* start, const int64_t* end);   // pileup.cu
void py2_set_order_hashes(const long long* hashes, int64_t n, std::vector<int32_t>& order);   // decode.cu
int py2_set_order_device(Ctx* c, int64_t n, const int64_t* d_hashes, int32_t* d_order_out);   // py2order.cu
bool inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, uint32_t* scratch);   // inflate.cpp
size_t inflate_scratch_words();
int select_columns(Ctx* c, int64_t n_sel, const int64_t* h_cols, int drop_empty);   // matrix.cu
int merge_rank_csc(Ctx* c, int world, int64_t n_cols, const int64_t* d_indptr_all, const int64_t* d_rank_off, const int64_t* d_indices_all,
                   const int32_t* d_data_all, int64_t total_nnz);                         // matrix.cu
int dump_window_counts(Ctx* c, int contig, int32_t* d_W);                                // pileup.cu
int compute_max_len(Ctx* c);                                                             // decode.cu
int fit_threshold_arrays(Ctx* c, const int64_t* d_values, const int64_t* d_weights, int64_t n, double odds_ratio);   // em.cu
void py2_set_order(const std::vector<std::string>& keys, std::vector<int64_t>& order);   // decode.cu
int encode_text(Ctx* c, int64_t n, const int32_t* d_chrom, const int32_t* d_start, const int32_t* d_end, const uint32_t* d_seq,
                const int32_t* d_dup, char* d_out, int64_t cap, int64_t* nbytes);        // decode.cu

static int upload_i64(Ctx* c, DevBuf& b, const std::vector<int64_t>& v) {
    CRATAC_TRY(b.reserve(sizeof(int64_t) * std::max<size_t>(v.size(), 1)));
    CRATAC_CUDA(cudaMemcpyAsync(b.p, v.data(), sizeof(int64_t) * v.size(), cudaMemcpyHostToDevice, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    return CRATAC_OK;
}

static int fetch_peak_count(Ctx* c) {
    if (!c->peaks_on_device_count) return CRATAC_OK;
    CRATAC_TRY(sync_status(c));
    c->n_peaks = c->h_status[ST_N_PEAKS];
    c->peaks_on_device_count = false;
    return CRATAC_OK;
}

}  // namespace cratac

using namespace cratac;

struct cratac_ctx { Ctx c; };

extern "C" {

const char* cratac_last_error(void) { return g_err; }
int cratac_abi_version(void) { return 1; }

int cratac_create(int device, int n_contigs, const char* const* contig_names, const int64_t* contig_lens, const int32_t* is_primary,
                  cratac_ctx** out) {
    if (!out || n_contigs <= 0 || !contig_names || !contig_lens) { set_error("cratac_create: bad arguments"); return CRATAC_ERR_ARG; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available (%s); libcratac has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return CRATAC_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (have %d)", device, ndev); return CRATAC_ERR_ARG; }
    CRATAC_CUDA(cudaSetDevice(device));
    cratac_ctx* h = new cratac_ctx();
    Ctx* c = &h->c;
    c->device = device;
    c->n_contigs = n_contigs;
    c->tile_bases = pileup_tile_bases();
    c->bc_table_cap = 1 << 20;   // 16 MiB of slots; grows x4 when more than half full
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); delete h; return CRATAC_ERR_CUDA; }
    int64_t base = 0, tiles = 0, pidx = 0;
    for (int i = 0; i < n_contigs; i++) {
        c->contig_names.push_back(contig_names[i]);
        c->contig_lens.push_back(contig_lens[i]);
        c->contig_base.push_back(base);
        c->contig_tile_off.push_back(tiles);
        c->contig_pidx_off.push_back(pidx);
        base += contig_lens[i] + CONTIG_GAP;
        bool primary = is_primary ? is_primary[i] != 0 : true;
        if (primary && contig_lens[i] > 0) tiles += (contig_lens[i] + c->tile_bases - 1) / c->tile_bases;
        pidx += ((contig_lens[i] + HALF_WINDOW + 1) >> POS_INDEX_SHIFT) + 3;
    }
    c->contig_base.push_back(base);
    c->contig_tile_off.push_back(tiles);
    c->contig_pidx_off.push_back(pidx);
    int rc = CRATAC_OK;
    if ((rc = c->d_status.reserve(sizeof(int64_t) * ST_SLOTS)) != CRATAC_OK) { delete h; return rc; }
    cudaMemsetAsync(c->d_status.p, 0, sizeof(int64_t) * ST_SLOTS, c->stream);
    if (cudaMallocHost(&c->h_status, sizeof(int64_t) * ST_SLOTS) != cudaSuccess) { set_error("cudaMallocHost failed"); delete h; return CRATAC_ERR_CUDA; }
    if (cudaMallocHost(&c->h_sp_scalars, sizeof(int64_t) * 8) != cudaSuccess) { set_error("cudaMallocHost failed"); delete h; return CRATAC_ERR_CUDA; }
    memset(c->h_status, 0, sizeof(int64_t) * ST_SLOTS);
    if ((rc = upload_i64(c, c->d_contig_lens, c->contig_lens)) || (rc = upload_i64(c, c->d_contig_base, c->contig_base)) ||
        (rc = upload_i64(c, c->d_contig_tile_off, c->contig_tile_off)) || (rc = upload_i64(c, c->d_contig_pidx_off, c->contig_pidx_off))) {
        delete h;
        return rc;
    }
    *out = h;
    return CRATAC_OK;
}

void cratac_destroy(cratac_ctx* h) {
    if (!h) return;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->d_contig_lens, &c->d_contig_base, &c->d_contig_tile_off, &c->d_contig_pidx_off, &c->d_contig_table, &c->d_chrom,
                      &c->d_start, &c->d_end, &c->d_bc, &c->d_dup, &c->d_contig_frag_off, &c->d_pos_index, &c->d_text_own, &c->d_tile_counts,
                      &c->d_tile_offsets, &c->d_bc_keys, &c->d_bc_first, &c->d_bc_textoff, &c->d_slot_to_dense, &c->d_dense_key, &c->d_dense_first,
                      &c->d_dense_textoff, &c->d_dense_to_col, &c->d_col_to_dense, &c->d_bc_bitmap, &c->d_hist, &c->d_hist_values, &c->d_hist_weights,
                      &c->d_em_params, &c->d_em_work, &c->d_chain, &c->d_tile_summary, &c->d_run_start, &c->d_run_end, &c->d_bridge, &c->d_fill,
                      &c->d_peak_chrom, &c->d_peak_start, &c->d_peak_end, &c->d_peak_row, &c->d_peak_off, &c->d_bc_hits, &c->d_seg_off,
                      &c->d_cursor, &c->d_hits, &c->d_out_row, &c->d_out_cnt, &c->d_nnz_col, &c->d_indptr, &c->d_indices, &c->d_data,
                      &c->d_status, &c->d_scan_tmp, &c->d_scan_tmp_bc, &c->d_global_c2d, &c->d_nnz_ordered, &c->d_dense_bckey, &c->d_tile_desc,
                      &c->d_union, &c->d_col_keys, &c->d_sp_state, &c->d_sp_scalars, &c->d_sp_irr,
                      &c->d_bg_W, &c->d_bg_tiles, &c->d_bg_start, &c->d_bg_end, &c->d_bg_count,
                      &c->d_sib_perm, &c->d_sib_seg, &c->d_sib_tmp, &c->d_sib_hist, &c->d_slot_seg};
    for (DevBuf* b : bufs) b->release();
    if (c->h_status) cudaFreeHost(c->h_status);
    if (c->h_sp_scalars) cudaFreeHost(c->h_sp_scalars);
    if (c->bc_thread.joinable()) c->bc_thread.join();
    if (c->h_spec) cudaFreeHost(const_cast<int64_t*>(c->h_spec));
    if (c->spec_stream) cudaStreamDestroy(c->spec_stream);
    if (c->count_stream) cudaStreamDestroy(c->count_stream);
    if (c->count_event) cudaEventDestroy(c->count_event);
    if (c->fit_event) cudaEventDestroy(c->fit_event);
    if (c->spec_done_event) cudaEventDestroy(c->spec_done_event);
    if (c->h_ring) cudaFreeHost(c->h_ring);
    if (c->h_union_hash) cudaFreeHost(c->h_union_hash);
    if (c->h_union_perm) cudaFreeHost(c->h_union_perm);
    if (c->union_event) cudaEventDestroy(c->union_event);
    if (c->h_bc_hash) cudaFreeHost(c->h_bc_hash);
    if (c->h_bc_order) cudaFreeHost(c->h_bc_order);
    if (c->h_c2d_pin) cudaFreeHost(c->h_c2d_pin);
    if (c->bc_event) cudaEventDestroy(c->bc_event);
    if (c->bc_stream) cudaStreamDestroy(c->bc_stream);
    cudaStreamDestroy(c->stream);
    delete h;
}

int cratac_set_option(cratac_ctx* h, const char* key, int64_t value) {
    if (!h || !key) return CRATAC_ERR_ARG;
    Ctx* c = &h->c;
    if (!strcmp(key, "barcode_order")) {
        c->bc_order = (int)value;
        if (c->bc_is_slot) { c->bc_ready = false; c->bc_strings_ready = false; }
        return CRATAC_OK;
    }
    if (!strcmp(key, "barcode_table_slots")) {
        int64_t cap = 16;
        while (cap < value) cap <<= 1;
        c->bc_table_cap = (int)cap;
        return CRATAC_OK;
    }
    if (!strcmp(key, "run_capacity")) { c->run_cap = value; return CRATAC_OK; }
    if (!strcmp(key, "em_fast")) { c->em_fast = (int)value; return CRATAC_OK; }
    if (!strcmp(key, "defer_sync")) { c->defer_sync = (int)value; return CRATAC_OK; }
    if (!strcmp(key, "stream_decode")) { c->stream_decode = (int)value; return CRATAC_OK; }
    if (!strcmp(key, "stream_chunk_tiles")) { c->stream_chunk_tiles = value; return CRATAC_OK; }
    if (!strcmp(key, "py2_device_min")) { c->py2_device_min = value; return CRATAC_OK; }
    if (!strcmp(key, "union_device_min")) { c->union_device_min = value; return CRATAC_OK; }
    if (!strcmp(key, "decode_kernel")) { c->decode_kernel = (int)value; return CRATAC_OK; }
    if (!strcmp(key, "decode_pipeline")) { c->decode_pipeline = (int)value; return CRATAC_OK; }
    if (!strcmp(key, "speculate")) { c->speculate = (int)value; return CRATAC_OK; }
    if (!strcmp(key, "spec_iter")) { c->spec_iter = (int)value; return CRATAC_OK; }
    if (!strcmp(key, "file_chunk_bytes")) { c->file_chunk_bytes = value; return CRATAC_OK; }
    if (!strcmp(key, "decode_profile")) { c->decode_profile = (int)value; return CRATAC_OK; }
    if (!strcmp(key, "decode_mode")) { c->decode_mode = (int)value; return CRATAC_OK; }
    if (!strcmp(key, "decode_ctas_per_sm")) { c->sp_ctas_per_sm = (int)value; return CRATAC_OK; }
    if (!strcmp(key, "profile")) {
        if (value && !c->ev[0]) for (int i = 0; i < 2 * SG_COUNT; i++) cudaEventCreate(&c->ev[i]);
        c->profile = value != 0;
        return CRATAC_OK;
    }
    set_error("unknown option '%s'", key);
    return CRATAC_ERR_ARG;
}

int cratac_decode_profile(cratac_ctx* h, int64_t* cycles_out, int n) {
    Ctx* c = &h->c;
    if (!cycles_out || n < 1) { set_error("cratac_decode_profile: bad argument"); return CRATAC_ERR_ARG; }
    CRATAC_CUDA(cudaSetDevice(c->device));
    for (int i = 0; i < n; i++) cycles_out[i] = 0;
    if (!c->d_sp_scalars.p) return CRATAC_OK;
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    CRATAC_CUDA(cudaMemcpy(cycles_out, c->d_sp_scalars.as<int64_t>() + 8, sizeof(int64_t) * (size_t)std::min(n, 10), cudaMemcpyDeviceToHost));
    return CRATAC_OK;
}

int cratac_fit_cell_mixture(cratac_ctx* h, const int64_t* values, const int64_t* weights, int64_t n_bins, double odds_ratio, double out[8]) {
    if (!h || !values || !weights || !out) { set_error("cratac_fit_cell_mixture: null argument"); return CRATAC_ERR_ARG; }
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    return cell_fit(c, values, weights, n_bins, odds_ratio, out);
}

// ---- sibling scans of the fragments (siblings.cu)
int cratac_low_targeting(cratac_ctx* h, int64_t n_regions, const int32_t* chrom, const int32_t* start, const int32_t* end, int n_paddings,
                         const int64_t* paddings, int64_t* out) {
    if (!h || n_regions < 0 || (n_regions > 0 && (!chrom || !start || !end)) || !out || (n_paddings > 0 && !paddings)) { set_error("cratac_low_targeting: bad argument"); return CRATAC_ERR_ARG; }
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    return low_targeting(c, n_regions, chrom, start, end, n_paddings, paddings, out);
}

int cratac_insert_sizes(cratac_ctx* h, int64_t n_sel, const int64_t* cols, int exclude_non_primary, int64_t* table_out) {
    if (!h || n_sel < 0 || (n_sel > 0 && (!cols || !table_out))) { set_error("cratac_insert_sizes: bad argument"); return CRATAC_ERR_ARG; }
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    std::vector<int32_t> prim((size_t)c->n_contigs);
    for (int k = 0; k < c->n_contigs; k++) prim[(size_t)k] = c->contig_tile_off[(size_t)k + 1] > c->contig_tile_off[(size_t)k] ? 1 : 0;
    return insert_sizes(c, n_sel, cols, prim.data(), exclude_non_primary, table_out);
}

int cratac_counts_by_barcode(cratac_ctx* h, const int64_t* track_off, const int32_t* r_chrom, const int32_t* r_start, const int32_t* r_end,
                             const int8_t* r_minus, int only_contig, int n_species, const int32_t* species_of_contig, int rel_half, int64_t* out,
                             int64_t* relpos) {
    if (!h || !track_off || !out || !relpos) { set_error("cratac_counts_by_barcode: bad argument"); return CRATAC_ERR_ARG; }
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    return counts_by_barcode(c, track_off, r_chrom, r_start, r_end, r_minus, only_contig, n_species, species_of_contig, rel_half, out, relpos);
}

int cratac_speculation_stats(cratac_ctx* h, int64_t* kept, int64_t* redone) {
    if (kept) *kept = h->c.spec_hits;
    if (redone) *redone = h->c.spec_misses;
    return CRATAC_OK;
}

int64_t cratac_speculation_second_guesses(cratac_ctx* h) { return h->c.spec_second; }

int64_t cratac_pileup_tile_bases(void) { return pileup_tile_bases(); }

int64_t cratac_decode_declined(cratac_ctx* h, int reset) {
    int64_t n = h->c.decode_declined;
    if (reset) h->c.decode_declined = 0;
    return n;
}

int64_t cratac_launch_count(cratac_ctx* h, int reset) {
    int64_t n = h->c.launches;
    if (reset) h->c.launches = 0;
    return n;
}

static const char* const kStageNames[SG_COUNT] = {"decode_count_newlines", "decode_parse", "fragment_index", "pileup_window_hist", "hist_compact",
                                                 "em_fit", "pileup_call_peaks", "merge_runs", "overlap_count", "overlap_scatter",
                                                 "reduce_small", "reduce_large", "csc_build", "barcode_table"};

int cratac_num_stages(void) { return SG_COUNT; }
const char* cratac_stage_name(int id) { return (id >= 0 && id < SG_COUNT) ? kStageNames[id] : ""; }

// milliseconds of each stage's kernels in the last run (CUDA events on the context stream); -1 = not run
int cratac_stage_times(cratac_ctx* h, float* ms_out) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < SG_COUNT; i++) {
        ms_out[i] = -1.0f;
        if (c->profile && c->ev_used[i]) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, c->ev[2 * i], c->ev[2 * i + 1]) == cudaSuccess) ms_out[i] = ms;
            c->ev_used[i] = false;
        }
    }
    return CRATAC_OK;
}

// the CUDA stream all kernels of this context are launched on (cudaStream_t as an integer)
int cratac_stream(cratac_ctx* h, uint64_t* stream_out) {
    *stream_out = (uint64_t)(uintptr_t)h->c.stream;
    return CRATAC_OK;
}

int cratac_synchronize(cratac_ctx* h) {
    CRATAC_CUDA(cudaSetDevice(h->c.device));
    CRATAC_CUDA(cudaStreamSynchronize(h->c.stream));
    return CRATAC_OK;
}

// ------------------------------------------------------------------------------------------ decode
int cratac_decode_device(cratac_ctx* h, const char* d_text, int64_t nbytes) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (nbytes < 0 || (nbytes > 0 && !d_text)) { set_error("cratac_decode_device: bad buffer"); return CRATAC_ERR_ARG; }
    c->bc_is_slot = true;
    CRATAC_TRY(decode_text(c, d_text, nbytes));
    return CRATAC_OK;
}

int cratac_decode_host(cratac_ctx* h, const char* text, int64_t nbytes) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (nbytes < 0 || (nbytes > 0 && !text)) { set_error("cratac_decode_host: bad buffer"); return CRATAC_ERR_ARG; }
    c->bc_is_slot = true;
    if (nbytes > 0 && (c->stream_decode == 2 || (c->stream_decode == 1 && nbytes >= ((int64_t)64 << 20)))) {
        // large text: chunked copy overlapped with the decode of the chunks already on the device
        const int rc = decode_text_streamed(c, text, nbytes);
        if (rc != CRATAC_ERR_CAPACITY) return rc;
        // estimate or table too small: the text is on the device by now, decode it the exact way
        return decode_text(c, c->d_text_own.as<char>(), nbytes);
    }
    CRATAC_TRY(c->d_text_own.reserve((size_t)nbytes + 16));
    if (nbytes > 0) CRATAC_CUDA(cudaMemcpyAsync(c->d_text_own.p, text, (size_t)nbytes, cudaMemcpyHostToDevice, c->stream));
    return decode_text(c, c->d_text_own.as<char>(), nbytes);
}

int cratac_set_fragments(cratac_ctx* h, int64_t n, const int32_t* chrom, const int32_t* start, const int32_t* end, const int32_t* bc,
                         const int32_t* dup, int64_t n_barcodes, const char* const* barcodes) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (n < 0 || n_barcodes < 0 || (n > 0 && (!chrom || !start || !end || !bc))) { set_error("cratac_set_fragments: bad arguments"); return CRATAC_ERR_ARG; }
    size_t nn = (size_t)std::max<int64_t>(n, 1);
    CRATAC_TRY(c->d_chrom.reserve(4 * nn)); CRATAC_TRY(c->d_start.reserve(4 * nn)); CRATAC_TRY(c->d_end.reserve(4 * nn));
    CRATAC_TRY(c->d_bc.reserve(4 * nn)); CRATAC_TRY(c->d_dup.reserve(4 * nn));
    if (n > 0) {
        CRATAC_CUDA(cudaMemcpyAsync(c->d_chrom.p, chrom, 4 * nn, cudaMemcpyHostToDevice, c->stream));
        CRATAC_CUDA(cudaMemcpyAsync(c->d_start.p, start, 4 * nn, cudaMemcpyHostToDevice, c->stream));
        CRATAC_CUDA(cudaMemcpyAsync(c->d_end.p, end, 4 * nn, cudaMemcpyHostToDevice, c->stream));
        CRATAC_CUDA(cudaMemcpyAsync(c->d_bc.p, bc, 4 * nn, cudaMemcpyHostToDevice, c->stream));
        if (dup) CRATAC_CUDA(cudaMemcpyAsync(c->d_dup.p, dup, 4 * nn, cudaMemcpyHostToDevice, c->stream));
        else CRATAC_CUDA(cudaMemsetAsync(c->d_dup.p, 0, 4 * nn, c->stream));
    }
    c->n_fragments = n;
    c->n_barcodes = n_barcodes;
    c->dense_nfrag_valid = false;          // counted on demand by peak_matrix_device
    c->bc_is_slot = false;
    c->barcodes.clear();
    for (int64_t j = 0; j < n_barcodes; j++) c->barcodes.push_back(barcodes ? barcodes[j] : "");
    size_t nb = (size_t)std::max<int64_t>(n_barcodes, 1);
    std::vector<int32_t> ident(nb);
    for (size_t j = 0; j < nb; j++) ident[j] = (int32_t)j;
    CRATAC_TRY(c->d_col_to_dense.reserve(4 * nb)); CRATAC_TRY(c->d_dense_to_col.reserve(4 * nb));
    CRATAC_CUDA(cudaMemcpyAsync(c->d_col_to_dense.p, ident.data(), 4 * nb, cudaMemcpyHostToDevice, c->stream));
    CRATAC_CUDA(cudaMemcpyAsync(c->d_dense_to_col.p, ident.data(), 4 * nb, cudaMemcpyHostToDevice, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    c->bc_ready = true;
    c->bc_strings_ready = true;
    c->h_c2d.assign(ident.begin(), ident.begin() + n_barcodes);
    CRATAC_TRY(compute_max_len(c));
    return finalize_fragments(c);
}

int cratac_num_fragments(cratac_ctx* h, int64_t* n_fragments, int64_t* n_barcodes) {
    if (n_fragments) *n_fragments = h->c.n_fragments;
    if (n_barcodes) *n_barcodes = h->c.n_barcodes;
    return CRATAC_OK;
}

int cratac_get_fragments(cratac_ctx* h, int32_t* chrom, int32_t* start, int32_t* end, int32_t* bc, int32_t* dup) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    size_t nn = (size_t)c->n_fragments;
    if (nn == 0) return CRATAC_OK;
    if (chrom) CRATAC_CUDA(cudaMemcpyAsync(chrom, c->d_chrom.p, 4 * nn, cudaMemcpyDeviceToHost, c->stream));
    if (start) CRATAC_CUDA(cudaMemcpyAsync(start, c->d_start.p, 4 * nn, cudaMemcpyDeviceToHost, c->stream));
    if (end) CRATAC_CUDA(cudaMemcpyAsync(end, c->d_end.p, 4 * nn, cudaMemcpyDeviceToHost, c->stream));
    if (dup) CRATAC_CUDA(cudaMemcpyAsync(dup, c->d_dup.p, 4 * nn, cudaMemcpyDeviceToHost, c->stream));
    if (bc) {
        CRATAC_TRY(build_barcodes(c));
        CRATAC_CUDA(cudaMemcpyAsync(bc, c->d_bc.p, 4 * nn, cudaMemcpyDeviceToHost, c->stream));
        if (c->bc_is_slot) {
            std::vector<int32_t> s2d((size_t)c->bc_table_cap), d2c((size_t)std::max<int64_t>(c->n_barcodes, 1));
            CRATAC_CUDA(cudaMemcpyAsync(s2d.data(), c->d_slot_to_dense.p, 4 * s2d.size(), cudaMemcpyDeviceToHost, c->stream));
            CRATAC_CUDA(cudaMemcpyAsync(d2c.data(), c->d_dense_to_col.p, 4 * d2c.size(), cudaMemcpyDeviceToHost, c->stream));
            CRATAC_CUDA(cudaStreamSynchronize(c->stream));
            for (size_t i = 0; i < nn; i++) bc[i] = d2c[s2d[bc[i]]];
        }
    }
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    return CRATAC_OK;
}

int cratac_max_barcode_len(cratac_ctx* h, int64_t* len) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    CRATAC_TRY(build_barcode_strings(c));
    size_t m = 0;
    for (const std::string& s : c->barcodes) m = std::max(m, s.size());
    *len = (int64_t)m;
    return CRATAC_OK;
}

int cratac_get_barcodes(cratac_ctx* h, char* out, int64_t stride) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    CRATAC_TRY(build_barcode_strings(c));
    for (size_t j = 0; j < c->barcodes.size(); j++) {
        if ((int64_t)c->barcodes[j].size() + 1 > stride) { set_error("barcode stride %lld too small", (long long)stride); return CRATAC_ERR_ARG; }
        memset(out + j * stride, 0, (size_t)stride);
        memcpy(out + j * stride, c->barcodes[j].data(), c->barcodes[j].size());
    }
    return CRATAC_OK;
}

// ------------------------------------------------------------------------------------------ cut sites
int cratac_count_cut_sites(cratac_ctx* h, int64_t* values, int64_t* weights, int64_t cap, int64_t* n_bins) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    CRATAC_TRY(window_histogram(c));
    CRATAC_TRY(sync_status(c));
    int64_t nb = c->h_status[ST_N_BINS];
    if (n_bins) *n_bins = nb;
    int64_t k = std::min(nb, cap);
    if (k > 0 && values) CRATAC_CUDA(cudaMemcpyAsync(values, c->d_hist_values.p, 8 * (size_t)k, cudaMemcpyDeviceToHost, c->stream));
    if (k > 0 && weights) CRATAC_CUDA(cudaMemcpyAsync(weights, c->d_hist_weights.p, 8 * (size_t)k, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    return CRATAC_OK;
}

int cratac_window_counts(cratac_ctx* h, int contig, int32_t* W) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (contig < 0 || contig >= c->n_contigs) { set_error("contig index out of range"); return CRATAC_ERR_ARG; }
    if (c->contig_tile_off[contig + 1] == c->contig_tile_off[contig]) { set_error("contig %d is not a primary contig", contig); return CRATAC_ERR_ARG; }
    int64_t L = c->contig_lens[contig];
    DevBuf d_W;
    CRATAC_TRY(d_W.reserve(4 * (size_t)std::max<int64_t>(L, 1)));
    int rc = dump_window_counts(c, contig, d_W.as<int32_t>());
    if (rc == CRATAC_OK && cudaMemcpyAsync(W, d_W.p, 4 * (size_t)L, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = CRATAC_ERR_CUDA;
    if (rc == CRATAC_OK) rc = sync_status(c);
    d_W.release();
    return rc;
}

int cratac_bedgraph_rows(cratac_ctx* h, int contig, int64_t* start, int64_t* end, int64_t* count, int64_t cap, int64_t* n_rows) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (contig < 0 || contig >= c->n_contigs) { set_error("contig index out of range"); return CRATAC_ERR_ARG; }
    if (c->contig_tile_off[contig + 1] == c->contig_tile_off[contig]) { set_error("contig %d is not a primary contig", contig); return CRATAC_ERR_ARG; }
    // rows are run-length encoded on the device (csrc/bedgraph.cu) and kept there: a first call with cap = 0 sizes the
    // caller's arrays, the second one only copies
    CRATAC_TRY(bedgraph_rows_device(c, contig));
    const int64_t n = c->bg_rows;
    if (n_rows) *n_rows = n;
    if (n == 0 || cap < n || !start || !end || !count) return sync_status(c);
    // 12 bytes per row cross PCIe; widened to the int64 of the interface on all host cores
    std::vector<int32_t> tmp[3];
    const void* src[3] = {c->d_bg_start.p, c->d_bg_end.p, c->d_bg_count.p};
    int64_t* dst[3] = {start, end, count};
    for (int k = 0; k < 3; k++) {
        tmp[k].resize((size_t)n);
        CRATAC_CUDA(cudaMemcpyAsync(tmp[k].data(), src[k], 4 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    }
    CRATAC_TRY(sync_status(c));
    const int n_thr = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)std::thread::hardware_concurrency(), n / (1 << 16) + 1));
    std::vector<std::thread> pool;
    auto widen = [&](int t) {
        const int64_t i0 = n * t / n_thr, i1 = n * (t + 1) / n_thr;
        for (int k = 0; k < 3; k++) {
            const int32_t* sp = tmp[k].data();
            int64_t* dp = dst[k];
            for (int64_t i = i0; i < i1; i++) dp[i] = sp[i];
        }
    };
    for (int t = 1; t < n_thr; t++) pool.emplace_back(widen, t);
    widen(0);
    for (auto& t : pool) t.join();
    return CRATAC_OK;
}

// ------------------------------------------------------------------------------------------ threshold
static int read_fit(Ctx* c, double params_out[6], int* has_params, int64_t* threshold_out, int64_t* em_it, int64_t* inner_it) {
    CRATAC_TRY(sync_status(c));
    double p[6];
    CRATAC_CUDA(cudaMemcpyAsync(p, c->d_em_params.p, sizeof(p), cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    if (params_out) memcpy(params_out, p, sizeof(p));
    if (has_params) *has_params = (int)c->h_status[ST_HAS_PARAMS];
    if (threshold_out) *threshold_out = c->h_status[ST_THRESHOLD];
    if (em_it) *em_it = c->h_status[ST_EM_ITERS];
    if (inner_it) *inner_it = c->h_status[ST_EM_INNER];
    return CRATAC_OK;
}

int cratac_fit_threshold(cratac_ctx* h, const int64_t* values, const int64_t* weights, int64_t n_bins, double odds_ratio,
                         double params_out[6], int* has_params, int64_t* threshold_out, int64_t* em_iterations, int64_t* inner_iterations) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (n_bins < 0 || n_bins >= HIST_CAP) { set_error("cratac_fit_threshold: bad bin count"); return CRATAC_ERR_ARG; }
    CRATAC_TRY(c->d_hist_values.reserve(sizeof(int64_t) * HIST_CAP));
    CRATAC_TRY(c->d_hist_weights.reserve(sizeof(int64_t) * HIST_CAP));
    if (n_bins > 0) {
        CRATAC_CUDA(cudaMemcpyAsync(c->d_hist_values.p, values, 8 * (size_t)n_bins, cudaMemcpyHostToDevice, c->stream));
        CRATAC_CUDA(cudaMemcpyAsync(c->d_hist_weights.p, weights, 8 * (size_t)n_bins, cudaMemcpyHostToDevice, c->stream));
    }
    CRATAC_TRY(fit_threshold_arrays(c, c->d_hist_values.as<int64_t>(), c->d_hist_weights.as<int64_t>(), n_bins, odds_ratio));
    return read_fit(c, params_out, has_params, threshold_out, em_iterations, inner_iterations);
}

int cratac_fit_threshold_device(cratac_ctx* h, double odds_ratio, double params_out[6], int* has_params, int64_t* threshold_out) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    CRATAC_TRY(fit_threshold_device(c, odds_ratio));
    if (c->defer_sync) return CRATAC_OK;     // queued only: cratac_fit_threshold_result picks the numbers up
    return read_fit(c, params_out, has_params, threshold_out, nullptr, nullptr);
}

int cratac_fit_threshold_result(cratac_ctx* h, double params_out[6], int* has_params, int64_t* threshold_out) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    return read_fit(c, params_out, has_params, threshold_out, nullptr, nullptr);
}

// ------------------------------------------------------------------------------------------ peaks
static int call_peaks_with_retry(Ctx* c) {
    for (int attempt = 0; attempt < 4; attempt++) {
        CRATAC_TRY(call_peaks_device(c));
        int rc = sync_status(c);
        if (rc == CRATAC_ERR_CAPACITY && c->run_cap < ((int64_t)1 << 30)) { c->run_cap <<= 2; continue; }
        if (rc != CRATAC_OK) return rc;
        c->n_peaks = c->h_status[ST_N_PEAKS];
        c->peaks_on_device_count = false;
        return CRATAC_OK;
    }
    set_error("run list overflow");
    return CRATAC_ERR_CAPACITY;
}

int cratac_call_peaks(cratac_ctx* h, int64_t threshold, int64_t* n_peaks) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (threshold >= 0) {
        int64_t* st = c->d_status.as<int64_t>();
        CRATAC_CUDA(cudaMemcpyAsync(&st[ST_THRESHOLD], &threshold, 8, cudaMemcpyHostToDevice, c->stream));
        CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    }
    CRATAC_TRY(call_peaks_with_retry(c));
    if (n_peaks) *n_peaks = c->n_peaks;
    return CRATAC_OK;
}

// DETECT_PEAKS.main on bedgraph rows: `n` signal intervals (count >= threshold) sorted by (contig, start)
int cratac_peaks_from_rows(cratac_ctx* h, int64_t n, const int32_t* chrom, const int64_t* start, const int64_t* end, int64_t* n_peaks) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    CRATAC_TRY(merge_runs_from_host(c, n, chrom, start, end));
    CRATAC_TRY(fetch_peak_count(c));
    if (n_peaks) *n_peaks = c->n_peaks;
    return CRATAC_OK;
}

int cratac_get_peaks(cratac_ctx* h, int32_t* chrom, int32_t* start, int32_t* end) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    CRATAC_TRY(fetch_peak_count(c));
    size_t n = (size_t)c->n_peaks;
    if (n == 0) return CRATAC_OK;
    // device arrays are sorted by (contig, start); row ids restore the caller's order for user peaks
    std::vector<int32_t> pc(n), ps(n), pe(n), pr(n);
    CRATAC_CUDA(cudaMemcpyAsync(pc.data(), c->d_peak_chrom.p, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaMemcpyAsync(ps.data(), c->d_peak_start.p, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaMemcpyAsync(pe.data(), c->d_peak_end.p, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaMemcpyAsync(pr.data(), c->d_peak_row.p, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; i++) {
        size_t r = (size_t)pr[i];
        if (chrom) chrom[r] = pc[i];
        if (start) start[r] = ps[i];
        if (end) end[r] = pe[i];
    }
    return CRATAC_OK;
}

int cratac_set_peaks(cratac_ctx* h, int64_t n, const int32_t* chrom, const int32_t* start, const int32_t* end) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (n < 0 || (n > 0 && (!chrom || !start || !end))) { set_error("cratac_set_peaks: bad arguments"); return CRATAC_ERR_ARG; }
    std::vector<int32_t> order((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        order[i] = (int32_t)i;
        if (chrom[i] < 0 || chrom[i] >= c->n_contigs) { set_error("peak %lld: contig index out of range", (long long)i); return CRATAC_ERR_ARG; }
    }
    std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
        if (chrom[x] != chrom[y]) return chrom[x] < chrom[y];
        if (start[x] != start[y]) return start[x] < start[y];
        return end[x] < end[y];
    });
    std::vector<int32_t> pc((size_t)std::max<int64_t>(n, 1)), ps(pc.size()), pe(pc.size());
    std::vector<int64_t> off((size_t)c->n_contigs + 1, 0);
    for (int64_t i = 0; i < n; i++) {
        int32_t o = order[i];
        pc[i] = chrom[o]; ps[i] = start[o]; pe[i] = end[o];
        off[chrom[o] + 1]++;
        // tenkit Regions merges overlapping regions on load (tenkit/regions.py:91-110); the merged name then
        // misses in peaks_dict (generate_peak_matrix/__init__.py:114 -> KeyError).  Touching peaks stay separate.
        if (i > 0 && pc[i - 1] == pc[i] && ps[i] < pe[i - 1]) {
            set_error("peaks %d and %d overlap (the reference fails with KeyError on overlapping peaks)", order[i - 1], o);
            return CRATAC_ERR_PEAKS_OVERLAP;
        }
    }
    for (int k = 0; k < c->n_contigs; k++) off[k + 1] += off[k];
    size_t nn = pc.size();
    CRATAC_TRY(c->d_peak_chrom.reserve(4 * nn)); CRATAC_TRY(c->d_peak_start.reserve(4 * nn)); CRATAC_TRY(c->d_peak_end.reserve(4 * nn));
    CRATAC_TRY(c->d_peak_row.reserve(4 * nn)); CRATAC_TRY(c->d_peak_off.reserve(8 * off.size()));
    if (n > 0) {
        CRATAC_CUDA(cudaMemcpyAsync(c->d_peak_chrom.p, pc.data(), 4 * nn, cudaMemcpyHostToDevice, c->stream));
        CRATAC_CUDA(cudaMemcpyAsync(c->d_peak_start.p, ps.data(), 4 * nn, cudaMemcpyHostToDevice, c->stream));
        CRATAC_CUDA(cudaMemcpyAsync(c->d_peak_end.p, pe.data(), 4 * nn, cudaMemcpyHostToDevice, c->stream));
        CRATAC_CUDA(cudaMemcpyAsync(c->d_peak_row.p, order.data(), 4 * nn, cudaMemcpyHostToDevice, c->stream));
    }
    CRATAC_CUDA(cudaMemcpyAsync(c->d_peak_off.p, off.data(), 8 * off.size(), cudaMemcpyHostToDevice, c->stream));
    int64_t* st = c->d_status.as<int64_t>();
    CRATAC_CUDA(cudaMemcpyAsync(&st[ST_N_PEAKS], &n, 8, cudaMemcpyHostToDevice, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    c->n_peaks = n;
    c->peaks_on_device_count = false;
    return CRATAC_OK;
}

// ------------------------------------------------------------------------------------------ matrix
int cratac_peak_matrix(cratac_ctx* h, int64_t* nnz) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    CRATAC_TRY(peak_matrix_device(c));
    CRATAC_TRY(sync_status(c));
    if (c->peaks_on_device_count) { c->n_peaks = c->h_status[ST_N_PEAKS]; c->peaks_on_device_count = false; }
    if (nnz) *nnz = c->nnz;
    return CRATAC_OK;
}

int cratac_peak_matrix_begin(cratac_ctx* h) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    return peak_matrix_hits(c);
}

int cratac_peak_matrix_finish(cratac_ctx* h, int64_t* nnz) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    CRATAC_TRY(peak_matrix_csc(c));
    CRATAC_TRY(sync_status(c));
    if (c->peaks_on_device_count) { c->n_peaks = c->h_status[ST_N_PEAKS]; c->peaks_on_device_count = false; }
    if (nnz) *nnz = c->nnz;
    return CRATAC_OK;
}

int cratac_union_columns(cratac_ctx* h, int world, int64_t max_barcodes, const int64_t* d_gathered, const int64_t* n_fragments, const int64_t* n_barcodes,
                         uint64_t stream, int64_t* n_cols) {
    if (!h || world < 1 || max_barcodes < 0 || !n_fragments || !n_barcodes || !n_cols) { set_error("cratac_union_columns: bad argument"); return CRATAC_ERR_ARG; }
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    return union_columns(c, world, max_barcodes, d_gathered, n_fragments, n_barcodes, stream ? (cudaStream_t)(uintptr_t)stream : c->stream, n_cols);
}

int cratac_column_keys_device(cratac_ctx* h, const int64_t** d_keys, int64_t* n_cols) {
    Ctx* c = &h->c;
    if (!c->global_cols) { set_error("cratac_column_keys_device: no global column map"); return CRATAC_ERR_ARG; }
    if (d_keys) *d_keys = c->d_col_keys.as<int64_t>();
    if (n_cols) *n_cols = c->n_global_cols;
    return CRATAC_OK;
}

int cratac_get_matrix(cratac_ctx* h, int64_t* indptr, int64_t* indices, int32_t* data) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (indptr) CRATAC_CUDA(cudaMemcpyAsync(indptr, c->d_indptr.p, 8 * (size_t)(c->n_cols + 1), cudaMemcpyDeviceToHost, c->stream));
    if (indices && c->nnz > 0) CRATAC_CUDA(cudaMemcpyAsync(indices, c->d_indices.p, 8 * (size_t)c->nnz, cudaMemcpyDeviceToHost, c->stream));
    if (data && c->nnz > 0) CRATAC_CUDA(cudaMemcpyAsync(data, c->d_data.p, 4 * (size_t)c->nnz, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    return CRATAC_OK;
}

int cratac_matrix_device(cratac_ctx* h, const int64_t** d_indptr, const int64_t** d_indices, const int32_t** d_data) {
    Ctx* c = &h->c;
    if (d_indptr) *d_indptr = c->d_indptr.as<int64_t>();
    if (d_indices) *d_indices = c->d_indices.as<int64_t>();
    if (d_data) *d_data = c->d_data.as<int32_t>();
    return CRATAC_OK;
}

// ---- FILTER_PEAK_MATRIX -------------------------------------------------------------------------------------------
// the current matrix from host CSC arrays (the stage starts from raw_peak_bc_matrix.h5)
int cratac_set_matrix(cratac_ctx* h, int64_t n_cols, const int64_t* indptr, const int64_t* indices, const int32_t* data) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (n_cols < 0 || !indptr) { set_error("cratac_set_matrix: bad arguments"); return CRATAC_ERR_ARG; }
    const int64_t nnz = indptr[n_cols];
    if (indptr[0] != 0 || nnz < 0 || (nnz > 0 && (!indices || !data))) { set_error("cratac_set_matrix: bad indptr"); return CRATAC_ERR_ARG; }
    for (int64_t j = 0; j < n_cols; j++)
        if (indptr[j + 1] < indptr[j]) { set_error("cratac_set_matrix: indptr decreases at column %lld", (long long)j); return CRATAC_ERR_ARG; }
    CRATAC_TRY(c->d_indptr.reserve(8 * (size_t)(n_cols + 1)));
    CRATAC_TRY(c->d_indices.reserve(8 * (size_t)std::max<int64_t>(nnz, 1))); CRATAC_TRY(c->d_data.reserve(4 * (size_t)std::max<int64_t>(nnz, 1)));
    CRATAC_CUDA(cudaMemcpyAsync(c->d_indptr.p, indptr, 8 * (size_t)(n_cols + 1), cudaMemcpyHostToDevice, c->stream));
    if (nnz > 0) {
        CRATAC_CUDA(cudaMemcpyAsync(c->d_indices.p, indices, 8 * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
        CRATAC_CUDA(cudaMemcpyAsync(c->d_data.p, data, 4 * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    }
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    c->n_cols = n_cols; c->nnz = nnz;
    return CRATAC_OK;
}

int cratac_select_columns(cratac_ctx* h, int64_t n_sel, const int64_t* cols, int drop_empty, int64_t* n_cols_out, int64_t* nnz_out) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (n_sel < 0 || (n_sel > 0 && !cols)) { set_error("cratac_select_columns: bad arguments"); return CRATAC_ERR_ARG; }
    CRATAC_TRY(select_columns(c, n_sel, cols, drop_empty));
    if (n_cols_out) *n_cols_out = c->n_sel_cols;
    if (nnz_out) *nnz_out = c->sel_nnz;
    return CRATAC_OK;
}

int cratac_get_selected(cratac_ctx* h, int64_t* indptr, int64_t* indices, int32_t* data, int64_t* kept) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (indptr) CRATAC_CUDA(cudaMemcpyAsync(indptr, c->d_sel_indptr.p, 8 * (size_t)(c->n_sel_cols + 1), cudaMemcpyDeviceToHost, c->stream));
    if (indices && c->sel_nnz > 0) CRATAC_CUDA(cudaMemcpyAsync(indices, c->d_sel_indices.p, 8 * (size_t)c->sel_nnz, cudaMemcpyDeviceToHost, c->stream));
    if (data && c->sel_nnz > 0) CRATAC_CUDA(cudaMemcpyAsync(data, c->d_sel_data.p, 4 * (size_t)c->sel_nnz, cudaMemcpyDeviceToHost, c->stream));
    if (kept && c->n_sel_cols > 0) CRATAC_CUDA(cudaMemcpyAsync(kept, c->d_sel_kept.p, 8 * (size_t)c->n_sel_cols, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    return CRATAC_OK;
}

int cratac_hist_device(cratac_ctx* h, const uint64_t** d_hist, int64_t* cap) {
    if (d_hist) *d_hist = h->c.d_hist.as<uint64_t>();
    if (cap) *cap = HIST_CAP;
    return CRATAC_OK;
}

// ------------------------------------------------------------------------------------------ whole path
// Peaks + matrix started from the threshold the fit shows after `spec_iter` EM iterations, on a second stream, while the
// fit (one CTA) runs to convergence on the context's own: the 147 other SMs would idle for those milliseconds.  The result
// stands only if the fit ends on the very same threshold; otherwise (the estimate still moved, or the fit re-seeded itself,
// peaks.py:183-191) peaks + matrix are simply done again.  -> 1: speculative result is in place, 0: not used.
static int speculate_peaks_and_matrix(Ctx* c, int* used) {
    *used = 0;
    if (!c->spec_armed) return CRATAC_OK;
    // wait for the tentative threshold or for the fit itself, whichever comes first
    bool tentative = false;
    for (;;) {
        if (c->h_spec[1] == c->spec_seq) { tentative = true; break; }
        const cudaError_t q = cudaEventQuery(c->fit_event);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) { set_error("fit failed: %s", cudaGetErrorString(q)); return CRATAC_ERR_CUDA; }
    }
    if (!tentative || cudaEventQuery(c->fit_event) == cudaSuccess) return CRATAC_OK;   // nothing left to hide behind
    int64_t tent = c->h_spec[0];
    cudaStream_t own = c->stream;
    c->stream = c->spec_stream; c->thr_slot = ST_THRESHOLD_SPEC;
    int rc = call_peaks_device(c);
    if (rc == CRATAC_OK) rc = peak_matrix_device(c);
    // the fit may correct itself while it runs on (second guess, em_fast.cu): the pass is queued again behind the first
    // one, still beside the fit
    for (;;) {
        if (c->h_spec[3] == c->spec_seq) {
            tent = c->h_spec[2];
            c->thr_slot = ST_THRESHOLD_SPEC2;
            c->spec_second++;
            rc = call_peaks_device(c);
            if (rc == CRATAC_OK) rc = peak_matrix_device(c);
            break;
        }
        const cudaError_t q = cudaEventQuery(c->fit_event);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) { c->stream = own; c->thr_slot = ST_THRESHOLD; set_error("fit failed: %s", cudaGetErrorString(q)); return CRATAC_ERR_CUDA; }
    }
    if (rc == CRATAC_OK) rc = sync_status(c);
    else cudaStreamSynchronize(c->spec_stream);     // whatever was queued must not run into the pass that replaces it
    c->stream = own; c->thr_slot = ST_THRESHOLD;
    c->host_mark("speculative_peaks_matrix");
    // the status words are shared with the fit: an error it raised meanwhile is the run's error, not the pass's
    if (rc == CRATAC_ERR_EM_ZERODIV || rc == CRATAC_ERR_EM_INDEX) return rc;
    // the verdict: the fit's own threshold
    int64_t* st = c->d_status.as<int64_t>();
    CRATAC_CUDA(cudaEventSynchronize(c->fit_event));
    CRATAC_CUDA(cudaMemcpyAsync(&c->h_status[ST_THRESHOLD], &st[ST_THRESHOLD], 8, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    if (rc == CRATAC_OK && c->h_status[ST_THRESHOLD] == tent) { *used = 1; c->spec_hits++; }
    else c->spec_misses++;
    return CRATAC_OK;     // an error of the speculative pass is not an error of the run: the pass is redone
}

// the next fit launched publishes a tentative threshold (stream, event and the host-mapped words are made on first use)
static int arm_speculation(Ctx* c) {
    if (!c->spec_stream) {
        CRATAC_CUDA(cudaStreamCreateWithFlags(&c->spec_stream, cudaStreamNonBlocking));
        CRATAC_CUDA(cudaEventCreateWithFlags(&c->fit_event, cudaEventDisableTiming));
        void* hp = nullptr;
        CRATAC_CUDA(cudaHostAlloc(&hp, 64, cudaHostAllocMapped));
        c->h_spec = reinterpret_cast<volatile int64_t*>(hp);
        c->h_spec[0] = 0; c->h_spec[1] = 0; c->h_spec[2] = 0; c->h_spec[3] = 0;
    }
    c->spec_seq++;
    c->spec_armed = true;
    return CRATAC_OK;
}

// ---- the same speculation as building blocks (multi-GPU step, cratac/multi.py): every rank fits the same all-reduced
//      histogram, so the tentative threshold -- and whether one is published at all -- is the same on every rank
int cratac_fit_speculative_launch(cratac_ctx* h, double odds_ratio) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (c->em_fast < 2) { set_error("speculation needs the register-resident fit kernel (em_fast = 2)"); return CRATAC_ERR_ARG; }
    CRATAC_TRY(arm_speculation(c));
    const int rc = fit_threshold_device(c, odds_ratio);
    c->spec_armed = false;
    CRATAC_TRY(rc);
    CRATAC_CUDA(cudaEventRecord(c->fit_event, c->stream));
    return CRATAC_OK;
}

// waits for the tentative threshold or for the end of the fit; *published = 1 and *threshold set iff the fit published one
// (a property of the histogram, not of timing: a fit of fewer than spec_iter iterations publishes none)
int cratac_fit_tentative(cratac_ctx* h, int* published, int64_t* threshold) {
    Ctx* c = &h->c;
    if (!c->fit_event || !published || !threshold) { set_error("cratac_fit_tentative: no speculative fit in flight"); return CRATAC_ERR_ARG; }
    for (;;) {
        if (c->h_spec[1] == c->spec_seq) break;
        const cudaError_t q = cudaEventQuery(c->fit_event);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) { set_error("fit failed: %s", cudaGetErrorString(q)); return CRATAC_ERR_CUDA; }
    }
    *published = c->h_spec[1] == c->spec_seq ? 1 : 0;
    *threshold = *published ? c->h_spec[0] : -1;
    return CRATAC_OK;
}

// waits for the fit's second guess or for the end of the fit; *published = 1 and *threshold set iff the fit corrected its
// tentative threshold (again a property of the histogram: the same on every rank of a sharded run)
int cratac_fit_revised(cratac_ctx* h, int* published, int64_t* threshold) {
    Ctx* c = &h->c;
    if (!c->fit_event || !published || !threshold) { set_error("cratac_fit_revised: no speculative fit in flight"); return CRATAC_ERR_ARG; }
    for (;;) {
        if (c->h_spec[3] == c->spec_seq) break;
        const cudaError_t q = cudaEventQuery(c->fit_event);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) { set_error("fit failed: %s", cudaGetErrorString(q)); return CRATAC_ERR_CUDA; }
    }
    *published = c->h_spec[3] == c->spec_seq ? 1 : 0;
    *threshold = *published ? c->h_spec[2] : -1;
    if (*published) c->spec_second++;
    return CRATAC_OK;
}

// on = 1: the library's calls are queued on the speculation stream and the peak kernels read the tentative threshold
// (on = 2: the fit's second guess); on = 0: back to the context's own stream and the fitted threshold.
// *stream_out: the stream now in use.
int cratac_speculation_scope(cratac_ctx* h, int on, uint64_t* stream_out) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (on) {
        if (!c->spec_stream) { set_error("cratac_speculation_scope: no speculative fit was launched"); return CRATAC_ERR_ARG; }
        if (!c->own_stream_saved) { c->own_stream_saved = c->stream; c->stream = c->spec_stream; }
        c->thr_slot = on == 2 ? ST_THRESHOLD_SPEC2 : ST_THRESHOLD_SPEC;
        // the column map may still be on its way on the side stream of cratac_union_columns
        if (c->union_event) CRATAC_CUDA(cudaStreamWaitEvent(c->spec_stream, c->union_event, 0));
    } else {
        if (c->own_stream_saved) {
            // whatever follows on the context's own stream sees the speculative results
            if (!c->spec_done_event) CRATAC_CUDA(cudaEventCreateWithFlags(&c->spec_done_event, cudaEventDisableTiming));
            CRATAC_CUDA(cudaEventRecord(c->spec_done_event, c->spec_stream));
            CRATAC_CUDA(cudaStreamWaitEvent(c->own_stream_saved, c->spec_done_event, 0));
            c->stream = c->own_stream_saved; c->own_stream_saved = nullptr;
        }
        c->thr_slot = ST_THRESHOLD;
    }
    if (stream_out) *stream_out = (uint64_t)(uintptr_t)c->stream;
    return CRATAC_OK;
}

static int run_tail(Ctx* c, double odds_ratio, cratac_run_result* r) {
    build_barcodes_async(c);
    CRATAC_TRY(window_histogram(c));
    c->spec_armed = false;
    if (c->speculate && c->em_fast >= 2 && !c->defer_sync) CRATAC_TRY(arm_speculation(c));
    CRATAC_TRY(fit_threshold_device(c, odds_ratio));
    int used = 0;
    if (c->spec_armed) {
        CRATAC_CUDA(cudaEventRecord(c->fit_event, c->stream));
        // the speculative kernels read what the kernels before the fit wrote (histogram pass: tile bounds, summaries)
        CRATAC_TRY(speculate_peaks_and_matrix(c, &used));
        c->spec_armed = false;
    }
    if (!used) {
        CRATAC_TRY(call_peaks_device(c));
        c->host_mark("launch_hist_fit_peaks");
        CRATAC_TRY(peak_matrix_device(c));
        c->host_mark("peak_matrix_total");
    }
    int rc = sync_status(c);
    c->host_mark("final_sync");
    if (rc == CRATAC_ERR_CAPACITY && c->run_cap < ((int64_t)1 << 30)) {
        // the run list overflowed: grow it and redo peaks + matrix (the fitted threshold is still on the device)
        c->run_cap <<= 2;
        CRATAC_TRY(call_peaks_with_retry(c));
        c->peaks_on_device_count = true;
        CRATAC_TRY(peak_matrix_device(c));
        rc = sync_status(c);
    }
    if (rc != CRATAC_OK) return rc;
    c->n_peaks = c->h_status[ST_N_PEAKS];
    c->peaks_on_device_count = false;
    if (r) {
        r->n_fragments = c->n_fragments; r->n_barcodes = c->n_barcodes; r->n_bins = c->h_status[ST_N_BINS];
        r->threshold = c->h_status[ST_THRESHOLD]; r->has_params = c->h_status[ST_HAS_PARAMS]; r->n_peaks = c->n_peaks; r->nnz = c->nnz;
        r->em_iterations = c->h_status[ST_EM_ITERS]; r->em_inner_iterations = c->h_status[ST_EM_INNER];
        r->n_hits = c->n_hits;
        CRATAC_CUDA(cudaMemcpyAsync(r->params, c->d_em_params.p, sizeof(double) * 6, cudaMemcpyDeviceToHost, c->stream));
        CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    }
    return CRATAC_OK;
}

int cratac_run(cratac_ctx* h, const char* text, int64_t nbytes, int text_is_device, double odds_ratio, cratac_run_result* result) {
    h->c.host_reset();
    if (text_is_device) CRATAC_TRY(cratac_decode_device(h, text, nbytes));
    else CRATAC_TRY(cratac_decode_host(h, text, nbytes));
    h->c.host_mark("decode_total");
    return run_tail(&h->c, odds_ratio, result);
}

// cycles per phase of the last threshold fit (profiling aid): e_step, m_sums, suffix, bracket, nodes, verify,
// doubling, descent, direct -- 9 doubles
int cratac_em_debug(cratac_ctx* h, double* out) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    CRATAC_CUDA(cudaMemcpyAsync(out, c->d_em_params.as<double>() + 8, 11 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    return CRATAC_OK;
}

// host wall-clock phases of the last cratac_run as "name=ms;name=ms;..." (option "profile" = 1)
int cratac_host_times(cratac_ctx* h, char* out, int64_t cap) {
    std::string s;
    for (auto& kv : h->c.host_ms) { char b[96]; snprintf(b, sizeof(b), "%s=%.3f;", kv.first.c_str(), kv.second); s += b; }
    if ((int64_t)s.size() + 1 > cap) return CRATAC_ERR_ARG;
    memcpy(out, s.c_str(), s.size() + 1);
    return CRATAC_OK;
}

int cratac_run_from_fragments(cratac_ctx* h, double odds_ratio, cratac_run_result* result) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    return run_tail(c, odds_ratio, result);
}

// ------------------------------------------------------------------------------------------ multi-GPU building blocks
// (one process per GPU; the exchanges themselves are NCCL calls made by the host code on these device buffers)
int cratac_status_device(cratac_ctx* h, int64_t** d_status, int64_t* n_slots) {
    if (d_status) *d_status = h->c.d_status.as<int64_t>();
    if (n_slots) *n_slots = ST_SLOTS;
    return CRATAC_OK;
}

// pileup + window + histogram for this shard's contigs; the histogram (cratac_hist_device) and the largest
// count (status slot 6) stay on the device for the all-reduce
int cratac_window_histogram(cratac_ctx* h) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    CRATAC_TRY(window_histogram(c));
    return c->defer_sync ? CRATAC_OK : sync_status(c);
}

// re-derive the compact (value, weight) arrays from the (all-reduced) device histogram
int cratac_compact_histogram(cratac_ctx* h) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    CRATAC_TRY(compact_histogram(c));
    return c->defer_sync ? CRATAC_OK : sync_status(c);
}

// barcode dictionary of this shard by dense id: 64-bit keys, first line (shard-local), CPython-2.7 hash
int cratac_barcode_table_device(cratac_ctx* h, const uint64_t** d_keys, const uint64_t** d_first, const int64_t** d_py2hash, int64_t* n) {
    Ctx* c = &h->c;
    if (!c->bc_is_slot) { set_error("barcode table is only available after a decode"); return CRATAC_ERR_ARG; }
    if (d_keys) *d_keys = c->d_dense_bckey.as<uint64_t>();
    if (d_first) *d_first = c->d_dense_first.as<uint64_t>();
    if (d_py2hash) *d_py2hash = c->d_dense_key.as<int64_t>();
    if (n) *n = c->n_barcodes;
    return CRATAC_OK;
}

// matrix columns = a global barcode list: d_col_to_dense[j] = this shard's dense barcode id of column j, or -1
int cratac_set_column_map(cratac_ctx* h, int64_t n_cols, const int32_t* d_col_to_dense) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (n_cols < 0) { c->global_cols = false; c->n_global_cols = 0; return CRATAC_OK; }
    CRATAC_TRY(c->d_global_c2d.reserve(4 * (size_t)std::max<int64_t>(n_cols, 1)));
    if (n_cols > 0) CRATAC_CUDA(cudaMemcpyAsync(c->d_global_c2d.p, d_col_to_dense, 4 * (size_t)n_cols, cudaMemcpyDeviceToDevice, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    c->global_cols = true;
    c->n_global_cols = n_cols;
    return CRATAC_OK;
}

// this shard's peaks are rows [row_base, row_base + n_peaks) of a matrix with total_rows rows
int cratac_set_row_offset(cratac_ctx* h, int64_t row_base, int64_t total_rows) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    CRATAC_TRY(fetch_peak_count(c));
    c->row_range = total_rows;
    return offset_peak_rows(c, row_base);
}

// rank 0: column blocks of all shards (gathered into contiguous device buffers) -> the final CSC of this context
int cratac_merge_rank_csc(cratac_ctx* h, int world, int64_t n_cols, const int64_t* d_indptr_all, const int64_t* d_rank_off,
                          const int64_t* d_indices_all, const int32_t* d_data_all, int64_t total_nnz) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    return merge_rank_csc(c, world, n_cols, d_indptr_all, d_rank_off, d_indices_all, d_data_all, total_nnz);
}

// CPython-2.7 list(set) order from the string hashes alone (distinct keys, insertion order)
int cratac_py2_set_order_hashes(int64_t n, const int64_t* hashes, int32_t* order_out) {
    std::vector<int32_t> order;
    py2_set_order_hashes(reinterpret_cast<const long long*>(hashes), n, order);
    if ((int64_t)order.size() != n) { set_error("cratac_py2_set_order_hashes: internal error"); return CRATAC_ERR_INTERNAL; }
    for (int64_t i = 0; i < n; i++) order_out[i] = order[i];
    return CRATAC_OK;
}

// the same order computed on the device (device pointers; returns when the order is complete)
int cratac_py2_set_order_hashes_device(cratac_ctx* h, int64_t n, const int64_t* d_hashes, int32_t* d_order_out, uint64_t stream) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    cudaStream_t saved = c->stream;
    if (stream) c->stream = (cudaStream_t)(uintptr_t)stream;   // e.g. a side stream while the fit occupies the context's own
    const int rc = py2_set_order_device(c, n, d_hashes, d_order_out);
    c->stream = saved;
    return rc;
}

// ------------------------------------------------------------------------------------------ host helpers
int cratac_py2_set_order(int64_t n, const char* const* keys, int64_t* order_out) {
    std::vector<std::string> k((size_t)n);
    for (int64_t i = 0; i < n; i++) k[i] = keys[i];
    std::vector<int64_t> order;
    py2_set_order(k, order);
    if ((int64_t)order.size() != n) { set_error("cratac_py2_set_order: keys are not distinct"); return CRATAC_ERR_ARG; }
    for (int64_t i = 0; i < n; i++) order_out[i] = order[i];
    return CRATAC_OK;
}

int cratac_encode_text_device(cratac_ctx* h, int64_t n, const int32_t* d_chrom, const int32_t* d_start, const int32_t* d_end,
                              const uint32_t* d_bc_seq, const int32_t* d_dup, char* d_out, int64_t cap, int64_t* nbytes) {
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    return encode_text(c, n, d_chrom, d_start, d_end, d_bc_seq, d_dup, d_out, cap, nbytes);
}

// read-only mapping of an input file: no copy, no zero-filled staging buffer, pages come in as the worker threads touch them
struct MappedFile {
    const unsigned char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open_path(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        n = (size_t)st.st_size;
        if (n == 0) return true;
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        p = static_cast<const unsigned char*>(m);
        madvise(m, n, MADV_WILLNEED);
        return true;
    }
    ~MappedFile() {
        if (p) munmap(const_cast<unsigned char*>(p), n);
        if (fd >= 0) ::close(fd);
    }
};

// block decoder statistics / selection (process-wide): blocks decoded by inflate.cpp, blocks handed to zlib
static std::atomic<long long> g_inflate_fast(0), g_inflate_zlib(0);
static std::atomic<int> g_inflate_mode(1);          // 1: own decoder first, zlib when it declines; 0: zlib only


// Header of the BGZF block at buf[0..avail): total block size (BSIZE + 1 of the 'BC' extra subfield), extra-field length
// and ISIZE.  Every bound a damaged file could break is checked here: the block must hold its own header, extra field and
// the 8-byte trailer, lie inside the buffer, and inflate to at most 64 KiB (the BGZF limit).  0: not a BGZF block / damaged.
static bool bgzf_block_header(const unsigned char* b, size_t avail, size_t* bsize_out, unsigned* xlen_out, uint32_t* isize_out) {
    if (avail < 18 + 8) return false;
    if (!(b[0] == 31 && b[1] == 139 && b[2] == 8 && (b[3] & 4))) return false;
    const unsigned xlen = b[10] | (b[11] << 8);
    const size_t xend = 12 + (size_t)xlen;
    if (xend + 8 > avail) return false;
    long bsize = -1;
    for (size_t q = 12; q + 4 <= xend;) {
        const unsigned slen = b[q + 2] | (b[q + 3] << 8);
        if (q + 4 + slen > xend) return false;
        if (b[q] == 'B' && b[q + 1] == 'C' && slen == 2) bsize = (b[q + 4] | (b[q + 5] << 8)) + 1;
        q += 4 + slen;
    }
    if (bsize < 0 || (size_t)bsize < xend + 8 || (size_t)bsize > avail) return false;
    uint32_t isize;
    memcpy(&isize, &b[bsize - 4], 4);
    if (isize > 65536u) return false;
    *bsize_out = (size_t)bsize; *xlen_out = xlen; *isize_out = isize;
    return true;
}

// one BGZF block: raw DEFLATE `in` -> exactly `isize` bytes at `dst`; the CRC-32 of the trailer (in[csize..csize+4)) is checked like `gzip -d` does
static bool inflate_block(z_stream* zs, uint32_t* scratch, const unsigned char* in, size_t csize, unsigned char* dst, uint32_t isize) {
    if (g_inflate_mode.load(std::memory_order_relaxed) && inflate_raw(in, csize, dst, isize, scratch)) {
        g_inflate_fast.fetch_add(1, std::memory_order_relaxed);
        return true;
    }
    g_inflate_zlib.fetch_add(1, std::memory_order_relaxed);
    if (isize == 0 && csize == 0) return true;
    inflateReset(zs);
    zs->next_in = const_cast<unsigned char*>(in); zs->avail_in = (uInt)csize;
    zs->next_out = dst; zs->avail_out = isize;
    const int rc = inflate(zs, Z_FINISH);
    return rc == Z_STREAM_END && zs->avail_out == 0;
}
static bool inflate_block_checked(z_stream* zs, uint32_t* scratch, const unsigned char* in, size_t csize, unsigned char* dst, uint32_t isize) {
    if (!inflate_block(zs, scratch, in, csize, dst, isize)) return false;
    uint32_t want;
    memcpy(&want, in + csize, 4);                       // trailer: CRC32, ISIZE
    return (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, isize) == want;
}

int cratac_inflate_stats(int64_t* fast_blocks, int64_t* zlib_blocks, int reset) {
    if (fast_blocks) *fast_blocks = g_inflate_fast.load();
    if (zlib_blocks) *zlib_blocks = g_inflate_zlib.load();
    if (reset) { g_inflate_fast = 0; g_inflate_zlib = 0; }
    return CRATAC_OK;
}

int cratac_inflate_mode(int mode) {
    g_inflate_mode = mode ? 1 : 0;
    return CRATAC_OK;
}

// BGZF (or plain multi-member gzip) inflate.  BGZF blocks carry their compressed size in the 'BC' extra
// field and their inflated size in the trailer, so the blocks inflate independently on `threads` threads.
int cratac_inflate_file(const char* path, int threads, char* out, int64_t cap, int64_t* nbytes) {
    MappedFile mf;
    if (!mf.open_path(path)) { set_error("cannot open %s", path); return CRATAC_ERR_IO; }
    const long fsz = (long)mf.n;
    const unsigned char* buf = mf.p;
    struct Blk { size_t off, csize; int64_t out_off; uint32_t isize; };
    std::vector<Blk> blocks;
    size_t p = 0;
    int64_t total = 0;
    bool bgzf = true;
    while (p < (size_t)fsz) {
        size_t bsize; unsigned xlen; uint32_t isize;
        if (!bgzf_block_header(&buf[p], (size_t)fsz - p, &bsize, &xlen, &isize)) { bgzf = false; break; }
        blocks.push_back(Blk{p + 12 + xlen, bsize - 12 - xlen - 8, total, isize});
        total += isize;
        p += bsize;
    }
    if (bgzf && p == (size_t)fsz) {
        *nbytes = total;
        if (cap == 0 || !out) return CRATAC_OK;
        if (cap < total) { set_error("inflate buffer too small: need %lld bytes", (long long)total); return CRATAC_ERR_ARG; }
        if (threads < 1) threads = 1;
        std::atomic<size_t> next(0);
        std::atomic<int> failed(0);
        auto work = [&]() {
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) { failed = 1; return; }
            std::vector<uint32_t> scratch(inflate_scratch_words());
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= blocks.size()) break;
                const Blk& k = blocks[i];
                if (!inflate_block_checked(&zs, scratch.data(), &buf[k.off], k.csize, reinterpret_cast<unsigned char*>(out) + k.out_off, k.isize)) { failed = 1; break; }
            }
            inflateEnd(&zs);
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; t++) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
        if (failed) { set_error("corrupt BGZF block in %s", path); return CRATAC_ERR_IO; }
        return CRATAC_OK;
    }
    // plain gzip (possibly multi-member) or uncompressed text: sequential
    if (fsz >= 2 && buf[0] == 31 && buf[1] == 139) {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, 15 + 16) != Z_OK) { set_error("zlib init failed"); return CRATAC_ERR_IO; }
        std::vector<unsigned char> scratch(1 << 20);
        zs.next_in = const_cast<unsigned char*>(buf);
        zs.avail_in = 0;
        size_t fed = 0;                                   // input is handed over 1 GiB at a time (avail_in is 32 bits)
        int64_t produced = 0;
        for (;;) {
            if (zs.avail_in == 0 && fed < (size_t)fsz) {
                const size_t take = std::min<size_t>((size_t)fsz - fed, (size_t)1 << 30);
                zs.next_in = const_cast<unsigned char*>(buf) + fed;
                zs.avail_in = (uInt)take;
                fed += take;
            }
            bool counting = (cap == 0 || !out);
            zs.next_out = counting ? scratch.data() : reinterpret_cast<unsigned char*>(out) + produced;
            int64_t room = counting ? (int64_t)scratch.size() : cap - produced;
            zs.avail_out = (uInt)std::min<int64_t>(room, 1 << 30);
            uInt before = zs.avail_out;
            int rc = inflate(&zs, Z_NO_FLUSH);
            produced += before - zs.avail_out;
            if (rc == Z_STREAM_END) {
                if (zs.avail_in == 0 && fed >= (size_t)fsz) break;
                inflateReset(&zs);
                continue;
            }
            if (rc == Z_BUF_ERROR && zs.avail_in == 0 && fed < (size_t)fsz) continue;      // needs the next slice of input
            if (rc != Z_OK) { inflateEnd(&zs); set_error("corrupt gzip stream in %s", path); return CRATAC_ERR_IO; }
            if (!counting && produced >= cap && zs.avail_in > 0) { inflateEnd(&zs); set_error("inflate buffer too small"); return CRATAC_ERR_ARG; }
        }
        inflateEnd(&zs);
        *nbytes = produced;
        return CRATAC_OK;
    }
    *nbytes = fsz;
    if (cap == 0 || !out) return CRATAC_OK;
    if (cap < fsz) { set_error("buffer too small"); return CRATAC_ERR_ARG; }
    memcpy(out, buf, (size_t)fsz);
    return CRATAC_OK;
}

// Inflate the BGZF virtual-offset range [voff_begin, voff_end) of a file (voff = compressed offset of a block << 16 |
// offset inside the inflated block: the unit of a tabix index, lib/python/tools/io.py:86 reads contigs through
// pysam.TabixFile the same way).  Only the bytes of the blocks involved are read; they inflate on `threads` threads.
int cratac_inflate_range(const char* path, int threads, uint64_t voff_begin, uint64_t voff_end, char* out, int64_t cap, int64_t* nbytes) {
    if (!path || !nbytes) { set_error("cratac_inflate_range: null argument"); return CRATAC_ERR_ARG; }
    *nbytes = 0;
    if (voff_end <= voff_begin) return CRATAC_OK;
    const uint64_t c_begin = voff_begin >> 16, c_end = voff_end >> 16;
    const uint32_t u_begin = (uint32_t)(voff_begin & 0xFFFF), u_end = (uint32_t)(voff_end & 0xFFFF);
    MappedFile mf;
    if (!mf.open_path(path)) { set_error("cannot open %s", path); return CRATAC_ERR_IO; }
    const uint64_t fsz = (uint64_t)mf.n;
    if (c_begin >= fsz) { set_error("virtual offset beyond the end of %s", path); return CRATAC_ERR_ARG; }
    // the last block is needed only when some of its bytes are inside the range
    const uint64_t read_end = std::min<uint64_t>(fsz, u_end > 0 ? c_end + 65536 : c_end);
    const unsigned char* buf = mf.p + c_begin;                    // the blocks of the range, read through the mapping
    const size_t buf_size = (size_t)(read_end - c_begin);
    struct Blk { size_t off, csize; int64_t out_off; uint32_t isize, lo, hi; };   // [lo, hi) of the inflated block is kept
    std::vector<Blk> blocks;
    size_t p = 0;
    int64_t total = 0;
    while (c_begin + p < c_end || (c_begin + p == c_end && u_end > 0)) {
        size_t bsize; unsigned xlen; uint32_t isize;
        if (p >= buf_size || !bgzf_block_header(&buf[p], buf_size - p, &bsize, &xlen, &isize)) {
            set_error("truncated or damaged BGZF block at offset %llu of %s", (unsigned long long)(c_begin + p), path);
            return CRATAC_ERR_IO;
        }
        const uint64_t at = c_begin + p;
        uint32_t lo = at == c_begin ? u_begin : 0u, hi = at == c_end ? u_end : isize;
        if (hi > isize || lo > hi) { set_error("virtual offset beyond its block in %s", path); return CRATAC_ERR_ARG; }
        blocks.push_back(Blk{p + 12 + xlen, bsize - 12 - xlen - 8, total, isize, lo, hi});
        total += hi - lo;
        p += bsize;
    }
    *nbytes = total;
    if (cap == 0 || !out) return CRATAC_OK;
    if (cap < total) { set_error("inflate buffer too small: need %lld bytes", (long long)total); return CRATAC_ERR_ARG; }
    if (threads < 1) threads = 1;
    std::atomic<size_t> next(0);
    std::atomic<int> failed(0);
    auto work = [&]() {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) { failed = 1; return; }
        std::vector<unsigned char> part;
        std::vector<uint32_t> scratch(inflate_scratch_words());
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= blocks.size()) break;
            const Blk& k = blocks[i];
            const bool whole = k.lo == 0 && k.hi == k.isize;
            if (!whole) part.resize(k.isize);
            unsigned char* dst = whole ? reinterpret_cast<unsigned char*>(out) + k.out_off : part.data();
            if (!inflate_block_checked(&zs, scratch.data(), &buf[k.off], k.csize, dst, k.isize)) { failed = 1; break; }
            if (!whole && k.hi > k.lo) memcpy(out + k.out_off, part.data() + k.lo, k.hi - k.lo);
        }
        inflateEnd(&zs);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (failed) { set_error("corrupt BGZF block in %s", path); return CRATAC_ERR_IO; }
    return CRATAC_OK;
}


// ---- fragments.tsv.gz straight into the decoder -------------------------------------------------------------------
// BgzfRing: the BGZF blocks of a file are inflated on host threads into a ring of pinned chunks, in file order; the
// streamed decode (decode.cu) copies a chunk to the device and parses it as soon as all of its blocks are there, while the
// threads are already filling the next slots.  Inflate of chunk i+1, the PCIe copy of chunk i and the parse of chunk i-1
// overlap; a slot is refilled once the copy that read it has completed.  Replaces the `gzip -c -d` pipe and the per-line
// split of lib/python/tools/io.py:197-217,75-92 for the whole file.
namespace {
struct BgzfBlk { size_t off, csize; int64_t out_off; uint32_t isize; };

struct BgzfRing : cratac::HostTextSource {
    const unsigned char* file = nullptr;
    std::vector<BgzfBlk> blocks;
    int64_t total = 0, chunk = 0;
    int slots = 0;
    char* ring = nullptr;                       // pinned, slots * chunk bytes
    char last = '\n';
    std::vector<std::thread> pool;
    std::mutex mu;
    std::condition_variable cv_done, cv_room;
    std::vector<unsigned char> done;            // per block
    size_t next_block = 0, done_prefix = 0;     // blocks [0, done_prefix) are inflated
    int64_t released = 0;                       // text bytes below this offset have left the ring (their slot may be refilled)
    bool failed = false, stop = false;
    std::vector<cudaEvent_t> ev;                // per slot: the copy that read it
    std::vector<int64_t> ev_end;                // text offset that copy ends at (0: none pending)

    char* at(int64_t o) const { return ring + ((o / chunk) % slots) * chunk + (o % chunk); }

    void worker() {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) { std::lock_guard<std::mutex> g(mu); failed = true; cv_done.notify_all(); return; }
        std::vector<uint32_t> scratch(cratac::inflate_scratch_words());
        std::vector<unsigned char> tmp(65536);
        for (;;) {
            size_t i;
            {
                std::unique_lock<std::mutex> g(mu);
                if (stop || failed || next_block >= blocks.size()) break;
                i = next_block;
                const BgzfBlk& k = blocks[i];
                // room: the block's bytes must fit into slots whose previous content has been copied out
                cv_room.wait(g, [&] { return stop || failed || k.out_off + (int64_t)k.isize <= released + (int64_t)slots * chunk; });
                if (stop || failed) break;
                if (i != next_block) continue;                 // another thread took it while this one waited
                next_block++;
            }
            const BgzfBlk& k = blocks[i];
            bool ok;
            const bool straddles = k.isize > 0 && (k.out_off / chunk) != ((k.out_off + (int64_t)k.isize - 1) / chunk);
            if (!straddles) {
                ok = inflate_block_checked(&zs, scratch.data(), file + k.off, k.csize, reinterpret_cast<unsigned char*>(at(k.out_off)), k.isize);
            } else {
                ok = inflate_block_checked(&zs, scratch.data(), file + k.off, k.csize, tmp.data(), k.isize);
                if (ok) {
                    const int64_t first = chunk - (k.out_off % chunk);
                    memcpy(at(k.out_off), tmp.data(), (size_t)first);
                    memcpy(at(k.out_off + first), tmp.data() + first, (size_t)(k.isize - first));
                }
            }
            {
                std::lock_guard<std::mutex> g(mu);
                if (!ok) failed = true;
                done[i] = 1;
                while (done_prefix < blocks.size() && done[done_prefix]) done_prefix++;
            }
            cv_done.notify_all();
        }
        inflateEnd(&zs);
    }

    void poll_copies(bool block_on_oldest) {
        // completed copies free their slots (in order)
        for (;;) {
            int best = -1;
            for (int s = 0; s < slots; s++) if (ev_end[s] > 0 && (best < 0 || ev_end[s] < ev_end[best])) best = s;
            if (best < 0) return;
            cudaError_t q = block_on_oldest ? cudaEventSynchronize(ev[best]) : cudaEventQuery(ev[best]);
            block_on_oldest = false;
            if (q != cudaSuccess) { if (q != cudaErrorNotReady) { std::lock_guard<std::mutex> g(mu); failed = true; } return; }
            {
                std::lock_guard<std::mutex> g(mu);
                released = std::max(released, ev_end[best]);
            }
            ev_end[best] = 0;
            cv_room.notify_all();
        }
    }

    char last_byte() override { return last; }

    const char* acquire(int64_t b0, int64_t b1) override {
        for (;;) {
            poll_copies(false);
            std::unique_lock<std::mutex> g(mu);
            const int64_t ready = done_prefix < blocks.size() ? blocks[done_prefix].out_off : total;
            if (failed) return nullptr;
            if (ready >= b1) return at(b0);
            // the writers may be waiting for room that only a finished copy gives: wake up regularly and look
            cv_done.wait_for(g, std::chrono::microseconds(200));
        }
    }

    int queued(int64_t b0, int64_t b1, cudaStream_t copy_stream) override {
        const int s = (int)((b0 / chunk) % slots);
        if (cudaEventRecord(ev[s], copy_stream) != cudaSuccess) return CRATAC_ERR_CUDA;
        ev_end[s] = b1;
        return CRATAC_OK;
    }

    void shutdown() {
        { std::lock_guard<std::mutex> g(mu); stop = true; }
        cv_room.notify_all(); cv_done.notify_all();
        for (auto& t : pool) t.join();
        pool.clear();
        for (auto e : ev) cudaEventDestroy(e);
        ev.clear();
    }
    ~BgzfRing() override { shutdown(); }
};
}  // namespace

int cratac_decode_file(cratac_ctx* h, const char* path, int threads) {
    if (!h || !path) { set_error("cratac_decode_file: null argument"); return CRATAC_ERR_ARG; }
    Ctx* c = &h->c;
    CRATAC_CUDA(cudaSetDevice(c->device));
    if (threads < 1) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    MappedFile mf;
    if (!mf.open_path(path)) { set_error("cannot open %s", path); return CRATAC_ERR_IO; }
    BgzfRing src;
    src.file = mf.p;
    size_t p = 0;
    bool bgzf = mf.n > 0;
    while (p < mf.n) {
        size_t bsize; unsigned xlen; uint32_t isize;
        if (!bgzf_block_header(mf.p + p, mf.n - p, &bsize, &xlen, &isize)) { bgzf = false; break; }
        src.blocks.push_back(BgzfBlk{p + 12 + xlen, bsize - 12 - xlen - 8, src.total, isize});
        src.total += isize;
        p += bsize;
    }
    if (!bgzf) {
        // plain gzip / text (or a damaged BGZF header, which the sequential reader reports): no block index to run in
        // parallel on -- inflate the whole file, then the host-text path
        int64_t n = 0;
        CRATAC_TRY(cratac_inflate_file(path, threads, nullptr, 0, &n));
        std::vector<char> text((size_t)std::max<int64_t>(n, 1));
        CRATAC_TRY(cratac_inflate_file(path, threads, text.data(), n, &n));
        return cratac_decode_host(h, text.data(), n);
    }
    c->bc_is_slot = true;
    const int64_t nbytes = src.total;
    // the last byte of the text decides whether a terminator is implied: inflate the last non-empty block first
    for (size_t i = src.blocks.size(); i-- > 0;) {
        const BgzfBlk& k = src.blocks[i];
        if (k.isize == 0) continue;
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) { set_error("zlib init failed"); return CRATAC_ERR_IO; }
        std::vector<uint32_t> scratch(inflate_scratch_words());
        std::vector<unsigned char> tmp(k.isize);
        const bool ok = inflate_block_checked(&zs, scratch.data(), mf.p + k.off, k.csize, tmp.data(), k.isize);
        inflateEnd(&zs);
        if (!ok) { set_error("corrupt BGZF block in %s", path); return CRATAC_ERR_IO; }
        src.last = (char)tmp[k.isize - 1];
        break;
    }
    src.chunk = decode_stream_chunk_bytes(c, std::max<int64_t>(c->file_chunk_bytes, 3 * 65536));   // a block (<= 64 KiB) crosses at most one slot border
    src.slots = 4;
    const size_t ring_bytes = (size_t)src.slots * (size_t)src.chunk;
    if (c->h_ring_cap < ring_bytes) {
        if (c->h_ring) cudaFreeHost(c->h_ring);
        c->h_ring = nullptr; c->h_ring_cap = 0;
        if (cudaMallocHost(&c->h_ring, ring_bytes) != cudaSuccess) { set_error("cudaMallocHost failed for the inflate ring (%zu bytes)", ring_bytes); return CRATAC_ERR_CUDA; }
        c->h_ring_cap = ring_bytes;
    }
    src.ring = c->h_ring;
    src.done.assign(src.blocks.size(), 0);
    src.ev.resize(src.slots);
    src.ev_end.assign(src.slots, 0);
    for (int s = 0; s < src.slots; s++) CRATAC_CUDA(cudaEventCreateWithFlags(&src.ev[s], cudaEventDisableTiming));
    for (int t = 0; t < threads; t++) src.pool.emplace_back([&src] { src.worker(); });
    int rc = nbytes > 0 ? decode_text_streamed_from(c, &src, nbytes, src.chunk) : decode_text(c, c->d_text_own.as<char>(), 0);
    const bool bad = src.failed;
    src.shutdown();
    if (bad) { set_error("corrupt BGZF block in %s", path); return CRATAC_ERR_IO; }
    if (rc == CRATAC_ERR_CAPACITY) rc = decode_text(c, c->d_text_own.as<char>(), nbytes);   // the text is on the device: exact path with a larger table
    return rc;
}

int cratac_run_file(cratac_ctx* h, const char* path, int threads, double odds_ratio, cratac_run_result* result) {
    if (!h || !path) { set_error("cratac_run_file: null argument"); return CRATAC_ERR_ARG; }
    h->c.host_reset();
    CRATAC_TRY(cratac_decode_file(h, path, threads));
    h->c.host_mark("decode_file_total");
    return run_tail(&h->c, odds_ratio, result);
}

// HDF5 chunk filters on all host cores: chunk k = elements [k * chunk_elems, (k + 1) * chunk_elems) of `data` (the last one
// padded with zeros to full size), byte-shuffled (item_size) and deflated (zlib stream) into out + k * stride;
// sizes[k] = its filtered length.  What h5py does one chunk at a time for compression='gzip', shuffle=True
// (cellranger/matrix.py:441-447); cratac/h5.py lays the chunks out and indexes them.
int cratac_h5_filter_chunks(const void* data, int64_t n_elems, int item_size, int64_t chunk_elems, int level, int threads, unsigned char* out,
                            int64_t stride, int64_t* sizes) {
    if (n_elems < 0 || item_size < 1 || chunk_elems < 1 || (n_elems > 0 && (!data || !out || !sizes))) { set_error("cratac_h5_filter_chunks: bad argument"); return CRATAC_ERR_ARG; }
    const int64_t chunk_bytes = chunk_elems * item_size;
    if (stride < (int64_t)compressBound((uLong)chunk_bytes)) { set_error("cratac_h5_filter_chunks: stride below compressBound"); return CRATAC_ERR_ARG; }
    const int64_t n_chunks = (n_elems + chunk_elems - 1) / chunk_elems;
    if (threads < 1) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::atomic<int64_t> next(0);
    std::atomic<int> bad(0);
    auto work = [&]() {
        std::vector<unsigned char> shuf((size_t)chunk_bytes);
        for (;;) {
            const int64_t k = next.fetch_add(1);
            if (k >= n_chunks) break;
            const int64_t e0 = k * chunk_elems, ne = std::min(chunk_elems, n_elems - e0);
            const unsigned char* src = static_cast<const unsigned char*>(data) + e0 * item_size;
            if (ne < chunk_elems) memset(shuf.data(), 0, (size_t)chunk_bytes);
            for (int j = 0; j < item_size; j++) {
                unsigned char* dst = shuf.data() + (size_t)j * (size_t)chunk_elems;
                for (int64_t i = 0; i < ne; i++) dst[i] = src[i * item_size + j];
            }
            uLongf out_len = (uLongf)stride;
            if (compress2(out + k * stride, &out_len, shuf.data(), (uLong)chunk_bytes, level) != Z_OK) { bad = 1; break; }
            sizes[k] = (int64_t)out_len;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (bad) { set_error("deflate failed"); return CRATAC_ERR_IO; }
    return CRATAC_OK;
}

// BGZF writer: `nbytes` of text -> blocks of at most 0xff00 input bytes, each a gzip member with the 'BC' extra field, closed
// by the empty EOF block; compressed on `threads` threads, written in order.  (The container mark_duplicates writes
// fragments.tsv.gz in, mark_duplicates/__init__.py:399; here it builds the synthetic inputs of bench.py and the tests.)
int cratac_write_bgzf(const char* path, const char* data, int64_t nbytes, int level, int threads) {
    if (!path || nbytes < 0 || (nbytes > 0 && !data)) { set_error("cratac_write_bgzf: bad argument"); return CRATAC_ERR_ARG; }
    FILE* f = fopen(path, "wb");
    if (!f) { set_error("cannot open %s for writing", path); return CRATAC_ERR_IO; }
    if (threads < 1) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const int64_t BLOCK_IN = 0xff00;
    const int64_t n_blocks = (nbytes + BLOCK_IN - 1) / BLOCK_IN;
    const int64_t per_task = 64;                                        // blocks per task
    const int64_t n_tasks = (n_blocks + per_task - 1) / per_task;
    bool failed = false;
    for (int64_t wave = 0; wave < n_tasks && !failed; wave += 4 * (int64_t)threads) {
        const int64_t in_wave = std::min<int64_t>(4 * (int64_t)threads, n_tasks - wave);
        std::vector<std::vector<unsigned char>> outs((size_t)in_wave);
        std::atomic<int64_t> next(0);
        std::atomic<int> bad(0);
        auto work = [&]() {
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad = 1; return; }
            for (;;) {
                const int64_t t = next.fetch_add(1);
                if (t >= in_wave) break;
                std::vector<unsigned char>& out = outs[(size_t)t];
                out.reserve((size_t)(per_task * 66000));
                for (int64_t b = (wave + t) * per_task; b < std::min(n_blocks, (wave + t + 1) * per_task); b++) {
                    const int64_t o = b * BLOCK_IN, len = std::min(BLOCK_IN, nbytes - o);
                    const size_t at = out.size();
                    out.resize(at + 18 + 65535 + 8);
                    deflateReset(&zs);
                    zs.next_in = reinterpret_cast<unsigned char*>(const_cast<char*>(data + o)); zs.avail_in = (uInt)len;
                    zs.next_out = out.data() + at + 18; zs.avail_out = 65535 - 26;
                    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { bad = 1; break; }
                    const size_t clen = (65535 - 26) - zs.avail_out;
                    const unsigned bsize = (unsigned)(clen + 25);       // BSIZE = block size - 1
                    const unsigned char head[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0, (unsigned char)(bsize & 255), (unsigned char)(bsize >> 8)};
                    memcpy(out.data() + at, head, 18);
                    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), reinterpret_cast<const unsigned char*>(data + o), (uInt)len);
                    const uint32_t isz = (uint32_t)len;
                    memcpy(out.data() + at + 18 + clen, &crc, 4);
                    memcpy(out.data() + at + 18 + clen + 4, &isz, 4);
                    out.resize(at + 18 + clen + 8);
                }
                if (bad) break;
            }
            deflateEnd(&zs);
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; t++) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
        if (bad) { failed = true; break; }
        for (auto& o : outs) if (!o.empty() && fwrite(o.data(), 1, o.size(), f) != o.size()) { failed = true; break; }
    }
    static const unsigned char eof_block[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (!failed && fwrite(eof_block, 1, 28, f) != 28) failed = true;
    if (fclose(f) != 0) failed = true;
    if (failed) { set_error("writing BGZF to %s failed", path); return CRATAC_ERR_IO; }
    return CRATAC_OK;
}

// Matrix Market body: "row col value\n" (1-based) for every stored entry, column by column -- the per-entry Python loop
// of CountMatrix.save_mex (cellranger/matrix.py:812-814).  `header` (the three header lines, already formatted) is written
// first.  Columns are cut into chunks of about 1 M entries; `threads` chunks are formatted at a time and written in order.
int cratac_write_mtx(const char* path, const char* header, int64_t n_cols, const int64_t* indptr, const int64_t* indices,
                     const int32_t* data, int threads) {
    if (!path || !indptr || n_cols < 0) { set_error("cratac_write_mtx: bad argument"); return CRATAC_ERR_ARG; }
    FILE* f = fopen(path, "wb");
    if (!f) { set_error("cannot open %s for writing", path); return CRATAC_ERR_IO; }
    if (header && fputs(header, f) < 0) { fclose(f); set_error("write failed on %s", path); return CRATAC_ERR_IO; }
    if (threads < 1) threads = 1;
    // chunk boundaries in columns
    const int64_t target = 1 << 20;
    std::vector<int64_t> cuts(1, 0);
    for (int64_t j = 0; j < n_cols;) {
        const int64_t want = indptr[j] + target;
        int64_t k = std::upper_bound(indptr + j + 1, indptr + n_cols + 1, want) - indptr;   // first column whose END exceeds want
        if (k > n_cols) k = n_cols;
        if (k <= j) k = j + 1;
        cuts.push_back(k);
        j = k;
    }
    const size_t n_chunks = cuts.size() - 1;
    auto put_uint = [](char* p, uint64_t v) -> char* {
        char tmp[24];
        int n = 0;
        do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (n) *p++ = tmp[--n];
        return p;
    };
    bool failed = false;
    for (size_t wave = 0; wave < n_chunks && !failed; wave += (size_t)threads) {
        const size_t in_wave = std::min<size_t>((size_t)threads, n_chunks - wave);
        std::vector<std::unique_ptr<char[]>> bufs(in_wave);
        std::vector<size_t> used(in_wave, 0);
        auto work = [&](size_t w) {
            const int64_t c0 = cuts[wave + w], c1 = cuts[wave + w + 1];
            const int64_t n_ent = indptr[c1] - indptr[c0];
            bufs[w].reset(new char[(size_t)n_ent * 44 + 16]);       // 20 + 1 + 10 + 1 + 11 + 1 bytes at most per line; not zero-filled
            char* p = bufs[w].get();
            for (int64_t j = c0; j < c1; j++) {
                for (int64_t k = indptr[j]; k < indptr[j + 1]; k++) {
                    p = put_uint(p, (uint64_t)(indices[k] + 1));
                    *p++ = ' ';
                    p = put_uint(p, (uint64_t)(j + 1));
                    *p++ = ' ';
                    int64_t v = data[k];
                    if (v < 0) { *p++ = '-'; v = -v; }
                    p = put_uint(p, (uint64_t)v);
                    *p++ = '\n';
                }
            }
            used[w] = (size_t)(p - bufs[w].get());
        };
        std::vector<std::thread> pool;
        for (size_t w = 1; w < in_wave; w++) pool.emplace_back(work, w);
        work(0);
        for (auto& t : pool) t.join();
        for (size_t w = 0; w < in_wave; w++)
            if (used[w] && fwrite(bufs[w].get(), 1, used[w], f) != used[w]) failed = true;
    }
    if (fclose(f) != 0) failed = true;
    if (failed) { set_error("write failed on %s", path); return CRATAC_ERR_IO; }
    return CRATAC_OK;
}

// ---- the cut_sites bedgraph, the file between COUNT_CUT_SITES and DETECT_PEAKS (hundreds of millions of rows genome-wide)
// rows "contig\tstart\tend\tcount\n" of one contig, appended to `path` (count_cut_sites/__init__.py:38-48 writes them one
// Python format() at a time); formatted in chunks of 1 M rows on `threads` threads, written in order.
int cratac_write_bedgraph(const char* path, int append, const char* contig, int64_t n, const int64_t* start, const int64_t* end,
                          const int64_t* count, int threads) {
    if (!path || !contig || n < 0) { set_error("cratac_write_bedgraph: bad argument"); return CRATAC_ERR_ARG; }
    FILE* f = fopen(path, append ? "ab" : "wb");
    if (!f) { set_error("cannot open %s for writing", path); return CRATAC_ERR_IO; }
    if (threads < 1) threads = 1;
    const size_t name_len = strlen(contig);
    const int64_t chunk = 1 << 20;
    const int64_t n_chunks = (n + chunk - 1) / chunk;
    auto put_int = [](char* p, int64_t x) -> char* {
        char tmp[24];
        int k = 0;
        uint64_t v = x < 0 ? (uint64_t)(-x) : (uint64_t)x;
        if (x < 0) *p++ = '-';
        do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (k) *p++ = tmp[--k];
        return p;
    };
    bool failed = false;
    for (int64_t wave = 0; wave < n_chunks && !failed; wave += threads) {
        const int64_t in_wave = std::min<int64_t>(threads, n_chunks - wave);
        std::vector<std::unique_ptr<char[]>> bufs((size_t)in_wave);
        std::vector<size_t> used((size_t)in_wave, 0);
        auto work = [&](int64_t w) {
            const int64_t r0 = (wave + w) * chunk, r1 = std::min<int64_t>(n, r0 + chunk);
            bufs[(size_t)w].reset(new char[(size_t)(r1 - r0) * (name_len + 3 * 21 + 4) + 16]);
            char* p = bufs[(size_t)w].get();
            for (int64_t r = r0; r < r1; r++) {
                memcpy(p, contig, name_len); p += name_len;
                *p++ = '\t'; p = put_int(p, start[r]);
                *p++ = '\t'; p = put_int(p, end[r]);
                *p++ = '\t'; p = put_int(p, count[r]);
                *p++ = '\n';
            }
            used[(size_t)w] = (size_t)(p - bufs[(size_t)w].get());
        };
        std::vector<std::thread> pool;
        for (int64_t w = 1; w < in_wave; w++) pool.emplace_back(work, w);
        work(0);
        for (auto& t : pool) t.join();
        for (int64_t w = 0; w < in_wave; w++)
            if (used[(size_t)w] && fwrite(bufs[(size_t)w].get(), 1, used[(size_t)w], f) != used[(size_t)w]) failed = true;
    }
    if (fclose(f) != 0) failed = true;
    if (failed) { set_error("write failed on %s", path); return CRATAC_ERR_IO; }
    return CRATAC_OK;
}

// Rows of a bedgraph whose count is >= threshold (all rows when has_threshold is 0): the `for track in f` loop of
// lib/python/utils.py:105-115 + the integer test of detect_peaks/__init__.py:49, on `threads` threads over a read-only
// mapping of the file.  The arrays are allocated here (free them with cratac_free_host); chrom = index into contig_names.
int cratac_read_bedgraph(const char* path, int n_contigs, const char* const* contig_names, int has_threshold, double threshold,
                         int threads, int32_t** chrom_out, int64_t** start_out, int64_t** end_out, int64_t* n_out) {
    if (!path || !chrom_out || !start_out || !end_out || !n_out) { set_error("cratac_read_bedgraph: null argument"); return CRATAC_ERR_ARG; }
    *chrom_out = nullptr; *start_out = nullptr; *end_out = nullptr; *n_out = 0;
    MappedFile mf;
    if (!mf.open_path(path)) { set_error("cannot open %s", path); return CRATAC_ERR_IO; }
    const char* base = reinterpret_cast<const char*>(mf.p);
    const size_t size = mf.n;
    if (threads < 1) threads = 1;
    std::unordered_map<std::string, int> ids;
    for (int i = 0; i < n_contigs; i++) ids.emplace(contig_names[i], i);
    // cut the file into ranges that start right after a newline
    const size_t n_ranges = size < ((size_t)1 << 20) ? 1 : (size_t)threads * 4;
    std::vector<size_t> cut(n_ranges + 1, size);
    cut[0] = 0;
    for (size_t r = 1; r < n_ranges; r++) {
        size_t p = size / n_ranges * r;
        while (p < size && base[p] != '\n') p++;
        cut[r] = p < size ? p + 1 : size;
    }
    struct Part { std::vector<int32_t> c; std::vector<int64_t> s, e; long long bad_line = -1; std::string bad_name; };
    std::vector<Part> parts(n_ranges);
    std::atomic<size_t> next(0);
    auto work = [&]() {
        for (;;) {
            const size_t r = next.fetch_add(1);
            if (r >= n_ranges) break;
            Part& out = parts[r];
            const char* p = base + cut[r];
            const char* const stop = base + cut[r + 1];
            const char* last_name = nullptr;
            size_t last_len = 0;
            int last_id = -1;
            while (p < stop) {
                const char* line = p;
                const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(stop - p)));
                const char* eol = nl ? nl : stop;
                p = nl ? nl + 1 : stop;
                if (eol == line) { out.bad_line = (long long)(line - base); break; }      // `a, b, c, d = "".split()` raises as well
                const char* t1 = static_cast<const char*>(memchr(line, '\t', (size_t)(eol - line)));
                const char* t2 = t1 ? static_cast<const char*>(memchr(t1 + 1, '\t', (size_t)(eol - t1 - 1))) : nullptr;
                const char* t3 = t2 ? static_cast<const char*>(memchr(t2 + 1, '\t', (size_t)(eol - t2 - 1))) : nullptr;
                if (!t3 || memchr(t3 + 1, '\t', (size_t)(eol - t3 - 1))) { out.bad_line = (long long)(line - base); break; }
                auto to_int = [](const char* b, const char* e, int64_t* v) -> bool {
                    while (b < e && (*b == ' ' || *b == '\r')) b++;
                    while (e > b && (e[-1] == ' ' || e[-1] == '\r')) e--;
                    bool neg = false;
                    if (b < e && (*b == '-' || *b == '+')) { neg = *b == '-'; b++; }
                    if (b == e || e - b > 18) return false;
                    int64_t x = 0;
                    for (; b < e; b++) { if (*b < '0' || *b > '9') return false; x = x * 10 + (*b - '0'); }
                    *v = neg ? -x : x;
                    return true;
                };
                int64_t s, e, cnt;
                if (!to_int(t1 + 1, t2, &s) || !to_int(t2 + 1, t3, &e) || !to_int(t3 + 1, eol, &cnt)) { out.bad_line = (long long)(line - base); break; }
                if (has_threshold && (double)cnt < threshold) continue;
                const size_t len = (size_t)(t1 - line);
                if (!(last_name && len == last_len && memcmp(last_name, line, len) == 0)) {
                    auto it = ids.find(std::string(line, len));
                    if (it == ids.end()) { out.bad_line = (long long)(line - base); out.bad_name.assign(line, len); break; }
                    last_name = line; last_len = len; last_id = it->second;
                }
                out.c.push_back(last_id); out.s.push_back(s); out.e.push_back(e);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    size_t total = 0;
    for (const Part& q : parts) {
        if (q.bad_line >= 0) {
            if (!q.bad_name.empty()) set_error("bedgraph %s: contig '%s' (byte %lld) is not in the reference", path, q.bad_name.c_str(), q.bad_line);
            else set_error("bedgraph %s: malformed row at byte %lld", path, q.bad_line);
            return CRATAC_ERR_PARSE;
        }
        total += q.c.size();
    }
    int32_t* c = static_cast<int32_t*>(malloc(std::max<size_t>(total, 1) * sizeof(int32_t)));
    int64_t* s = static_cast<int64_t*>(malloc(std::max<size_t>(total, 1) * sizeof(int64_t)));
    int64_t* e = static_cast<int64_t*>(malloc(std::max<size_t>(total, 1) * sizeof(int64_t)));
    if (!c || !s || !e) { free(c); free(s); free(e); set_error("out of memory reading %s", path); return CRATAC_ERR_IO; }
    size_t at = 0;
    for (const Part& q : parts) {
        if (q.c.empty()) continue;
        memcpy(c + at, q.c.data(), q.c.size() * sizeof(int32_t));
        memcpy(s + at, q.s.data(), q.s.size() * sizeof(int64_t));
        memcpy(e + at, q.e.data(), q.e.size() * sizeof(int64_t));
        at += q.c.size();
    }
    *chrom_out = c; *start_out = s; *end_out = e; *n_out = (int64_t)total;
    return CRATAC_OK;
}

void cratac_free_host(void* p) { free(p); }

}  // extern "C"
