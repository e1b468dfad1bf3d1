// Shared declarations for libcratac (sm_100a).  Internal header: the public C ABI is include/cratac.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/cratac.h"

namespace cratac {

// ---------------------------------------------------------------- constants of the path
constexpr int HALF_WINDOW = 200;       // lib/python/constants.py:68  WINDOW_SIZE // 2
constexpr int MERGE_DIST = 500;        // lib/python/constants.py:70  PEAK_MERGE_DISTANCE
constexpr int DEFAULT_THRESHOLD = 15;  // lib/python/analysis/peaks.py:17
constexpr int POS_INDEX_SHIFT = 10;    // one fragment-index entry per 1024 genome bases
constexpr int64_t CONTIG_GAP = 1024;   // spacing between contigs in the global run coordinate (> MERGE_DIST)
constexpr int HIST_CAP = 1 << 20;      // largest representable window count
constexpr int NUM_SMS = 148;

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
#define CRATAC_CUDA(call)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            cratac::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return CRATAC_ERR_CUDA;                                                            \
        }                                                                                      \
    } while (0)
#define CRATAC_TRY(call)                \
    do {                                \
        int _rc = (call);               \
        if (_rc != CRATAC_OK) return _rc; \
    } while (0)

// device-side status word layout (ctx->d_status, int64[16])
enum StatusSlot {
    ST_ERROR = 0,        // first error code raised on device (0 = none)
    ST_ERROR_ARG = 1,    // detail (line number, tile id ...)
    ST_N_FRAGMENTS = 2,
    ST_N_BARCODES = 3,
    ST_MAX_POS_LEN = 4,  // max(end - start, 0)
    ST_MAX_NEG_LEN = 5,  // max(start - end, 0)
    ST_MAX_COUNT = 6,    // largest window count seen
    ST_N_BINS = 7,
    ST_N_RUNS = 8,
    ST_N_FILLS = 9,
    ST_N_PEAKS = 10,
    ST_N_HITS = 11,
    ST_NNZ = 12,
    ST_THRESHOLD = 13,
    ST_HAS_PARAMS = 14,
    ST_EM_ITERS = 15,
    ST_EM_INNER = 16,
    ST_TICKET = 17,
    ST_TICKET2 = 18,
    ST_EM_DIRECT = 19,   // direct evaluations of the dispersion fixed-point map
    ST_EM_MODE = 20,     // 0 = fit pending, 1 = left to the generic kernel, 2 = done by the register-resident kernel
    ST_THRESHOLD_SPEC = 21,   // threshold after the first `spec_iter` EM iterations (speculative peak calling, api.cu run_tail)
    ST_THRESHOLD_SPEC2 = 22,  // the fit's second guess: the threshold its crossing is converging to, when that is not the first (em_fast.cu)
    ST_SLOTS = 24
};

// per-kernel timers (cudaEvent pairs on the context stream; only recorded when profiling is on)
enum StageId {
    SG_DECODE_COUNT = 0, SG_DECODE_PARSE, SG_FRAG_INDEX, SG_PILEUP_HIST, SG_HIST_COMPACT, SG_EM_FIT, SG_PILEUP_PEAKS, SG_MERGE_RUNS,
    SG_OVERLAP_COUNT, SG_OVERLAP_SCATTER, SG_REDUCE_SMALL, SG_REDUCE_LARGE, SG_CSC_BUILD, SG_BARCODE_TABLE, SG_COUNT
};

// growable device buffer (grow-only; keeps repeated steps free of cudaMalloc)
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // reference contigs (genome.fa.fai order)
    int n_contigs = 0;
    std::vector<std::string> contig_names;
    std::vector<int64_t> contig_lens;
    std::vector<int64_t> contig_base;      // global run coordinate of base 0 of each contig
    std::vector<int64_t> contig_tile_off;  // first genome tile of each contig (+ total)
    std::vector<int64_t> contig_pidx_off;  // first pos-index entry of each contig (+ total)
    DevBuf d_contig_lens, d_contig_base, d_contig_tile_off, d_contig_pidx_off, d_contig_table;
    int contig_table_cap = 0;
    int tile_bases = 0;                    // genome bases owned by one pileup tile

    // fragments (SoA, device)
    int64_t n_fragments = 0;
    DevBuf d_chrom, d_start, d_end, d_bc, d_dup;
    DevBuf d_contig_frag_off;              // int64[n_contigs+1]
    DevBuf d_pos_index;                    // int64 per 1024 bases: first fragment with start >= k * 1024
    DevBuf d_tile_desc;                    // per pileup tile: contig, extent, fragment range
    bool tile_desc_valid = false;
    bool tile_bound_valid = false;         // d_tile_bound / d_tile_summary hold the histogram pass's per-tile results for these fragments
    DevBuf d_tile_bound;
    // text
    const char* d_text = nullptr;          // borrowed or owned device text of the last decode
    int64_t text_bytes = 0;
    DevBuf d_text_own;
    DevBuf d_tile_counts, d_tile_offsets;  // decode line bookkeeping
    // barcodes
    int bc_table_cap = 0;                  // slots (power of two)
    DevBuf d_bc_keys, d_bc_first, d_bc_textoff, d_slot_to_dense;
    int64_t n_barcodes = 0;
    DevBuf d_dense_key, d_dense_first, d_dense_textoff, d_dense_to_col, d_col_to_dense, d_bc_bitmap;
    long long* h_bc_hash = nullptr;        // pinned: py2 hashes in first-appearance order
    int32_t* h_bc_order = nullptr;         // pinned: dense id of the k-th barcode to appear
    int32_t* h_c2d_pin = nullptr;          // pinned: column -> dense id
    size_t h_bc_cap = 0;
    cudaEvent_t bc_event = nullptr;
    const long long* d_hash_first = nullptr;   // device copies of the two mirrors above (inside d_bc_bitmap)
    const int32_t* d_order_first = nullptr;
    int64_t union_device_min = 20000;      // option "union_device_min": same for the union of the shards' barcodes (multi.cu), where the host
                                           // simulation would be the critical path of the step
    int64_t py2_device_min = 300000;       // option "py2_device_min": from this many barcodes on, the column order is worked out on the device
    std::vector<int32_t> h_c2d;
    bool bc_strings_ready = false;
    std::vector<std::string> barcodes;     // matrix column order
    int bc_order = CRATAC_BC_ORDER_PY2SET;
    bool bc_ready = false;
    std::thread bc_thread;                 // host simulation of the column order (build_barcodes_async)
    volatile bool bc_thread_done = false;
    DevBuf d_py2_rank;                     // first-appearance rank of every matrix column (device column order)
    DevBuf d_bc_nfrag, d_dense_nfrag;      // fragments per table slot / per dense barcode (int32)
    bool dense_nfrag_valid = false;        // d_dense_nfrag matches the current fragments (decode fills it; set_fragments counts on demand)
    bool bc_is_slot = false;               // d_bc holds hash-table slots (decode) vs dense ids (set_fragments)
    // histogram / threshold
    DevBuf d_hist;                         // uint64[HIST_CAP]
    DevBuf d_p2_scratch;                   // tables and probe state of py2_set_order_device (py2order.cu)
    DevBuf d_hist_values, d_hist_weights;  // compact, ascending
    DevBuf d_em_params;                    // double[8]
    DevBuf d_em_work;
    // peak calling
    DevBuf d_chain;                        // uint64 per tile
    DevBuf d_tile_summary;                 // int4-ish per tile
    DevBuf d_run_start, d_run_end, d_bridge, d_fill;
    int64_t run_cap = 0;
    DevBuf d_peak_chrom, d_peak_start, d_peak_end, d_peak_row, d_peak_off;  // sorted by (contig,start)
    int64_t n_peaks = 0;
    int64_t peak_cap = 0;
    bool peaks_on_device_count = false;    // n_peaks lives in d_status[ST_N_PEAKS] until fetched
    // matrix
    DevBuf d_bc_hits, d_seg_off, d_cursor, d_hits, d_out_row, d_out_cnt, d_nnz_col;
    DevBuf d_slot_seg;                     // segment offsets and cursors indexed by hash-table slot + dense id -> slot (overlap pass after a decode)
    DevBuf d_indptr, d_indices, d_data;
    DevBuf d_sel_indptr, d_sel_indices, d_sel_data, d_sel_kept, d_sel_scratch;   // result of select_columns (FILTER_PEAK_MATRIX)
    int64_t n_sel_cols = 0, sel_nnz = 0;
    int64_t nnz = 0;
    int64_t n_hits = 0;
    int64_t n_cols = 0;                    // columns of the current matrix (= n_barcodes unless a global column map is set)
    // multi-GPU: global column map / row range
    bool global_cols = false;
    int64_t n_global_cols = 0;
    int64_t row_range = 0;                 // rows of the global matrix (0 = the local peak count)
    DevBuf d_global_c2d, d_nnz_ordered, d_dense_bckey;
    // sibling scans (siblings.cu): fragment indices grouped by barcode in file order, segment offsets, sort scratch
    DevBuf d_sib_perm, d_sib_seg, d_sib_tmp, d_sib_hist;
    bool sib_valid = false;
    char* h_ring = nullptr;                // pinned ring the BGZF blocks of cratac_decode_file are inflated into
    size_t h_ring_cap = 0;
    int64_t file_chunk_bytes = (int64_t)32 << 20;   // option "file_chunk_bytes": text bytes per ring slot / H2D copy / decode launch
    DevBuf d_bg_W, d_bg_tiles, d_bg_start, d_bg_end, d_bg_count;   // bedgraph rows of one contig (bedgraph.cu)
    int64_t bg_rows = 0;
    int bg_contig = -1;
    bool bg_valid = false;                 // d_bg_* hold the rows of bg_contig for the current fragments
    DevBuf d_union, d_col_keys;            // scratch of union_columns (multi.cu); 64-bit barcode keys in global column order
    long long* h_union_hash = nullptr;     // pinned mirrors for the host simulation of the set order
    int32_t* h_union_perm = nullptr;
    size_t h_union_cap = 0;
    cudaEvent_t union_event = nullptr;
    // misc
    DevBuf d_status;                       // int64[ST_SLOTS]
    DevBuf d_scan_tmp;
    DevBuf d_scan_tmp_bc;                  // the same for the barcode-order thread (its scans run beside the main thread's)
    cudaStream_t bc_stream = nullptr;      // stream of that thread
    int64_t* h_status = nullptr;           // pinned mirror
    std::atomic<int64_t> launches{0};      // kernels launched since last reset (bench "gpu_launches"); the barcode thread counts too
    cudaStream_t copy_stream = nullptr;    // chunked H2D of a host text (decode_text_streamed)
    cudaEvent_t copy_event = nullptr;
    cudaStream_t count_stream = nullptr;   // high priority: newline counts + scans of the next chunk under the parse of the current one
    cudaEvent_t count_event = nullptr;
    int decode_pipeline = 0;               // option "decode_pipeline": 1 = device-resident texts of 64 MiB and more are decoded chunk by chunk on two
                                           // streams (2: any size); measured slower than one count + one parse launch: off
    DevBuf d_stream_scalars;               // running line base + chunk total of the streamed decode
    // single-pass decode (k_parse_stream): look-back words per tile, {ticket, declined ranges, line base}, declined ranges
    DevBuf d_sp_state, d_sp_scalars, d_sp_irr;
    int64_t sp_irr_cap = 0;
    int64_t* h_sp_scalars = nullptr;       // pinned mirror of d_sp_scalars (4 words)
    int decode_kernel = 0;                 // option "decode_kernel": 0 (default) two passes (count + parse); 1 single pass (k_parse_stream) with
                                           // 16 KiB tiles, 2 with 32 KiB tiles -- measured slower on B200, kept selectable
    int sp_ctas_per_sm = 0;                // option "decode_ctas_per_sm": cap on resident CTAs of k_parse_stream (0: occupancy)
    int decode_mode = 0;                   // option "decode_mode": experiment bits of k_parse_stream (StreamArgs::mode)
    int decode_profile = 0;                // option "decode_profile": per-phase cycle counters of k_parse_stream (cratac_decode_profile)
    int64_t decode_declined = 0;           // text ranges the single-pass kernel handed to the general kernel (since creation)
    int stream_decode = 1;                 // option "stream_decode": 0 off, 1 overlap the H2D copy of a host text >= 64 MiB with its decode, 2 always
    int64_t stream_chunk_tiles = 0;        // option "stream_chunk_tiles": text tiles per chunk (0: 128 MiB worth)
    // speculation: peaks + matrix are started from the threshold the fit shows after a few EM iterations, on a second
    // stream, while the fit (one SM) runs on; kept when the fit ends on the same threshold, redone otherwise
    int speculate = 1;                     // option "speculate"
    int spec_iter = 32;                    // option "spec_iter": EM iteration after which the tentative threshold is published
    int thr_slot = ST_THRESHOLD;           // status slot the peak kernels read their threshold from
    cudaStream_t spec_stream = nullptr;
    cudaStream_t own_stream_saved = nullptr;   // the context's stream while cratac_speculation_scope(1) is in force
    cudaEvent_t fit_event = nullptr;
    cudaEvent_t spec_done_event = nullptr;
    volatile int64_t* h_spec = nullptr;    // pinned + mapped: {tentative threshold, sequence number, second guess, sequence number} written by the fit kernel
    int64_t spec_seq = 0;
    bool spec_armed = false;               // the fit being launched publishes a tentative threshold
    int64_t spec_hits = 0, spec_misses = 0, spec_second = 0;   // runs kept / redone / runs in which the fit's second guess started a second pass
    int defer_sync = 0;                    // 1: window_histogram / compact_histogram / fit_threshold_device only queue their work
    int em_fast = 2;                       // 0 plain fixed point, 1 accelerated (em.cu), 2 register-resident kernel first (em_fast.cu)
    bool profile = false;
    cudaEvent_t ev[2 * SG_COUNT] = {};
    bool ev_used[SG_COUNT] = {};
    // host wall-clock marks (profiling only): time since the previous mark is charged to `name`
    std::vector<std::pair<std::string, double>> host_ms;
    std::chrono::steady_clock::time_point host_t0;
    void host_reset() { if (profile) { host_ms.clear(); host_t0 = std::chrono::steady_clock::now(); } }
    void host_mark(const char* name) {
        if (!profile) return;
        auto t = std::chrono::steady_clock::now();
        host_ms.emplace_back(name, std::chrono::duration<double, std::milli>(t - host_t0).count());
        host_t0 = t;
    }
    void stage_begin(int id) { if (profile) { cudaEventRecord(ev[2 * id], stream); } }
    void stage_end(int id) { if (profile) { cudaEventRecord(ev[2 * id + 1], stream); ev_used[id] = true; } }
};

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ void raise_error(int64_t* status, int code, int64_t arg) {
    if (atomicCAS(reinterpret_cast<unsigned long long*>(&status[ST_ERROR]), 0ULL, (unsigned long long)code) == 0ULL)
        status[ST_ERROR_ARG] = arg;
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---------------------------------------------------------------- scan.cu
// exclusive prefix sum of n int64 (in -> out, may alias); total written to *d_total if non-null.
int exclusive_scan_i64(Ctx* c, const int64_t* d_in, int64_t* d_out, int64_t n, int64_t* d_total);
int exclusive_scan_i32_to_i64(Ctx* c, const int32_t* d_in, int64_t* d_out, int64_t n, int64_t* d_total);

// A helper thread of the library (build_barcodes_async: the column order of very many barcodes, rebuilt on the device)
// queues its kernels on a stream of its own and scans with scratch of its own; the functions it shares with the main
// thread (scan.cu, py2order.cu) ask here which ones to use.
struct Lane { cudaStream_t stream; DevBuf* scan_tmp; };
extern thread_local Lane* tl_lane;
inline cudaStream_t lane_stream(Ctx* c) { return tl_lane ? tl_lane->stream : c->stream; }
inline DevBuf& lane_scan_tmp(Ctx* c) { return tl_lane ? *tl_lane->scan_tmp : c->d_scan_tmp; }

// Where the streamed decode takes its host text from: a buffer that is all there (cratac_decode_host), or a ring of
// pinned chunks that BGZF blocks are being inflated into on host threads (cratac_decode_file).
struct HostTextSource {
    virtual ~HostTextSource() {}
    virtual char last_byte() = 0;                                  // last byte of the whole text ('\n' when it is empty)
    virtual const char* acquire(int64_t b0, int64_t b1) = 0;       // blocks until text bytes [b0, b1) are in host memory; contiguous pointer (null: failed)
    virtual int queued(int64_t b0, int64_t b1, cudaStream_t copy_stream) = 0;   // their H2D copy has been queued on copy_stream
};

// ---------------------------------------------------------------- stage entry points (host side)
int decode_text(Ctx* c, const char* d_text, int64_t nbytes);          // decode.cu
int decode_text_streamed(Ctx* c, const char* h_text, int64_t nbytes); // decode.cu: host text, H2D overlapped with the decode
int64_t decode_stream_chunk_bytes(Ctx* c, int64_t chunk_bytes_arg);   // decode.cu: the chunk size decode_text_streamed_from will use
int decode_text_streamed_from(Ctx* c, HostTextSource* src, int64_t nbytes, int64_t chunk_bytes);   // decode.cu: same, chunks supplied as they become available
int finalize_fragments(Ctx* c);                                       // decode.cu: contig offsets, pos index, checks
void build_barcodes_async(Ctx* c);                                    // decode.cu: start the host half on a thread of its own
int build_barcodes(Ctx* c);                                           // decode.cu: matrix column order (host half)
int build_barcode_strings(Ctx* c);                                    // decode.cu: barcode strings in column order (lazy)
int window_histogram(Ctx* c);                                         // pileup.cu: K2+K3 fused
int call_peaks_device(Ctx* c);                                        // pileup.cu: K5 (threshold from d_status)
int fit_threshold_device(Ctx* c, double odds_ratio);                  // em.cu: K4 on compact histogram
int cell_fit(Ctx* c, const int64_t* h_values, const int64_t* h_weights, int64_t n, double odds_ratio, double out[8]);   // em.cu: cell-caller mixture
int peak_matrix_device(Ctx* c);                                       // matrix.cu: K6-K8
int peak_matrix_hits(Ctx* c);                                         // matrix.cu: K6 + K7 (no columns needed)
int peak_matrix_csc(Ctx* c);                                          // matrix.cu: K8
int low_targeting(Ctx* c, int64_t n_regions, const int32_t* chrom, const int32_t* start, const int32_t* end, int n_pad, const int64_t* paddings,
                  int64_t* out);                                      // siblings.cu
int insert_sizes(Ctx* c, int64_t n_sel, const int64_t* cols, const int32_t* h_primary, int exclude_non_primary, int64_t* table_out);
int counts_by_barcode(Ctx* c, const int64_t* track_off, const int32_t* r_chrom, const int32_t* r_start, const int32_t* r_end, const int8_t* r_minus,
                      int only_contig, int n_species, const int32_t* species_of_contig, int rel_half, int64_t* out, int64_t* relpos);
int bedgraph_rows_device(Ctx* c, int contig);                         // bedgraph.cu: K5b
int union_columns(Ctx* c, int world, int64_t max_b, const int64_t* d_gathered, const int64_t* h_nfrag, const int64_t* h_nbc, cudaStream_t s,
                  int64_t* n_cols_out);                               // multi.cu
int sync_status(Ctx* c);                                              // api.cu: D2H of status words + error check

}  // namespace cratac
