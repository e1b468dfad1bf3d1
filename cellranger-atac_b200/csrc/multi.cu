// Multi-GPU: the barcode universe of all shards -> global matrix columns, on the device.
//
// Replaces, for a run sharded by genome range, the `set` of generate_peak_matrix/__init__.py:44 that the reference builds by
// reading the whole fragments file: every rank decodes only its contig range, so it knows its own barcodes (64-bit key,
// first line inside the shard, CPython-2.7 hash).  The ranks all-gather those triples (the transport is the caller's:
// NCCL through torch.distributed) and every rank runs the same union here:
//
//   k_union_insert    open-addressing table over all (rank, barcode) pairs: atomicCAS on the key, atomicMin on the first
//                     GLOBAL line (shard-local first line + lines of the lower ranks)
//   k_union_compact   occupied slots -> dense "unique" ids (any order)
//   first-line rank   the first lines of distinct barcodes are distinct fragment indices: bitmap + popcount prefix ranks
//                     them without a sort = insertion order of the reference's set
//   column order      CPython-2.7 set order of the hashes in insertion order (host simulation or py2order.cu) or the
//                     insertion order itself
//   k_union_local     this rank's barcodes looked up in the table -> column of every local barcode
//
// Everything is queued on the stream the caller names (a side stream: the mixture fit occupies the context's own).
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace cratac {

void py2_set_order_hashes(const long long* hashes, int64_t n, std::vector<int32_t>& order);        // decode.cu
int py2_set_order_device(Ctx* c, int64_t n, const int64_t* d_hashes, int32_t* d_order_out);        // py2order.cu

constexpr unsigned long long U_EMPTY = 0xFFFFFFFFFFFFFFFFULL;
struct __align__(16) USlot { unsigned long long key; unsigned long long first; };

__device__ __forceinline__ unsigned long long u_mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

__global__ void k_union_init(USlot* __restrict__ tab, int64_t cap) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap) { USlot e; e.key = U_EMPTY; e.first = U_EMPTY; tab[s] = e; }
}

// gathered: [world][3][max_b] = keys, shard-local first lines, CPython hashes
__global__ void k_union_insert(const int64_t* __restrict__ gathered, const int64_t* __restrict__ n_bc, const int64_t* __restrict__ line_base, int world,
                               int64_t max_b, USlot* __restrict__ tab, long long* __restrict__ tab_hash, unsigned long long mask, int64_t* status) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)world * max_b) return;
    const int r = (int)(idx / max_b);
    const int64_t i = idx - (int64_t)r * max_b;
    if (i >= n_bc[r]) return;
    const int64_t* g = gathered + (int64_t)r * 3 * max_b;
    const unsigned long long key = (unsigned long long)g[i];
    const unsigned long long first = (unsigned long long)(g[max_b + i] + line_base[r]);
    unsigned long long slot = u_mix64(key) & mask;
    for (unsigned long long probes = 0; probes <= mask; probes++, slot = (slot + 1) & mask) {
        unsigned long long k = tab[slot].key;
        if (k == U_EMPTY) k = atomicCAS(&tab[slot].key, U_EMPTY, key);
        if (k == U_EMPTY || k == key) {
            atomicMin(&tab[slot].first, first);
            tab_hash[slot] = g[2 * max_b + i];          // the same value from every rank that holds the barcode
            return;
        }
    }
    raise_error(status, CRATAC_ERR_CAPACITY, idx);
}

__global__ void k_union_compact(const USlot* __restrict__ tab, const long long* __restrict__ tab_hash, int64_t cap, unsigned long long* counter,
                                unsigned long long* __restrict__ uniq_key, unsigned long long* __restrict__ uniq_first, long long* __restrict__ uniq_hash,
                                int32_t* __restrict__ slot_to_uniq) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap) return;
    const USlot e = tab[s];
    if (e.key == U_EMPTY) { slot_to_uniq[s] = -1; return; }
    const unsigned long long u = atomicAdd(counter, 1ULL);
    slot_to_uniq[s] = (int32_t)u;
    uniq_key[u] = e.key; uniq_first[u] = e.first; uniq_hash[u] = tab_hash[s];
}

__global__ void k_union_mark(const unsigned long long* __restrict__ uniq_first, int64_t n, unsigned int* __restrict__ bitmap) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const unsigned long long f = uniq_first[u];
    atomicOr(&bitmap[f >> 5], 1u << (f & 31));
}
__global__ void k_union_popc(const unsigned int* __restrict__ bitmap, int64_t n_words, int32_t* __restrict__ cnt) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < n_words) cnt[w] = __popc(bitmap[w]);
}
// rank r of unique barcode u in insertion order; hash_by_rank feeds the set-order simulation
__global__ void k_union_rank(const unsigned long long* __restrict__ uniq_first, const long long* __restrict__ uniq_hash, int64_t n,
                             const unsigned int* __restrict__ bitmap, const int64_t* __restrict__ word_prefix, int32_t* __restrict__ uniq_of_rank,
                             long long* __restrict__ hash_by_rank) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const unsigned long long f = uniq_first[u];
    const int64_t r = word_prefix[f >> 5] + __popc(bitmap[f >> 5] & ((1u << (f & 31)) - 1u));
    uniq_of_rank[r] = (int32_t)u;
    hash_by_rank[r] = uniq_hash[u];
}
// column j holds the barcode of insertion rank rank_of_col[j] (null: j itself)
__global__ void k_union_columns(const int32_t* __restrict__ rank_of_col, const int32_t* __restrict__ uniq_of_rank, const unsigned long long* __restrict__ uniq_key,
                                int64_t n, int32_t* __restrict__ col_of_uniq, int64_t* __restrict__ col_keys, int32_t* __restrict__ c2d) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int32_t u = uniq_of_rank[rank_of_col ? rank_of_col[j] : (int32_t)j];
    col_of_uniq[u] = (int32_t)j;
    col_keys[j] = (int64_t)uniq_key[u];
    c2d[j] = -1;
}
__global__ void k_union_local(const unsigned long long* __restrict__ local_keys, int64_t nb, const USlot* __restrict__ tab, unsigned long long mask,
                              const int32_t* __restrict__ slot_to_uniq, const int32_t* __restrict__ col_of_uniq, int32_t* __restrict__ c2d, int64_t* status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const unsigned long long key = local_keys[i];
    unsigned long long slot = u_mix64(key) & mask;
    for (unsigned long long probes = 0; probes <= mask; probes++, slot = (slot + 1) & mask) {
        const unsigned long long k = tab[slot].key;
        if (k == key) { c2d[col_of_uniq[slot_to_uniq[slot]]] = (int32_t)i; return; }
        if (k == U_EMPTY) break;
    }
    raise_error(status, CRATAC_ERR_INTERNAL, -400 - i);
}

// see include/cratac.h: cratac_union_columns
int union_columns(Ctx* c, int world, int64_t max_b, const int64_t* d_gathered, const int64_t* h_nfrag, const int64_t* h_nbc, cudaStream_t s,
                  int64_t* n_cols_out) {
    int64_t total_bc = 0, total_frag = 0;
    std::vector<int64_t> meta(2 * (size_t)world);                       // n_bc[world] | line_base[world]
    for (int r = 0; r < world; r++) {
        if (h_nbc[r] < 0 || h_nbc[r] > max_b || h_nfrag[r] < 0) { set_error("cratac_union_columns: bad counts for rank %d", r); return CRATAC_ERR_ARG; }
        meta[r] = h_nbc[r];
        meta[world + r] = total_frag;
        total_bc += h_nbc[r];
        total_frag += h_nfrag[r];
    }
    int64_t cap = 1024;
    while (cap < 2 * total_bc) cap <<= 1;
    const int64_t n_words = (total_frag + 31) / 32 + 1;
    const size_t tb = (size_t)std::max<int64_t>(total_bc, 1);
    // scratch layout
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_tab = take(16 * (size_t)cap), o_hash = take(8 * (size_t)cap), o_s2u = take(4 * (size_t)cap), o_key = take(8 * tb), o_first = take(8 * tb),
                 o_uhash = take(8 * tb), o_bitmap = take(4 * (size_t)n_words), o_cnt = take(4 * (size_t)n_words), o_prefix = take(8 * (size_t)n_words),
                 o_uor = take(4 * tb), o_hbr = take(8 * tb), o_perm = take(4 * tb), o_cou = take(4 * tb), o_meta = take(16 * (size_t)world), o_counter = take(8);
    CRATAC_TRY(c->d_union.reserve(off));
    char* B = c->d_union.as<char>();
    USlot* tab = reinterpret_cast<USlot*>(B + o_tab);
    long long* tab_hash = reinterpret_cast<long long*>(B + o_hash);
    int32_t* slot_to_uniq = reinterpret_cast<int32_t*>(B + o_s2u);
    unsigned long long* uniq_key = reinterpret_cast<unsigned long long*>(B + o_key);
    unsigned long long* uniq_first = reinterpret_cast<unsigned long long*>(B + o_first);
    long long* uniq_hash = reinterpret_cast<long long*>(B + o_uhash);
    unsigned int* bitmap = reinterpret_cast<unsigned int*>(B + o_bitmap);
    int32_t* word_cnt = reinterpret_cast<int32_t*>(B + o_cnt);
    int64_t* word_prefix = reinterpret_cast<int64_t*>(B + o_prefix);
    int32_t* uniq_of_rank = reinterpret_cast<int32_t*>(B + o_uor);
    long long* hash_by_rank = reinterpret_cast<long long*>(B + o_hbr);
    int32_t* perm = reinterpret_cast<int32_t*>(B + o_perm);
    int32_t* col_of_uniq = reinterpret_cast<int32_t*>(B + o_cou);
    int64_t* d_meta = reinterpret_cast<int64_t*>(B + o_meta);
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(B + o_counter);
    int64_t* st = c->d_status.as<int64_t>();

    // the library's scan helper works on the context's current stream: point it at the caller's for the duration
    cudaStream_t saved = c->stream;
    c->stream = s;
    struct Restore { Ctx* c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{c, saved};

    CRATAC_CUDA(cudaMemcpyAsync(d_meta, meta.data(), 16 * (size_t)world, cudaMemcpyHostToDevice, s));
    CRATAC_CUDA(cudaMemsetAsync(counter, 0, 8, s));
    CRATAC_CUDA(cudaMemsetAsync(bitmap, 0, 4 * (size_t)n_words, s));
    k_union_init<<<(unsigned)((cap + 255) / 256), 256, 0, s>>>(tab, cap);
    const int64_t pairs = (int64_t)world * max_b;
    if (pairs > 0)
        k_union_insert<<<(unsigned)((pairs + 255) / 256), 256, 0, s>>>(d_gathered, d_meta, d_meta + world, world, max_b, tab, tab_hash, (unsigned long long)(cap - 1), st);
    k_union_compact<<<(unsigned)((cap + 255) / 256), 256, 0, s>>>(tab, tab_hash, cap, counter, uniq_key, uniq_first, uniq_hash, slot_to_uniq);
    c->launches += 3;
    CRATAC_CUDA(cudaGetLastError());
    unsigned long long h_n = 0;
    CRATAC_CUDA(cudaMemcpyAsync(&h_n, counter, 8, cudaMemcpyDeviceToHost, s));
    CRATAC_CUDA(cudaStreamSynchronize(s));
    const int64_t n_cols = (int64_t)h_n;
    *n_cols_out = n_cols;
    const size_t nn = (size_t)std::max<int64_t>(n_cols, 1);
    CRATAC_TRY(c->d_global_c2d.reserve(4 * nn));
    CRATAC_TRY(c->d_col_keys.reserve(8 * nn));
    if (n_cols > 0) {
        const unsigned g = (unsigned)((n_cols + 255) / 256);
        k_union_mark<<<g, 256, 0, s>>>(uniq_first, n_cols, bitmap);
        k_union_popc<<<(unsigned)((n_words + 255) / 256), 256, 0, s>>>(bitmap, n_words, word_cnt);
        c->launches += 2;
        CRATAC_TRY(exclusive_scan_i32_to_i64(c, word_cnt, word_prefix, n_words, nullptr));
        k_union_rank<<<g, 256, 0, s>>>(uniq_first, uniq_hash, n_cols, bitmap, word_prefix, uniq_of_rank, hash_by_rank);
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
        const int32_t* rank_of_col = nullptr;
        if (c->bc_order == CRATAC_BC_ORDER_PY2SET) {
            if (n_cols >= c->union_device_min) {
                CRATAC_TRY(py2_set_order_device(c, n_cols, reinterpret_cast<const int64_t*>(hash_by_rank), perm));
            } else {
                // host simulation of the set (sequential, ~50 ns per key): runs while the GPU is busy with the fit
                if (c->h_union_cap < nn) {
                    if (c->h_union_hash) cudaFreeHost(c->h_union_hash);
                    if (c->h_union_perm) cudaFreeHost(c->h_union_perm);
                    const size_t want = nn + nn / 2 + 1024;
                    if (cudaMallocHost(&c->h_union_hash, 8 * want) != cudaSuccess || cudaMallocHost(&c->h_union_perm, 4 * want) != cudaSuccess) {
                        set_error("cudaMallocHost failed for the column-order mirrors");
                        return CRATAC_ERR_CUDA;
                    }
                    c->h_union_cap = want;
                }
                CRATAC_CUDA(cudaMemcpyAsync(c->h_union_hash, hash_by_rank, 8 * (size_t)n_cols, cudaMemcpyDeviceToHost, s));
                CRATAC_CUDA(cudaStreamSynchronize(s));
                std::vector<int32_t> order;
                py2_set_order_hashes(c->h_union_hash, n_cols, order);
                std::copy(order.begin(), order.end(), c->h_union_perm);
                CRATAC_CUDA(cudaMemcpyAsync(perm, c->h_union_perm, 4 * (size_t)n_cols, cudaMemcpyHostToDevice, s));
            }
            rank_of_col = perm;
        } else if (c->bc_order != CRATAC_BC_ORDER_FIRST_SEEN) {
            set_error("cratac_union_columns: barcode order %d is not available across shards", c->bc_order);
            return CRATAC_ERR_ARG;
        }
        k_union_columns<<<g, 256, 0, s>>>(rank_of_col, uniq_of_rank, uniq_key, n_cols, col_of_uniq, c->d_col_keys.as<int64_t>(), c->d_global_c2d.as<int32_t>());
        c->launches++;
        const int64_t nb = c->n_barcodes;
        if (nb > 0) {
            k_union_local<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(c->d_dense_bckey.as<unsigned long long>(), nb, tab, (unsigned long long)(cap - 1), slot_to_uniq,
                                                                      col_of_uniq, c->d_global_c2d.as<int32_t>(), st);
            c->launches++;
        }
        CRATAC_CUDA(cudaGetLastError());
    }
    // the matrix build on the context's own stream reads the map: order it behind this stream
    if (!c->union_event) CRATAC_CUDA(cudaEventCreateWithFlags(&c->union_event, cudaEventDisableTiming));
    CRATAC_CUDA(cudaEventRecord(c->union_event, s));
    if (s != saved) CRATAC_CUDA(cudaStreamWaitEvent(saved, c->union_event, 0));
    c->global_cols = true;
    c->n_global_cols = n_cols;
    return CRATAC_OK;
}

}  // namespace cratac
