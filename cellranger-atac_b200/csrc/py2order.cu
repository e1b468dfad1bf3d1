// Matrix column order on the device: the iteration order of a CPython-2.7 set built by inserting distinct str
// keys one by one (generate_peak_matrix/__init__.py:47 takes `list(set(barcodes))`), from the keys' hashes in
// insertion order.  Same result as the host simulation py2_set_order_hashes (decode.cu), which is sequential
// (~10 ms for 200k barcodes, per rank, on the critical path of a multi-GPU step).
//
// A set of that era is an open-addressing table: slot i = hash & mask, then i = 5 i + perturb + 1,
// perturb >>= 5 (Objects/setobject.c, set_lookkey_string); it is rebuilt 4x (2x above 50000 entries) larger
// when fill * 3 >= size * 2, re-inserting the old table in slot order.  Each table is therefore the result of
// inserting a known SEQUENCE of keys (old slot order, then the new keys) one after the other.  Sequential
// insertion puts every key into the first slot of its probe sequence that no EARLIER key of the sequence ends
// up in -- which is a fixed point that can be reached in parallel: a slot holds the smallest sequence position
// that has asked for it (atomicMin), a key that finds a smaller position in its slot moves on along its probe
// sequence, and the rounds repeat until nothing moves.  Slot values only ever decrease, so a key never has to
// come back to a slot it has left.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace cratac {

constexpr uint32_t P2_EMPTY = 0xFFFFFFFFu;

// seq[t] = key (index into hashes) at sequence position t
__global__ void k_p2_begin(const int64_t* __restrict__ hashes, const int32_t* __restrict__ seq, int m, unsigned long long mask,
                           unsigned long long* __restrict__ st_i, unsigned long long* __restrict__ st_p, uint32_t* __restrict__ cand) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const unsigned long long h = (unsigned long long)hashes[seq[t]];
    st_i[t] = h & mask;
    st_p[t] = h;
    cand[t] = (uint32_t)(h & mask);
}

__global__ void k_p2_round(int m, unsigned long long mask, unsigned long long* __restrict__ st_i, unsigned long long* __restrict__ st_p,
                           uint32_t* __restrict__ cand, uint32_t* __restrict__ table, int* __restrict__ changed) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    uint32_t c = cand[t];
    if (table[c] == (uint32_t)t) return;            // holding its slot (for now)
    unsigned long long i = st_i[t], p = st_p[t];
    bool moved = false;
    for (int guard = 0; guard < (1 << 20); guard++) {
        const uint32_t prev = atomicMin(&table[c], (uint32_t)t);
        if (prev >= (uint32_t)t) break;             // the slot is mine unless an earlier key still comes for it
        i = i * 5ULL + p + 1ULL;                    // set_lookkey_string: i = (i << 2) + i + perturb + 1
        p >>= 5;
        c = (uint32_t)(i & mask);
        moved = true;
    }
    if (moved) { st_i[t] = i; st_p[t] = p; cand[t] = c; }
    *changed = 1;
}

__global__ void k_p2_flags(const uint32_t* __restrict__ table, int64_t size, int32_t* __restrict__ flags) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < size) flags[s] = table[s] != P2_EMPTY;
}

__global__ void k_p2_collect(const uint32_t* __restrict__ table, const int64_t* __restrict__ pos, int64_t size, const int32_t* __restrict__ seq,
                             int32_t* __restrict__ out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < size && table[s] != P2_EMPTY) out[pos[s]] = seq[table[s]];
}

__global__ void k_p2_iota(int32_t* __restrict__ seq, int from, int to) {
    const int k = from + blockIdx.x * blockDim.x + threadIdx.x;
    if (k < to) seq[k] = k;
}

int py2_set_order_device(Ctx* c, int64_t n, const int64_t* d_hashes, int32_t* d_order_out) {
    if (n <= 0) return CRATAC_OK;
    if (n >= (int64_t)1 << 30) { set_error("py2_set_order_device: too many keys"); return CRATAC_ERR_ARG; }
    cudaStream_t s = lane_stream(c);
    // largest table: the growth rule applied to n keys
    uint64_t max_size = 8;
    {
        uint64_t size = 8, fill = 0;
        for (int64_t used = 0; used < n;) {
            const uint64_t room = (2 * size + 2) / 3;                       // smallest fill with fill * 3 >= size * 2
            const uint64_t take = std::min<uint64_t>((uint64_t)(n - used), room > fill ? room - fill : 1);
            used += (int64_t)take; fill += take;
            if (fill * 3 >= size * 2) {
                const uint64_t minused = (uint64_t)used > 50000 ? (uint64_t)used * 2 : (uint64_t)used * 4;
                size = 8;
                while (size <= minused) size <<= 1;
                fill = (uint64_t)used;
            }
            max_size = std::max(max_size, size);
        }
    }
    DevBuf& B = c->d_p2_scratch;
    const size_t nn = (size_t)n, ts = (size_t)max_size;
    // layout: table u32[ts] | flags i32[ts] | pos i64[ts + 1] | st_i u64[n] | st_p u64[n] | cand u32[n] | seqA i32[n] | seqB i32[n] | changed
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_table = take(4 * ts), o_flags = take(4 * ts), o_pos = take(8 * (ts + 1)), o_i = take(8 * nn), o_p = take(8 * nn),
                 o_cand = take(4 * nn), o_a = take(4 * nn), o_b = take(4 * nn), o_chg = take(256);
    CRATAC_TRY(B.reserve(off));
    char* base = B.as<char>();
    uint32_t* table = reinterpret_cast<uint32_t*>(base + o_table);
    int32_t* flags = reinterpret_cast<int32_t*>(base + o_flags);
    int64_t* pos = reinterpret_cast<int64_t*>(base + o_pos);
    unsigned long long* st_i = reinterpret_cast<unsigned long long*>(base + o_i);
    unsigned long long* st_p = reinterpret_cast<unsigned long long*>(base + o_p);
    uint32_t* cand = reinterpret_cast<uint32_t*>(base + o_cand);
    int32_t* seq = reinterpret_cast<int32_t*>(base + o_a);
    int32_t* nxt = reinterpret_cast<int32_t*>(base + o_b);
    int* changed = reinterpret_cast<int*>(base + o_chg);

    uint64_t size = 8, fill = 0;
    int64_t used = 0;
    bool pending = true;                        // a table has to be built (new keys or a rebuild)
    while (pending) {
        // keys of this table: the previous slot order seq[0, used) and the new keys used .. m - 1
        const uint64_t room = (2 * size + 2) / 3;
        const int64_t fresh = std::min<int64_t>(n - used, room > fill ? (int64_t)(room - fill) : (n > used ? 1 : 0));
        const int64_t m = used + fresh;
        if (fresh > 0) {
            k_p2_iota<<<(unsigned)((fresh + 255) / 256), 256, 0, s>>>(seq, (int)used, (int)m);
            c->launches++;
        }
        const unsigned long long mask = size - 1;
        CRATAC_CUDA(cudaMemsetAsync(table, 0xFF, 4 * (size_t)size, s));
        const unsigned g = (unsigned)((m + 255) / 256);
        k_p2_begin<<<g, 256, 0, s>>>(d_hashes, seq, (int)m, mask, st_i, st_p, cand);
        c->launches++;
        // rounds are idempotent once the fixed point is reached: queue four at a time, then look at the flag of the
        // last one (one host round trip per batch instead of per round)
        for (int batch = 0;; batch++) {
            if (batch > 1024) { set_error("py2_set_order_device: no fixed point"); return CRATAC_ERR_INTERNAL; }
            for (int r = 0; r < 4; r++) {
                if (r == 3) CRATAC_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), s));
                k_p2_round<<<g, 256, 0, s>>>((int)m, mask, st_i, st_p, cand, table, changed);
            }
            c->launches += 4;
            int h_changed = 0;
            CRATAC_CUDA(cudaMemcpyAsync(&h_changed, changed, sizeof(int), cudaMemcpyDeviceToHost, s));
            CRATAC_CUDA(cudaStreamSynchronize(s));
            if (!h_changed) break;
        }
        // slot order of this table
        const unsigned gs = (unsigned)((size + 255) / 256);
        k_p2_flags<<<gs, 256, 0, s>>>(table, (int64_t)size, flags);
        CRATAC_TRY(exclusive_scan_i32_to_i64(c, flags, pos, (int64_t)size, pos + size));
        const bool last_keys = m == n;
        fill += (uint64_t)fresh;
        used = m;
        const bool grow = fill * 3 >= size * 2;
        int32_t* out = (last_keys && !grow) ? d_order_out : nxt;
        k_p2_collect<<<gs, 256, 0, s>>>(table, pos, (int64_t)size, seq, out);
        c->launches += 2;
        CRATAC_CUDA(cudaGetLastError());
        if (last_keys && !grow) break;
        std::swap(seq, nxt);
        if (grow) {
            const uint64_t minused = (uint64_t)used > 50000 ? (uint64_t)used * 2 : (uint64_t)used * 4;
            size = 8;
            while (size <= minused) size <<= 1;
            fill = (uint64_t)used;
        }
        pending = true;
    }
    CRATAC_CUDA(cudaStreamSynchronize(s));
    return CRATAC_OK;
}

}  // namespace cratac
