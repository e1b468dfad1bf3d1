// K1: fragments.tsv text -> SoA int32 (chrom, start, end, barcode slot, dup) + barcode dictionary.
// Replaces lib/python/tools/io.py:75-79 / :82-92 (one str.split + 3 int() per line on the host).
//
// Layout: the text is cut into fixed 16 KiB tiles.  A line belongs to the tile that holds its
// terminating '\n'.  Pass A counts newlines per tile (128-bit loads), a scan turns the counts into
// line offsets, pass B stages each tile (+256 B look-back for the line that straddles the tile
// start) into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier), turns the newline and tab
// bits of the tile into ordered position lists, and parses one line per thread (its 4 tabs are 4
// consecutive list entries).  A host text is copied in chunks and decoded chunk by chunk under the copy
// (decode_text_streamed).  Barcodes of the 10x form [ACGT]{16}-<int> pack
// losslessly into 64 bits; any other byte string falls back to a 63-bit hash.  Both go through one
// open-addressing table in HBM/L2 (atomicCAS) that also records the first line of each barcode.
#include <algorithm>
#include <string.h>

#include "common.cuh"
#include "parse_swar.cuh"

namespace cratac {

constexpr int DEC_TILE = 16384;
constexpr int DEC_LOOKBACK = 256;
constexpr int DEC_THREADS = 256;
constexpr int DEC_MAX_LINES = DEC_TILE / 8;   // a valid line has >= 10 bytes
constexpr int DEC_MAX_TABS = 3072;               // tab list capacity (lines beyond it take the general path)
constexpr unsigned long long BC_EMPTY = 0xFFFFFFFFFFFFFFFFULL;

struct ContigEntry { unsigned long long hash; int32_t id; int32_t used; };
// barcode table slot: key and first line share one 16-byte slot (one sector per probe)
struct __align__(16) BcSlot { unsigned long long key; unsigned long long first; };

__global__ void k_init_table(BcSlot* __restrict__ tab, int cap) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap) { BcSlot e; e.key = BC_EMPTY; e.first = 0x7F7F7F7F7F7F7F7FULL; tab[s] = e; }
}

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// ------------------------------------------------------------------ pass A: newline counts
__global__ void __launch_bounds__(256) k_count_newlines(const char* __restrict__ text, int64_t nbytes, int64_t n_eff,
                                                         int32_t* __restrict__ tile_counts, int64_t tile0, int tile_bytes) {
    // one CTA per tile (tile_bytes: a multiple of 16, DEC_TILE on the two-pass path); 256 threads x 16 B per round
    int64_t tile_base = ((int64_t)blockIdx.x + tile0) * tile_bytes;
    int cnt = 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(text) & 15) == 0);
    for (int k = 0; k * (256 * 16) < tile_bytes; k++) {
        if ((k * 256 + (int)threadIdx.x) * 16 >= tile_bytes) break;
        int64_t off = tile_base + ((int64_t)k * 256 + threadIdx.x) * 16;
        if (aligned && off + 16 <= nbytes) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(text + off));
            cnt += __popc(__vcmpeq4(v.x, 0x0a0a0a0au)) + __popc(__vcmpeq4(v.y, 0x0a0a0a0au)) +
                   __popc(__vcmpeq4(v.z, 0x0a0a0a0au)) + __popc(__vcmpeq4(v.w, 0x0a0a0a0au));
        } else {
            int c8 = 0;
            for (int j = 0; j < 16; j++) {
                int64_t p = off + j;
                if (p < nbytes) c8 += (text[p] == '\n');
                else if (p < n_eff) c8 += 1;   // virtual terminator after an unterminated last line
            }
            cnt += c8 * 8;
        }
    }
    cnt >>= 3;  // __vcmpeq4 sets 8 bits per matching byte
    __shared__ int warp_cnt[8];
#pragma unroll
    for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < 8; w++) s += warp_cnt[w];
        tile_counts[blockIdx.x + tile0] = s;
    }
}

// ------------------------------------------------------------------ TMA bulk copy helpers (sm_90+/sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
// bounded wait: a mis-programmed copy must surface as an error, never as a hung GPU
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t phase) {
    for (int spin = 0; spin < (1 << 22); spin++)
        if (mbar_try_wait(bar, phase)) return true;
    return false;
}

// ------------------------------------------------------------------ pass B: parse
struct DecodeArgs {
    const char* text;
    int64_t nbytes, n_eff;
    const int64_t* tile_offsets;     // exclusive scan of tile_counts
    const int32_t* tile_counts;
    const ContigEntry* contig_table;
    int contig_mask;
    int32_t *chrom, *start, *end, *bc, *dup;
    BcSlot* bc_tab;
    int32_t* bc_nfrag;                // fragments per table slot (sizes the matrix's per-barcode hit segments: no count pass)
    long long* bc_textoff;
    int bc_mask;
    int64_t* status;
    int use_tma;
    // streamed decode (host text copied in chunks): tiles of this launch start at tile0, line numbers at *line_base
    int64_t tile0;
    const int64_t* line_base;
    int64_t line_cap;                 // capacity of the output arrays when the total is not known in advance (0: exact)
    // the two-pass path uses DEC_TILE-byte tiles numbered by blockIdx; as the general-path fallback of k_parse_stream the
    // kernel is given the list of byte ranges that kernel declined (each with its line count and first output line)
    const struct IrrTile* irr;        // null: tile = blockIdx.x + tile0
};
// one range of text handed from k_parse_stream to k_parse_tiles: [byte_base, byte_base + nbytes) (nbytes a multiple of
// 16, at most DEC_TILE), the lines whose '\n' lies inside it, the output index of the first of them, and how many of
// them are already committed
struct IrrTile { int64_t byte_base; int64_t out_base; int32_t nbytes; int32_t n_lines; int32_t skip; int32_t pad; };
static_assert(sizeof(IrrTile) == 32, "IrrTile layout");

// The host rebuilds a table that ends more than half full four times as large and decodes again (decode_text), so a launch
// that has already crossed that line only needs to end: a long probe sequence checks the key counter and gives up with
// CRATAC_ERR_CAPACITY.  Without this a table with fewer slots than the file has barcodes is probed from end to end by every
// line with a new key (1.7 M barcodes against the initial 1 M slots: 215 s for 160 M lines instead of 12 ms).
__device__ __forceinline__ bool bc_table_crowded(const int64_t* status, int bc_mask) {
    return 2 * *reinterpret_cast<const volatile long long*>(&status[ST_N_BARCODES]) > (long long)bc_mask + 1;
}
__device__ __forceinline__ unsigned byte_at(const uint32_t* words, int p) { return (words[p >> 2] >> ((p & 3) * 8)) & 0xFFu; }

__device__ __forceinline__ unsigned long long barcode_key_words(const uint32_t* words, int p, int n) {
    // fast path: [ACGT]{16}-[0-9]{1,9} without a leading zero in the group (so that "-01" cannot alias "-1")
    if (n >= 18 && n <= 26 && byte_at(words, p + 16) == '-' && !(n > 18 && byte_at(words, p + 17) == '0')) {
        bool ok = true;
        unsigned seq = (swar_pack4(swar_get4(words, p), &ok) << 24) | (swar_pack4(swar_get4(words, p + 4), &ok) << 16) |
                       (swar_pack4(swar_get4(words, p + 8), &ok) << 8) | swar_pack4(swar_get4(words, p + 12), &ok);
        int32_t grp = 0;
        ok = ok && swar_parse_uint(words, p + 17, n - 17, &grp);
        if (ok) return ((unsigned long long)(unsigned)grp << 32) | seq;
    }
    unsigned long long h = mix64(swar_fnv1a(words, p, n) ^ ((unsigned long long)n << 56));
    h = (h >> 1) | 0x8000000000000000ULL;
    if (h == BC_EMPTY) h ^= 2ULL;
    return h;
}

__device__ __forceinline__ bool sp_parse_uint(const uint32_t* words, int p, int e, int32_t* out);   // below: 32-bit arithmetic only

__global__ void __launch_bounds__(DEC_THREADS) k_parse_tiles(DecodeArgs a) {
    // text tile with 16 B of slack on either side: the word windows of the field parsers may reach before the first byte
    // (a number ending within the first 8 bytes) and past the last one
    __shared__ __align__(128) uint32_t words_raw[4 + (DEC_LOOKBACK + DEC_TILE) / 4 + 4];
    uint32_t* const words = words_raw + 4;
    __shared__ uint32_t lb_masks[DEC_LOOKBACK / 16];   // look-back chunks: newline bits (low half), tab bits (high half)
    __shared__ uint16_t nlpos[DEC_MAX_LINES];          // newline positions of the tile region, in order
    __shared__ uint16_t nltabs[DEC_MAX_LINES];         // tabs of the tile region before each of them
    __shared__ uint16_t tabpos[DEC_MAX_TABS];          // tab positions of the tile region, in order
    __shared__ int warp_tot[DEC_THREADS / 32];
    __shared__ int s_first_begin;
    __shared__ __align__(8) uint64_t bar;
    unsigned char* buf = reinterpret_cast<unsigned char*>(words);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    int64_t tile_base, out_base;
    int TB, n_lines, skip = 0;
    if (a.irr) {
        const IrrTile e = a.irr[blockIdx.x];
        tile_base = e.byte_base; out_base = e.out_base; TB = e.nbytes; n_lines = e.n_lines; skip = e.skip;
    } else {
        const int64_t tile = (int64_t)blockIdx.x + a.tile0;
        tile_base = tile * DEC_TILE; TB = DEC_TILE; n_lines = a.tile_counts[tile];
        out_base = a.tile_offsets[tile] + (a.line_base ? *a.line_base : 0);
    }
    if (n_lines == 0) return;   // uniform per CTA
    const int64_t src0 = tile_base - DEC_LOOKBACK;

    // ---- stage bytes [src0, tile_base + TB) into shared memory: one TMA bulk copy when the tile is interior
    const bool full = a.use_tma && src0 >= 0 && tile_base + TB <= a.nbytes && (src0 & 15) == 0;
    if (full) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar, DEC_LOOKBACK + TB);
            tma_load_1d(buf, a.text + src0, DEC_LOOKBACK + TB, &bar);
        }
        if (!mbar_wait(&bar, 0)) { if (tid == 0) raise_error(a.status, CRATAC_ERR_INTERNAL, -100 - tile_base / DEC_TILE); return; }
    } else {
        for (int i = tid; i < DEC_LOOKBACK + TB; i += DEC_THREADS) {
            int64_t p = src0 + i;
            unsigned char ch = '\n';
            if (p >= 0 && p < a.nbytes) ch = (unsigned char)a.text[p];
            buf[i] = ch;   // p < 0 and p >= nbytes read as '\n' (file start / virtual terminator)
        }
        __syncthreads();
    }

    // ---- newline / tab bits of my 64 tile bytes (thread t owns 4 chunks of 16; the first 16 threads also take
    //      one look-back chunk each), then ordered position lists of both through one packed block scan
    constexpr int PER_THREAD = DEC_TILE / DEC_THREADS;   // 64 contiguous bytes per thread
    unsigned long long mask = 0, tmask = 0;
    {
        const int c0 = DEC_LOOKBACK / 16 + tid * (PER_THREAD / 16);
        // a thread's four chunks are 64 bytes apart from its neighbour's: taken in the same order by every lane, the eight
        // lanes of a quarter-warp would meet in two groups of banks (16-way conflicts, measured: 46 % of the kernel's
        // shared-memory wavefronts).  Rotating the order by lane / 2 spreads a quarter-warp over all 32 banks.
        const int rot = (lane >> 1) & 3;
#pragma unroll
        for (int k = 0; k < PER_THREAD / 16; k++) {
            const int kk = (k + rot) & 3;
            const uint4 v = *reinterpret_cast<const uint4*>(words + 4 * (c0 + kk));
            mask |= (unsigned long long)swar_eqmask16(v.x, v.y, v.z, v.w, 0x0a0a0a0au) << (16 * kk);
            tmask |= (unsigned long long)swar_eqmask16(v.x, v.y, v.z, v.w, 0x09090909u) << (16 * kk);
        }
        if (tid < DEC_LOOKBACK / 16) {
            const uint4 v = *reinterpret_cast<const uint4*>(words + 4 * tid);
            lb_masks[tid] = swar_eqmask16(v.x, v.y, v.z, v.w, 0x0a0a0a0au) | (swar_eqmask16(v.x, v.y, v.z, v.w, 0x09090909u) << 16);
        }
    }
    // bytes at or beyond n_eff are padding newlines of the last tile: they terminate nothing; bytes at or beyond
    // the tile's own TB bytes are not staged at all
    {
        int64_t lim = min(a.n_eff, tile_base + TB) - (tile_base + (int64_t)tid * PER_THREAD);
        if (lim < PER_THREAD) {
            const unsigned long long keep = lim <= 0 ? 0ULL : ((1ULL << lim) - 1ULL);
            mask &= keep; tmask &= keep;
        }
    }
    const int my_cnt = __popcll(mask) | (__popcll(tmask) << 16);   // both counts stay below 2^15
    int incl = my_cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; w++) woff += warp_tot[w];
    {
        const int excl = woff + incl - my_cnt;
        int idx = excl & 0xFFFF, tidx = excl >> 16;
        const int pos0 = DEC_LOOKBACK + tid * PER_THREAD;
        unsigned long long m = mask;
        while (m) {
            const int j = __ffsll((long long)m) - 1;
            m &= m - 1;
            if (idx < DEC_MAX_LINES) {
                nlpos[idx] = (uint16_t)(pos0 + j);
                nltabs[idx] = (uint16_t)(tidx + __popcll(tmask & ((1ULL << j) - 1ULL)));
            }
            idx++;
        }
        m = tmask;
        while (m) {
            const int j = __ffsll((long long)m) - 1;
            m &= m - 1;
            if (tidx < DEC_MAX_TABS) tabpos[tidx] = (uint16_t)(pos0 + j);
            tidx++;
        }
    }
    if (tid == 0) {
        // start of the first line: one past the last '\n' before the tile (or the file start)
        int begin = -1;
        for (int c = DEC_LOOKBACK / 16 - 1; c >= 0; c--) {
            unsigned nl = lb_masks[c] & 0xFFFFu;
            if (nl) { begin = c * 16 + (31 - __clz(nl)) + 1; break; }
        }
        if (tile_base == 0) begin = DEC_LOOKBACK;
        s_first_begin = begin;
    }
    __syncthreads();
    if (n_lines > DEC_MAX_LINES || s_first_begin < 0) {
        if (tid == 0) raise_error(a.status, CRATAC_ERR_PARSE, out_base);
        return;
    }

    // ---- one line per thread: tabs from the chunk masks, numbers and barcode a word at a time
    int loc_maxpos = 0, loc_maxneg = 0;
    unsigned long long ckey_cache = 0xFFFFFFFFFFFFFFFFULL;
    int32_t cid_cache = -1;
    for (int li = tid + skip; li < n_lines; li += DEC_THREADS) {
        int b = li == 0 ? s_first_begin : nlpos[li - 1] + 1;
        int e = nlpos[li];
        int64_t line_no = out_base + li;
        int t0 = 0, t1 = 0, t2 = 0, t3 = 0;   // scalars, not an indexed array: keeps them out of local memory
        int nt = 0;
        // Common case: the line lies in the tile region, has exactly 4 tabs and nothing for strip() to remove:
        // its tab positions are 4 consecutive entries of the tile's tab list.
        if (li > 0) {
            const int kt = nltabs[li - 1];
            if (nltabs[li] - kt == 4 && kt + 4 <= DEC_MAX_TABS && e - b >= 9 && byte_at(words, b) > ' ' && byte_at(words, e - 1) > ' ') {
                t0 = tabpos[kt]; t1 = tabpos[kt + 1]; t2 = tabpos[kt + 2]; t3 = tabpos[kt + 3];
                nt = 4;
            }
        }
        if (nt == 0) {
            // general path (first line of the tile, blanks to strip, wrong number of fields): byte by byte
            // strip(): leading / trailing blanks, '\r'
            while (b < e) { unsigned ch = byte_at(words, b); if (ch == ' ' || ch == '\t' || ch == '\r') b++; else break; }
            while (e > b) { unsigned ch = byte_at(words, e - 1); if (ch == ' ' || ch == '\t' || ch == '\r') e--; else break; }
            for (int p = b; p < e; p++) {
                if (byte_at(words, p) == '\t') {
                    if (nt == 0) t0 = p; else if (nt == 1) t1 = p; else if (nt == 2) t2 = p; else if (nt == 3) t3 = p;
                    nt++;
                }
            }
        }
        if (nt != 4) { raise_error(a.status, CRATAC_ERR_PARSE, line_no); continue; }
        // barcode key first: the load of its table slot (a random 16-byte access, usually the longest latency of
        // the line) is in flight while the numbers and the contig name are parsed
        const unsigned long long key = barcode_key_words(words, t2 + 1, t3 - t2 - 1);
        unsigned slot = (unsigned)(mix64(key)) & (unsigned)a.bc_mask;
        ulonglong2 sl = *reinterpret_cast<const ulonglong2*>(&a.bc_tab[slot]);   // key and first line in one 16-byte load
        int32_t v_start = 0, v_end = 0, v_dup = 0;
        bool ok = sp_parse_uint(words, t0 + 1, t1, &v_start) & sp_parse_uint(words, t1 + 1, t2, &v_end) & (t0 > b) & (t3 > t2 + 1);
        if (e - t3 == 2) { const unsigned dg = byte_at(words, t3 + 1) - 0x30u; v_dup = (int32_t)dg; ok &= dg <= 9u; }   // one digit nearly always
        else ok &= sp_parse_uint(words, t3 + 1, e, &v_dup);
        if (!ok) { raise_error(a.status, CRATAC_ERR_PARSE, line_no); continue; }
        // contig id: the name of the previous line of this thread, else the table
        const unsigned long long ch = swar_contig_key(words, b, t0 - b);
        int32_t cid = -1;
        if (ch == ckey_cache) {
            cid = cid_cache;
        } else {
            for (int s = (int)(swar_key_slot(ch) & (unsigned)a.contig_mask), probes = 0; probes <= a.contig_mask; s = (s + 1) & a.contig_mask, probes++) {
                const int4 raw = __ldg(reinterpret_cast<const int4*>(&a.contig_table[s]));   // small read-only table: L1
                const unsigned long long eh = (unsigned long long)(unsigned)raw.x | ((unsigned long long)(unsigned)raw.y << 32);
                if (!raw.w) break;
                if (eh == ch) { cid = raw.z; break; }
            }
            if (cid >= 0) { ckey_cache = ch; cid_cache = cid; }
        }
        if (cid < 0) { raise_error(a.status, CRATAC_ERR_CONTIG, line_no); continue; }
        // barcode slot
        int probes = 0;
        bool placed = false;
        unsigned long long seen_first = 0xFFFFFFFFFFFFFFFFULL;
        while (probes <= a.bc_mask) {
            // a key never changes once written, so a cached read that already shows our key is final;
            // anything else is re-checked by the CAS
            if (probes > 0) sl = *reinterpret_cast<const ulonglong2*>(&a.bc_tab[slot]);
            unsigned long long k = sl.x;
            seen_first = sl.y;
            if (k == BC_EMPTY) {
                unsigned long long old = atomicCAS(&a.bc_tab[slot].key, BC_EMPTY, key);
                if (old == BC_EMPTY) {
                    long long off = src0 + (t2 + 1);
                    a.bc_textoff[slot] = (off << 8) | (long long)min(t3 - t2 - 1, 255);
                    atomicAdd(reinterpret_cast<unsigned long long*>(&a.status[ST_N_BARCODES]), 1ULL);
                    placed = true;
                    break;
                }
                k = old;
            }
            if (k == key) { placed = true; break; }
            slot = (slot + 1) & (unsigned)a.bc_mask;
            probes++;
            if ((probes & 63) == 0 && bc_table_crowded(a.status, a.bc_mask)) break;
        }
        if (!placed) { raise_error(a.status, CRATAC_ERR_CAPACITY, line_no); continue; }
        if (a.line_cap > 0 && line_no >= a.line_cap) { raise_error(a.status, CRATAC_ERR_CAPACITY, -7); continue; }   // more lines than estimated
        // `seen_first` may be stale (cached) but only ever too large: the atomicMin decides
        if (seen_first > (unsigned long long)line_no) atomicMin(&a.bc_tab[slot].first, (unsigned long long)line_no);
        atomicAdd(&a.bc_nfrag[slot], 1);
        a.chrom[line_no] = cid;
        a.start[line_no] = v_start;
        a.end[line_no] = v_end;
        a.bc[line_no] = (int32_t)slot;
        a.dup[line_no] = v_dup;
        int d = v_end - v_start;
        loc_maxpos = max(loc_maxpos, d);
        loc_maxneg = max(loc_maxneg, -d);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        loc_maxpos = max(loc_maxpos, __shfl_xor_sync(0xffffffffu, loc_maxpos, d));
        loc_maxneg = max(loc_maxneg, __shfl_xor_sync(0xffffffffu, loc_maxneg, d));
    }
    if (lane == 0) {
        if (loc_maxpos > 0) atomicMax(reinterpret_cast<long long*>(&a.status[ST_MAX_POS_LEN]), (long long)loc_maxpos);
        if (loc_maxneg > 0) atomicMax(reinterpret_cast<long long*>(&a.status[ST_MAX_NEG_LEN]), (long long)loc_maxneg);
    }
}

// ------------------------------------------------------------------ single pass: k_parse_stream
// One read of the text.  Persistent CTAs take tiles from a ticket counter, so tiles start in index order; the next tile is
// already on its way into the second shared-memory buffer (TMA bulk copy) while the current one is parsed.  A tile is
// [t * TB, (t + 1) * TB) with SP_LB look-back bytes in front of it; a line belongs to the tile that holds its '\n'.
//
//  1. separator bits (bytes below 0x0b: tab, newline and, never in a valid file, 0x00-0x08) of every 16-byte chunk, three
//     integer instructions per word and one multiply to gather four flags.  Thread i reads its own CPT consecutive chunks
//     (the chunk order is rotated by lane so the 128-bit shared loads of a quarter-warp fall in different banks) and the
//     16-bit masks stay in its registers;
//  2. popcounts of those masks -> block scan -> ONE ordered list of separator positions in shared memory.
//     In a well-formed tile the separators of a line are five consecutive entries (four tabs, the newline) and the entry
//     before them is the previous line's newline.  The five bytes are checked per line (and the tabs of the unfinished last
//     line), so the structure is not assumed but verified; the look-back gives the number of tabs of the first line that lie
//     before the tile.
//  3. the tile's line count is published (decoupled look-back: aggregate first, inclusive prefix once known) and the last
//     warp resolves the tile's first output line while the other warps already parse.
//  4. groups of SP_THREADS lines: fields parsed a word at a time into registers, `bad` OR-reduced over the CTA, then the
//     barcode table insert and the five coalesced stores.
// Anything outside the common shape (a line longer than the look-back, a stray control byte, blanks for strip() to remove,
// a field that does not parse, an unknown contig) is not handled here: the tile -- or what is left of it -- is handed to
// k_parse_tiles (the general kernel above), which also raises the errors.  Nothing is committed for a line before its whole
// group has parsed, so the hand-over is exact.
constexpr int SP_THREADS = 256;
constexpr int SP_LB = 64;                       // look-back bytes = the chunks of thread 0
constexpr int SP_LOOK = 8;                     // look-back words per lane and round
constexpr int SP_PAD = 16;                      // slack before and after the staged bytes (word windows may over-read)
constexpr unsigned long long SP_FLAG_AGG = 1ULL << 62, SP_FLAG_INC = 2ULL << 62, SP_VALUE_MASK = (1ULL << 62) - 1ULL;

struct StreamArgs {
    DecodeArgs d;                 // text, tables, outputs, status (tile arrays unused)
    int tile_bytes;               // TB: multiple of 16, SP_LB + TB <= SP_THREADS * CPT * 16
    int64_t n_tiles;              // tiles of this launch: global tiles tile0 .. tile0 + n_tiles - 1
    int64_t tile0;
    unsigned long long* state;    // [n_tiles] look-back words of this launch (zeroed)
    unsigned long long* ticket;   // zeroed
    const int64_t* line_base;     // lines before this launch
    int64_t* line_base_out;       // written by the last tile: lines before the next launch (a different word: earlier tiles may still read line_base)
    IrrTile* irr;                 // declined ranges, appended
    unsigned long long* n_irr;
    int64_t irr_cap;
    unsigned long long* prof;     // null, or SPP_COUNT cycle counters
    int mode;                     // bit 0: the next tile's ticket is taken when the tile starts (no prefetch of its bytes);
                                  // bit 1: line numbers come from a count pass (d.tile_offsets): no publish, no look-back
};

// The look-back words carry their whole message in one aligned 64-bit value (flag + count), so they are read and written
// relaxed: no fence, no L1 invalidation per spin (an acquire load at gpu scope costs both, on every lane, every spin).
__device__ __forceinline__ unsigned long long sp_ld_state(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sp_st_state(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// per-phase cycle counters (thread 0 / the look-back warp's lane 0), only when StreamArgs::prof is set (tools/decode_bench.py)
enum { SPP_WAIT = 0, SPP_CLASSIFY, SPP_LIST, SPP_VERIFY, SPP_LOOKBACK, SPP_PARSE, SPP_COMMIT, SPP_TILES, SPP_ROUNDS, SPP_RETRIES, SPP_COUNT };
#define SP_PROF(slot)                                                                                 \
    do {                                                                                              \
        if (PROF && tid == 0) { const long long _t = clock64(); atomicAdd(&a.prof[slot], (unsigned long long)(_t - prof_t)); prof_t = _t; } \
    } while (0)

// separator flags (bit 7 of every byte below 0x0b), exact for all 256 byte values
__device__ __forceinline__ uint32_t sp_sepflags(uint32_t w) {
    const uint32_t x = (w & 0x7f7f7f7fu) + 0x75757575u;      // bit 7 set iff the low 7 bits are >= 0x0b
    return ~(x | w) & 0x80808080u;
}
// the four flags (bits 7, 15, 23, 31) gathered into the top nibble: no two partial products meet
__device__ __forceinline__ uint32_t sp_gather(uint32_t f) { return f * 0x00204081u; }
__device__ __forceinline__ uint32_t sp_sepmask16(uint4 v) {
    uint32_t m = sp_gather(sp_sepflags(v.w)) >> 28;
    m = __funnelshift_l(sp_gather(sp_sepflags(v.z)), m, 4);
    m = __funnelshift_l(sp_gather(sp_sepflags(v.y)), m, 4);
    m = __funnelshift_l(sp_gather(sp_sepflags(v.x)), m, 4);
    return m;
}

// decimal field of 1..10 digits ending at byte e (exclusive), 32-bit arithmetic only: the last 8 bytes are loaded as two
// words, bytes in front of the field are replaced by '0', all 8 are checked to be digits and folded 1 -> 2 -> 4 -> 8 digits
__device__ __forceinline__ bool sp_parse_uint(const uint32_t* words, int p, int e, int32_t* out) {
    const int len = e - p;
    if (len < 1 || len > 10) return false;
    uint32_t lo = swar_get4(words, e - 8), hi = swar_get4(words, e - 4);
    const int s = 8 * (8 - len);                                   // bits of (hi:lo) in front of the field (when len < 8)
    if (s > 0) {
        const uint32_t mlo = s >= 32 ? 0xFFFFFFFFu : ((1u << s) - 1u);
        const uint32_t mhi = s > 32 ? ((1u << (s - 32)) - 1u) : 0u;
        lo = (lo & ~mlo) | (0x30303030u & mlo);
        hi = (hi & ~mhi) | (0x30303030u & mhi);
    }
    const uint32_t xl = lo ^ 0x30303030u, xh = hi ^ 0x30303030u;   // digits -> 0..9
    uint32_t bad = ((xl + 0x76767676u) | xl | (xh + 0x76767676u) | xh) & 0x80808080u;
    uint32_t a = (xl * 2561u) >> 8 & 0x00FF00FFu;                  // 10 * d0 + d1 per byte pair
    a = (a * 6553601u) >> 16;                                      // 4 digits
    uint32_t b = (xh * 2561u) >> 8 & 0x00FF00FFu;
    b = (b * 6553601u) >> 16;
    uint32_t val = a * 10000u + b;
    if (len > 8) {                                                 // 9 or 10 digits: the leading one or two by hand
        const uint32_t w = swar_get4(words, p);
        uint32_t d0 = (w & 0xFFu) - 0x30u, head = d0;
        bad |= (d0 > 9u) ? 1u : 0u;
        if (len == 10) {
            const uint32_t d1 = ((w >> 8) & 0xFFu) - 0x30u;
            bad |= (d1 > 9u) ? 1u : 0u;
            head = d0 * 10u + d1;
            if (head > 21u || (head == 21u && val > 47483647u)) bad |= 1u;   // above INT32_MAX
        }
        val += head * 100000000u;
    }
    *out = (int32_t)val;
    return bad == 0u;
}

template <int CPT, bool PROF, int MINB>
__global__ void __launch_bounds__(SP_THREADS, MINB) k_parse_stream(StreamArgs a) {
    constexpr int REGION_MAX = SP_THREADS * CPT * 16;              // look-back + tile bytes
    constexpr int MAX_SEPS = REGION_MAX / 6;                       // 5 separators per line of >= 30 bytes (typical: 45)
    extern __shared__ __align__(128) unsigned char sp_smem[];
    uint32_t* bufw0 = reinterpret_cast<uint32_t*>(sp_smem);
    uint32_t* bufw1 = (a.mode & 4) ? bufw0 : bufw0 + (SP_PAD + REGION_MAX + SP_PAD) / 4;            // one-tile mode: a single buffer
    uint16_t* seplist = reinterpret_cast<uint16_t*>(bufw1 + (SP_PAD + REGION_MAX + SP_PAD) / 4);   // [MAX_SEPS]
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int warp_tot[SP_THREADS / 32];
    __shared__ int64_t s_tile[2];
    __shared__ int s_r, s_irregular, s_nlb;
    __shared__ long long s_excl;
    __shared__ int s_nl_parts[3];

    const DecodeArgs& d = a.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int TB = a.tile_bytes;
    const int region = SP_LB + TB;
    const int n_chunks = region >> 4;
    if (tid == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto tile_is_tma = [&](int64_t t) -> bool {     // t: tile of this launch
        const int64_t g = t + a.tile0;
        return d.use_tma && g * TB - SP_LB >= 0 && (g + 1) * TB <= d.nbytes;
    };
    auto issue = [&](int64_t t, int b) {            // thread 0
        uint32_t* w = b ? bufw1 : bufw0;
        mbar_expect_tx(&bar[b], (uint32_t)region);
        tma_load_1d(reinterpret_cast<unsigned char*>(w) + SP_PAD, d.text + (t + a.tile0) * TB - SP_LB, (uint32_t)region, &bar[b]);
    };
    auto next_tile = [&](int64_t prev) -> int64_t {   // thread 0; prev < 0: the CTA's first tile
        if (a.mode & 4) return prev < 0 ? (int64_t)blockIdx.x : a.n_tiles;          // one tile per CTA (grid = tiles): no prefetch, one buffer
        if (a.mode & 2) return prev < 0 ? (int64_t)blockIdx.x : prev + gridDim.x;   // no tile waits on another: static round-robin
        return (int64_t)atomicAdd(a.ticket, 1ULL);
    };
    if (tid == 0) {
        const int64_t t = next_tile(-1);
        s_tile[0] = t;
        if (t < a.n_tiles && tile_is_tma(t)) issue(t, 0);
    }
    __syncthreads();

    uint32_t phase0 = 0, phase1 = 0;
    int cur = 0;
    unsigned long long ckey_cache = 0xFFFFFFFFFFFFFFFFULL;   // last contig name seen by this thread and its id
    int32_t cid_cache = -1;
    int loc_maxpos = 0, loc_maxneg = 0;

    long long prof_t = PROF ? clock64() : 0;
    for (;;) {
        const int64_t t = s_tile[cur];
        if (t >= a.n_tiles) break;
        const int64_t gt = t + a.tile0;
        const int64_t tile_base = gt * TB;
        uint32_t* const bufw = cur ? bufw1 : bufw0;
        const uint32_t* const words = bufw + SP_PAD / 4;            // byte position 0 = first look-back byte
        unsigned char* const bytes = reinterpret_cast<unsigned char*>(bufw) + SP_PAD;
        if (tid == 0 && !(a.mode & 1)) {                            // next tile: ticket now, bytes in flight while this one is parsed
            const int64_t tn = next_tile(t);
            s_tile[cur ^ 1] = tn;
            if (tn < a.n_tiles && tile_is_tma(tn)) issue(tn, cur ^ 1);
        }
        if (tile_is_tma(t)) {
            const bool ok = mbar_wait(&bar[cur], cur ? phase1 : phase0);
            if (cur) phase1 ^= 1; else phase0 ^= 1;
            if (!ok) { if (tid == 0) raise_error(d.status, CRATAC_ERR_INTERNAL, -100 - gt); return; }
        } else {
            const int64_t src0 = tile_base - SP_LB;
            for (int i = tid; i < region; i += SP_THREADS) {
                const int64_t p = src0 + i;
                bytes[i] = (p >= 0 && p < d.nbytes) ? (unsigned char)d.text[p] : (unsigned char)'\n';   // before the file / virtual terminator
            }
            __syncthreads();
        }

        SP_PROF(SPP_WAIT);
        // ---- 1. separator masks of my CPT consecutive 16-byte chunks (64 contiguous bytes per 4 chunks: the 128-bit shared
        //         loads of a quarter-warp collide 4 ways, which costs less than a trip of the masks through shared memory
        //         and the barrier that goes with it)
        const int64_t lim = d.n_eff - (tile_base - SP_LB);           // region bytes at or beyond it are padding
        unsigned long long mk[CPT / 4];
        int my_cnt = 0;
#pragma unroll
        for (int q = 0; q < CPT / 4; q++) {
            unsigned long long m64 = 0;
#pragma unroll
            for (int k0 = 0; k0 < 4; k0++) {
                const int k = (k0 + ((lane >> 1) & 3)) & 3;          // rotated per lane: a quarter-warp's 128-bit loads cover all banks
                const int c = tid * CPT + 4 * q + k;
                uint32_t m = 0;
                if (c < n_chunks) {
                    const uint4 v = *reinterpret_cast<const uint4*>(words + 4 * c);
                    m = sp_sepmask16(v);
                    if (lim < region) {                              // last tile only
                        const int64_t keep = lim - 16 * (int64_t)c;
                        m = keep <= 0 ? 0u : (keep >= 16 ? m : (m & ((1u << keep) - 1u)));
                    }
                }
                m64 |= (unsigned long long)m << (16 * k);
            }
            mk[q] = m64;
            my_cnt += __popcll(m64);
        }
        SP_PROF(SPP_CLASSIFY);

        // ---- 2. ordered separator list
        int incl = my_cnt;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, dd);
            if (lane >= dd) incl += y;
        }
        if (lane == 31) warp_tot[warp] = incl;
        if (tid == 0) {
            // first line of the tile: tabs after the last newline of the look-back (thread 0's chunks ARE the look-back
            // when CPT == 4; with CPT == 8 the look-back is the first half of them)
            unsigned long long lb = mk[0];
            const int my_cnt_lb = __popcll(lb);                      // separators of the look-back (the first SP_LB bytes)
            int r = 0, found = 0, irregular = 0;
            while (lb) {
                const int j = 63 - __clzll((long long)lb);
                lb &= ~(1ULL << j);
                const unsigned ch = bytes[j];
                if (ch == '\n') { found = 1; break; }
                if (ch != '\t' || ++r > 4) { irregular = 1; break; }
            }
            if (!found) irregular = 1;                               // a line longer than the look-back
            s_r = r; s_irregular = irregular; s_nlb = my_cnt_lb;
        }
        __syncthreads();
        int woff = 0, s_all = 0;
#pragma unroll
        for (int w = 0; w < SP_THREADS / 32; w++) { const int x = warp_tot[w]; if (w < warp) woff += x; s_all += x; }
        {
            int idx = woff + incl - my_cnt;
#pragma unroll
            for (int q = 0; q < CPT / 4; q++) {
                unsigned long long m = mk[q];
                const int pos0 = (tid * CPT + 4 * q) * 16;
                while (m) {
                    const int j = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    if (idx < MAX_SEPS) seplist[idx] = (uint16_t)(pos0 + j);
                    idx++;
                }
            }
        }
        __syncthreads();
        SP_PROF(SPP_LIST);
        const int n_lb = s_nlb;
        const int first = n_lb - s_r;                                // list index of the first line's first separator
        const int q_seps = s_all - first;
        const int n_lines = q_seps / 5;
        const int rem = q_seps - 5 * n_lines;
        bool bad = s_irregular != 0 || s_all > MAX_SEPS;

        // ---- structure check: tab tab tab tab newline for every line, tabs only after the last newline
        if (!bad) {
            for (int li = tid; li < n_lines; li += SP_THREADS) {
                const int k = first + 5 * li;
                const unsigned b0 = bytes[seplist[k]], b1 = bytes[seplist[k + 1]], b2 = bytes[seplist[k + 2]], b3 = bytes[seplist[k + 3]], b4 = bytes[seplist[k + 4]];
                if ((b0 | (b1 << 8) | (b2 << 16) | (b3 << 24)) != 0x09090909u || b4 != '\n') bad = true;
            }
            if (tid == SP_THREADS - 1)
                for (int k = 0; k < rem; k++) if (bytes[seplist[first + 5 * n_lines + k]] != '\t') bad = true;
        }
        const bool irregular = __syncthreads_or(bad ? 1 : 0) != 0;
        SP_PROF(SPP_VERIFY);

        // ---- exact line count of a declined tile: newlines of the tile region, per half (k_parse_tiles takes <= DEC_TILE bytes)
        int tile_lines = n_lines;
        const int half = ((TB / 2) + 15) & ~15;                      // first sub-range [0, half), second [half, TB)
        if (irregular) {
            int c0 = 0, c1 = 0;
            for (int i = tid; i < TB; i += SP_THREADS) {
                const int64_t p = tile_base + i;
                const bool nl = p < d.n_eff && bytes[SP_LB + i] == '\n';
                if (nl) { if (i < half) c0++; else c1++; }
            }
#pragma unroll
            for (int dd = 16; dd; dd >>= 1) { c0 += __shfl_xor_sync(0xffffffffu, c0, dd); c1 += __shfl_xor_sync(0xffffffffu, c1, dd); }
            if (tid == 0) { s_nl_parts[0] = 0; s_nl_parts[1] = 0; }
            __syncthreads();
            if (lane == 0) { atomicAdd(&s_nl_parts[0], c0); atomicAdd(&s_nl_parts[1], c1); }
            __syncthreads();
            tile_lines = s_nl_parts[0] + s_nl_parts[1];
        }

        // ---- 3. publish the count, resolve the first output line (last warp), meanwhile parse
        if (a.mode & 2) {
            if (tid == 0) s_excl = d.tile_offsets[gt];
        } else if (warp == SP_THREADS / 32 - 1) {
            if (lane == 0) sp_st_state(&a.state[t], SP_FLAG_AGG | (unsigned long long)tile_lines);
            long long excl = 0;
            const long long lb_t0 = PROF ? clock64() : 0;
            if (t == 0) {
                excl = *a.line_base;
            } else {
                // Every tile of a wave looks back at the same time and none of them has its inclusive prefix yet, so the walk
                // goes back through the whole wave (hundreds of tiles): each lane takes SP_LOOK consecutive words per round
                // (independent loads, one memory round trip for 32 * SP_LOOK tiles).  Lane 0 holds the nearest words.
                int64_t look = t - 1;
                for (;;) {
                    long long lane_sum = 0;
                    bool lane_inc = false, lane_stuck = false;
                    unsigned long long v[SP_LOOK];
                    const int64_t top = look - (int64_t)lane * SP_LOOK;
#pragma unroll
                    for (int j = 0; j < SP_LOOK; j++) {
                        const int64_t idx = top - j;
                        v[j] = idx >= 0 ? sp_ld_state(&a.state[idx])
                                        : (idx == -1 ? (SP_FLAG_INC | (unsigned long long)*a.line_base) : SP_FLAG_AGG);   // virtual tile before the launch
                    }
#pragma unroll
                    for (int j = 0; j < SP_LOOK; j++) {
                        const int64_t idx = top - j;
                        int spins = 0;
                        while ((v[j] >> 62) == 0 && ++spins < (1 << 22)) { __nanosleep(200); v[j] = sp_ld_state(&a.state[idx]); }
                        if (PROF && spins) atomicAdd(&a.prof[SPP_RETRIES], (unsigned long long)spins);
                        if ((v[j] >> 62) == 0) lane_stuck = true;
                        if (!lane_inc) { lane_sum += (long long)(v[j] & SP_VALUE_MASK); lane_inc = (v[j] >> 62) == 2; }
                    }
                    if (PROF && lane == 0) atomicAdd(&a.prof[SPP_ROUNDS], 1ULL);
                    const unsigned has_inc = __ballot_sync(0xffffffffu, lane_inc);
                    const unsigned stuck = __ballot_sync(0xffffffffu, lane_stuck);
                    if (stuck) { if (lane == 0) raise_error(d.status, CRATAC_ERR_INTERNAL, -200 - gt); excl = -1; break; }
                    long long contrib = lane_sum;
                    if (has_inc) {
                        const int fp = __ffs(has_inc) - 1;           // nearest lane that met an inclusive prefix
                        if (lane > fp) contrib = 0;
                    }
#pragma unroll
                    for (int dd = 16; dd; dd >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, dd);
                    excl += contrib;
                    if (has_inc) break;
                    look -= 32 * SP_LOOK;
                }
            }
            if (PROF && lane == 0) atomicAdd(&a.prof[SPP_LOOKBACK], (unsigned long long)(clock64() - lb_t0));
            if (lane == 0) {
                if (excl >= 0) sp_st_state(&a.state[t], SP_FLAG_INC | (unsigned long long)(excl + tile_lines));
                s_excl = excl;
                if (excl >= 0 && t == a.n_tiles - 1) *a.line_base_out = excl + tile_lines;
            }
        }

        if (irregular) {
            __syncthreads();                                         // s_excl
            if (tid == 0 && s_excl >= 0) {
                const int n_parts = TB > DEC_TILE ? 2 : 1;
                const unsigned long long at = atomicAdd(a.n_irr, (unsigned long long)n_parts);
                if ((int64_t)(at + n_parts) <= a.irr_cap) {
                    if (n_parts == 1) {
                        a.irr[at] = IrrTile{tile_base, s_excl, TB, tile_lines, 0, 0};
                    } else {
                        a.irr[at] = IrrTile{tile_base, s_excl, half, s_nl_parts[0], 0, 0};
                        a.irr[at + 1] = IrrTile{tile_base + half, s_excl + s_nl_parts[0], TB - half, s_nl_parts[1], 0, 0};
                    }
                } else {
                    raise_error(d.status, CRATAC_ERR_INTERNAL, -300 - gt);
                }
            }
            if (s_excl < 0) return;
            cur ^= 1;
            if ((a.mode & 1) && tid == 0) {
                const int64_t tn = next_tile(t);
                s_tile[cur] = tn;
                if (tn < a.n_tiles && tile_is_tma(tn)) issue(tn, cur);
            }
            __syncthreads();
            continue;
        }

        // ---- 4. groups of SP_THREADS lines
        bool published_seen = false;
        for (int base = 0; base < n_lines; base += SP_THREADS) {
            const int li = base + tid;
            bool ok = true;
            int32_t v_start = 0, v_end = 0, v_dup = 0, cid = -1;
            unsigned long long key = 0;
            int bc_pos = 0, bc_len = 0;
            unsigned pre_slot = 0;
            ulonglong2 pre = make_ulonglong2(0ULL, 0ULL);
            if (li < n_lines) {
                const int k = first + 5 * li;
                const int b = seplist[k - 1] + 1;
                const int t0 = seplist[k], t1 = seplist[k + 1], t2 = seplist[k + 2], t3 = seplist[k + 3], e = seplist[k + 4];
                bc_pos = t2 + 1; bc_len = t3 - t2 - 1;
                key = barcode_key_words(words, bc_pos, bc_len);
                // the barcode's table slot (a random 16-byte access) is requested now: it arrives while the numbers parse
                pre_slot = (unsigned)(mix64(key)) & (unsigned)d.bc_mask;
                pre = *reinterpret_cast<const ulonglong2*>(&d.bc_tab[pre_slot]);
                ok = sp_parse_uint(words, t0 + 1, t1, &v_start) & sp_parse_uint(words, t1 + 1, t2, &v_end) & (t0 > b) & (bc_len > 0);
                {                                                    // the duplicate count is a single digit nearly always
                    const int dl = e - t3 - 1;
                    if (dl == 1) { const unsigned dg = (unsigned)bytes[t3 + 1] - 0x30u; v_dup = (int32_t)dg; ok &= dg <= 9u; }
                    else ok &= sp_parse_uint(words, t3 + 1, e, &v_dup);
                }
                const unsigned long long ch = swar_contig_key(words, b, t0 - b);
                if (ch == ckey_cache) {
                    cid = cid_cache;
                } else {
                    for (int s = (int)(swar_key_slot(ch) & (unsigned)d.contig_mask), probes = 0; probes <= d.contig_mask; s = (s + 1) & d.contig_mask, probes++) {
                        const int4 raw = __ldg(reinterpret_cast<const int4*>(&d.contig_table[s]));
                        const unsigned long long eh = (unsigned long long)(unsigned)raw.x | ((unsigned long long)(unsigned)raw.y << 32);
                        if (!raw.w) break;
                        if (eh == ch) { cid = raw.z; break; }
                    }
                    if (cid >= 0) { ckey_cache = ch; cid_cache = cid; }
                }
                ok &= cid >= 0;
            }
            const bool group_bad = __syncthreads_or(ok ? 0 : 1) != 0;  // also orders s_excl before its first use
            SP_PROF(SPP_PARSE);
            if (!published_seen) { published_seen = true; if (s_excl < 0) return; }
            const long long out0 = s_excl;
            if (group_bad) {
                // the lines before this group are committed; the general kernel takes the rest (and raises what is an error)
                if (tid == 0) {
                    const int n_parts = TB > DEC_TILE ? 2 : 1;
                    const unsigned long long at = atomicAdd(a.n_irr, (unsigned long long)n_parts);
                    if ((int64_t)(at + n_parts) <= a.irr_cap) {
                        if (n_parts == 1) {
                            a.irr[at] = IrrTile{tile_base, out0, TB, n_lines, base, 0};
                        } else {
                            // newlines of the first half: list entries below SP_LB + half that are line ends
                            int n0 = 0;
                            for (int q = 0; q < n_lines; q++) if (seplist[first + 5 * q + 4] < SP_LB + half) n0++; else break;
                            a.irr[at] = IrrTile{tile_base, out0, half, n0, min(base, n0), 0};
                            a.irr[at + 1] = IrrTile{tile_base + half, out0 + n0, TB - half, n_lines - n0, max(base - n0, 0), 0};
                        }
                    } else {
                        raise_error(d.status, CRATAC_ERR_INTERNAL, -300 - gt);
                    }
                }
                break;
            }
            if (li < n_lines) {
                const long long line_no = out0 + li;
                if (d.line_cap > 0 && line_no >= d.line_cap) {
                    raise_error(d.status, CRATAC_ERR_CAPACITY, -7);  // more lines than estimated
                } else {
                    unsigned slot = pre_slot;
                    int probes = 0;
                    bool placed = false;
                    unsigned long long seen_first = 0xFFFFFFFFFFFFFFFFULL;
                    while (probes <= d.bc_mask) {
                        // a key never changes once written, so the early read is final when it shows our key; anything else is re-checked
                        const ulonglong2 sl = probes == 0 ? pre : *reinterpret_cast<const ulonglong2*>(&d.bc_tab[slot]);
                        unsigned long long kk = sl.x;
                        seen_first = sl.y;
                        if (kk == BC_EMPTY) {
                            const unsigned long long old = atomicCAS(&d.bc_tab[slot].key, BC_EMPTY, key);
                            if (old == BC_EMPTY) {
                                const long long off = tile_base - SP_LB + bc_pos;
                                d.bc_textoff[slot] = (off << 8) | (long long)min(bc_len, 255);
                                atomicAdd(reinterpret_cast<unsigned long long*>(&d.status[ST_N_BARCODES]), 1ULL);
                                placed = true;
                                break;
                            }
                            kk = old;
                        }
                        if (kk == key) { placed = true; break; }
                        slot = (slot + 1) & (unsigned)d.bc_mask;
                        probes++;
                        if ((probes & 63) == 0 && bc_table_crowded(d.status, d.bc_mask)) break;
                    }
                    if (!placed) {
                        raise_error(d.status, CRATAC_ERR_CAPACITY, line_no);
                    } else {
                        if (seen_first > (unsigned long long)line_no) atomicMin(&d.bc_tab[slot].first, (unsigned long long)line_no);
                        atomicAdd(&d.bc_nfrag[slot], 1);
                        d.chrom[line_no] = cid;
                        d.start[line_no] = v_start;
                        d.end[line_no] = v_end;
                        d.bc[line_no] = (int32_t)slot;
                        d.dup[line_no] = v_dup;
                        const int dl = v_end - v_start;
                        loc_maxpos = max(loc_maxpos, dl);
                        loc_maxneg = max(loc_maxneg, -dl);
                    }
                }
            }
        }
        if (n_lines == 0) { __syncthreads(); if (s_excl < 0) return; }
        cur ^= 1;
        if ((a.mode & 1) && tid == 0) {                              // ticket taken when the tile starts: tiles start in index order
            const int64_t tn = next_tile(t);
            s_tile[cur] = tn;
            if (tn < a.n_tiles && tile_is_tma(tn)) issue(tn, cur);
        }
        __syncthreads();                                             // buffers, lists and s_* are reused by the next tile
        SP_PROF(SPP_COMMIT);
        if (PROF && tid == 0) atomicAdd(&a.prof[SPP_TILES], 1ULL);
    }
#pragma unroll
    for (int dd = 16; dd; dd >>= 1) {
        loc_maxpos = max(loc_maxpos, __shfl_xor_sync(0xffffffffu, loc_maxpos, dd));
        loc_maxneg = max(loc_maxneg, __shfl_xor_sync(0xffffffffu, loc_maxneg, dd));
    }
    if (lane == 0) {
        if (loc_maxpos > 0) atomicMax(reinterpret_cast<long long*>(&d.status[ST_MAX_POS_LEN]), (long long)loc_maxpos);
        if (loc_maxneg > 0) atomicMax(reinterpret_cast<long long*>(&d.status[ST_MAX_NEG_LEN]), (long long)loc_maxneg);
    }
}

// ------------------------------------------------------------------ fragment bookkeeping
// chrom[] must be non-decreasing and start[] non-decreasing inside a contig.
__global__ void __launch_bounds__(256) k_check_sorted(const int32_t* __restrict__ chrom, const int32_t* __restrict__ start, int64_t n, int64_t* status) {
    // grid-stride over groups of 4 fragments (128-bit loads; the arrays come from cudaMalloc, so group starts are aligned)
    const int64_t n4 = (n + 3) >> 2;
    int64_t bad = -1;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = g << 2;
        int c[5], s[5];
        if (i0 + 4 <= n) {
            const int4 cv = *reinterpret_cast<const int4*>(chrom + i0), sv = *reinterpret_cast<const int4*>(start + i0);
            c[1] = cv.x; c[2] = cv.y; c[3] = cv.z; c[4] = cv.w; s[1] = sv.x; s[2] = sv.y; s[3] = sv.z; s[4] = sv.w;
        } else {
            for (int k = 0; k < 4; k++) { const int64_t i = i0 + k < n ? i0 + k : n - 1; c[k + 1] = chrom[i]; s[k + 1] = start[i]; }
        }
        c[0] = i0 > 0 ? chrom[i0 - 1] : c[1];
        s[0] = i0 > 0 ? start[i0 - 1] : s[1];
#pragma unroll
        for (int k = 1; k <= 4; k++)
            if (i0 + k - 1 < n && (c[k] < c[k - 1] || (c[k] == c[k - 1] && s[k] < s[k - 1])) && bad < 0) bad = i0 + k - 1;
    }
    if (bad >= 0) raise_error(status, CRATAC_ERR_UNSORTED, bad);
}

__global__ void k_contig_frag_offsets(const int32_t* __restrict__ chrom, int64_t n, int n_contigs, int64_t* __restrict__ off) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_contigs) return;
    int64_t lo = 0, hi = n;   // first i with chrom[i] >= c
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (chrom[mid] < c) lo = mid + 1; else hi = mid;
    }
    off[c] = lo;
}

// pos_index[pidx_off[c] + k] = first fragment i of contig c with start[i] >= k * 1024  (absolute i)
__global__ void k_pos_index(const int32_t* __restrict__ start, const int64_t* __restrict__ frag_off,
                            const int64_t* __restrict__ pidx_off, int n_contigs, int64_t total, int64_t* __restrict__ pos_index) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    int lo_c = 0, hi_c = n_contigs;   // contig of entry g: last c with pidx_off[c] <= g
    while (hi_c - lo_c > 1) {
        int mid = (lo_c + hi_c) >> 1;
        if (pidx_off[mid] <= g) lo_c = mid; else hi_c = mid;
    }
    int c = lo_c;
    int64_t target = (g - pidx_off[c]) << POS_INDEX_SHIFT;
    int64_t lo = frag_off[c], hi = frag_off[c + 1];
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)start[mid] < target) lo = mid + 1; else hi = mid;
    }
    pos_index[g] = lo;
}

__global__ void k_max_len(const int32_t* __restrict__ start, const int32_t* __restrict__ end, int64_t n, int64_t* status) {
    int mp = 0, mn = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int d = end[i] - start[i];
        mp = max(mp, d); mn = max(mn, -d);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        mp = max(mp, __shfl_xor_sync(0xffffffffu, mp, d));
        mn = max(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    }
    if ((threadIdx.x & 31) == 0) {
        if (mp > 0) atomicMax(reinterpret_cast<long long*>(&status[ST_MAX_POS_LEN]), (long long)mp);
        if (mn > 0) atomicMax(reinterpret_cast<long long*>(&status[ST_MAX_NEG_LEN]), (long long)mn);
    }
}

// ------------------------------------------------------------------ barcode table -> dense ids
__global__ void k_table_compact(const BcSlot* __restrict__ tab, const int32_t* __restrict__ nfrag, int32_t* __restrict__ dense_nfrag,
                                const long long* __restrict__ textoff, int cap, int32_t* __restrict__ slot_to_dense,
                                unsigned long long* __restrict__ dense_first, long long* __restrict__ dense_textoff,
                                unsigned long long* __restrict__ dense_bckey, unsigned long long* counter) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap) return;
    const BcSlot e = tab[s];
    if (e.key == BC_EMPTY) { slot_to_dense[s] = -1; return; }
    unsigned long long d = atomicAdd(counter, 1ULL);
    slot_to_dense[s] = (int32_t)d;
    dense_first[d] = e.first;
    dense_textoff[d] = textoff[s];
    dense_bckey[d] = e.key;
    dense_nfrag[d] = nfrag[s];
}

__global__ void k_gather_barcode_text(const char* __restrict__ text, const long long* __restrict__ dense_textoff, int64_t n,
                                      int stride, char* __restrict__ out) {
    int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    long long packed = dense_textoff[d];
    long long off = packed >> 8;
    int len = (int)(packed & 255);
    char* o = out + d * stride;
    for (int i = 0; i < stride; i++) o[i] = (i < len) ? text[off + i] : 0;
}

// CPython-2.7 string_hash of every barcode (Objects/stringobject.c), straight from the text
__global__ void k_barcode_py2hash(const char* __restrict__ text, const long long* __restrict__ dense_textoff, int64_t n, long long* __restrict__ out) {
    int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    long long packed = dense_textoff[d];
    const unsigned char* p = reinterpret_cast<const unsigned char*>(text) + (packed >> 8);
    int len = (int)(packed & 255);
    unsigned long long x = 0;
    if (len > 0) {
        x = (unsigned long long)p[0] << 7;
        for (int i = 0; i < len; i++) x = (1000003ULL * x) ^ (unsigned long long)p[i];
        x ^= (unsigned long long)len;
        if (x == 0xFFFFFFFFFFFFFFFFULL) x = 0xFFFFFFFFFFFFFFFEULL;
    }
    out[d] = (long long)x;
}

// first-appearance rank without a sort: the first lines of the barcodes are distinct fragment indices, so a
// bitmap over the fragments + a prefix popcount ranks them.
__global__ void k_mark_first(const unsigned long long* __restrict__ dense_first, int64_t n, unsigned int* __restrict__ bitmap) {
    int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    unsigned long long f = dense_first[d];
    atomicOr(&bitmap[f >> 5], 1u << (f & 31));
}
__global__ void k_popc_words(const unsigned int* __restrict__ bitmap, int64_t n_words, int32_t* __restrict__ cnt) {
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < n_words) cnt[w] = __popc(bitmap[w]);
}
__global__ void k_rank_first(const unsigned long long* __restrict__ dense_first, const long long* __restrict__ dense_hash, int64_t n,
                             const unsigned int* __restrict__ bitmap, const int64_t* __restrict__ word_prefix,
                             int32_t* __restrict__ order_f, long long* __restrict__ hash_f) {
    int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    unsigned long long f = dense_first[d];
    int64_t r = word_prefix[f >> 5] + __popc(bitmap[f >> 5] & ((1u << (f & 31)) - 1u));
    order_f[r] = (int32_t)d;           // the r-th barcode to appear in the file is dense id d
    hash_f[r] = dense_hash[d];
}
// col_to_dense[j] = dense id of the barcode whose first-appearance rank is rank_of_col[j]
__global__ void k_compose_order(const int32_t* __restrict__ rank_of_col, const int32_t* __restrict__ order_first, int64_t n, int32_t* __restrict__ col_to_dense) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) col_to_dense[j] = order_first[rank_of_col[j]];
}

__global__ void k_invert_order(const int32_t* __restrict__ col_to_dense, int64_t n, int32_t* __restrict__ dense_to_col) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dense_to_col[col_to_dense[j]] = (int32_t)j;
}

// ------------------------------------------------------------------ host side
static int build_contig_table(Ctx* c) {
    int cap = 16;
    while (cap < 4 * c->n_contigs) cap <<= 1;
    std::vector<ContigEntry> tab(cap);
    memset(tab.data(), 0, sizeof(ContigEntry) * cap);
    for (int i = 0; i < c->n_contigs; i++) {
        const std::string& nm = c->contig_names[i];
        std::vector<uint32_t> padded(nm.size() / 4 + 4, 0u);
        memcpy(padded.data(), nm.data(), nm.size());
        unsigned long long h = swar_contig_key(padded.data(), 0, (int)nm.size());
        int s = (int)(swar_key_slot(h) & (unsigned)(cap - 1));
        while (tab[s].used) {
            if (tab[s].hash == h) { set_error("contig names '%s' and '%s' collide", nm.c_str(), c->contig_names[tab[s].id].c_str()); return CRATAC_ERR_ARG; }
            s = (s + 1) & (cap - 1);
        }
        tab[s].hash = h; tab[s].id = i; tab[s].used = 1;
    }
    CRATAC_TRY(c->d_contig_table.reserve(sizeof(ContigEntry) * cap));
    CRATAC_CUDA(cudaMemcpyAsync(c->d_contig_table.p, tab.data(), sizeof(ContigEntry) * cap, cudaMemcpyHostToDevice, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    c->contig_table_cap = cap;
    return CRATAC_OK;
}

// device half of the barcode dictionary, queued right after the parse kernel: hash-table slots -> dense ids
// (slot_to_dense, needed by the overlap kernels), CPython-2.7 hashes, first-appearance order; the small
// results travel to pinned host memory asynchronously and an event marks their arrival.
static int barcode_prepare_device(Ctx* c) {
    if (c->bc_thread.joinable()) c->bc_thread.join();       // a column order still being simulated for the previous fragments
    int64_t* st = c->d_status.as<int64_t>();
    const int64_t nb = c->n_barcodes, n = c->n_fragments;
    size_t nn = (size_t)std::max<int64_t>(nb, 1);
    CRATAC_TRY(c->d_dense_first.reserve(8 * nn)); CRATAC_TRY(c->d_dense_textoff.reserve(8 * nn));
    CRATAC_TRY(c->d_dense_key.reserve(8 * nn));                 // py2 hashes by dense id
    CRATAC_TRY(c->d_dense_bckey.reserve(8 * nn));               // 64-bit barcode keys by dense id
    CRATAC_TRY(c->d_dense_nfrag.reserve(4 * nn));               // fragments per dense barcode
    c->dense_nfrag_valid = true;
    CRATAC_TRY(c->d_dense_to_col.reserve(4 * nn)); CRATAC_TRY(c->d_col_to_dense.reserve(4 * nn));
    const int64_t n_words = (n + 31) / 32 + 1;
    CRATAC_TRY(c->d_bc_bitmap.reserve(16 * (size_t)n_words + 12 * nn + 64));
    unsigned int* bitmap = c->d_bc_bitmap.as<unsigned int>();
    int32_t* word_cnt = reinterpret_cast<int32_t*>(bitmap + n_words);
    int64_t* word_prefix = reinterpret_cast<int64_t*>(word_cnt + n_words);   // 2 * n_words int32 => 8-byte aligned
    long long* hash_f = reinterpret_cast<long long*>(word_prefix + n_words);
    int32_t* order_f = reinterpret_cast<int32_t*>(hash_f + nn);
    c->d_hash_first = hash_f;
    c->d_order_first = order_f;
    cudaStream_t s = c->stream;
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_TICKET], 0, sizeof(int64_t), s));
    c->stage_begin(SG_BARCODE_TABLE);
    k_table_compact<<<(c->bc_table_cap + 255) / 256, 256, 0, s>>>(c->d_bc_keys.as<BcSlot>(), c->d_bc_nfrag.as<int32_t>(), c->d_dense_nfrag.as<int32_t>(),
                                                               c->d_bc_textoff.as<long long>(), c->bc_table_cap, c->d_slot_to_dense.as<int32_t>(),
                                                               c->d_dense_first.as<unsigned long long>(), c->d_dense_textoff.as<long long>(),
                                                               c->d_dense_bckey.as<unsigned long long>(), reinterpret_cast<unsigned long long*>(&st[ST_TICKET]));
    c->launches++;
    if (nb > 0) {
        unsigned g = (unsigned)((nb + 255) / 256);
        k_barcode_py2hash<<<g, 256, 0, s>>>(c->d_text, c->d_dense_textoff.as<long long>(), nb, c->d_dense_key.as<long long>());
        CRATAC_CUDA(cudaMemsetAsync(bitmap, 0, 4 * (size_t)n_words, s));
        k_mark_first<<<g, 256, 0, s>>>(c->d_dense_first.as<unsigned long long>(), nb, bitmap);
        k_popc_words<<<(unsigned)((n_words + 255) / 256), 256, 0, s>>>(bitmap, n_words, word_cnt);
        c->launches += 3;
        CRATAC_CUDA(cudaGetLastError());
        CRATAC_TRY(exclusive_scan_i32_to_i64(c, word_cnt, word_prefix, n_words, nullptr));
        k_rank_first<<<g, 256, 0, s>>>(c->d_dense_first.as<unsigned long long>(), c->d_dense_key.as<long long>(), nb, bitmap, word_prefix, order_f, hash_f);
        c->launches++;
    }
    c->stage_end(SG_BARCODE_TABLE);
    CRATAC_CUDA(cudaGetLastError());
    // pinned host mirrors
    if (c->h_bc_cap < nn) {
        if (c->h_bc_hash) cudaFreeHost(c->h_bc_hash);
        if (c->h_bc_order) cudaFreeHost(c->h_bc_order);
        if (c->h_c2d_pin) cudaFreeHost(c->h_c2d_pin);
        size_t cap = nn + nn / 2 + 1024;
        if (cudaMallocHost(&c->h_bc_hash, 8 * cap) != cudaSuccess || cudaMallocHost(&c->h_bc_order, 4 * cap) != cudaSuccess ||
            cudaMallocHost(&c->h_c2d_pin, 4 * cap) != cudaSuccess) {
            set_error("cudaMallocHost failed for the barcode mirrors");
            return CRATAC_ERR_CUDA;
        }
        c->h_bc_cap = cap;
    }
    if (nb > 0) {
        CRATAC_CUDA(cudaMemcpyAsync(c->h_bc_hash, hash_f, 8 * (size_t)nb, cudaMemcpyDeviceToHost, s));
        CRATAC_CUDA(cudaMemcpyAsync(c->h_bc_order, order_f, 4 * (size_t)nb, cudaMemcpyDeviceToHost, s));
    }
    if (!c->bc_event) CRATAC_CUDA(cudaEventCreateWithFlags(&c->bc_event, cudaEventDisableTiming));
    CRATAC_CUDA(cudaEventRecord(c->bc_event, s));
    c->bc_ready = false;
    c->bc_strings_ready = false;
    return CRATAC_OK;
}

// ---- host side of k_parse_stream
static int sp_cpt(const Ctx* c) { return c->decode_kernel == 2 ? 8 : 4; }
static int sp_tile_bytes(const Ctx* c) { return SP_THREADS * sp_cpt(c) * 16 - SP_LB; }
static size_t sp_smem_bytes(int cpt) {
    const size_t region_max = (size_t)SP_THREADS * cpt * 16;
    return 2 * (SP_PAD + region_max + SP_PAD) + 2 * (region_max / 6) + 64;
}

// queue one launch over the tiles [tile0, tile0 + n_tiles) of the text described by `base`
static int sp_launch(Ctx* c, const DecodeArgs& base, int64_t tile0, int64_t n_tiles, const int64_t* line_base, int64_t* line_base_out) {
    const int cpt = sp_cpt(c);
    StreamArgs a;
    a.d = base;
    a.tile_bytes = sp_tile_bytes(c);
    a.n_tiles = n_tiles; a.tile0 = tile0;
    a.state = c->d_sp_state.as<unsigned long long>() + tile0;
    a.ticket = c->d_sp_scalars.as<unsigned long long>();
    a.n_irr = c->d_sp_scalars.as<unsigned long long>() + 1;
    a.irr = c->d_sp_irr.as<IrrTile>();
    a.irr_cap = c->sp_irr_cap;
    a.line_base = line_base; a.line_base_out = line_base_out;
    a.prof = c->decode_profile ? c->d_sp_scalars.as<unsigned long long>() + 8 : nullptr;
    a.mode = c->decode_mode;
    if (a.mode & 2) {
        // line numbers from a count pass over the same tiles (A/B of the look-back; one launch only: offsets are absolute)
        CRATAC_TRY(c->d_tile_counts.reserve(sizeof(int32_t) * (size_t)n_tiles));
        CRATAC_TRY(c->d_tile_offsets.reserve(sizeof(int64_t) * (size_t)(n_tiles + 1)));
        c->stage_begin(SG_DECODE_COUNT);
        k_count_newlines<<<(unsigned)n_tiles, 256, 0, c->stream>>>(base.text, base.nbytes, base.n_eff, c->d_tile_counts.as<int32_t>(), tile0, a.tile_bytes);
        c->stage_end(SG_DECODE_COUNT);
        CRATAC_TRY(exclusive_scan_i32_to_i64(c, c->d_tile_counts.as<int32_t>(), c->d_tile_offsets.as<int64_t>(), n_tiles, line_base_out));
        a.d.tile_offsets = c->d_tile_offsets.as<int64_t>();
        c->launches++;
    }
    CRATAC_CUDA(cudaMemsetAsync(a.state, 0, 8 * (size_t)n_tiles, c->stream));
    CRATAC_CUDA(cudaMemsetAsync(a.ticket, 0, 8, c->stream));
    const bool one_tile = (a.mode & 4) != 0;
    const size_t smem = sp_smem_bytes(cpt) - (one_tile ? (SP_PAD + (size_t)SP_THREADS * cpt * 16 + SP_PAD) : 0);
    void (*kern)(StreamArgs) = one_tile ? (cpt == 8 ? k_parse_stream<8, false, 3> : k_parse_stream<4, false, 6>)
                                        : cpt == 8 ? (a.prof ? k_parse_stream<8, true, 2> : k_parse_stream<8, false, 2>)
                                                   : (a.prof ? k_parse_stream<4, true, 4> : k_parse_stream<4, false, 5>);
    CRATAC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CRATAC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SP_THREADS, smem));
    if (per_sm < 1) { set_error("k_parse_stream does not fit on an SM"); return CRATAC_ERR_INTERNAL; }
    if (c->sp_ctas_per_sm > 0) per_sm = std::min(per_sm, c->sp_ctas_per_sm);
    const unsigned grid = one_tile ? (unsigned)n_tiles : (unsigned)std::min<int64_t>(n_tiles, (int64_t)NUM_SMS * per_sm);   // persistent: every CTA resident, tiles wait on earlier tiles
    kern<<<grid, SP_THREADS, smem, c->stream>>>(a);
    c->launches++;
    CRATAC_CUDA(cudaGetLastError());
    return CRATAC_OK;
}

// buffers of the single-pass decode for a text of n_eff bytes
static int sp_reserve(Ctx* c, int64_t n_eff, int64_t* n_tiles_out) {
    const int TB = sp_tile_bytes(c);
    int64_t n_tiles = (n_eff + TB - 1) / TB;
    if (n_tiles == 0) n_tiles = 1;
    CRATAC_TRY(c->d_sp_state.reserve(8 * (size_t)n_tiles));
    CRATAC_TRY(c->d_sp_scalars.reserve(256));
    c->sp_irr_cap = 2 * n_tiles;
    CRATAC_TRY(c->d_sp_irr.reserve(sizeof(IrrTile) * (size_t)c->sp_irr_cap));
    CRATAC_CUDA(cudaMemsetAsync(c->d_sp_scalars.p, 0, 256, c->stream));   // ticket, declined ranges, line base, cycle counters
    *n_tiles_out = n_tiles;
    return CRATAC_OK;
}

// the ranges k_parse_stream declined go through the general kernel; n_irr is on the host by now
static int sp_run_declined(Ctx* c, DecodeArgs a, int64_t n_irr) {
    if (n_irr <= 0) return CRATAC_OK;
    if (n_irr > c->sp_irr_cap) { set_error("declined-tile list overflow"); return CRATAC_ERR_INTERNAL; }
    a.irr = c->d_sp_irr.as<IrrTile>();
    a.line_base = nullptr; a.tile0 = 0;
    k_parse_tiles<<<(unsigned)n_irr, DEC_THREADS, 0, c->stream>>>(a);
    c->launches++;
    c->decode_declined += n_irr;
    CRATAC_CUDA(cudaGetLastError());
    return CRATAC_OK;
}

static int decode_error(Ctx* c, int64_t err) {
    const int64_t arg = c->h_status[ST_ERROR_ARG];
    if (err == CRATAC_ERR_PARSE) set_error("malformed fragments line %lld (expected contig\\tstart\\tstop\\tbarcode\\tdup_count)", (long long)(arg + 1));
    else if (err == CRATAC_ERR_CONTIG) set_error("fragments line %lld: contig is not in the reference", (long long)(arg + 1));
    else if (err == CRATAC_ERR_CAPACITY) set_error("barcode table overflow at line %lld", (long long)(arg + 1));
    else set_error("device error %lld (arg %lld) in decode", (long long)err, (long long)arg);
    c->n_fragments = 0;
    return (int)err;
}

static int decode_text_two_pass(Ctx* c, const char* d_text, int64_t nbytes);
static int decode_text_streamed_two_pass(Ctx* c, HostTextSource* src, int64_t nbytes, int64_t chunk_bytes, const char* d_resident);

// Device text -> fragments in ONE pass over the text (k_parse_stream): the number of lines is not known in advance, so the
// output arrays are sized from an estimate (a fragments.tsv line is rarely under 24 bytes); when that or the barcode table
// does not hold, the table grows and / or the exact two-pass path takes over.
int decode_text(Ctx* c, const char* d_text, int64_t nbytes) {
    if (c->decode_kernel == 0) {
        if (nbytes > 0 && (c->decode_pipeline == 2 || (c->decode_pipeline == 1 && nbytes >= ((int64_t)64 << 20)))) {
            // large text: count passes of the next chunk under the parse of the current one
            struct ResidentText : HostTextSource {
                char last;
                char last_byte() override { return last; }
                const char* acquire(int64_t, int64_t) override { return nullptr; }
                int queued(int64_t, int64_t, cudaStream_t) override { return CRATAC_OK; }
            } src;
            src.last = '\n';
            CRATAC_CUDA(cudaMemcpyAsync(&src.last, d_text + nbytes - 1, 1, cudaMemcpyDeviceToHost, c->stream));
            CRATAC_CUDA(cudaStreamSynchronize(c->stream));
            const int rc = decode_text_streamed_two_pass(c, &src, nbytes, 0, d_text);
            if (rc != CRATAC_ERR_CAPACITY) return rc;       // estimate or table too small: the exact path below
        }
        return decode_text_two_pass(c, d_text, nbytes);
    }
    if (c->contig_table_cap == 0) CRATAC_TRY(build_contig_table(c));
    c->d_text = d_text;
    c->text_bytes = nbytes;
    c->bc_ready = false;
    c->barcodes.clear();
    int64_t* st = c->d_status.as<int64_t>();
    char last = '\n';
    if (nbytes > 0) CRATAC_CUDA(cudaMemcpyAsync(&last, d_text + nbytes - 1, 1, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    const int64_t n_eff = nbytes + (last != '\n' ? 1 : 0);
    int64_t n_tiles = 0;
    CRATAC_TRY(sp_reserve(c, n_eff, &n_tiles));
    const size_t held = std::min({c->d_chrom.cap, c->d_start.cap, c->d_end.cap, c->d_bc.cap, c->d_dup.cap}) / 4;
    const int64_t line_cap = std::max<int64_t>(nbytes / 24 + 4096, (int64_t)held);
    {
        const size_t nn = (size_t)line_cap;
        CRATAC_TRY(c->d_chrom.reserve(4 * nn)); CRATAC_TRY(c->d_start.reserve(4 * nn)); CRATAC_TRY(c->d_end.reserve(4 * nn));
        CRATAC_TRY(c->d_bc.reserve(4 * nn)); CRATAC_TRY(c->d_dup.reserve(4 * nn));
    }
    int64_t* line_base = c->d_sp_scalars.as<int64_t>() + 2;
    for (int attempt = 0;; attempt++) {
        const size_t cap = (size_t)c->bc_table_cap;
        CRATAC_TRY(c->d_bc_keys.reserve(sizeof(BcSlot) * cap));
        CRATAC_TRY(c->d_bc_textoff.reserve(8 * cap)); CRATAC_TRY(c->d_slot_to_dense.reserve(4 * cap));
        CRATAC_TRY(c->d_bc_nfrag.reserve(4 * cap));
        CRATAC_CUDA(cudaMemsetAsync(c->d_bc_nfrag.p, 0, 4 * cap, c->stream));
        k_init_table<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>(c->d_bc_keys.as<BcSlot>(), (int)cap);
        c->launches++;
        CRATAC_CUDA(cudaMemsetAsync(&st[ST_ERROR], 0, sizeof(int64_t) * 2, c->stream));
        CRATAC_CUDA(cudaMemsetAsync(&st[ST_N_BARCODES], 0, sizeof(int64_t) * 3, c->stream));
        CRATAC_CUDA(cudaMemsetAsync(c->d_sp_scalars.p, 0, 64, c->stream));
        DecodeArgs a;
        a.text = d_text; a.nbytes = nbytes; a.n_eff = n_eff;
        a.tile_offsets = nullptr; a.tile_counts = nullptr;
        a.contig_table = c->d_contig_table.as<ContigEntry>(); a.contig_mask = c->contig_table_cap - 1;
        a.chrom = c->d_chrom.as<int32_t>(); a.start = c->d_start.as<int32_t>(); a.end = c->d_end.as<int32_t>();
        a.bc = c->d_bc.as<int32_t>(); a.dup = c->d_dup.as<int32_t>();
        a.bc_tab = c->d_bc_keys.as<BcSlot>(); a.bc_nfrag = c->d_bc_nfrag.as<int32_t>();
        a.bc_textoff = c->d_bc_textoff.as<long long>(); a.bc_mask = c->bc_table_cap - 1;
        a.status = st;
        a.use_tma = ((reinterpret_cast<uintptr_t>(d_text) & 15) == 0) ? 1 : 0;
        a.tile0 = 0; a.line_base = nullptr; a.line_cap = line_cap; a.irr = nullptr;
        c->stage_begin(SG_DECODE_PARSE);
        CRATAC_TRY(sp_launch(c, a, 0, n_tiles, line_base, line_base + 1));
        c->stage_end(SG_DECODE_PARSE);
        CRATAC_CUDA(cudaMemcpyAsync(c->h_status, st, sizeof(int64_t) * ST_SLOTS, cudaMemcpyDeviceToHost, c->stream));
        CRATAC_CUDA(cudaMemcpyAsync(c->h_sp_scalars, c->d_sp_scalars.p, 32, cudaMemcpyDeviceToHost, c->stream));
        CRATAC_CUDA(cudaStreamSynchronize(c->stream));
        c->host_mark("decode_stream_sync");
        int64_t err = c->h_status[ST_ERROR];
        if (err == 0 && c->h_sp_scalars[1] > 0) {
            CRATAC_TRY(sp_run_declined(c, a, c->h_sp_scalars[1]));
            CRATAC_CUDA(cudaMemcpyAsync(c->h_status, st, sizeof(int64_t) * ST_SLOTS, cudaMemcpyDeviceToHost, c->stream));
            CRATAC_CUDA(cudaStreamSynchronize(c->stream));
            err = c->h_status[ST_ERROR];
        }
        const int64_t nb = c->h_status[ST_N_BARCODES];
        if (err == CRATAC_ERR_CAPACITY && c->h_status[ST_ERROR_ARG] == -7) return decode_text_two_pass(c, d_text, nbytes);   // more lines than estimated
        const bool crowded = nb * 2 > (int64_t)c->bc_table_cap;
        if ((err == CRATAC_ERR_CAPACITY || (err == 0 && crowded)) && attempt < 6 && c->bc_table_cap < (1 << 28)) {
            c->bc_table_cap <<= 2;   // grow and decode again
            continue;
        }
        if (err != 0) return decode_error(c, err);
        c->n_fragments = c->h_sp_scalars[3];
        c->n_barcodes = nb;
        CRATAC_CUDA(cudaMemcpyAsync(&st[ST_N_FRAGMENTS], line_base + 1, sizeof(int64_t), cudaMemcpyDeviceToDevice, c->stream));
        break;
    }
    CRATAC_TRY(finalize_fragments(c));
    return barcode_prepare_device(c);
}

static int decode_text_two_pass(Ctx* c, const char* d_text, int64_t nbytes) {
    if (c->contig_table_cap == 0) CRATAC_TRY(build_contig_table(c));
    c->d_text = d_text;
    c->text_bytes = nbytes;
    c->bc_ready = false;
    c->barcodes.clear();
    int64_t* st = c->d_status.as<int64_t>();
    // virtual newline after an unterminated last line
    char last = '\n';
    if (nbytes > 0) CRATAC_CUDA(cudaMemcpyAsync(&last, d_text + nbytes - 1, 1, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    int64_t n_eff = nbytes + (last != '\n' ? 1 : 0);
    int64_t n_tiles = (n_eff + DEC_TILE - 1) / DEC_TILE;
    if (n_tiles == 0) n_tiles = 1;
    CRATAC_TRY(c->d_tile_counts.reserve(sizeof(int32_t) * (size_t)n_tiles));
    CRATAC_TRY(c->d_tile_offsets.reserve(sizeof(int64_t) * (size_t)(n_tiles + 1)));
    c->stage_begin(SG_DECODE_COUNT);
    k_count_newlines<<<(unsigned)n_tiles, 256, 0, c->stream>>>(d_text, nbytes, n_eff, c->d_tile_counts.as<int32_t>(), 0, DEC_TILE);
    c->stage_end(SG_DECODE_COUNT);
    c->launches++;
    CRATAC_CUDA(cudaGetLastError());
    CRATAC_TRY(exclusive_scan_i32_to_i64(c, c->d_tile_counts.as<int32_t>(), c->d_tile_offsets.as<int64_t>(), n_tiles, &st[ST_N_FRAGMENTS]));
    CRATAC_CUDA(cudaMemcpyAsync(&c->h_status[ST_N_FRAGMENTS], &st[ST_N_FRAGMENTS], sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    int64_t n = c->h_status[ST_N_FRAGMENTS];
    c->n_fragments = n;
    c->host_mark("decode_count_sync");
    size_t nn = (size_t)std::max<int64_t>(n, 1);
    CRATAC_TRY(c->d_chrom.reserve(4 * nn)); CRATAC_TRY(c->d_start.reserve(4 * nn)); CRATAC_TRY(c->d_end.reserve(4 * nn));
    CRATAC_TRY(c->d_bc.reserve(4 * nn)); CRATAC_TRY(c->d_dup.reserve(4 * nn));

    for (int attempt = 0;; attempt++) {
        size_t cap = (size_t)c->bc_table_cap;
        CRATAC_TRY(c->d_bc_keys.reserve(sizeof(BcSlot) * cap));
        CRATAC_TRY(c->d_bc_textoff.reserve(8 * cap)); CRATAC_TRY(c->d_slot_to_dense.reserve(4 * cap));
        CRATAC_TRY(c->d_bc_nfrag.reserve(4 * cap));
        CRATAC_CUDA(cudaMemsetAsync(c->d_bc_nfrag.p, 0, 4 * cap, c->stream));
        k_init_table<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>(c->d_bc_keys.as<BcSlot>(), (int)cap);
        c->launches++;
        CRATAC_CUDA(cudaMemsetAsync(&st[ST_ERROR], 0, sizeof(int64_t) * 2, c->stream));
        CRATAC_CUDA(cudaMemsetAsync(&st[ST_N_BARCODES], 0, sizeof(int64_t) * 3, c->stream));  // barcodes, maxpos, maxneg
        DecodeArgs a;
        a.text = d_text; a.nbytes = nbytes; a.n_eff = n_eff;
        a.tile_offsets = c->d_tile_offsets.as<int64_t>(); a.tile_counts = c->d_tile_counts.as<int32_t>();
        a.contig_table = c->d_contig_table.as<ContigEntry>(); a.contig_mask = c->contig_table_cap - 1;
        a.chrom = c->d_chrom.as<int32_t>(); a.start = c->d_start.as<int32_t>(); a.end = c->d_end.as<int32_t>();
        a.bc = c->d_bc.as<int32_t>(); a.dup = c->d_dup.as<int32_t>();
        a.bc_tab = c->d_bc_keys.as<BcSlot>(); a.bc_nfrag = c->d_bc_nfrag.as<int32_t>();
        a.bc_textoff = c->d_bc_textoff.as<long long>(); a.bc_mask = c->bc_table_cap - 1;
        a.status = st;
        a.use_tma = ((reinterpret_cast<uintptr_t>(d_text) & 15) == 0) ? 1 : 0;
        a.tile0 = 0; a.line_base = nullptr; a.line_cap = 0; a.irr = nullptr;
        c->stage_begin(SG_DECODE_PARSE);
        k_parse_tiles<<<(unsigned)n_tiles, DEC_THREADS, 0, c->stream>>>(a);
        c->stage_end(SG_DECODE_PARSE);
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
        CRATAC_CUDA(cudaMemcpyAsync(c->h_status, st, sizeof(int64_t) * ST_SLOTS, cudaMemcpyDeviceToHost, c->stream));
        CRATAC_CUDA(cudaStreamSynchronize(c->stream));
        c->host_mark("decode_parse_sync");
        int64_t err = c->h_status[ST_ERROR];
        int64_t nb = c->h_status[ST_N_BARCODES];
        bool crowded = nb * 2 > (int64_t)c->bc_table_cap;
        if ((err == CRATAC_ERR_CAPACITY || (err == 0 && crowded)) && attempt < 6 && c->bc_table_cap < (1 << 28)) {
            c->bc_table_cap <<= 2;   // grow and decode again
            continue;
        }
        if (err != 0) {
            int64_t arg = c->h_status[ST_ERROR_ARG];
            if (err == CRATAC_ERR_PARSE) set_error("malformed fragments line %lld (expected contig\\tstart\\tstop\\tbarcode\\tdup_count)", (long long)(arg + 1));
            else if (err == CRATAC_ERR_CONTIG) set_error("fragments line %lld: contig is not in the reference", (long long)(arg + 1));
            else if (err == CRATAC_ERR_CAPACITY) set_error("barcode table overflow at line %lld", (long long)(arg + 1));
            else set_error("device error %lld (arg %lld) in decode", (long long)err, (long long)arg);
            c->n_fragments = 0;
            return (int)err;
        }
        c->n_barcodes = nb;
        break;
    }
    CRATAC_TRY(finalize_fragments(c));
    return barcode_prepare_device(c);
}

__global__ void k_add_i64(int64_t* __restrict__ acc, const int64_t* __restrict__ x) { *acc += *x; }

// Host text -> fragments with the PCIe copy overlapped: the text is copied in chunks on a copy stream and every chunk is
// counted, scanned (line offsets relative to the chunk, plus a running base kept on the device) and parsed as soon as it
// has arrived, while the next chunk is in flight.  The number of lines is not known before the last chunk, so the output
// arrays are sized from an estimate; returns CRATAC_ERR_CAPACITY when anything did not fit (table or estimate) -- the
// caller then falls back to copy-all-then-decode.
static int decode_text_streamed_two_pass(Ctx* c, HostTextSource* src, int64_t nbytes, int64_t chunk_bytes, const char* d_resident);

namespace {
struct PlainHostText : HostTextSource {
    const char* p; int64_t n;
    PlainHostText(const char* p_, int64_t n_) : p(p_), n(n_) {}
    char last_byte() override { return n > 0 ? p[n - 1] : '\n'; }
    const char* acquire(int64_t b0, int64_t) override { return p + b0; }
    int queued(int64_t, int64_t, cudaStream_t) override { return CRATAC_OK; }
};
}  // namespace

// bytes per chunk the streamed decode will ask its source for, given the caller's wish (0: the default)
int64_t decode_stream_chunk_bytes(Ctx* c, int64_t chunk_bytes_arg) {
    const int64_t tile = c->decode_kernel == 0 ? DEC_TILE : sp_tile_bytes(c);
    const int64_t want = chunk_bytes_arg > 0 ? chunk_bytes_arg : (c->stream_chunk_tiles > 0 ? c->stream_chunk_tiles * DEC_TILE : (int64_t)(128 << 20));
    return std::max<int64_t>(want / tile, 1) * tile;
}

int decode_text_streamed(Ctx* c, const char* h_text, int64_t nbytes) {
    PlainHostText src(h_text, nbytes);
    return decode_text_streamed_from(c, &src, nbytes, 0);
}

// Same with the single-pass kernel: one k_parse_stream launch per chunk, the running line base stays on the device.
int decode_text_streamed_from(Ctx* c, HostTextSource* src, int64_t nbytes, int64_t chunk_bytes_arg) {
    if (c->decode_kernel == 0) return decode_text_streamed_two_pass(c, src, nbytes, chunk_bytes_arg, nullptr);
    if (c->contig_table_cap == 0) CRATAC_TRY(build_contig_table(c));
    CRATAC_TRY(c->d_text_own.reserve((size_t)nbytes + 16));
    char* d_text = c->d_text_own.as<char>();
    c->d_text = d_text;
    c->text_bytes = nbytes;
    c->bc_ready = false;
    c->barcodes.clear();
    int64_t* st = c->d_status.as<int64_t>();
    if (!c->copy_stream) {
        CRATAC_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CRATAC_CUDA(cudaEventCreateWithFlags(&c->copy_event, cudaEventDisableTiming));
    }
    const int64_t n_eff = nbytes + ((nbytes > 0 && src->last_byte() != '\n') ? 1 : 0);
    int64_t n_tiles = 0;
    CRATAC_TRY(sp_reserve(c, n_eff, &n_tiles));
    const int TB = sp_tile_bytes(c);
    const size_t held = std::min({c->d_chrom.cap, c->d_start.cap, c->d_end.cap, c->d_bc.cap, c->d_dup.cap}) / 4;
    const int64_t line_cap = std::max<int64_t>(nbytes / 24 + 4096, (int64_t)held);
    {
        const size_t nn = (size_t)line_cap;
        CRATAC_TRY(c->d_chrom.reserve(4 * nn)); CRATAC_TRY(c->d_start.reserve(4 * nn)); CRATAC_TRY(c->d_end.reserve(4 * nn));
        CRATAC_TRY(c->d_bc.reserve(4 * nn)); CRATAC_TRY(c->d_dup.reserve(4 * nn));
    }
    const size_t cap = (size_t)c->bc_table_cap;
    CRATAC_TRY(c->d_bc_keys.reserve(sizeof(BcSlot) * cap));
    CRATAC_TRY(c->d_bc_textoff.reserve(8 * cap)); CRATAC_TRY(c->d_slot_to_dense.reserve(4 * cap));
    CRATAC_TRY(c->d_bc_nfrag.reserve(4 * cap));
    CRATAC_CUDA(cudaMemsetAsync(c->d_bc_nfrag.p, 0, 4 * cap, c->stream));
    k_init_table<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>(c->d_bc_keys.as<BcSlot>(), (int)cap);
    c->launches++;
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_ERROR], 0, sizeof(int64_t) * 2, c->stream));
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_N_BARCODES], 0, sizeof(int64_t) * 3, c->stream));
    int64_t* line_base = c->d_sp_scalars.as<int64_t>() + 2;
    DecodeArgs a;
    a.text = d_text; a.nbytes = nbytes; a.n_eff = n_eff;
    a.tile_offsets = nullptr; a.tile_counts = nullptr;
    a.contig_table = c->d_contig_table.as<ContigEntry>(); a.contig_mask = c->contig_table_cap - 1;
    a.chrom = c->d_chrom.as<int32_t>(); a.start = c->d_start.as<int32_t>(); a.end = c->d_end.as<int32_t>();
    a.bc = c->d_bc.as<int32_t>(); a.dup = c->d_dup.as<int32_t>();
    a.bc_tab = c->d_bc_keys.as<BcSlot>(); a.bc_nfrag = c->d_bc_nfrag.as<int32_t>();
    a.bc_textoff = c->d_bc_textoff.as<long long>(); a.bc_mask = c->bc_table_cap - 1;
    a.status = st;
    a.use_tma = ((reinterpret_cast<uintptr_t>(d_text) & 15) == 0) ? 1 : 0;
    a.tile0 = 0; a.line_base = nullptr; a.line_cap = line_cap; a.irr = nullptr;
    const int64_t chunk_bytes = chunk_bytes_arg > 0 ? chunk_bytes_arg : (c->stream_chunk_tiles > 0 ? c->stream_chunk_tiles * DEC_TILE : (int64_t)(128 << 20));   // default: 128 MiB of text per chunk
    const int64_t CHUNK_TILES = std::max<int64_t>(chunk_bytes / TB, 1);
    int n_launches = 0;
    c->stage_begin(SG_DECODE_PARSE);
    for (int64_t t0 = 0; t0 < n_tiles; t0 += CHUNK_TILES) {
        const int64_t t1 = std::min(n_tiles, t0 + CHUNK_TILES);
        const int64_t b0 = t0 * TB, b1 = std::min(nbytes, t1 * TB);
        if (b1 > b0) {
            const char* hp = src->acquire(b0, b1);
            if (!hp) return CRATAC_ERR_IO;
            CRATAC_CUDA(cudaMemcpyAsync(d_text + b0, hp, (size_t)(b1 - b0), cudaMemcpyHostToDevice, c->copy_stream));
            CRATAC_TRY(src->queued(b0, b1, c->copy_stream));
        }
        CRATAC_CUDA(cudaEventRecord(c->copy_event, c->copy_stream));
        CRATAC_CUDA(cudaStreamWaitEvent(c->stream, c->copy_event, 0));
        CRATAC_TRY(sp_launch(c, a, t0, t1 - t0, line_base + (n_launches & 1), line_base + ((n_launches + 1) & 1)));
        n_launches++;
    }
    c->stage_end(SG_DECODE_PARSE);
    const int total_word = 2 + (n_launches & 1);       // the word the last launch wrote
    CRATAC_CUDA(cudaMemcpyAsync(c->h_status, st, sizeof(int64_t) * ST_SLOTS, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaMemcpyAsync(c->h_sp_scalars, c->d_sp_scalars.p, 32, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    c->host_mark("decode_streamed_sync");
    int64_t err = c->h_status[ST_ERROR];
    if (err == 0 && c->h_sp_scalars[1] > 0) {
        CRATAC_TRY(sp_run_declined(c, a, c->h_sp_scalars[1]));
        CRATAC_CUDA(cudaMemcpyAsync(c->h_status, st, sizeof(int64_t) * ST_SLOTS, cudaMemcpyDeviceToHost, c->stream));
        CRATAC_CUDA(cudaStreamSynchronize(c->stream));
        err = c->h_status[ST_ERROR];
    }
    const int64_t nb = c->h_status[ST_N_BARCODES];
    if (err == CRATAC_ERR_CAPACITY || (err == 0 && nb * 2 > (int64_t)c->bc_table_cap)) return CRATAC_ERR_CAPACITY;   // caller: exact path with a larger table
    if (err != 0) return decode_error(c, err);
    c->n_fragments = c->h_sp_scalars[total_word];
    c->n_barcodes = nb;
    CRATAC_CUDA(cudaMemcpyAsync(&st[ST_N_FRAGMENTS], c->d_sp_scalars.as<int64_t>() + total_word, sizeof(int64_t), cudaMemcpyDeviceToDevice, c->stream));
    CRATAC_TRY(finalize_fragments(c));
    return barcode_prepare_device(c);
}

// Chunk by chunk on two streams: the newline count + scan of chunk k + 1 (HBM-bound, a high-priority side stream) run
// under the parse of chunk k (issue-bound, the context's stream); for a host text the H2D copy of the chunk comes first.
// d_resident: the text is already in HBM at this address (no copies, src is only asked for the last byte).
static int decode_text_streamed_two_pass(Ctx* c, HostTextSource* src, int64_t nbytes, int64_t chunk_bytes_arg, const char* d_resident) {
    if (c->contig_table_cap == 0) CRATAC_TRY(build_contig_table(c));
    if (!d_resident) CRATAC_TRY(c->d_text_own.reserve((size_t)nbytes + 16));
    char* d_text = d_resident ? const_cast<char*>(d_resident) : c->d_text_own.as<char>();
    c->d_text = d_text;
    c->text_bytes = nbytes;
    c->bc_ready = false;
    c->barcodes.clear();
    int64_t* st = c->d_status.as<int64_t>();
    if (!c->copy_stream) {
        CRATAC_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CRATAC_CUDA(cudaEventCreateWithFlags(&c->copy_event, cudaEventDisableTiming));
    }
    const int64_t n_eff = nbytes + ((nbytes > 0 && src->last_byte() != '\n') ? 1 : 0);
    int64_t n_tiles = (n_eff + DEC_TILE - 1) / DEC_TILE;
    if (n_tiles == 0) n_tiles = 1;
    CRATAC_TRY(c->d_tile_counts.reserve(sizeof(int32_t) * (size_t)n_tiles));
    CRATAC_TRY(c->d_tile_offsets.reserve(sizeof(int64_t) * (size_t)(n_tiles + 1)));
    const size_t held = std::min({c->d_chrom.cap, c->d_start.cap, c->d_end.cap, c->d_bc.cap, c->d_dup.cap}) / 4;
    const int64_t line_cap = std::max<int64_t>(nbytes / 24 + 4096, (int64_t)held);   // a fragments.tsv line is rarely under 24 bytes
    {
        const size_t nn = (size_t)line_cap;
        CRATAC_TRY(c->d_chrom.reserve(4 * nn)); CRATAC_TRY(c->d_start.reserve(4 * nn)); CRATAC_TRY(c->d_end.reserve(4 * nn));
        CRATAC_TRY(c->d_bc.reserve(4 * nn)); CRATAC_TRY(c->d_dup.reserve(4 * nn));
    }
    const size_t cap = (size_t)c->bc_table_cap;
    CRATAC_TRY(c->d_bc_keys.reserve(sizeof(BcSlot) * cap));
    CRATAC_TRY(c->d_bc_textoff.reserve(8 * cap)); CRATAC_TRY(c->d_slot_to_dense.reserve(4 * cap));
    CRATAC_TRY(c->d_bc_nfrag.reserve(4 * cap));
    CRATAC_CUDA(cudaMemsetAsync(c->d_bc_nfrag.p, 0, 4 * cap, c->stream));
    const int64_t CHUNK_TILES = chunk_bytes_arg > 0 ? std::max<int64_t>(chunk_bytes_arg / DEC_TILE, 1)
                                                    : (c->stream_chunk_tiles > 0 ? c->stream_chunk_tiles : (int64_t)(128 << 20) / DEC_TILE);   // default: 128 MiB of text per chunk
    const int64_t n_chunks = (n_tiles + CHUNK_TILES - 1) / CHUNK_TILES;
    CRATAC_TRY(c->d_stream_scalars.reserve(8 * (size_t)(n_chunks + 4)));
    int64_t* line_base = c->d_stream_scalars.as<int64_t>();
    int64_t* chunk_totals = line_base + 2;                   // one word per chunk: the counts run ahead of the parses
    if (!c->count_stream) {
        int lo = 0, hi = 0;
        CRATAC_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CRATAC_CUDA(cudaStreamCreateWithPriority(&c->count_stream, cudaStreamNonBlocking, hi));
        CRATAC_CUDA(cudaEventCreateWithFlags(&c->count_event, cudaEventDisableTiming));
    }
    k_init_table<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>(c->d_bc_keys.as<BcSlot>(), (int)cap);
    c->launches++;
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_ERROR], 0, sizeof(int64_t) * 2, c->stream));
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_N_BARCODES], 0, sizeof(int64_t) * 3, c->stream));
    CRATAC_CUDA(cudaMemsetAsync(line_base, 0, 16, c->stream));
    DecodeArgs a;
    a.text = d_text; a.nbytes = nbytes; a.n_eff = n_eff;
    a.tile_offsets = c->d_tile_offsets.as<int64_t>(); a.tile_counts = c->d_tile_counts.as<int32_t>();
    a.contig_table = c->d_contig_table.as<ContigEntry>(); a.contig_mask = c->contig_table_cap - 1;
    a.chrom = c->d_chrom.as<int32_t>(); a.start = c->d_start.as<int32_t>(); a.end = c->d_end.as<int32_t>();
    a.bc = c->d_bc.as<int32_t>(); a.dup = c->d_dup.as<int32_t>();
    a.bc_tab = c->d_bc_keys.as<BcSlot>(); a.bc_nfrag = c->d_bc_nfrag.as<int32_t>();
    a.bc_textoff = c->d_bc_textoff.as<long long>(); a.bc_mask = c->bc_table_cap - 1;
    a.status = st;
    a.use_tma = ((reinterpret_cast<uintptr_t>(d_text) & 15) == 0) ? 1 : 0;
    a.line_base = line_base; a.line_cap = line_cap; a.irr = nullptr;
    c->stage_begin(SG_DECODE_PARSE);
    // the side stream starts where the context's stream stands now (whatever produced the text, the table reset above)
    CRATAC_CUDA(cudaEventRecord(c->count_event, c->stream));
    CRATAC_CUDA(cudaStreamWaitEvent(c->count_stream, c->count_event, 0));
    cudaStream_t own = c->stream;
    int64_t k = 0;
    for (int64_t t0 = 0; t0 < n_tiles; t0 += CHUNK_TILES, k++) {
        const int64_t t1 = std::min(n_tiles, t0 + CHUNK_TILES);
        const int64_t b0 = t0 * DEC_TILE, b1 = std::min(nbytes, t1 * DEC_TILE);
        if (!d_resident) {
            if (b1 > b0) {
                const char* hp = src->acquire(b0, b1);
                if (!hp) return CRATAC_ERR_IO;
                CRATAC_CUDA(cudaMemcpyAsync(d_text + b0, hp, (size_t)(b1 - b0), cudaMemcpyHostToDevice, c->copy_stream));
                CRATAC_TRY(src->queued(b0, b1, c->copy_stream));
            }
            CRATAC_CUDA(cudaEventRecord(c->copy_event, c->copy_stream));
            CRATAC_CUDA(cudaStreamWaitEvent(c->count_stream, c->copy_event, 0));
        }
        const unsigned nt = (unsigned)(t1 - t0);
        k_count_newlines<<<nt, 256, 0, c->count_stream>>>(d_text, nbytes, n_eff, c->d_tile_counts.as<int32_t>(), t0, DEC_TILE);
        c->stream = c->count_stream;                        // the scan helper queues on the context's current stream
        const int rc_scan = exclusive_scan_i32_to_i64(c, c->d_tile_counts.as<int32_t>() + t0, c->d_tile_offsets.as<int64_t>() + t0, t1 - t0, chunk_totals + k);
        c->stream = own;
        CRATAC_TRY(rc_scan);
        CRATAC_CUDA(cudaEventRecord(c->count_event, c->count_stream));
        CRATAC_CUDA(cudaStreamWaitEvent(c->stream, c->count_event, 0));      // (captures this record: the event is recorded again for the next chunk)
        a.tile0 = t0;
        k_parse_tiles<<<nt, DEC_THREADS, 0, c->stream>>>(a);
        k_add_i64<<<1, 1, 0, c->stream>>>(line_base, chunk_totals + k);
        c->launches += 3;
        CRATAC_CUDA(cudaGetLastError());
    }
    c->stage_end(SG_DECODE_PARSE);
    CRATAC_CUDA(cudaMemcpyAsync(&st[ST_N_FRAGMENTS], line_base, sizeof(int64_t), cudaMemcpyDeviceToDevice, c->stream));
    CRATAC_CUDA(cudaMemcpyAsync(&c->h_status[ST_N_FRAGMENTS], line_base, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaMemcpyAsync(c->h_status, st, sizeof(int64_t) * 2, cudaMemcpyDeviceToHost, c->stream));   // error, arg
    CRATAC_CUDA(cudaMemcpyAsync(&c->h_status[ST_N_BARCODES], &st[ST_N_BARCODES], sizeof(int64_t) * 3, cudaMemcpyDeviceToHost, c->stream));
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    c->host_mark("decode_streamed_sync");
    const int64_t err = c->h_status[ST_ERROR];
    const int64_t nb = c->h_status[ST_N_BARCODES];
    if (err == CRATAC_ERR_CAPACITY || (err == 0 && nb * 2 > (int64_t)c->bc_table_cap)) return CRATAC_ERR_CAPACITY;   // caller: exact two-pass path
    if (err != 0) {
        int64_t arg = c->h_status[ST_ERROR_ARG];
        if (err == CRATAC_ERR_PARSE) set_error("malformed fragments line %lld (expected contig\\tstart\\tstop\\tbarcode\\tdup_count)", (long long)(arg + 1));
        else if (err == CRATAC_ERR_CONTIG) set_error("fragments line %lld: contig is not in the reference", (long long)(arg + 1));
        else set_error("device error %lld (arg %lld) in decode", (long long)err, (long long)arg);
        c->n_fragments = 0;
        return (int)err;
    }
    c->n_fragments = c->h_status[ST_N_FRAGMENTS];
    c->n_barcodes = nb;
    CRATAC_TRY(finalize_fragments(c));
    return barcode_prepare_device(c);
}

int finalize_fragments(Ctx* c) {
    int64_t n = c->n_fragments;
    c->tile_desc_valid = false;
    c->tile_bound_valid = false;
    c->bg_valid = false;
    c->sib_valid = false;
    int64_t* st = c->d_status.as<int64_t>();
    int nc = c->n_contigs;
    CRATAC_TRY(c->d_contig_frag_off.reserve(sizeof(int64_t) * (nc + 2)));
    int64_t total_pidx = c->contig_pidx_off[nc];
    CRATAC_TRY(c->d_pos_index.reserve(sizeof(int64_t) * (size_t)total_pidx));
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_ERROR], 0, sizeof(int64_t) * 2, c->stream));
    c->stage_begin(SG_FRAG_INDEX);
    if (n > 0) {
        k_check_sorted<<<(unsigned)std::min<int64_t>((n / 4 + 255) / 256 + 1, (int64_t)NUM_SMS * 16), 256, 0, c->stream>>>(c->d_chrom.as<int32_t>(), c->d_start.as<int32_t>(), n, st);
        c->launches++;
    }
    k_contig_frag_offsets<<<(nc + 1 + 127) / 128, 128, 0, c->stream>>>(c->d_chrom.as<int32_t>(), n, nc, c->d_contig_frag_off.as<int64_t>());
    c->launches++;
    k_pos_index<<<(unsigned)((total_pidx + 255) / 256), 256, 0, c->stream>>>(c->d_start.as<int32_t>(), c->d_contig_frag_off.as<int64_t>(),
                                                                             c->d_contig_pidx_off.as<int64_t>(), nc, total_pidx,
                                                                             c->d_pos_index.as<int64_t>());
    c->stage_end(SG_FRAG_INDEX);
    c->launches++;
    CRATAC_CUDA(cudaGetLastError());
    CRATAC_TRY(sync_status(c));
    return CRATAC_OK;
}

int compute_max_len(Ctx* c) {
    int64_t* st = c->d_status.as<int64_t>();
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_MAX_POS_LEN], 0, sizeof(int64_t) * 2, c->stream));
    if (c->n_fragments > 0) {
        k_max_len<<<NUM_SMS * 4, 256, 0, c->stream>>>(c->d_start.as<int32_t>(), c->d_end.as<int32_t>(), c->n_fragments, st);
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
    }
    return CRATAC_OK;
}

// ---- CPython 2.7 string hash / set order (see cratac_py2_set_order in api.cu)
// CPython 2.7 set of distinct str keys given by their hashes in insertion order; returns the insertion index
// of the key in every occupied slot, in slot order = list(set) order.  Distinct keys never compare equal, so the
// probe sequence depends on the hashes only (Objects/setobject.c: set_lookkey_string, set_insert_clean,
// set_add_key, set_table_resize; PySet_MINSIZE 8, PERTURB_SHIFT 5).
void py2_set_order_hashes(const long long* hashes, int64_t n, std::vector<int32_t>& order) {
    // The table keeps (hash, index) together so that a resize reads the old table front to back without going
    // back to hashes[], and the home slot of the element 16 positions ahead is prefetched: the simulation is
    // bound by cache misses of random probes into a table of a few MB, not by arithmetic.
    struct Ent { uint64_t h; int64_t k; };
    constexpr int AHEAD = 16;
    std::vector<Ent> table(8, Ent{0, -1});
    uint64_t mask = 7, used = 0, fill = 0;
    for (int64_t k = 0; k < n; k++) {
        if (k + AHEAD < n) __builtin_prefetch(&table[(uint64_t)hashes[k + AHEAD] & mask], 1, 1);
        const uint64_t h = (uint64_t)hashes[k];
        uint64_t i = h & mask, perturb = h;
        while (table[i & mask].k >= 0) { i = i * 5 + perturb + 1; perturb >>= 5; }
        table[i & mask] = Ent{h, k};
        used++; fill++;
        if (fill * 3 >= (mask + 1) * 2) {
            uint64_t minused = used > 50000 ? used * 2 : used * 4, newsize = 8;
            while (newsize <= minused) newsize <<= 1;
            std::vector<Ent> nt(newsize, Ent{0, -1});
            const uint64_t m2 = newsize - 1;
            const size_t old_n = table.size();
            for (size_t t = 0; t < old_n; t++) {
                if (t + AHEAD < old_n && table[t + AHEAD].k >= 0) __builtin_prefetch(&nt[table[t + AHEAD].h & m2], 1, 1);
                const Ent e = table[t];
                if (e.k < 0) continue;
                uint64_t q = e.h & m2, pp = e.h;
                while (nt[q & m2].k >= 0) { q = q * 5 + pp + 1; pp >>= 5; }
                nt[q & m2] = e;
            }
            table.swap(nt);
            mask = m2;
            fill = used;
        }
    }
    order.clear();
    order.reserve((size_t)n);
    for (const Ent& e : table) if (e.k >= 0) order.push_back((int32_t)e.k);
}

static inline long long py2_hash(const std::string& s) {
    if (s.empty()) return 0;
    uint64_t x = (uint64_t)(unsigned char)s[0] << 7;
    for (unsigned char ch : s) x = (1000003ULL * x) ^ ch;
    x ^= (uint64_t)s.size();
    long long h = (long long)x;
    return h == -1 ? -2 : h;
}

// string front end (cratac_py2_set_order): equal strings collapse first, then the hash-only simulation applies
void py2_set_order(const std::vector<std::string>& keys, std::vector<int64_t>& order) {
    std::vector<long long> h(keys.size());
    for (size_t k = 0; k < keys.size(); k++) h[k] = py2_hash(keys[k]);
    std::vector<int32_t> o;
    py2_set_order_hashes(h.data(), (int64_t)keys.size(), o);
    order.assign(o.begin(), o.end());
}

static int fetch_barcode_strings(Ctx* c, std::vector<std::string>& names) {
    const int64_t nb = c->n_barcodes;
    size_t nn = (size_t)std::max<int64_t>(nb, 1);
    const int stride = 64;
    DevBuf d_txt;
    CRATAC_TRY(d_txt.reserve(nn * stride));
    std::vector<long long> textoff(nn);
    std::vector<char> txt(nn * stride);
    cudaStream_t s = c->stream;
    if (nb > 0) {
        k_gather_barcode_text<<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(c->d_text, c->d_dense_textoff.as<long long>(), nb, stride, d_txt.as<char>());
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
    }
    CRATAC_CUDA(cudaMemcpyAsync(txt.data(), d_txt.p, nn * stride, cudaMemcpyDeviceToHost, s));
    CRATAC_CUDA(cudaMemcpyAsync(textoff.data(), c->d_dense_textoff.p, 8 * nn, cudaMemcpyDeviceToHost, s));
    CRATAC_CUDA(cudaStreamSynchronize(s));
    d_txt.release();
    names.resize((size_t)nb);
    for (int64_t d = 0; d < nb; d++) {
        int len = (int)(textoff[d] & 255);
        if (len >= stride) { set_error("barcode longer than %d bytes", stride - 1); return CRATAC_ERR_PARSE; }
        names[d].assign(txt.data() + d * stride, (size_t)len);
    }
    return CRATAC_OK;
}

// host half: matrix column order.  Only the hashes (PY2SET) or nothing (FIRST_SEEN) cross to the host; the
// caller queues GPU work first so that this overlaps it.
int py2_set_order_device(Ctx* c, int64_t n, const int64_t* d_hashes, int32_t* d_order_out);   // py2order.cu

// The host simulation of the CPython-2.7 set (sequential, ~50 ns per key) on a thread of its own, started as soon as the
// hashes are on their way to the host (right after the decode): by the time the CSC build asks for the column order the
// answer is there, whatever the main thread was waiting for in between.
void build_barcodes_async(Ctx* c) {
    if (c->bc_ready || c->bc_thread.joinable()) return;
    const int64_t nb = c->n_barcodes;
    if (c->bc_order != CRATAC_BC_ORDER_PY2SET || nb <= 0 || !c->bc_event) return;
    c->bc_thread_done = false;
    if (nb >= c->py2_device_min && c->d_hash_first) {
        // very many barcodes: the set is rebuilt on the device (py2order.cu: a dozen table generations, each a few kernel
        // rounds and a host look at a flag -- 10 ms for 2 M keys).  On this thread, a stream and scan scratch of its own, it
        // runs beside the histogram pass and the fit instead of between the reduce and the CSC build.
        if (!c->bc_stream && cudaStreamCreateWithFlags(&c->bc_stream, cudaStreamNonBlocking) != cudaSuccess) { c->bc_stream = nullptr; return; }
        if (c->d_py2_rank.reserve(4 * (size_t)nb) != CRATAC_OK) return;
        c->bc_thread = std::thread([c, nb] {
            if (cudaSetDevice(c->device) != cudaSuccess || cudaEventSynchronize(c->bc_event) != cudaSuccess) return;
            Lane lane{c->bc_stream, &c->d_scan_tmp_bc};
            tl_lane = &lane;
            const int rc = py2_set_order_device(c, nb, reinterpret_cast<const int64_t*>(c->d_hash_first), c->d_py2_rank.as<int32_t>());
            tl_lane = nullptr;
            if (rc != CRATAC_OK) return;                       // the main thread does it again and reports the error
            const unsigned g = (unsigned)((nb + 255) / 256);
            k_compose_order<<<g, 256, 0, c->bc_stream>>>(c->d_py2_rank.as<int32_t>(), c->d_order_first, nb, c->d_col_to_dense.as<int32_t>());
            k_invert_order<<<g, 256, 0, c->bc_stream>>>(c->d_col_to_dense.as<int32_t>(), nb, c->d_dense_to_col.as<int32_t>());
            c->launches += 2;
            if (cudaMemcpyAsync(c->h_c2d_pin, c->d_col_to_dense.p, 4 * (size_t)nb, cudaMemcpyDeviceToHost, c->bc_stream) != cudaSuccess) return;
            if (cudaStreamSynchronize(c->bc_stream) != cudaSuccess) return;
            c->bc_thread_done = true;
        });
        return;
    }
    c->bc_thread = std::thread([c, nb] {
        if (cudaSetDevice(c->device) != cudaSuccess || cudaEventSynchronize(c->bc_event) != cudaSuccess) return;
        std::vector<int32_t> order;
        py2_set_order_hashes(c->h_bc_hash, nb, order);
        for (int64_t j = 0; j < nb; j++) c->h_c2d_pin[j] = c->h_bc_order[order[j]];
        c->bc_thread_done = true;
    });
}

int build_barcodes(Ctx* c) {
    bool from_thread = false;
    if (c->bc_thread.joinable()) { c->bc_thread.join(); from_thread = c->bc_thread_done && !c->bc_ready; }
    if (c->bc_ready) return CRATAC_OK;
    const int64_t nb = c->n_barcodes;
    cudaStream_t s = c->stream;
    CRATAC_CUDA(cudaEventSynchronize(c->bc_event));
    int32_t* c2d = c->h_c2d_pin;    // pinned: the upload below does not wait for the stream
    if (from_thread && c->bc_order == CRATAC_BC_ORDER_PY2SET && nb >= c->py2_device_min && nb > 0 && c->d_hash_first) {
        // the thread rebuilt the set on the device: both maps are in HBM (its stream was synchronised), c2d is their host copy
        c->h_c2d.assign(c2d, c2d + nb);
        c->bc_ready = true;
        return CRATAC_OK;
    } else if (from_thread && c->bc_order == CRATAC_BC_ORDER_PY2SET) {
        // c2d was filled by build_barcodes_async
    } else if (c->bc_order == CRATAC_BC_ORDER_FIRST_SEEN) {
        for (int64_t j = 0; j < nb; j++) c2d[j] = c->h_bc_order[j];
    } else if (c->bc_order == CRATAC_BC_ORDER_SORTED) {
        std::vector<std::string> names;
        CRATAC_TRY(fetch_barcode_strings(c, names));
        for (int64_t j = 0; j < nb; j++) c2d[j] = (int32_t)j;
        std::sort(c2d, c2d + nb, [&](int32_t x, int32_t y) { return names[x] < names[y]; });
    } else if (nb >= c->py2_device_min && nb > 0 && c->d_hash_first) {
        // many barcodes: the sequential host simulation (~50 ns per key) would outlast the GPU work it hides behind,
        // so the set is rebuilt on the device (py2order.cu), behind whatever the stream already holds
        CRATAC_TRY(c->d_py2_rank.reserve(4 * (size_t)nb));
        CRATAC_TRY(py2_set_order_device(c, nb, reinterpret_cast<const int64_t*>(c->d_hash_first), c->d_py2_rank.as<int32_t>()));
        const unsigned g = (unsigned)((nb + 255) / 256);
        k_compose_order<<<g, 256, 0, s>>>(c->d_py2_rank.as<int32_t>(), c->d_order_first, nb, c->d_col_to_dense.as<int32_t>());
        k_invert_order<<<g, 256, 0, s>>>(c->d_col_to_dense.as<int32_t>(), nb, c->d_dense_to_col.as<int32_t>());
        c->launches += 2;
        CRATAC_CUDA(cudaGetLastError());
        CRATAC_CUDA(cudaMemcpyAsync(c2d, c->d_col_to_dense.p, 4 * (size_t)nb, cudaMemcpyDeviceToHost, s));
        CRATAC_CUDA(cudaStreamSynchronize(s));
        c->h_c2d.assign(c2d, c2d + nb);
        c->bc_ready = true;
        return CRATAC_OK;
    } else {
        std::vector<int32_t> order;
        py2_set_order_hashes(c->h_bc_hash, nb, order);
        for (int64_t j = 0; j < nb; j++) c2d[j] = c->h_bc_order[order[j]];
    }
    c->h_c2d.assign(c2d, c2d + nb);
    if (nb > 0) {
        CRATAC_CUDA(cudaMemcpyAsync(c->d_col_to_dense.p, c2d, 4 * (size_t)nb, cudaMemcpyHostToDevice, s));
        k_invert_order<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(c->d_col_to_dense.as<int32_t>(), nb, c->d_dense_to_col.as<int32_t>());
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
    }
    c->bc_ready = true;
    return CRATAC_OK;
}

// barcode strings in column order (lazy: only the getters need them)
int build_barcode_strings(Ctx* c) {
    CRATAC_TRY(build_barcodes(c));
    if (c->bc_strings_ready) return CRATAC_OK;
    std::vector<std::string> names;
    CRATAC_TRY(fetch_barcode_strings(c, names));
    c->barcodes.resize((size_t)c->n_barcodes);
    for (int64_t j = 0; j < c->n_barcodes; j++) c->barcodes[j] = names[c->h_c2d[j]];
    c->bc_strings_ready = true;
    return CRATAC_OK;
}

// ------------------------------------------------------------------ synthetic text encoder (bench / tests)
__device__ __forceinline__ int dec_digits(uint32_t v) {
    int n = 1;
    while (v >= 10) { v /= 10; n++; }
    return n;
}
__device__ __forceinline__ char* put_uint(char* p, uint32_t v) {
    int n = dec_digits(v);
    for (int i = n - 1; i >= 0; i--) { p[i] = (char)('0' + v % 10); v /= 10; }
    return p + n;
}

struct EncArgs {
    int64_t n;
    const int32_t *chrom, *start, *end, *dup;
    const uint32_t* seq;
    const char* names;     // 32 bytes per contig, NUL padded
    const int32_t* name_len;
    int64_t* line_off;     // in: lengths (pass 0 writes), then exclusive offsets
    char* out;
};

__global__ void k_encode_len(EncArgs a) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    a.line_off[i] = a.name_len[a.chrom[i]] + 1 + dec_digits((uint32_t)a.start[i]) + 1 + dec_digits((uint32_t)a.end[i]) + 1 + 18 + 1 +
                    dec_digits((uint32_t)a.dup[i]) + 1;
}

__global__ void k_encode_write(EncArgs a) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    char* p = a.out + a.line_off[i];
    int c = a.chrom[i];
    for (int k = 0; k < a.name_len[c]; k++) *p++ = a.names[c * 32 + k];
    *p++ = '\t';
    p = put_uint(p, (uint32_t)a.start[i]); *p++ = '\t';
    p = put_uint(p, (uint32_t)a.end[i]); *p++ = '\t';
    uint32_t s = a.seq[i];
    const char L[4] = {'A', 'C', 'T', 'G'};   // inverse of (ch >> 1) & 3
#pragma unroll
    for (int k = 0; k < 16; k++) *p++ = L[(s >> (30 - 2 * k)) & 3u];
    *p++ = '-'; *p++ = '1'; *p++ = '\t';
    p = put_uint(p, (uint32_t)a.dup[i]);
    *p++ = '\n';
}

int encode_text(Ctx* c, int64_t n, const int32_t* d_chrom, const int32_t* d_start, const int32_t* d_end, const uint32_t* d_seq,
                const int32_t* d_dup, char* d_out, int64_t cap, int64_t* nbytes) {
    std::vector<char> names((size_t)c->n_contigs * 32, 0);
    std::vector<int32_t> lens((size_t)c->n_contigs);
    for (int i = 0; i < c->n_contigs; i++) {
        if (c->contig_names[i].size() > 31) { set_error("contig name too long for the encoder"); return CRATAC_ERR_ARG; }
        memcpy(&names[(size_t)i * 32], c->contig_names[i].data(), c->contig_names[i].size());
        lens[i] = (int32_t)c->contig_names[i].size();
    }
    DevBuf d_names, d_lens, d_off;
    CRATAC_TRY(d_names.reserve(names.size())); CRATAC_TRY(d_lens.reserve(4 * lens.size())); CRATAC_TRY(d_off.reserve(8 * (size_t)(n + 1)));
    CRATAC_CUDA(cudaMemcpyAsync(d_names.p, names.data(), names.size(), cudaMemcpyHostToDevice, c->stream));
    CRATAC_CUDA(cudaMemcpyAsync(d_lens.p, lens.data(), 4 * lens.size(), cudaMemcpyHostToDevice, c->stream));
    EncArgs a;
    a.n = n; a.chrom = d_chrom; a.start = d_start; a.end = d_end; a.dup = d_dup; a.seq = d_seq;
    a.names = d_names.as<char>(); a.name_len = d_lens.as<int32_t>(); a.line_off = d_off.as<int64_t>(); a.out = d_out;
    int64_t total = 0;
    if (n > 0) {
        k_encode_len<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(a);
        CRATAC_CUDA(cudaGetLastError());
        CRATAC_TRY(exclusive_scan_i64(c, a.line_off, a.line_off, n, a.line_off + n));
        CRATAC_CUDA(cudaMemcpyAsync(&total, a.line_off + n, 8, cudaMemcpyDeviceToHost, c->stream));
        CRATAC_CUDA(cudaStreamSynchronize(c->stream));
    }
    *nbytes = total;
    int rc = CRATAC_OK;
    if (d_out != nullptr && cap > 0) {
        if (total > cap) { set_error("encode buffer too small: need %lld", (long long)total); rc = CRATAC_ERR_ARG; }
        else if (n > 0) {
            k_encode_write<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(a);
            if (cudaGetLastError() != cudaSuccess) rc = CRATAC_ERR_CUDA;
            cudaStreamSynchronize(c->stream);
        }
    }
    d_names.release(); d_lens.release(); d_off.release();
    return rc;
}

}  // namespace cratac
