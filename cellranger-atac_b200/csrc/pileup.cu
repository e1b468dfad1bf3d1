// K2+K3 (fused) and K5: cut-site pileup, +-200 bp window sum, window-count histogram, threshold runs,
// 500 bp merge.  Replaces count_cut_sites/__init__.py:94-104 and utils.py:105-115 +
// detect_peaks/__init__.py:40-62.
//
// The dense per-base count array of the reference (int32[contig_len], 12.4 GB for GRCh38) never
// exists in HBM.  The genome is cut into tiles of TILE_T bases; a persistent CTA takes tiles from
// a ticket counter, finds the tile's fragments through a 1024-base position index (fragments are
// start-sorted), scatters both cut sites of every fragment into a shared-memory count array with
// shared-memory atomics, turns it into an inclusive prefix sum in place and reads
//     W[p] = S[p+200] - S[p-201]            (window of count_cut_sites/__init__.py:100-101)
// for the bases it owns.  Mode HIST bins W>0 (block-local bins, run-length aggregated, flushed once
// per CTA); mode PEAKS emits the boundaries of the runs W >= threshold into CTA-private regions, tile by tile
// (k_gather_events restores genome order from the per-tile counts: no CTA waits for another) plus what the zero-gap rule of
// write_chrom_bedgraph (count_cut_sites/__init__.py:36-37) needs; mode DUMP writes W (tests, bedgraph).
// Mode HIST also leaves, per tile, the largest W of its own bases and the first / last non-zero window: mode PEAKS skips
// the tiles that cannot hold a run (largest W below the threshold) without staging them.
// HBM traffic: 8 B per fragment (start, end) x (1 + halo overhead); everything per-base is on chip.
#include <algorithm>

#include "common.cuh"

namespace cratac {

constexpr int PU_THREADS = 256;
constexpr int PU_CTAS_PER_SM = 4;
constexpr int PU_CHUNK = 44;                        // 4 x odd: thread-chunked 128-bit shared accesses are bank-conflict free
constexpr int PU_N = PU_THREADS * PU_CHUNK;         // shared count array entries (11264)
constexpr int PU_PAD = 64;                          // slack for vector over-reads past the last window
constexpr int PU_HALO_LO = HALF_WINDOW + 2;         // 202
constexpr int PU_HALO_HI = HALF_WINDOW + 1;         // 201 (positions up to b + 200)
constexpr int TILE_T = PU_N - PU_HALO_LO - PU_HALO_HI + 1;   // owned bases per tile (10862)
constexpr int PU_OWN_CHUNK = PU_CHUNK;              // owned bases per thread (multiple of 4)
constexpr int PU_HB = 2048;                         // block-local histogram bins
static_assert(PU_OWN_CHUNK * PU_THREADS >= TILE_T, "owned chunking must cover the tile");
static_assert(TILE_T + 402 <= PU_N, "tile + halos must fit the shared array");

enum { MODE_HIST = 0, MODE_PEAKS = 1, MODE_DUMP = 2 };

int pileup_tile_bases() { return TILE_T; }

// everything a CTA needs to start on a tile, precomputed once per fragment set (k_tile_desc)
struct __align__(16) TileDesc {
    int32_t contig, own;        // contig id, owned bases
    int64_t ta;                 // first owned base (contig coordinate)
    int64_t i_lo, i_hi;         // fragment index range whose cut sites can fall into the tile's window
};

struct PileupArgs {
    const TileDesc* desc;
    const int32_t* start;
    const int32_t* end;
    const int64_t* frag_off;        // per contig
    const int64_t* pos_index;
    const int64_t* pidx_off;        // per contig
    const int64_t* contig_lens;
    const int64_t* contig_base;     // global run coordinate
    const int64_t* tile_off;        // per contig (+total); non-primary contigs own no tiles
    int n_contigs;
    int64_t n_tiles;
    int64_t* status;
    unsigned long long* ghist;      // MODE_HIST
    // MODE_PEAKS
    int64_t* evt_start;             // per-CTA private regions of cta_cap entries: run boundaries in tile order
    int64_t* evt_end;
    int64_t cta_cap;
    int32_t* tile_ns;               // per tile: number of run starts / ends it emitted
    int32_t* tile_ne;
    int64_t* tile_src_s;            // per tile: where its events sit in evt_start / evt_end
    int64_t* tile_src_e;
    int thr_slot;                   // status slot holding the threshold (ST_THRESHOLD, or ST_THRESHOLD_SPEC for a speculative pass)
    int4* tile_summary;             // first_nz_pos, first_nz_val, last_nz_pos, last_nz_val
    int32_t* tile_bound;            // per tile: the largest window count of its own bases.  MODE_HIST writes it (and the summary);
                                    // MODE_PEAKS skips the tiles whose bound is below the threshold (null: no skipping)
    int64_t* fills;                 // pairs (g0, g1)
    int64_t fill_cap;
    // MODE_DUMP
    int32_t* dump_W;
    int dump_contig;
};

__device__ __forceinline__ int block_excl_scan_i32(int v, int* warp_tot, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < PU_THREADS / 32; w++) {   // 8 warps: every thread adds up the warp totals itself
        const int t = warp_tot[w];
        off += w < warp ? t : 0;
        tot += t;
    }
    if (total) *total = tot;
    int r = off + x - v;
    __syncthreads();   // warp_tot may be reused by the caller
    return r;
}

// W(k) = S[k + 402] - S[k + 1] for the bases k0 <= k < k1 of one thread (k0 a multiple of 4), four at a time
// from 128-bit shared loads: the two prefix streams are offset by 2 and 1 words from 16-byte alignment, so each
// group of four windows combines two consecutive vectors of each stream.
__global__ void k_tile_desc(PileupArgs a, TileDesc* __restrict__ out) {
    const int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= a.n_tiles) return;
    int lo = 0, hi = a.n_contigs;   // last c with tile_off[c] <= tile
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (a.tile_off[mid] <= tile) lo = mid; else hi = mid; }
    const int c = lo;
    const int64_t L = a.contig_lens[c];
    const int64_t ta = (tile - a.tile_off[c]) * (int64_t)TILE_T;
    const int64_t tb = ta + TILE_T < L ? ta + TILE_T : L;
    const int64_t base = ta - PU_HALO_LO;
    const int maxpos = (int)a.status[ST_MAX_POS_LEN], maxneg = (int)a.status[ST_MAX_NEG_LEN];
    const int64_t f0 = a.frag_off[c], f1 = a.frag_off[c + 1];
    const int64_t n_ent = a.pidx_off[c + 1] - a.pidx_off[c];
    int64_t lo_pos = base - maxpos; if (lo_pos < 0) lo_pos = 0;
    const int64_t hi_pos = base + PU_N + maxneg;                              // exclusive
    const int64_t e_lo = lo_pos >> POS_INDEX_SHIFT;
    const int64_t e_hi = (hi_pos + (1 << POS_INDEX_SHIFT) - 1) >> POS_INDEX_SHIFT;
    int64_t i_lo = e_lo >= n_ent ? f1 : a.pos_index[a.pidx_off[c] + e_lo];
    const int64_t i_hi = e_hi >= n_ent - 1 ? f1 : a.pos_index[a.pidx_off[c] + e_hi];
    if (i_lo < f0) i_lo = f0;
    TileDesc d;
    d.contig = c; d.own = (int)(tb - ta); d.ta = ta; d.i_lo = i_lo; d.i_hi = i_hi;
    out[tile] = d;
}

template <typename F>
__device__ __forceinline__ void for_each_window(const int* sc, int k0, int k1, F&& body) {
    if (k0 >= k1) return;
    const int4* U4 = reinterpret_cast<const int4*>(sc + k0 + 400);
    const int4* V4 = reinterpret_cast<const int4*>(sc + k0);
    int4 u = U4[0], v = V4[0];
    for (int k = k0, j = 1; k < k1; k += 4, j++) {
        const int4 un = U4[j], vn = V4[j];
        body(k, u.z - v.y);
        if (k + 1 < k1) body(k + 1, u.w - v.z);
        if (k + 2 < k1) body(k + 2, un.x - v.w);
        if (k + 3 < k1) body(k + 3, un.y - vn.x);
        u = un; v = vn;
    }
}

// the same for a chunk of exactly 4 * N4 bases: no bounds checks
template <int N4, typename F>
__device__ __forceinline__ void for_each_window_full(const int* sc, int k0, F&& body) {
    const int4* U4 = reinterpret_cast<const int4*>(sc + k0 + 400);
    const int4* V4 = reinterpret_cast<const int4*>(sc + k0);
    int4 u = U4[0], v = V4[0];
#pragma unroll
    for (int j = 1; j <= N4; j++) {
        const int4 un = U4[j], vn = V4[j];
        const int k = k0 + 4 * (j - 1);
        body(k, u.z - v.y);
        body(k + 1, u.w - v.z);
        body(k + 2, un.x - v.w);
        body(k + 3, un.y - vn.x);
        u = un; v = vn;
    }
}

template <int MODE>
__global__ void __launch_bounds__(PU_THREADS, PU_CTAS_PER_SM) k_pileup(PileupArgs a) {
    extern __shared__ __align__(16) int smem[];
    int* sc = smem;                       // PU_N (+ PU_PAD)
    int* hist_s = smem + PU_N + PU_PAD;   // PU_HB (MODE_HIST)
    __shared__ int warp_tot[PU_THREADS / 32];
    __shared__ int s_first_nz, s_last_nz, s_bound;
    __shared__ unsigned long long s_firstp, s_lastp;   // MODE_HIST: (absolute position << 32 | W) of the tile's first / last non-zero base
    __shared__ long long s_base_s, s_base_e;
    __shared__ int s_cur_s, s_cur_e;

    const int tid = threadIdx.x;
    if (MODE == MODE_HIST)
        for (int i = tid; i < PU_HB; i += PU_THREADS) hist_s[i] = 0;
    const int maxpos = (int)a.status[ST_MAX_POS_LEN];
    const int maxneg = (int)a.status[ST_MAX_NEG_LEN];
    long long thr_ll = a.status[a.thr_slot];
    const int thr = thr_ll < 1 ? 1 : (thr_ll > 0x7fffffffLL ? 0x7fffffff : (int)thr_ll);
    int loc_max = 0;
    if (MODE == MODE_PEAKS && tid == 0) { s_cur_s = 0; s_cur_e = 0; }

    // tiles come from a ticket counter; the ticket of the NEXT tile is taken while the current one is processed
    __shared__ long long s_next[2];   // double-buffered: one barrier per tile hand-over
    int par = 0;
    if (tid == 0) s_next[0] = (long long)atomicAdd(reinterpret_cast<unsigned long long*>(&a.status[ST_TICKET]), 1ULL);
    long long prev_tile = -1;         // MODE_HIST: tile whose non-zero summary is still in s_firstp / s_lastp
    auto flush_summary = [&]() {      // thread 0, right after a barrier that follows the tile's window walk
        int4 sm;
        sm.x = sm.z = -1; sm.y = sm.w = 0;
        if (s_lastp != 0ULL) {
            sm.x = (int)(s_firstp >> 32); sm.y = (int)(s_firstp & 0xFFFFFFFFULL);
            sm.z = (int)(s_lastp >> 32); sm.w = (int)(s_lastp & 0xFFFFFFFFULL);
        }
        a.tile_summary[prev_tile] = sm;
        a.tile_bound[prev_tile] = s_bound;     // the largest window count of the tile's own bases and the base before them (exact: the walk saw them all)
    };
    while (true) {
        __syncthreads();
        const long long tile = s_next[par];
        if (MODE == MODE_HIST && a.tile_bound && tid == 0 && prev_tile >= 0) flush_summary();
        if (tile >= a.n_tiles) break;
        if (tid == 0) {
            s_next[par ^ 1] = (long long)atomicAdd(reinterpret_cast<unsigned long long*>(&a.status[ST_TICKET]), 1ULL);
            s_first_nz = 0x7fffffff;
            s_last_nz = -1;
            s_bound = 0;
            s_firstp = 0xFFFFFFFFFFFFFFFFULL;
            s_lastp = 0ULL;
        }
        prev_tile = tile;
        par ^= 1;
        const TileDesc td = a.desc[tile];
        const int c = td.contig;
        const int64_t L = a.contig_lens[c];
        const int64_t ta = td.ta;
        const int64_t tb = ta + td.own;
        const int own = td.own;
        const int64_t base = ta - PU_HALO_LO;                                  // genome position of sc[0]

        if (MODE == MODE_DUMP && c != a.dump_contig) continue;
        // no window count of this tile can reach the threshold (bound from the histogram pass): nothing to emit, and the
        // tile's non-zero summary is already in place
        if (MODE == MODE_PEAKS && a.tile_bound && a.tile_bound[tile] < thr) continue;

        // ---- zero the count array
        // ---- issue the first fragment loads, zero the count array while they fly, then scatter
        {
            constexpr int PRE = 4;
            const int64_t i_lo = td.i_lo, i_hi = td.i_hi;
            int fs[PRE], fe[PRE];
#pragma unroll
            for (int q = 0; q < PRE; q++) {
                const int64_t i = i_lo + tid + (int64_t)q * PU_THREADS;
                fs[q] = i < i_hi ? a.start[i] : -0x40000000;
                fe[q] = i < i_hi ? a.end[i] : -0x40000000;
            }
            int4* z = reinterpret_cast<int4*>(sc);
#pragma unroll
            for (int r = tid; r < PU_N / 4; r += PU_THREADS) z[r] = make_int4(0, 0, 0, 0);
            __syncthreads();
#pragma unroll
            for (int q = 0; q < PRE; q++) {
                int64_t rs = (int64_t)fs[q] - base, re = (int64_t)fe[q] - base;
                if ((uint64_t)rs < (uint64_t)PU_N) atomicAdd(&sc[rs], 1);
                if ((uint64_t)re < (uint64_t)PU_N) atomicAdd(&sc[re], 1);
            }
            for (int64_t i = i_lo + tid + (int64_t)PRE * PU_THREADS; i < i_hi; i += PU_THREADS) {
                int64_t rs = (int64_t)a.start[i] - base;
                int64_t re = (int64_t)a.end[i] - base;
                if ((uint64_t)rs < (uint64_t)PU_N) atomicAdd(&sc[rs], 1);
                if ((uint64_t)re < (uint64_t)PU_N) atomicAdd(&sc[re], 1);
            }
        }
        __syncthreads();

        // ---- in-place inclusive prefix sum (thread-chunked, 128-bit accesses)
        {
            int4* mine = reinterpret_cast<int4*>(sc + tid * PU_CHUNK);
            int sum = 0;
#pragma unroll
            for (int j = 0; j < PU_CHUNK / 4; j++) { const int4 q = mine[j]; sum += (q.x + q.y) + (q.z + q.w); }
            int run = block_excl_scan_i32(sum, warp_tot, nullptr);
#pragma unroll
            for (int j = 0; j < PU_CHUNK / 4; j++) {
                int4 q = mine[j];
                q.x += run; q.y += q.x; q.z += q.y; q.w += q.z;
                run = q.w;
                mine[j] = q;
            }
        }
        __syncthreads();

        // W(k) for owned base ta + k:  S[k + 402] - S[k + 1]
        const int k0 = tid * PU_OWN_CHUNK;
        const int k1 = min(k0 + PU_OWN_CHUNK, own);

        if (MODE == MODE_HIST) {
            // run-length aggregated histogram update, branch-free on the common path: one predicated shared-memory
            // reduction per change of W (counts >= PU_HB go to the global histogram on a rare slow path)
            int cur = 0, runlen = 0;
            const uint32_t hist_addr = (uint32_t)__cvta_generic_to_shared(hist_s);
            auto flush_slow = [&](int v, int len) {
                if (v < HIST_CAP) atomicAdd(&a.ghist[v], (unsigned long long)len);
                else raise_error(a.status, CRATAC_ERR_CAPACITY, v);
            };
            // Common case: no window count of the tile can reach PU_HB (bounded from the prefix sums: every W of my
            // chunk is at most S[k0 + chunk + 401] - S[k0]) -> no range check, no global-max tracking, and bin 0 may
            // collect garbage (it is never read); the checked path handles the few tiles over the bound.
            const int hi = min(k0 + PU_OWN_CHUNK + 401, PU_N - 1);
            const int bound = k0 < own ? sc[hi] - sc[k0] : 0;
            const bool fast = !__syncthreads_or(bound >= PU_HB);
            int w_first = 0;          // W of my first base (captured by the walkers below at no cost)
            int wmax = 0;             // for the peak pass: the tile's largest window count.  The bound above is too loose for
                                      // that: over 256 chunks of pure background it reaches the threshold by noise alone, and
                                      // the peak pass then piles up tiles that hold no peak
            if (fast) {
                auto body = [&](int k, int w) {
                    if (k == k0) w_first = w;
                    wmax = max(wmax, w);
                    const bool changed = w != cur;
                    if (changed) asm volatile("red.shared::cta.add.u32 [%0], %1;" ::"r"(hist_addr + 4u * (uint32_t)cur), "r"(runlen) : "memory");
                    runlen = changed ? 1 : runlen + 1;
                    cur = w;
                };
                if (k1 - k0 == PU_OWN_CHUNK) for_each_window_full<PU_OWN_CHUNK / 4>(sc, k0, body);
                else for_each_window(sc, k0, k1, body);
                asm volatile("red.shared::cta.add.u32 [%0], %1;" ::"r"(hist_addr + 4u * (uint32_t)cur), "r"(runlen) : "memory");
            } else {
                for_each_window(sc, k0, k1, [&](int k, int w) {
                    if (k == k0) w_first = w;
                    const bool changed = w != cur;
                    if (changed && cur > 0) {
                        if (cur < PU_HB) asm volatile("red.shared::cta.add.u32 [%0], %1;" ::"r"(hist_addr + 4u * (uint32_t)cur), "r"(runlen) : "memory");
                        else flush_slow(cur, runlen);
                    }
                    loc_max = max(loc_max, w);
                    wmax = max(wmax, w);
                    runlen = changed ? 1 : runlen + 1;
                    cur = w;
                });
                if (cur > 0) {
                    if (cur < PU_HB) asm volatile("red.shared::cta.add.u32 [%0], %1;" ::"r"(hist_addr + 4u * (uint32_t)cur), "r"(runlen) : "memory");
                    else flush_slow(cur, runlen);
                }
            }
            if (a.tile_bound) {
                // the peak pass also closes, at the tile's first base, a run that ended on the last base of the tile before:
                // that base counts for this tile too
                if (tid == 0 && ta > 0) wmax = max(wmax, sc[401] - sc[0]);
                const int wm = __reduce_max_sync(0xffffffffu, wmax);
                if ((tid & 31) == 0 && wm > 0) atomicMax(&s_bound, wm);
                // first / last base of the tile with a non-zero window count and their counts, for the peak pass's
                // zero-gap logic.  Almost always my chunk's own ends, whose counts the walk has just produced (w_first,
                // cur); thread 0 writes the tile's summary after the next barrier (loop top).
                int f = 0x7fffffff, l = -1, wf = 0, wl = 0;
                if (k0 < k1) {
                    if (w_first != 0) { f = k0; wf = w_first; }
                    else {
                        int ff = k0;
                        while (ff < k1 && sc[ff + 402] - sc[ff + 1] == 0) ff++;
                        if (ff < k1) { f = ff; wf = sc[ff + 402] - sc[ff + 1]; }
                    }
                    if (f != 0x7fffffff) {
                        if (cur != 0) { l = k1 - 1; wl = cur; }
                        else {
                            int ll = k1 - 1;
                            while (sc[ll + 402] - sc[ll + 1] == 0) ll--;
                            l = ll; wl = sc[ll + 402] - sc[ll + 1];
                        }
                    }
                }
                const int mf = __reduce_min_sync(0xffffffffu, f), ml = __reduce_max_sync(0xffffffffu, l);
                if (ml >= 0) {
                    const int src_f = __ffs(__ballot_sync(0xffffffffu, f == mf)) - 1, src_l = __ffs(__ballot_sync(0xffffffffu, l == ml)) - 1;
                    const int vf = __shfl_sync(0xffffffffu, wf, src_f), vl = __shfl_sync(0xffffffffu, wl, src_l);
                    if ((tid & 31) == 0) {
                        atomicMin(&s_firstp, ((unsigned long long)(unsigned)(ta + mf) << 32) | (unsigned)vf);
                        atomicMax(&s_lastp, ((unsigned long long)(unsigned)(ta + ml) << 32) | (unsigned)vl);
                    }
                }
            }
        } else if (MODE == MODE_DUMP) {
            for_each_window(sc, k0, k1, [&](int k, int w) { a.dump_W[ta + k] = w; });
        } else {
            // ---- run boundaries of (W >= thr), first/last non-zero W, in-tile zero gaps
            int n_s = 0, n_e = 0;
            int first_nz = 0x7fffffff, last_nz = -1;
            int prev_w = 0;
            if (k0 < k1 && (k0 > 0 || ta > 0)) prev_w = sc[k0 + 401] - sc[k0];   // W at base ta + k0 - 1
            {
                int pw = prev_w;
                bool psig = pw >= thr, zero_drop = false;
                auto body = [&](int, int w) {
                    const bool sig = w >= thr;
                    const bool down = !sig && psig;
                    n_s += (sig && !psig);
                    n_e += down;
                    zero_drop |= down && w == 0;
                    psig = sig;
                    pw = w;
                };
                if (k1 - k0 == PU_OWN_CHUNK) for_each_window_full<PU_OWN_CHUNK / 4>(sc, k0, body);
                else for_each_window(sc, k0, k1, body);
                if (zero_drop) {
                    // W dropped from >= thr to 0 in one step somewhere in my chunk: zero gap of the bedgraph writer
                    // (rare; kept out of the unrolled loop).  Walk to the next non-zero base of this tile; equal
                    // value => the row spans the gap.
                    int pv = prev_w;
                    for (int k = k0; k < k1; k++) {
                        const int w = sc[k + 402] - sc[k + 1];
                        if (w == 0 && pv >= thr) {
                            int k2 = k + 1;
                            while (k2 < own && sc[k2 + 402] - sc[k2 + 1] == 0) k2++;
                            if (k2 < own && sc[k2 + 402] - sc[k2 + 1] == pv) {
                                unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long*>(&a.status[ST_N_FILLS]), 1ULL);
                                if ((int64_t)slot < a.fill_cap) {
                                    a.fills[2 * slot] = a.contig_base[c] + ta + k;
                                    a.fills[2 * slot + 1] = a.contig_base[c] + ta + k2;
                                } else raise_error(a.status, CRATAC_ERR_CAPACITY, -2);
                            }
                        }
                        pv = w;
                    }
                }
                // first / last base of my chunk with a non-zero window count (almost always the chunk's ends)
                if (k0 < k1) {
                    int f = k0, l = k1 - 1;
                    while (f < k1 && sc[f + 402] - sc[f + 1] == 0) f++;
                    if (f < k1) {
                        while (sc[l + 402] - sc[l + 1] == 0) l--;
                        first_nz = f; last_nz = l;
                    }
                }
                const int pw_last = pw;
                pw = pw_last;
                // the contig's last base closes an open run at L
                if (k1 == own && k0 < k1 && tb == L && pw >= thr) n_e += 1;
            }
            if (last_nz >= 0) { atomicMin(&s_first_nz, first_nz); atomicMax(&s_last_nz, last_nz); }
            // one block scan for both counters (a tile has fewer than 2^15 boundaries of either kind)
            static_assert(PU_N < (1 << 15), "packed boundary counters");
            int tot_p = 0;
            const int off_p = block_excl_scan_i32(n_s | (n_e << 16), warp_tot, &tot_p);
            const int off_s = off_p & 0xFFFF, off_e = off_p >> 16, tot_s = tot_p & 0xFFFF, tot_e = tot_p >> 16;
            // ---- this CTA appends the tile's boundaries to its private region (tile order is restored by
            //      k_gather_events from the per-tile counts; no CTA ever waits for another)
            if (tid == 0) {
                s_base_s = (long long)blockIdx.x * a.cta_cap + s_cur_s;
                s_base_e = (long long)blockIdx.x * a.cta_cap + s_cur_e;
                a.tile_ns[tile] = tot_s; a.tile_ne[tile] = tot_e;
                a.tile_src_s[tile] = s_base_s; a.tile_src_e[tile] = s_base_e;
                if ((long long)s_cur_s + tot_s > a.cta_cap || (long long)s_cur_e + tot_e > a.cta_cap) {
                    raise_error(a.status, CRATAC_ERR_CAPACITY, -1);
                    a.tile_ns[tile] = 0; a.tile_ne[tile] = 0;
                    s_base_s = -1;
                } else { s_cur_s += tot_s; s_cur_e += tot_e; }
                int4 sm;
                sm.x = sm.z = -1; sm.y = sm.w = 0;
                if (s_last_nz >= 0) {
                    int f = s_first_nz, l = s_last_nz;
                    sm.x = (int)(ta + f); sm.y = sc[f + 402] - sc[f + 1];
                    sm.z = (int)(ta + l); sm.w = sc[l + 402] - sc[l + 1];
                }
                a.tile_summary[tile] = sm;
            }
            __syncthreads();
            if (n_s + n_e > 0 && s_base_s >= 0) {
                const int64_t g0 = a.contig_base[c] + ta;
                int64_t os = s_base_s + off_s;
                int64_t oe = s_base_e + off_e;
                int pw = prev_w;
                for (int k = k0; k < k1; k++) {
                    int w = sc[k + 402] - sc[k + 1];
                    bool sig = w >= thr, psig = pw >= thr;
                    if (sig && !psig) a.evt_start[os++] = g0 + k;
                    if (!sig && psig) a.evt_end[oe++] = g0 + k;
                    pw = w;
                }
                if (k1 == own && tb == L && pw >= thr) a.evt_end[oe] = g0 + own;
            }
        }
    }

    if (MODE == MODE_HIST) {
        __syncthreads();
        for (int i = tid; i < PU_HB; i += PU_THREADS) {     // bin 0 is scratch of the unchecked path
            int v = hist_s[i];
            if (v && i > 0) { atomicAdd(&a.ghist[i], (unsigned long long)v); loc_max = max(loc_max, i); }
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) loc_max = max(loc_max, __shfl_xor_sync(0xffffffffu, loc_max, d));
        if ((tid & 31) == 0 && loc_max > 0) atomicMax(reinterpret_cast<long long*>(&a.status[ST_MAX_COUNT]), (long long)loc_max);
    }
}

// histogram -> compact (value, weight) arrays, ascending.  One CTA.
__global__ void __launch_bounds__(1024) k_hist_compact(const unsigned long long* __restrict__ ghist, int64_t* status,
                                                        int64_t* __restrict__ values, int64_t* __restrict__ weights, int64_t cap) {
    __shared__ int warp_tot[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t maxc = status[ST_MAX_COUNT];
    if (maxc >= HIST_CAP) maxc = HIST_CAP - 1;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 <= maxc; b0 += 1024) {
        int64_t v = b0 + tid;
        unsigned long long w = (v <= maxc && v > 0) ? ghist[v] : 0ULL;
        int flag = w != 0ULL;
        int x = flag;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int t = warp_tot[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += y; }
            warp_tot[lane] = t;
        }
        __syncthreads();
        int pos = s_base + (warp ? warp_tot[warp - 1] : 0) + x - flag;
        if (flag && pos < cap) { values[pos] = v; weights[pos] = (int64_t)w; }
        __syncthreads();
        if (tid == 0) s_base += warp_tot[31];
        __syncthreads();
    }
    if (tid == 0) status[ST_N_BINS] = s_base;
}

// ---- run list post-processing -----------------------------------------------------------------
// totals of the two per-tile scans must agree: every run has one start and one end
__global__ void k_finish_runs(int64_t* status) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (status[ST_N_RUNS] != status[ST_TICKET2]) raise_error(status, CRATAC_ERR_INTERNAL, -3);
    }
}

// one warp per tile: copy the tile's boundaries from the CTA-private region to their global rank
__global__ void __launch_bounds__(256) k_gather_events(const int64_t* __restrict__ evt, const int32_t* __restrict__ tile_n,
                                                        const int64_t* __restrict__ tile_src, const int64_t* __restrict__ tile_dst,
                                                        int64_t n_tiles, int64_t cap, int64_t* __restrict__ out, int64_t* status) {
    const int lane = threadIdx.x & 31;
    int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_tiles) return;
    int n = tile_n[t];
    int64_t src = tile_src[t], dst = tile_dst[t];
    for (int k = lane; k < n; k += 32) {
        if (dst + k < cap) out[dst + k] = evt[src + k];
        else raise_error(status, CRATAC_ERR_CAPACITY, -1);
    }
}

// zero gaps that span tile boundaries: last non-zero W of tile t vs first non-zero W of the next non-empty tile
__global__ void k_cross_tile_gaps(const int4* __restrict__ summary, const int64_t* __restrict__ tile_off, const int64_t* __restrict__ contig_base,
                                  int n_contigs, int64_t n_tiles, int64_t* status, int64_t* fills, int64_t fill_cap, int thr_slot) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    int4 me = summary[t];
    long long thr_ll = status[thr_slot];
    int thr = thr_ll < 1 ? 1 : (thr_ll > 0x7fffffffLL ? 0x7fffffff : (int)thr_ll);
    if (me.z < 0 || me.w < thr) return;
    int lo = 0, hi = n_contigs;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (tile_off[mid] <= t) lo = mid; else hi = mid; }
    int c = lo;
    int64_t t_end = tile_off[c + 1];
    for (int64_t u = t + 1; u < t_end; u++) {
        int4 nx = summary[u];
        if (nx.x < 0) continue;
        if (nx.x - me.z > 1 && nx.y == me.w) {
            unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long*>(&status[ST_N_FILLS]), 1ULL);
            if ((int64_t)slot < fill_cap) { fills[2 * slot] = contig_base[c] + me.z + 1; fills[2 * slot + 1] = contig_base[c] + nx.x; }
            else raise_error(status, CRATAC_ERR_CAPACITY, -2);
        }
        break;
    }
}

// bridge[i] = 1 when a filled zero gap joins run i and run i+1
__global__ void k_mark_bridges(const int64_t* __restrict__ fills, const int64_t* __restrict__ run_end, const int64_t* __restrict__ run_start,
                               int64_t* status, int32_t* bridge, int64_t fill_cap) {
    int64_t nf = status[ST_N_FILLS];
    if (nf > fill_cap) nf = fill_cap;
    int64_t n = status[ST_N_RUNS];
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
        int64_t g0 = fills[2 * f], g1 = fills[2 * f + 1];
        int64_t lo = 0, hi = n;   // first run with end >= g0
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (run_end[mid] < g0) lo = mid + 1; else hi = mid; }
        if (lo + 1 < n && run_end[lo] == g0 && run_start[lo + 1] == g1) bridge[lo] = 1;
        else raise_error(status, CRATAC_ERR_INTERNAL, -4);
    }
}

// head[i] = 1 when run i opens a new peak: detect_peaks/__init__.py:53 (pos - 500 <= window_end keeps merging)
__global__ void k_run_heads(const int64_t* __restrict__ run_start, const int64_t* __restrict__ run_end, const int32_t* __restrict__ bridge,
                            const int64_t* __restrict__ status, int64_t cap, int32_t* __restrict__ head) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    int64_t n = status[ST_N_RUNS];
    int h = 0;
    if (i < n) h = (i == 0) || (run_start[i] - MERGE_DIST > run_end[i - 1] && !bridge[i - 1]);
    head[i] = h;
}

__global__ void k_emit_peaks(const int64_t* __restrict__ run_start, const int64_t* __restrict__ run_end, const int32_t* __restrict__ head,
                             const int64_t* __restrict__ head_rank, const int64_t* __restrict__ contig_base, int n_contigs, int64_t* status,
                             int64_t cap, int32_t* __restrict__ pk_chrom, int32_t* __restrict__ pk_start, int32_t* __restrict__ pk_end,
                             int32_t* __restrict__ pk_row) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n = status[ST_N_RUNS];
    if (i >= n || i >= cap) return;
    int64_t k = head_rank[i] - (head[i] ? 0 : 1);   // exclusive rank of heads: run i belongs to peak (rank of its head)
    bool is_last = (i == n - 1) || head[i + 1];
    if (!head[i] && !is_last) return;
    int64_t g = head[i] ? run_start[i] : run_end[i];
    int lo = 0, hi = n_contigs;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (contig_base[mid] <= g - (head[i] ? 0 : 1)) lo = mid; else hi = mid; }
    if (head[i]) { pk_chrom[k] = lo; pk_start[k] = (int32_t)(run_start[i] - contig_base[lo]); pk_row[k] = (int32_t)k; }
    if (is_last) {
        int64_t ge = run_end[i];
        int lo2 = 0, hi2 = n_contigs;
        while (hi2 - lo2 > 1) { int mid = (lo2 + hi2) >> 1; if (contig_base[mid] <= ge - 1) lo2 = mid; else hi2 = mid; }
        pk_end[k] = (int32_t)(ge - contig_base[lo2]);
    }
    if (i == n - 1) status[ST_N_PEAKS] = k + 1;
}

__global__ void k_zero_peaks_if_no_runs(int64_t* status) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && status[ST_N_RUNS] == 0) status[ST_N_PEAKS] = 0;
}

// peak_off[c] = first peak (sorted by contig,start) of contig c; reads the peak count from the device
__global__ void k_peak_offsets(const int32_t* __restrict__ pk_chrom, const int64_t* __restrict__ status, int n_contigs, int64_t* __restrict__ peak_off) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_contigs) return;
    int64_t n = status[ST_N_PEAKS];
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (pk_chrom[mid] < c) lo = mid + 1; else hi = mid; }
    peak_off[c] = lo;
}

// ---- host launchers -----------------------------------------------------------------------------
int compute_max_len(Ctx* c);   // decode.cu

static int fill_args(Ctx* c, PileupArgs& a) {
    a.desc = c->d_tile_desc.as<TileDesc>();
    a.start = c->d_start.as<int32_t>(); a.end = c->d_end.as<int32_t>();
    a.frag_off = c->d_contig_frag_off.as<int64_t>(); a.pos_index = c->d_pos_index.as<int64_t>();
    a.pidx_off = c->d_contig_pidx_off.as<int64_t>(); a.contig_lens = c->d_contig_lens.as<int64_t>();
    a.contig_base = c->d_contig_base.as<int64_t>(); a.tile_off = c->d_contig_tile_off.as<int64_t>();
    a.n_contigs = c->n_contigs; a.n_tiles = c->contig_tile_off[c->n_contigs];
    a.status = c->d_status.as<int64_t>(); a.thr_slot = c->thr_slot;
    a.ghist = nullptr; a.evt_start = a.evt_end = nullptr; a.cta_cap = 0; a.tile_ns = a.tile_ne = nullptr;
    a.tile_src_s = a.tile_src_e = nullptr; a.tile_summary = nullptr; a.tile_bound = nullptr;
    a.fills = nullptr; a.fill_cap = 0; a.dump_W = nullptr; a.dump_contig = -1;
    return CRATAC_OK;
}

static size_t pileup_smem(int mode) { return sizeof(int) * (PU_N + PU_PAD + (mode == MODE_HIST ? PU_HB : 0)); }

// per-tile descriptors: valid for the current fragment set (recomputed whenever the fragments changed)
static int ensure_tile_desc(Ctx* c, PileupArgs& a) {
    if (c->tile_desc_valid) return CRATAC_OK;
    if (a.n_tiles > 0) {
        CRATAC_TRY(c->d_tile_desc.reserve(sizeof(TileDesc) * (size_t)a.n_tiles));
        a.desc = c->d_tile_desc.as<TileDesc>();
        k_tile_desc<<<(unsigned)((a.n_tiles + 255) / 256), 256, 0, c->stream>>>(a, c->d_tile_desc.as<TileDesc>());
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
    }
    c->tile_desc_valid = true;
    return CRATAC_OK;
}

template <int MODE>
static int launch_pileup(Ctx* c, PileupArgs& a) {
    static bool configured = false;
    if (!configured) {
        CRATAC_CUDA(cudaFuncSetAttribute(k_pileup<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pileup_smem(MODE)));
        configured = true;
    }
    CRATAC_TRY(ensure_tile_desc(c, a));
    int64_t* st = c->d_status.as<int64_t>();
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_TICKET], 0, sizeof(int64_t), c->stream));
    int64_t grid = std::min<int64_t>(a.n_tiles, PU_CTAS_PER_SM * NUM_SMS);
    if (grid <= 0) return CRATAC_OK;
    k_pileup<MODE><<<(unsigned)grid, PU_THREADS, pileup_smem(MODE), c->stream>>>(a);
    c->launches++;
    CRATAC_CUDA(cudaGetLastError());
    return CRATAC_OK;
}

int window_histogram(Ctx* c) {
    int64_t* st = c->d_status.as<int64_t>();
    CRATAC_TRY(c->d_hist.reserve(sizeof(unsigned long long) * HIST_CAP));
    CRATAC_TRY(c->d_hist_values.reserve(sizeof(int64_t) * HIST_CAP));
    CRATAC_TRY(c->d_hist_weights.reserve(sizeof(int64_t) * HIST_CAP));
    // only the part of the histogram that the previous run touched needs clearing, but 8 MB is cheap
    CRATAC_CUDA(cudaMemsetAsync(c->d_hist.p, 0, sizeof(unsigned long long) * HIST_CAP, c->stream));
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_MAX_COUNT], 0, sizeof(int64_t) * 2, c->stream));   // max count, n_bins
    PileupArgs a;
    fill_args(c, a);
    a.ghist = c->d_hist.as<unsigned long long>();
    {
        const size_t nt = (size_t)std::max<int64_t>(c->contig_tile_off[c->n_contigs], 1);
        CRATAC_TRY(c->d_tile_summary.reserve(16 * nt));
        CRATAC_TRY(c->d_tile_bound.reserve(4 * nt));
        a.tile_summary = c->d_tile_summary.as<int4>();
        a.tile_bound = c->d_tile_bound.as<int32_t>();
        c->tile_bound_valid = true;           // (cleared when the fragments change)
    }
    c->stage_begin(SG_PILEUP_HIST);
    CRATAC_TRY(launch_pileup<MODE_HIST>(c, a));
    c->stage_end(SG_PILEUP_HIST);
    c->stage_begin(SG_HIST_COMPACT);
    k_hist_compact<<<1, 1024, 0, c->stream>>>(a.ghist, st, c->d_hist_values.as<int64_t>(), c->d_hist_weights.as<int64_t>(), HIST_CAP);
    c->stage_end(SG_HIST_COMPACT);
    c->launches++;
    CRATAC_CUDA(cudaGetLastError());
    return CRATAC_OK;
}

// re-run the histogram compaction (after a cross-rank all-reduce of d_hist / ST_MAX_COUNT)
int compact_histogram(Ctx* c) {
    int64_t* st = c->d_status.as<int64_t>();
    c->stage_begin(SG_HIST_COMPACT);
    k_hist_compact<<<1, 1024, 0, c->stream>>>(c->d_hist.as<unsigned long long>(), st, c->d_hist_values.as<int64_t>(), c->d_hist_weights.as<int64_t>(), HIST_CAP);
    c->stage_end(SG_HIST_COMPACT);
    c->launches++;
    CRATAC_CUDA(cudaGetLastError());
    return CRATAC_OK;
}

__global__ void k_offset_rows(int32_t* __restrict__ pk_row, int64_t n, int32_t base) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pk_row[i] = (int32_t)i + base;
}
// matrix rows of this shard's peaks start at `base` in the global matrix
int offset_peak_rows(Ctx* c, int64_t base) {
    if (c->n_peaks > 0) {
        k_offset_rows<<<(unsigned)((c->n_peaks + 255) / 256), 256, 0, c->stream>>>(c->d_peak_row.as<int32_t>(), c->n_peaks, (int32_t)base);
        c->launches++;
        CRATAC_CUDA(cudaGetLastError());
    }
    return CRATAC_OK;
}

int dump_window_counts(Ctx* c, int contig, int32_t* d_W) {
    PileupArgs a;
    fill_args(c, a);
    a.dump_W = d_W; a.dump_contig = contig;
    return launch_pileup<MODE_DUMP>(c, a);
}

int call_peaks_device(Ctx* c) {
    int64_t* st = c->d_status.as<int64_t>();
    int64_t n_tiles = c->contig_tile_off[c->n_contigs];
    if (c->run_cap == 0) c->run_cap = 1 << 22;
    for (int attempt = 0;; attempt++) {
        int64_t cap = c->run_cap;
        const size_t nt = (size_t)std::max<int64_t>(n_tiles, 1);
        const int64_t n_cta = std::max<int64_t>(std::min<int64_t>(n_tiles, PU_CTAS_PER_SM * NUM_SMS), 1);
        const int64_t cta_cap = std::max<int64_t>(cap / n_cta, 64);
        // chain buffer layout: tile_ns | tile_ne (int32) then tile_src_s | tile_src_e | tile_dst_s | tile_dst_e (int64)
        CRATAC_TRY(c->d_chain.reserve(nt * (2 * 4 + 4 * 8)));
        CRATAC_TRY(c->d_tile_summary.reserve(16 * nt));
        CRATAC_TRY(c->d_run_start.reserve(8 * (size_t)cap + 8 * (size_t)(n_cta * cta_cap)));
        CRATAC_TRY(c->d_run_end.reserve(8 * (size_t)cap + 8 * (size_t)(n_cta * cta_cap)));
        CRATAC_TRY(c->d_bridge.reserve(4 * (size_t)cap * 2));          // bridge flags + head flags
        CRATAC_TRY(c->d_fill.reserve(16 * (size_t)cap / 16 + 8 * (size_t)cap));   // fills + head ranks
        CRATAC_TRY(c->d_peak_chrom.reserve(4 * (size_t)cap)); CRATAC_TRY(c->d_peak_start.reserve(4 * (size_t)cap));
        CRATAC_TRY(c->d_peak_end.reserve(4 * (size_t)cap)); CRATAC_TRY(c->d_peak_row.reserve(4 * (size_t)cap));
        CRATAC_TRY(c->d_peak_off.reserve(8 * (size_t)(c->n_contigs + 2)));
        c->peak_cap = cap;
        int64_t fill_cap = cap / 16;
        int64_t* fills = c->d_fill.as<int64_t>();
        int64_t* head_rank = fills + 2 * fill_cap;
        int32_t* bridge = c->d_bridge.as<int32_t>();
        int32_t* head = bridge + cap;
        int32_t* tile_ns = c->d_chain.as<int32_t>();
        int32_t* tile_ne = tile_ns + nt;
        int64_t* tile_src_s = reinterpret_cast<int64_t*>(tile_ne + nt);
        int64_t* tile_src_e = tile_src_s + nt;
        int64_t* tile_dst_s = tile_src_e + nt;
        int64_t* tile_dst_e = tile_dst_s + nt;
        CRATAC_CUDA(cudaMemsetAsync(c->d_chain.p, 0, nt * 8, c->stream));   // tile_ns, tile_ne
        CRATAC_CUDA(cudaMemsetAsync(bridge, 0, 4 * (size_t)cap, c->stream));
        CRATAC_CUDA(cudaMemsetAsync(&st[ST_N_RUNS], 0, sizeof(int64_t) * 3, c->stream));   // runs, fills, peaks
        // the error slot is NOT cleared here: an error of the fit queued before this pass (a mixture component without
        // weight) must still be there when the host reads the status; sync_status clears the slot when it reports one
        PileupArgs a;
        fill_args(c, a);
        int64_t* run_start = c->d_run_start.as<int64_t>();
        int64_t* run_end = c->d_run_end.as<int64_t>();
        a.evt_start = run_start + cap; a.evt_end = run_end + cap; a.cta_cap = cta_cap;
        a.tile_ns = tile_ns; a.tile_ne = tile_ne; a.tile_src_s = tile_src_s; a.tile_src_e = tile_src_e;
        a.tile_summary = c->d_tile_summary.as<int4>(); a.fills = fills; a.fill_cap = fill_cap;
        a.tile_bound = c->tile_bound_valid ? c->d_tile_bound.as<int32_t>() : nullptr;
        c->stage_begin(SG_PILEUP_PEAKS);
        CRATAC_TRY(launch_pileup<MODE_PEAKS>(c, a));
        c->stage_end(SG_PILEUP_PEAKS);
        c->stage_begin(SG_MERGE_RUNS);
        CRATAC_TRY(exclusive_scan_i32_to_i64(c, tile_ns, tile_dst_s, n_tiles, &st[ST_N_RUNS]));
        CRATAC_TRY(exclusive_scan_i32_to_i64(c, tile_ne, tile_dst_e, n_tiles, &st[ST_TICKET2]));
        k_finish_runs<<<1, 32, 0, c->stream>>>(st);
        if (n_tiles > 0) {
            unsigned gg = (unsigned)((n_tiles * 32 + 255) / 256);
            k_gather_events<<<gg, 256, 0, c->stream>>>(a.evt_start, tile_ns, tile_src_s, tile_dst_s, n_tiles, cap, run_start, st);
            k_gather_events<<<gg, 256, 0, c->stream>>>(a.evt_end, tile_ne, tile_src_e, tile_dst_e, n_tiles, cap, run_end, st);
            k_cross_tile_gaps<<<(unsigned)((n_tiles + 127) / 128), 128, 0, c->stream>>>(a.tile_summary, a.tile_off, a.contig_base, c->n_contigs, n_tiles, st,
                                                                                       fills, fill_cap, c->thr_slot);
            c->launches += 3;
        }
        k_mark_bridges<<<32, 128, 0, c->stream>>>(fills, run_end, run_start, st, bridge, fill_cap);
        k_run_heads<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>(run_start, run_end, bridge, st, cap, head);
        c->launches += 3;
        CRATAC_CUDA(cudaGetLastError());
        CRATAC_TRY(exclusive_scan_i32_to_i64(c, head, head_rank, cap, nullptr));
        k_zero_peaks_if_no_runs<<<1, 32, 0, c->stream>>>(st);
        k_emit_peaks<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>(run_start, run_end, head, head_rank, a.contig_base, c->n_contigs, st, cap,
                                                                           c->d_peak_chrom.as<int32_t>(), c->d_peak_start.as<int32_t>(),
                                                                           c->d_peak_end.as<int32_t>(), c->d_peak_row.as<int32_t>());
        k_peak_offsets<<<(c->n_contigs + 1 + 127) / 128, 128, 0, c->stream>>>(c->d_peak_chrom.as<int32_t>(), st, c->n_contigs, c->d_peak_off.as<int64_t>());
        c->stage_end(SG_MERGE_RUNS);
        c->launches += 3;
        CRATAC_CUDA(cudaGetLastError());
        // capacity overflow is only visible on the host: the caller checks status when it syncs.
        (void)attempt;
        break;
    }
    c->peaks_on_device_count = true;
    c->row_range = 0;
    return CRATAC_OK;
}

// Peaks from already-thresholded signal intervals (the DETECT_PEAKS stage receives bedgraph rows, not fragments):
// rows sorted by (contig, start), merged while next.start - 500 <= cur.end (detect_peaks/__init__.py:53).
int merge_runs_from_host(Ctx* c, int64_t n, const int32_t* chrom, const int64_t* start, const int64_t* end) {
    int64_t* st = c->d_status.as<int64_t>();
    int64_t cap = std::max<int64_t>(n, 1);
    CRATAC_TRY(c->d_run_start.reserve(8 * (size_t)cap)); CRATAC_TRY(c->d_run_end.reserve(8 * (size_t)cap));
    CRATAC_TRY(c->d_bridge.reserve(4 * (size_t)cap * 2));
    CRATAC_TRY(c->d_fill.reserve(8 * (size_t)cap));
    CRATAC_TRY(c->d_peak_chrom.reserve(4 * (size_t)cap)); CRATAC_TRY(c->d_peak_start.reserve(4 * (size_t)cap));
    CRATAC_TRY(c->d_peak_end.reserve(4 * (size_t)cap)); CRATAC_TRY(c->d_peak_row.reserve(4 * (size_t)cap));
    CRATAC_TRY(c->d_peak_off.reserve(8 * (size_t)(c->n_contigs + 2)));
    c->peak_cap = cap;
    std::vector<int64_t> gs((size_t)cap), ge((size_t)cap);
    for (int64_t i = 0; i < n; i++) {
        if (chrom[i] < 0 || chrom[i] >= c->n_contigs) { set_error("row %lld: contig index out of range", (long long)i); return CRATAC_ERR_ARG; }
        gs[i] = c->contig_base[chrom[i]] + start[i];
        ge[i] = c->contig_base[chrom[i]] + end[i];
        if (i > 0 && gs[i] < ge[i - 1]) { set_error("rows must be sorted by (contig, start) and disjoint (row %lld)", (long long)i); return CRATAC_ERR_UNSORTED; }
    }
    int32_t* bridge = c->d_bridge.as<int32_t>();
    int32_t* head = bridge + cap;
    int64_t* head_rank = c->d_fill.as<int64_t>();
    CRATAC_CUDA(cudaMemcpyAsync(c->d_run_start.p, gs.data(), 8 * (size_t)cap, cudaMemcpyHostToDevice, c->stream));
    CRATAC_CUDA(cudaMemcpyAsync(c->d_run_end.p, ge.data(), 8 * (size_t)cap, cudaMemcpyHostToDevice, c->stream));
    CRATAC_CUDA(cudaMemsetAsync(bridge, 0, 4 * (size_t)cap, c->stream));
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_N_RUNS], 0, sizeof(int64_t) * 3, c->stream));
    CRATAC_CUDA(cudaMemcpyAsync(&st[ST_N_RUNS], &n, 8, cudaMemcpyHostToDevice, c->stream));
    k_run_heads<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>(c->d_run_start.as<int64_t>(), c->d_run_end.as<int64_t>(), bridge, st, cap, head);
    c->launches++;
    CRATAC_TRY(exclusive_scan_i32_to_i64(c, head, head_rank, cap, nullptr));
    k_emit_peaks<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>(c->d_run_start.as<int64_t>(), c->d_run_end.as<int64_t>(), head, head_rank,
                                                                       c->d_contig_base.as<int64_t>(), c->n_contigs, st, cap,
                                                                       c->d_peak_chrom.as<int32_t>(), c->d_peak_start.as<int32_t>(),
                                                                       c->d_peak_end.as<int32_t>(), c->d_peak_row.as<int32_t>());
    k_peak_offsets<<<(c->n_contigs + 1 + 127) / 128, 128, 0, c->stream>>>(c->d_peak_chrom.as<int32_t>(), st, c->n_contigs, c->d_peak_off.as<int64_t>());
    c->launches += 2;
    CRATAC_CUDA(cudaGetLastError());
    CRATAC_CUDA(cudaStreamSynchronize(c->stream));   // gs / ge live on this stack frame
    c->peaks_on_device_count = true;
    c->row_range = 0;
    return CRATAC_OK;
}

}  // namespace cratac
