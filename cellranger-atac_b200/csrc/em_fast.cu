// K4, on-chip variant: the same fit as em.cu (peaks.py:28-193, em_fitting.py:10-87).  One CTA of 512 threads;
// the histogram (value, weight, log-factorial, background weight per bin) lives in shared memory for up to
// 4608 bins (global scratch beyond), as does the tail-weight table T(j), so an EM iteration of a typical
// histogram touches HBM not at all.  The E-step and the sums of the M-step are fused: responsibilities are
// never stored, only the background weight w*r1.  The dispersion fixed point uses the interpolate / compose /
// descend scheme described in em.cu, restructured so that every step uses the whole CTA:
//   * F = log G(exp x) at 64 Chebyshev-Gauss nodes + 16 check points in one pass (FP64-pipe bound)
//   * the interpolant is kept as node values and evaluated with the barycentric formula by groups of 8 lanes
//     (64 evaluations per pass): composing F^(2^L) with itself is one evaluation per node, no coefficient fit
//   * the stopping point x_c (where the step size crosses 1e-6) located once by 6 rounds of 64-way sectioning
//     on the level-0 interpolant; the descent then needs one evaluation per level
//   * the last iterations and the stopping test evaluated directly, exactly as the reference does.
// The control flow is written so that every large piece (E-step, direct iteration, jump) has ONE call site:
// the kernel is latency bound on a single SM and a body that overflows the instruction cache costs 2-3x.
// Histograms outside the limits, or any failed safety check, are left to k_em_fit (em.cu).
#include <math.h>

#include "common.cuh"
#include "nb_pmf.cuh"
#include "spec_rule.cuh"

namespace cratac {

constexpr int EF_THREADS = 512;
constexpr int EF_WARPS = EF_THREADS / 32;
constexpr int EF_NS = 4608;          // bins kept in shared memory (more: global scratch)
constexpr int EF_MAXN = 65536;       // bins this kernel handles at all
constexpr int EF_TS = 6144;          // tail-weight table entries in shared memory (more: global scratch)
constexpr int EF_MAXV = 1 << 20;
constexpr int CM = 64;          // Chebyshev nodes
constexpr int CL = 14;          // composition levels (2^13 = 8192 < 10000 iterations)
constexpr double EF_PI = 3.14159265358979323846;

struct EfArgs {
    const int64_t* values;
    const int64_t* weights;
    int64_t n_bins_host;
    int64_t* status;
    double* params_out;
    double odds_ratio;
    // global scratch for histograms that do not fit shared memory
    double* g_w; double* g_lg; double* g_o1; double* g_T;
    int* g_v;
    // speculation (api.cu run_tail): after EM iteration spec_iter of the first run the threshold of the current parameters
    // goes to status[ST_THRESHOLD_SPEC] and, behind a system-wide fence, to the host-mapped words spec_host[0..1].  Every
    // eighth iteration after that the fit extrapolates where its threshold is heading; once that is certain and is not the
    // first guess, the second guess goes to status[ST_THRESHOLD_SPEC2] and spec_host[2..3] (at most once per fit).
    volatile long long* spec_host;   // null: off
    long long spec_seq;
    int spec_iter;
};

struct EfSmem {
    double xn[CM + CM / 2];     // Chebyshev-Gauss nodes cos((i + 1/2) pi / N), descending: N = 64 at [0, 64), N = 32 at [64, 96)
    double wn[CM + CM / 2];     // their barycentric weights (-1)^i sin((i + 1/2) pi / N)
    double T[EF_TS];
    const double* Tp;           // tail-weight table in use (shared or global)
    double red[EF_WARPS * 8];
    double pts[2 * CM];
    double val[2 * CM];
    double lv[CL][CM];          // F^(2^L) at the nodes
    double wtot[EF_WARPS];
    double bcast[4];
    int flag;
    int ibc;
    long long cyc[12];
};

enum { EP_ESTEP = 0, EP_MSUMS, EP_TFILL, EP_BRACKET, EP_NODES, EP_XC, EP_DOUBLING, EP_DESCENT, EP_DIRECT, EP_NCOEF, EP_LEVELS, EP_COUNT };
#define EP_BEGIN() long long _ep_t0 = clock64()
#define EP_END(s, ph) do { long long _t = clock64(); if (threadIdx.x == 0) (s).cyc[ph] += _t - _ep_t0; _ep_t0 = _t; } while (0)

// the kernel's dynamic shared memory starts with the EfSmem block; taking it from the extern array (rather than
// through a reference parameter) keeps the accesses in non-inlined device functions as LDS / STS, not generic
extern __shared__ __align__(16) unsigned char ef_raw[];
__device__ __forceinline__ EfSmem& ef_smem() { return *reinterpret_cast<EfSmem*>(ef_raw); }

struct Th { double p_zero, mean_bg, alpha_bg, p_signal, f_zero, f_bg; };

template <int K>
__device__ __forceinline__ void ef_block_sum(double (&x)[K], double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < K; j++) {
#pragma unroll
        for (int d = 16; d; d >>= 1) x[j] += __shfl_xor_sync(0xffffffffu, x[j], d);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < K; j++) red[warp * K + j] = x[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; j++) {
        double v = lane < EF_WARPS ? red[lane * K + j] : 0.0;
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        x[j] = v;
    }
}

__device__ __forceinline__ double ef_block_max(double x, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d; d >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, d));
    __syncthreads();
    if (lane == 0) red[warp] = x;
    __syncthreads();
    double v = red[lane < EF_WARPS ? lane : 0];
#pragma unroll
    for (int d = 16; d; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

// 1 / y to ~1 ulp: hardware seed (rcp.approx.ftz.f64, relative error 2^-23), one Newton step and a final
// residual correction.  A quarter of the FP64-pipe work of an IEEE division; y = 0 gives inf / NaN.
__device__ __forceinline__ double ef_rcp(double y) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    r = r * fma(-y, r, 2.0);
    return fma(r, fma(-y, r, 1.0), r);
}

// log Gamma(x): Stirling series for x >= 24 (truncation error < 1e-18 there), library lgamma below.  The E-step
// evaluates lgamma(1/alpha + v) for every bin and iteration; the series is ~4x cheaper than the library routine.
__device__ __forceinline__ double ef_lgamma(double x) {
    if (x < 24.0) return lgamma(x);
    const double r = ef_rcp(x), r2 = r * r;
    // 1/12 - r2/360 + r2^2/1260 - r2^3/1680 + r2^4/1188 - r2^5 * 691/360360
    double p = fma(r2, -691.0 / 360360.0, 1.0 / 1188.0);
    p = fma(r2, p, -1.0 / 1680.0);
    p = fma(r2, p, 1.0 / 1260.0);
    p = fma(r2, p, -1.0 / 360.0);
    p = fma(r2, p, 1.0 / 12.0);
    return fma(x - 0.5, log(x), -x) + 0.91893853320467274178 + p * r;
}

// alpha j / (1 + alpha j) = 1 - 1 / (1 + alpha j)
// out of line: the saddle-point pmf is the rare regime (dispersion at its floor) and must not lengthen the E-step body
__device__ __noinline__ double ef_nb_saddle(double k, double n, double p) { return nb_pmf_saddle(k, n, p, 1.0 - p); }

// estimate_threshold (peaks.py:43-63): largest x in [0, 1000) with sig / bg < odds, -1 when there is none.  Out of line:
// called at the end of the fit and once inside the loop (speculation), and the loop body must stay small.
__device__ __noinline__ double ef_estimate_threshold(Th th, double odds_ratio, double* red) {
    const double f_sig = 1.0 - th.f_zero - th.f_bg;
    const double nn = 1.0 / th.alpha_bg, pp = 1.0 / (1.0 + th.alpha_bg * th.mean_bg);
    double best = -1.0;
    for (int xi = threadIdx.x; xi < 1000; xi += EF_THREADS) {
        double xv = (double)xi;
        double ysig = f_sig * (pow(1.0 - th.p_signal, xv) * th.p_signal);
        double nb = nn > NB_SADDLE_MIN ? ef_nb_saddle(xv, nn, pp)
                                       : exp(lgamma(nn + xv) - lgamma(xv + 1.0) - lgamma(nn) + nn * log(pp) + xv * log1p(-pp));
        double ybg = th.f_zero * (pow(1.0 - th.p_zero, xv) * th.p_zero) + th.f_bg * nb;
        if (ysig / ybg < odds_ratio) best = fmax(best, xv);
    }
    return ef_block_max(best, red);
}
__device__ __forceinline__ void ef_publish_tentative(const EfArgs& a, long long thr, int shot) {
    a.status[shot ? ST_THRESHOLD_SPEC2 : ST_THRESHOLD_SPEC] = thr;
    __threadfence_system();
    a.spec_host[2 * shot] = thr;
    __threadfence_system();
    a.spec_host[2 * shot + 1] = a.spec_seq;
}

// estimate_threshold as a real number: *t_out = the threshold t of the current parameters (largest x < 1000 with
// sig / bg < odds, -1 when there is none) and, returned, the x in [t, t + 1) at which the log ratio -- interpolated linearly
// between t and t + 1 -- crosses log(odds).  The integer moves in steps while EM converges; this moves smoothly, so its limit
// can be extrapolated.  Only speculation reads it: the fitted threshold itself comes from ef_estimate_threshold.
__device__ __noinline__ double ef_crossing(Th th, double odds_ratio, double* red, double* t_out) {
    const double f_sig = 1.0 - th.f_zero - th.f_bg;
    const double nn = 1.0 / th.alpha_bg, pp = 1.0 / (1.0 + th.alpha_bg * th.mean_bg);
    double ratio[2] = {-1.0, -1.0};
    double best = -1.0;
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int xi = threadIdx.x + q * EF_THREADS;
        if (xi >= 1000) continue;
        const double xv = (double)xi;
        const double ysig = f_sig * (pow(1.0 - th.p_signal, xv) * th.p_signal);
        const double nb = nn > NB_SADDLE_MIN ? ef_nb_saddle(xv, nn, pp)
                                             : exp(lgamma(nn + xv) - lgamma(xv + 1.0) - lgamma(nn) + nn * log(pp) + xv * log1p(-pp));
        const double ybg = th.f_zero * (pow(1.0 - th.p_zero, xv) * th.p_zero) + th.f_bg * nb;
        ratio[q] = ysig / ybg;
        if (ratio[q] < odds_ratio) best = fmax(best, xv);
    }
    best = ef_block_max(best, red);
    *t_out = best;
    if (best < 0.0) return -1.0;
    const int t = (int)best;
    __syncthreads();
    if (threadIdx.x == 0) { red[64] = -1.0; red[65] = -1.0; }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int xi = threadIdx.x + q * EF_THREADS;
        if (xi == t) red[64] = ratio[q];
        if (xi == t + 1 && xi < 1000) red[65] = ratio[q];
    }
    __syncthreads();
    const double r0 = red[64], r1 = red[65];
    __syncthreads();
    if (!(r0 > 0.0) || !(r1 > r0) || !(r1 < 1e300)) return best + 0.5;
    return best + (log(odds_ratio) - log(r0)) / (log(r1) - log(r0));
}

__device__ __forceinline__ double ef_frac(double alpha, double j) { return 1.0 - ef_rcp(fma(alpha, j, 1.0)); }

// G(alpha) (em_fitting.py:41-56 with the sums swapped) at NP points held in s.pts, EF_THREADS / NP threads per point
template <int NP>
__device__ __forceinline__ void ef_eval_points(EfSmem& s, int jmax, double mu, double N1, int n_active = NP) {
    constexpr int TPP = EF_THREADS / NP;
    const int grp = threadIdx.x / TPP, sub = threadIdx.x % TPP;
    const double alpha = s.pts[grp < n_active ? grp : 0];
    double p0 = 0.0, p1 = 0.0;
    int j = grp < n_active ? 1 + sub : jmax;
    double jd = (double)j;                  // carried as a double: no int -> double conversion per term
    for (; j + TPP < jmax; j += 2 * TPP, jd += (double)(2 * TPP)) {
        p0 = fma(ef_frac(alpha, jd), s.Tp[j], p0);
        p1 = fma(ef_frac(alpha, jd + (double)TPP), s.Tp[j + TPP], p1);
    }
    if (j < jmax) p0 = fma(ef_frac(alpha, jd), s.Tp[j], p0);
    double part = p0 + p1;
#pragma unroll
    for (int d = TPP / 2; d; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (sub == 0 && grp < n_active) s.val[grp] = log(1.0 + alpha * mu) / (mu - part / N1);
    __syncthreads();
}

// The degree-63 interpolant through the values f[0..64) at the Chebyshev-Gauss nodes, evaluated at t by a group
// of G lanes (G = 8, 16 or 32) with the barycentric formula of the second kind: lane s owns the nodes s, s+G, ...
// The terms are independent (no recurrence), so a group finishes in a few hundred cycles; no coefficient fit is
// needed, which makes composing two interpolants one evaluation per node.  Every lane returns the value.
template <int G, int N = CM>
__device__ __forceinline__ double ef_bary_group(const EfSmem& s, const double* f, double t) {
    constexpr int TAB = N == CM ? 0 : CM;
    const double* xn = s.xn + TAB;
    const double* wn = s.wn + TAB;
    const int sub = threadIdx.x & (G - 1);
    t = fmin(1.0, fmax(-1.0, t));
    double num = 0.0, den = 0.0;
#pragma unroll
    for (int m = 0; m < N / G; m++) {
        const int i = sub + G * m;
        double d = t - xn[i];
        d = d == 0.0 ? 1e-300 : d;                   // t is a node: its term swamps the others and f[i] comes out
        const double q = wn[i] * ef_rcp(d);
        num = fma(q, f[i], num);
        den += q;
    }
#pragma unroll
    for (int d = G / 2; d; d >>= 1) {
        num += __shfl_xor_sync(0xffffffffu, num, d);
        den += __shfl_xor_sync(0xffffffffu, den, d);
    }
    return num * ef_rcp(den);
}

// G(alpha) evaluated by the whole CTA (same value in every thread)
__device__ __forceinline__ double ef_G_block(EfSmem& s, int jmax, double mu, double N1, double al) {
    const double* T = s.Tp;
    double part[1] = {0.0};
    double jd = (double)(threadIdx.x + 1);
    for (int j = threadIdx.x + 1; j < jmax; j += EF_THREADS, jd += (double)EF_THREADS) part[0] = fma(ef_frac(al, jd), T[j], part[0]);
    ef_block_sum<1>(part, s.red);
    return log(1.0 + al * mu) / (mu - part[0] / N1);
}

// The part of the jump that works on the interpolant of F = log G(exp x) over [xa, xb] with N nodes.
// Returns 1 on success, 2 when the interpolation error check failed (more nodes may do), 0 otherwise.
template <int N>
__device__ __noinline__ int ef_jump_nodes(int jmax, double mu, double N1, double alpha0, double sg, double xa, double xb, int k_offset,
                                          double* alpha_k, int* k_done) {
    EfSmem& s = ef_smem();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double half = 0.5 * (xb - xa), mid = 0.5 * (xb + xa);
    const double inv_half = 1.0 / half;
    EP_BEGIN();
    // ---- F(x) = log G(exp x) at the N Chebyshev-Gauss nodes + N / 4 check points (extrema of T_N)
    constexpr int NCHK = N / 4;
    constexpr int WG = 32;                  // lanes of a single-warp evaluation
    const double* xn = s.xn + (N == CM ? 0 : CM);
    if (tid < N) s.pts[tid] = exp(mid + half * xn[tid]);
    else if (tid < N + NCHK) s.pts[tid] = exp(mid + half * cos(EF_PI * (double)(tid - N) / (double)NCHK));
    if (tid == 0) s.flag = 1;
    __syncthreads();
    ef_eval_points<2 * N>(s, jmax, mu, N1, N + NCHK);
    double fx = 0.0;
    if (tid < N + NCHK) { fx = log(s.val[tid]); if (!isfinite(fx)) s.flag = 0; }
    __syncthreads();
    if (tid < N) s.lv[0][tid] = fx;
    else if (tid < N + NCHK) s.val[tid] = fx;
    __syncthreads();
    const double* f0 = s.lv[0];
    if (tid < N - 1) {   // secant slopes in (0, 1): monotone, contracting
        double dx = half * (xn[tid] - xn[tid + 1]);
        double df = f0[tid] - f0[tid + 1];
        if (!(df > 0.0) || !(df < dx)) s.flag = 0;
    }
    constexpr int BG = EF_THREADS / N;      // lanes per interpolant evaluation in the node-parallel passes (8 or 16)
    const int grp = tid / BG, sub = tid % BG;
    if (grp < NCHK) {     // interpolant against the true F half-way between the nodes
        double t = cos(EF_PI * (double)grp / (double)NCHK);
        double err = fabs(ef_bary_group<BG, N>(s, f0, t) - s.val[N + grp]);
        if (sub == 0 && !(err < 2e-12)) s.flag = 2;      // interpolation error: worth another try with more nodes
    }
    __syncthreads();
    if (s.flag != 1) return s.flag == 2 ? 2 : 0;
    EP_END(s, EP_NODES);
    // ---- x_c: last point on the path (from x0 towards the bracket end) whose step is still >= eps * (1 + 1e-3);
    //      4 rounds of 64-way sectioning, 8 lanes per trial point.  x_c only has to be on the early side of the
    //      crossing: a resolution of 64^-4 of the path costs at most a fraction of one extra direct iteration.
    const double x0 = log(alpha0), xe = sg > 0.0 ? xb : xa;
    const double epsx = 1e-6 * (1.0 + 1e-3);
    double lo = 0.0, hi = 1.0;
    {
        constexpr int SG = 8, NS = EF_THREADS / SG;
        const int sgrp = tid / SG, ssub = tid % SG;
        for (int round = 0; round < 4; round++) {
            double u = lo + (hi - lo) * (double)(sgrp + 1) / (double)NS;
            double x = x0 + u * (xe - x0);
            double g = sg * (ef_bary_group<SG, N>(s, f0, (x - mid) * inv_half) - x);
            int cnt = __syncthreads_count(ssub == 0 && g >= epsx);
            double nlo = lo + (hi - lo) * (double)cnt / (double)NS;
            double nhi = lo + (hi - lo) * (double)min(cnt + 1, NS) / (double)NS;
            lo = nlo; hi = nhi;
        }
    }
    const double xc = x0 + lo * (xe - x0);
    if (!(sg * (xc - x0) > 0.0)) return 0;   // the stopping point is (almost) at x0: nothing to jump over
    EP_END(s, EP_XC);
    // ---- number of composition levels from a contraction estimate (step at x0, slope at the far end);
    //      the descent below is exact whatever the estimate
    int want_levels;
    if (warp == 0) {
        const double h = 1e-3 * half;
        const int g3 = lane / 8;             // 8-lane groups 0..2 evaluate at x0, xe, xe - h towards x0
        const double xq = g3 == 0 ? x0 : (g3 == 1 ? xe : xe - sg * h);
        const double fq = ef_bary_group<8, N>(s, f0, (xq - mid) * inv_half);
        const double fx0 = __shfl_sync(0xffffffffu, fq, 0), fe = __shfl_sync(0xffffffffu, fq, 8), fh = __shfl_sync(0xffffffffu, fq, 16);
        const double d0 = fabs(fx0 - x0);
        const double slope = fabs(fe - fh) / h;
        double kest = 16.0;
        if (slope > 0.0 && slope < 1.0 && d0 > epsx) kest = log(d0 / epsx) / -log(slope);
        if (!(kest >= 1.0)) kest = 1.0;
        if (kest > 16384.0) kest = 16384.0;
        int wl = (int)ceil(log2(kest * 1.25)) + 1;
        if (wl < 2) wl = 2;
        if (wl > CL) wl = CL;
        if (lane == 0) s.ibc = wl;
    }
    __syncthreads();
    want_levels = s.ibc;
    // ---- F^(2^L) at the nodes: one interpolant evaluation per node and level
    int n_levels = 1;
    while (n_levels < want_levels) {
        const double* fl = s.lv[n_levels - 1];
        double v = ef_bary_group<BG, N>(s, fl, (fl[grp] - mid) * inv_half);
        if (sub == 0) s.lv[n_levels][grp] = v;
        __syncthreads();
        n_levels++;
    }
    EP_END(s, EP_DOUBLING);
    if (tid == 0) s.cyc[EP_LEVELS] += n_levels;
    // ---- descent: keep every jump that still lands at or before x_c
    double x = x0;
    int k = 0;
    if (warp == 0) {
        for (int L = n_levels - 1; L >= 0; L--) {
            double y = ef_bary_group<WG, N>(s, s.lv[L], (x - mid) * inv_half);
            if (sg * (xc - y) >= 0.0 && k_offset + k + (1 << L) <= 9998) { x = y; k += 1 << L; }
        }
        if (lane == 0) { s.bcast[0] = x; s.ibc = k; }
    }
    __syncthreads();
    x = s.bcast[0]; k = s.ibc;
    __syncthreads();
    EP_END(s, EP_DESCENT);
    if (k == 0) return 0;
    *alpha_k = exp(x);
    *k_done = k;
    return 1;
}


// Jump along the fixed-point iteration.  On success *alpha_k is iterate number *k_done (k_done >= 1) counted
// from alpha0.  `guess` (> 0) is the previous M-step's result: the bracket end is first looked for right next to it.
__device__ __noinline__ bool ef_jump(int jmax, double mu, double N1, double alpha0, double a1, int k_offset, double guess,
                        double* alpha_k, int* k_done) {
    EfSmem& s = ef_smem();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double sg = a1 > alpha0 ? 1.0 : -1.0;
    EP_BEGIN();
    // ---- bracket end b: a point just beyond the fixed point (G(b) - b has the opposite sign of the travel)
    double b = 0.0;
    if (guess > 0.0 && sg * (guess - alpha0) > 0.0) {
        const double c1 = guess * (1.0 + 0.02 * sg);
        if (sg * (ef_G_block(s, jmax, mu, N1, c1) - c1) <= 0.0) b = c1;
        else {
            const double c2 = guess * (1.0 + 0.15 * sg);
            if (sg * (ef_G_block(s, jmax, mu, N1, c2) - c2) <= 0.0) b = c2;
        }
    }
    if (!(b > 0.0)) {
        if (tid < CM) s.pts[tid] = alpha0 * pow(1.1, sg * (double)(tid + 1));
        __syncthreads();
        ef_eval_points<CM>(s, jmax, mu, N1);
        int hit = -1;
        for (int i = 0; i < CM; i++)
            if (sg * (s.val[i] - s.pts[i]) <= 0.0) { hit = i; break; }
        if (hit < 0) return false;
        b = s.pts[hit];
        __syncthreads();
    }
    const double xa = log(fmin(alpha0, b)), xb = log(fmax(alpha0, b));
    const double half = 0.5 * (xb - xa);
    if (!(half > 0.0) || !isfinite(half)) return false;
    EP_END(s, EP_BRACKET);
    // 32 nodes first (half the evaluations of G and half the interpolation work); 64 when its error check says so
    // (16 nodes were tried: the error check fails on every M-step of the 250 M-fragment histogram)
    int rc = ef_jump_nodes<CM / 2>(jmax, mu, N1, alpha0, sg, xa, xb, k_offset, alpha_k, k_done);
    if (rc == 2) { if (tid == 0) s.cyc[EP_NCOEF] += 1; rc = ef_jump_nodes<CM>(jmax, mu, N1, alpha0, sg, xa, xb, k_offset, alpha_k, k_done); }
    return rc == 1;
}

__global__ void __launch_bounds__(EF_THREADS) k_em_fit_fast(EfArgs a) {
    EfSmem& s = ef_smem();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n64 = a.n_bins_host >= 0 ? a.n_bins_host : a.status[ST_N_BINS];
    if (n64 > EF_MAXN) { if (tid == 0) a.status[ST_EM_MODE] = 1; return; }
    const int n = (int)n64;
    // ---- per-bin arrays: shared memory when they fit, global scratch otherwise
    double* wS; double* lgS; double* o1S; int* vS;
    if (n <= EF_NS) {
        double* base = reinterpret_cast<double*>(ef_raw + ((sizeof(EfSmem) + 15) & ~(size_t)15));
        wS = base; lgS = base + EF_NS; o1S = base + 2 * EF_NS; vS = reinterpret_cast<int*>(base + 3 * EF_NS);
    } else { wS = a.g_w; lgS = a.g_lg; o1S = a.g_o1; vS = a.g_v; }
    if (tid < 12) s.cyc[tid] = 0;
    double tot[1] = {0.0};
    double vmax = 0.0;
    for (int i = tid; i < n; i += EF_THREADS) {
        const double v = (double)a.values[i], w = (double)a.weights[i];
        vS[i] = (int)a.values[i]; wS[i] = w; lgS[i] = lgamma(v + 1.0);
        tot[0] += w;
        vmax = fmax(vmax, v);
    }
    if (tid < CM + CM / 2) {
        const int nn = tid < CM ? CM : CM / 2, i = tid < CM ? tid : tid - CM;
        const double th = EF_PI * ((double)i + 0.5) / (double)nn;
        s.xn[tid] = cos(th);
        s.wn[tid] = (i & 1) ? -sin(th) : sin(th);
    }
    ef_block_sum<1>(tot, s.red);
    vmax = ef_block_max(vmax, s.red);
    if (n < 30 || tot[0] < 1e5) {   // peaks.py:176-177
        if (tid == 0) {
            a.status[ST_THRESHOLD] = DEFAULT_THRESHOLD; a.status[ST_HAS_PARAMS] = 0; a.status[ST_EM_ITERS] = 0; a.status[ST_EM_INNER] = 0;
            a.status[ST_EM_DIRECT] = 0; a.status[ST_EM_MODE] = 2;
            for (int k = 0; k < 6; k++) a.params_out[k] = 0.0;
        }
        return;
    }
    if (vmax + 2.0 > (double)EF_MAXV) { if (tid == 0) a.status[ST_EM_MODE] = 1; return; }
    double* Tw = (vmax + 2.0 <= (double)EF_TS) ? s.T : a.g_T;   // writable view of the tail-weight table
    if (tid == 0) s.Tp = Tw;
    __syncthreads();
    const int n_slabs = (n + EF_THREADS - 1) / EF_THREADS;

    // inclusive scan over the bins of f(i), forward (prefix) or backward (suffix), result handed to g(i, value)
    auto scan_bins = [&](bool backward, auto f, auto g) {
        double carry = 0.0;
        for (int sl = 0; sl < n_slabs; sl++) {
            const int slab = backward ? n_slabs - 1 - sl : sl;
            const int i = slab * EF_THREADS + tid;
            double inc = i < n ? f(i) : 0.0;
            if (!backward) {
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { double y = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += y; }
            } else {
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { double y = __shfl_down_sync(0xffffffffu, inc, d); if (lane + d < 32) inc += y; }
            }
            __syncthreads();
            if (lane == (backward ? 0 : 31)) s.wtot[warp] = inc;
            __syncthreads();
            double off = carry, slab_sum = 0.0;
            for (int q = 0; q < EF_WARPS; q++) {
                double t = s.wtot[q];
                if (backward ? q > warp : q < warp) off += t;
                slab_sum += t;
            }
            if (i < n) g(i, off + inc);
            carry += slab_sum;
        }
    };

    // ---- simple_threshold_estimate (peaks.py:28-40)
    int t0;
    {
        double cnt[3] = {0.0, 0.0, 0.0};   // bins <= 15, bins > 15, sum of v*w over bins > 15
        for (int i = tid; i < n; i += EF_THREADS) {
            const bool above = vS[i] > DEFAULT_THRESHOLD;
            cnt[0] += above ? 0.0 : 1.0;
            cnt[1] += above ? 1.0 : 0.0;
            cnt[2] += above ? (double)vS[i] * wS[i] : 0.0;
        }
        ef_block_sum<3>(cnt, s.red);
        const double total = cnt[2];
        double best = -1.0;
        scan_bins(false, [&](int i) { return vS[i] > DEFAULT_THRESHOLD ? (double)vS[i] * wS[i] : 0.0; },
                  [&](int i, double cum) { if (vS[i] > DEFAULT_THRESHOLD && cum / total <= 0.25) best = fmax(best, (double)vS[i]); });
        best = ef_block_max(best, s.red);
        if (best > 0.0) t0 = (int)best;
        else if (cnt[1] >= 2.0) t0 = (int)a.values[(int)cnt[0] + 1];   // count_values[1]
        else { if (tid == 0) raise_error(a.status, CRATAC_ERR_EM_INDEX, n); return; }
    }

    long long iters = 0, inner = 0, direct = 0;
    int prev_inner = 1 << 20;     // iterations of the previous dispersion fixed point (first one: assume long)
    double prev_alpha = 0.0;      // its result: where the next bracket is looked for first
    bool zerodiv = false;

    // ---- fused E-step (peaks.py:66-87) + sums of the M-step (peaks.py:90-128).  hard = the initial assignment
    //      of estimate_parameters (peaks.py:152-155).  Only the background weight o1 = w * r1 is stored.
    auto em_update = [&](bool hard, Th& th) -> bool {
        EP_BEGIN();
        double acc[6] = {0, 0, 0, 0, 0, 0};
        double vmax1 = 0.0;
        const double f_sig = 1.0 - th.f_zero - th.f_bg;
        double nn = 0.0, lgn = 0.0, nlogp = 0.0, l1p = 0.0, lz = 0.0, ls = 0.0, pp = 1.0;
        bool saddle = false;      // uniform over the CTA: large 1/alpha takes the cancellation-free form (nb_pmf.cuh)
        if (!hard) {
            nn = 1.0 / th.alpha_bg;
            pp = 1.0 / (1.0 + th.alpha_bg * th.mean_bg);
            saddle = nn > NB_SADDLE_MIN;
            lgn = lgamma(nn); nlogp = nn * log(pp); l1p = log1p(-pp);
            lz = log(1.0 - th.p_zero); ls = log(1.0 - th.p_signal);
        }
        for (int i = tid; i < n; i += EF_THREADS) {
            const double v = (double)vS[i], w = wS[i];
            double r0, r1, r2;
            if (hard) {
                r0 = v <= (double)DEFAULT_THRESHOLD ? 1.0 : 0.0;
                r1 = (v <= (double)t0 && v > (double)DEFAULT_THRESHOLD) ? 1.0 : 0.0;
                r2 = v > (double)t0 ? 1.0 : 0.0;
            } else {
                const double z = exp((v - 1.0) * lz) * th.p_zero * th.f_zero;
                const double b = (saddle ? ef_nb_saddle(v, nn, pp) : exp(ef_lgamma(nn + v) - lgS[i] - lgn + nlogp + v * l1p)) * th.f_bg;
                const double g = exp((v - 1.0) * ls) * th.p_signal * f_sig;
                const double t = z + b + g;
                if (t == 0.0) { r0 = r1 = r2 = 0.0; } else { const double it = ef_rcp(t); r0 = z * it; r1 = b * it; r2 = g * it; }
            }
            const double o0 = w * r0, o1 = w * r1, o2 = w * r2;
            acc[0] += o0; acc[1] += v * o0; acc[2] += o1; acc[3] += v * o1; acc[4] += o2; acc[5] += v * o2;
            if (r1 > 0.0) vmax1 = fmax(vmax1, v);
            o1S[i] = o1;
        }
        EP_END(s, EP_ESTEP);
        ef_block_sum<6>(acc, s.red);
        vmax1 = ef_block_max(vmax1, s.red);
        const double N0 = acc[0], K0 = acc[1], N1 = acc[2], M1 = acc[3], N2 = acc[4], K2 = acc[5];
        if (N0 == 0.0 || N2 == 0.0) { zerodiv = true; return false; }
        th.p_zero = N0 / (N0 + (K0 - N0));
        th.p_signal = N2 / (N2 + (K2 - N2));
        const double mu = M1 / N1;
        th.mean_bg = mu;
        const double FT = N0 + N1 + N2;
        th.f_zero = N0 / FT;
        th.f_bg = N1 / FT;
        if (N1 == 0.0) { th.alpha_bg = 1e-6; return true; }
        double sq[1] = {0.0};
        for (int i = tid; i < n; i += EF_THREADS) { double d = mu - (double)vS[i]; sq[0] += d * d * o1S[i]; }
        ef_block_sum<1>(sq, s.red);
        const double var = sq[0] / N1;
        double alpha = fmax(1e-6, (var - mu) / (mu * mu));
        if (vmax1 > 1e8 || !(var > mu)) { th.alpha_bg = alpha; return true; }
        EP_END(s, EP_MSUMS);
        // tail weights T(j) = sum_{i: v_i > j} w_i r1_i : bin i covers j in [v_{i-1}, v_i)
        // (thread t owns the contiguous bins [t * per, (t + 1) * per): one block-wide suffix scan of the thread sums,
        // then a serial walk down the thread's own bins)
        {
            const int per = (n + EF_THREADS - 1) / EF_THREADS;
            const int b0 = min(n, tid * per), b1 = min(n, b0 + per);
            double loc = 0.0;
            for (int i = b0; i < b1; i++) loc += o1S[i];
            double inc = loc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { double y = __shfl_down_sync(0xffffffffu, inc, d); if (lane + d < 32) inc += y; }
            __syncthreads();
            if (lane == 0) s.wtot[warp] = inc;
            __syncthreads();
            double run = inc - loc;
            for (int q = warp + 1; q < EF_WARPS; q++) run += s.wtot[q];
            for (int i = b1 - 1; i >= b0; i--) {
                run += o1S[i];
                const int jlo = i > 0 ? vS[i - 1] : 0;
                for (int j = jlo; j < vS[i]; j++) Tw[j] = run;
            }
        }
        __syncthreads();
        EP_END(s, EP_TFILL);
        const int jmax = (int)vmax1;
        const double* T = s.Tp;
        auto G_block = [&](double al) {
            double part[1] = {0.0};
            double jd = (double)(tid + 1);
            for (int j = tid + 1; j < jmax; j += EF_THREADS, jd += (double)EF_THREADS) part[0] = fma(ef_frac(al, jd), T[j], part[0]);
            ef_block_sum<1>(part, s.red);
            direct++;
            return log(1.0 + al * mu) / (mu - part[0] / N1);
        };
        // Direct iterations alpha <- G(alpha) from evaluation `itc` up to `limit` evaluations; true when the
        // reference's stopping test fired.  For short tables one warp runs the whole loop on its own (no block
        // barrier per iteration) and publishes the outcome; longer tables use the block-wide reduction.
        auto plain = [&](double& al, double& nv, int& itc, int limit) -> bool {
            bool stop = false;
            if (jmax <= 512) {
                if (warp == 0) {
                    long long nd = 0;
                    while (itc < limit) {
                        double part = 0.0, part2 = 0.0;
                        int j = lane + 1;
                        for (; j + 32 < jmax; j += 64) {
                            part = fma(ef_frac(al, (double)j), T[j], part);
                            part2 = fma(ef_frac(al, (double)(j + 32)), T[j + 32], part2);
                        }
                        if (j < jmax) part = fma(ef_frac(al, (double)j), T[j], part);
                        part += part2;
#pragma unroll
                        for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                        nv = log(1.0 + al * mu) / (mu - part / N1);
                        itc++; nd++;
                        if (2.0 * fabs(nv - al) / (nv + al) < 1e-6) { stop = true; break; }
                        al = nv;
                    }
                    if (lane == 0) { s.bcast[0] = al; s.bcast[1] = nv; s.bcast[2] = (double)itc; s.bcast[3] = stop ? 1.0 : 0.0; s.ibc = (int)nd; }
                }
                __syncthreads();
                al = s.bcast[0]; nv = s.bcast[1]; itc = (int)s.bcast[2]; stop = s.bcast[3] != 0.0;
                direct += s.ibc;
                __syncthreads();
            } else {
                while (itc < limit) {
                    nv = G_block(al);
                    itc++;
                    if (2.0 * fabs(nv - al) / (nv + al) < 1e-6) { stop = true; break; }
                    al = nv;
                }
            }
            return stop;
        };
        // Plain iterations first when the previous M-step was short; otherwise go to the jump almost at once.
        // A failed jump (safety check) is retried after 64 more plain iterations, closer to the fixed point.
        // One loop, one call site each for the direct iteration and the jump (code size).
        enum { PH_PLAIN, PH_DIRECTION, PH_FINAL };
        double newv = alpha, a_prev = alpha;
        int it = 0, phase = PH_PLAIN, attempt = 0;
        int limit = prev_inner <= 64 ? 96 : 0;
        for (;;) {
            const bool stopped = plain(alpha, newv, it, limit);
            if (stopped || phase == PH_FINAL) break;
            if (phase == PH_PLAIN) {
                if (attempt >= 8 || it >= 9000) { phase = PH_FINAL; limit = 10000; continue; }
                a_prev = alpha;                          // iterate number `it`
                phase = PH_DIRECTION; limit = it + 1;    // one more evaluation: direction of travel
                continue;
            }
            double ak = a_prev;
            int kd = 0;
            EP_END(s, EP_DIRECT);
            const bool jumped = ef_jump(jmax, mu, N1, a_prev, newv, it - 1, prev_alpha, &ak, &kd);
            _ep_t0 = clock64();
            attempt++;
            if (jumped) { alpha = ak; it = it - 1 + kd; phase = PH_FINAL; limit = 10000; }
            else { phase = PH_PLAIN; limit = it + 64; }
        }
        prev_inner = it;
        prev_alpha = newv;
        EP_END(s, EP_DIRECT);
        inner += it;
        th.alpha_bg = newv;
        return true;
    };

    auto moved = [](const Th& l, const Th& c) {
        const double eps = 1e-6;
        return (fabs(l.p_zero - c.p_zero) / c.p_zero > eps) || (fabs(l.mean_bg - c.mean_bg) / c.mean_bg > eps) ||
               (fabs(l.alpha_bg - c.alpha_bg) / c.alpha_bg > eps) || (fabs(l.p_signal - c.p_signal) / c.p_signal > eps) ||
               (fabs(l.f_zero - c.f_zero) / c.f_zero > eps) || (fabs(l.f_bg - c.f_bg) / c.f_bg > eps);
    };

    // estimate_parameters (peaks.py:144-170) inside the reseeding loop of estimate_final_threshold
    // (peaks.py:183-191), written as one loop with one em_update call site
    Th th;
    th.p_zero = th.mean_bg = th.alpha_bg = th.p_signal = th.f_zero = th.f_bg = 0.0;
    bool ok = true, hard = true;
    int tries = 0;
    SpecRule rule;                         // speculation: when to tell the host which threshold the fit is heading for
    for (;;) {
        Th last = th;
        bool have_last = false;
        int i = 0;
        for (;;) {
            const bool do_hard = hard;
            if (!do_hard) {
                if (!(i < 500 && (!have_last || moved(last, th)))) break;
                i++;
                last = th;
                have_last = true;
            }
            Th nt = th;
            if (!em_update(do_hard, nt)) { ok = false; break; }   // E-step with th, M-step into nt
            if (!do_hard && nt.p_zero < nt.p_signal) {   // normalize_params (peaks.py:131-141)
                double pz = nt.p_signal, ps = nt.p_zero, fz = 1.0 - (nt.f_zero + nt.f_bg);
                nt.p_zero = pz; nt.p_signal = ps; nt.f_zero = fz;
            }
            th = nt;
            hard = false;
            // speculation: every eighth iterate the rule of spec_rule.cuh sees the threshold of the current parameters and the
            // real-valued crossing behind it, and says whether a first or a second guess goes out to the host
            if (a.spec_host && tries == 0 && spec_rule_due(rule, i, a.spec_iter)) {
                double tent;
                const double cx = ef_crossing(th, a.odds_ratio, s.red, &tent);
                long long guess = 0;
                const int action = spec_rule_step(rule, i, a.spec_iter, tent, cx, &guess);
                if (action != SPEC_NOTHING && tid == 0) ef_publish_tentative(a, guess, action == SPEC_SECOND_GUESS ? 1 : 0);
            }
        }
        iters += i;
        if (!(ok && (1.0 / th.p_zero < th.mean_bg) && tries < 3)) break;
        Th seed;
        seed.p_zero = th.p_signal; seed.mean_bg = 1.0 / th.p_zero; seed.alpha_bg = th.alpha_bg; seed.p_signal = 1.0 / th.mean_bg;
        seed.f_zero = 1.0 - (th.f_zero + th.f_bg); seed.f_bg = th.f_bg;
        th = seed;
        tries++;
    }
    if (!ok) {
        if (tid == 0) { if (zerodiv) raise_error(a.status, CRATAC_ERR_EM_ZERODIV, n); a.status[ST_EM_MODE] = 2; }
        return;
    }
    // ---- estimate_threshold (peaks.py:43-63)
    const double best = ef_estimate_threshold(th, a.odds_ratio, s.red);
    if (tid == 0) {
        a.status[ST_THRESHOLD] = best < 0.0 ? 0 : (int64_t)best;
        a.status[ST_HAS_PARAMS] = best < 0.0 ? 2 : 1;
        a.status[ST_EM_ITERS] = iters; a.status[ST_EM_INNER] = inner; a.status[ST_EM_DIRECT] = direct; a.status[ST_EM_MODE] = 2;
        a.params_out[0] = th.p_zero; a.params_out[1] = th.mean_bg; a.params_out[2] = th.alpha_bg;
        a.params_out[3] = th.p_signal; a.params_out[4] = th.f_zero; a.params_out[5] = th.f_bg;
        for (int q = 0; q < EP_COUNT; q++) a.params_out[8 + q] = (double)s.cyc[q];
    }
}

static size_t ef_smem_bytes() { return ((sizeof(EfSmem) + 15) & ~(size_t)15) + (size_t)EF_NS * (3 * 8 + 4); }

int em_fast_launch(Ctx* c, const int64_t* d_values, const int64_t* d_weights, int64_t n_bins_host, double odds_ratio) {
    static bool configured = false;
    if (!configured) {
        CRATAC_CUDA(cudaFuncSetAttribute(k_em_fit_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ef_smem_bytes()));
        configured = true;
    }
    // global scratch shared with the generic kernel's work buffer (em.cu sizes it: >= 6 * 65536 + 2^20 doubles)
    EfArgs a;
    a.values = d_values; a.weights = d_weights; a.n_bins_host = n_bins_host; a.status = c->d_status.as<int64_t>();
    a.params_out = c->d_em_params.as<double>(); a.odds_ratio = odds_ratio;
    double* w = c->d_em_work.as<double>();
    a.g_w = w; a.g_lg = w + EF_MAXN; a.g_o1 = w + 2 * EF_MAXN; a.g_v = reinterpret_cast<int*>(w + 3 * EF_MAXN);
    a.g_T = w + 6 * (size_t)EF_MAXN + 2;      // the generic kernel's T region
    a.spec_host = c->spec_armed ? reinterpret_cast<volatile long long*>(c->h_spec) : nullptr;
    a.spec_seq = c->spec_seq; a.spec_iter = c->spec_iter;
    k_em_fit_fast<<<1, EF_THREADS, ef_smem_bytes(), c->stream>>>(a);
    c->launches++;
    CRATAC_CUDA(cudaGetLastError());
    return CRATAC_OK;
}

}  // namespace cratac
