// K4: the peak-caller mixture fit in one CTA, fp64.  Replaces lib/python/analysis/peaks.py:28-193
// (estimate_final_threshold and everything under it) and lib/python/analysis/em_fitting.py:10-87.
//
// Model: frac_zero * Geom(p_zero) + frac_bg * NB(mean_bg, alpha_bg) + frac_sig * Geom(p_signal) fitted to
// the histogram (value v_i, weight w_i).  The operation order follows the reference (SURVEY.md
// Appendix A); reductions are tree sums instead of numpy pairwise / Python sequential sums, so
// results agree to rounding, not bit for bit (stated tolerance: threshold identical, parameters
// 1e-6 relative = the reference's own convergence epsilon).
//
// The one algebraic change: the NB dispersion fixed point (em_fitting.py:41-56) evaluates
//     sum_i w_i * cumsum_j(alpha*j/(1+alpha*j))[v_i - 1]
// once per inner iteration.  Swapping the sums gives  sum_j alpha*j/(1+alpha*j) * T(j)  with
// T(j) = sum_{i: v_i > j} w_i, a suffix sum that only changes once per M-step; the inner iteration
// is then one reduction over max(v) terms and no scan.
#include <math.h>

#include "common.cuh"
#include "nb_pmf.cuh"

namespace cratac {

constexpr int EM_THREADS = 1024;
constexpr int EM_WARPS = EM_THREADS / 32;

struct EmArgs {
    const int64_t* values;
    const int64_t* weights;
    int64_t n_bins_host;       // < 0: read from status[ST_N_BINS]
    int64_t* status;
    double* params_out;        // [6]
    double odds_ratio;
    // work arrays
    double* vd; double* wd; double* R0; double* R1; double* R2; double* Q; double* T;
    int* first_gt;
    double* costab;            // [64][64] cos(k * theta_i), Chebyshev-Gauss nodes
    int fast_fixed_point;
    int64_t work_n, work_v;
};

template <int K>
__device__ __forceinline__ void block_sum(double (&x)[K], double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < K; j++) {
#pragma unroll
        for (int d = 16; d; d >>= 1) x[j] += __shfl_xor_sync(0xffffffffu, x[j], d);
    }
    __syncthreads();   // previous readers of sm are done
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < K; j++) sm[warp * K + j] = x[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; j++) {
        double s = 0.0;
        for (int w = 0; w < EM_WARPS; w++) s += sm[w * K + j];
        x[j] = s;
    }
}

__device__ __forceinline__ double block_max(double x, double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d; d >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, d));
    __syncthreads();
    if (lane == 0) sm[warp] = x;
    __syncthreads();
    double m = sm[0];
    for (int w = 1; w < EM_WARPS; w++) m = fmax(m, sm[w]);
    return m;
}

// profiling: cycles spent per phase (thread 0's clock), written to params_out[8 + phase]
enum { PH_ESTEP = 0, PH_MSUMS, PH_SUFFIX, PH_BRACKET, PH_NODES, PH_VERIFY, PH_DOUBLING, PH_DESCENT, PH_DIRECT, PH_COUNT };
__device__ long long g_em_cycles[PH_COUNT];
#define PH_BEGIN() long long _ph_t0 = clock64()
#define PH_END(ph) do { long long _t = clock64(); if (threadIdx.x == 0) g_em_cycles[ph] += _t - _ph_t0; _ph_t0 = _t; } while (0)

struct Theta { double p_zero, mean_bg, alpha_bg, p_signal, f_zero, f_bg; };

__device__ __forceinline__ double nb_pmf(double k, double n, double p) {
    // large size (dispersion near its 1e-6 floor): saddle-point form, the log-gamma terms would cancel to 1e-9 (nb_pmf.cuh)
    if (n > NB_SADDLE_MIN) return nb_pmf_saddle(k, n, p, 1.0 - p);
    // legacy scipy nbinom._pmf: exp(gammaln(n+k) - gammaln(k+1) - gammaln(n) + n*log(p) + k*log1p(-p))
    return exp(lgamma(n + k) - lgamma(k + 1.0) - lgamma(n) + n * log(p) + k * log1p(-p));
}

// suffix sums Q[i] = sum_{i' >= i} R1[i'] * w[i'] (block-wide, n arbitrary), Q[n] = 0
__device__ void suffix_sums(const EmArgs& a, int n, double* sm) {
    // process blocks of EM_THREADS from the top; carry in shared
    __shared__ double s_carry;
    __shared__ double s_warp[EM_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_carry = 0.0; a.Q[n] = 0.0; }
    __syncthreads();
    for (int hi = n; hi > 0; hi -= EM_THREADS) {
        int i = hi - 1 - tid;            // thread 0 takes the top element
        double x = (i >= 0) ? a.R1[i] * a.wd[i] : 0.0;
        double incl = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { double y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        double off = s_carry;
        for (int w = 0; w < warp; w++) off += s_warp[w];
        if (i >= 0) a.Q[i] = off + incl;
        __syncthreads();
        if (tid == EM_THREADS - 1) s_carry = off + incl;
        __syncthreads();
    }
    (void)sm;
}


// ------------------------------------------------------------------------------------------------
// Accelerated NB-dispersion fixed point.
//
// The reference iterates alpha <- G(alpha) up to 10000 times per M-step (em_fitting.py:69-74) and
// stops at the first iterate whose relative step is < 1e-6; with linear convergence that is
// hundreds of strictly sequential evaluations.  Here F(x) = log G(exp x) is interpolated once per
// M-step by a 64-node Chebyshev series on an interval that brackets every iterate (accuracy
// checked against direct evaluation, ~1e-14), the maps F^(2), F^(4), ... are built by composing
// the interpolant with itself, and a binary descent over those powers lands a few iterations
// short of the stopping index; the last iterations and the stopping test itself are evaluated
// directly, exactly as the reference does.  Any failed safety check falls back to the plain loop.
constexpr int CH_M = 64;
constexpr int CH_LEVELS = 14;

struct FixedPointSmem {
    double pts[CH_M];
    double val[CH_M];
    double coef[CH_LEVELS][CH_M];
    int ncoef[CH_LEVELS];
    int flag;
    int imax;
};

// G(alpha) of em_fitting.py:41-56 for 64 values of alpha at once: 16 threads per value
__device__ void eval_G_points(const EmArgs& a, int jmax, double mu, double N1, FixedPointSmem& f) {
    const int grp = threadIdx.x >> 4, sub = threadIdx.x & 15;
    const double alpha = f.pts[grp];
    double part = 0.0;
    for (int j = 1 + sub; j < jmax; j += 16) {
        double aj = alpha * (double)j;
        part += aj / (1.0 + aj) * a.T[j];
    }
#pragma unroll
    for (int d = 8; d; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (sub == 0) f.val[grp] = log(1.0 + alpha * mu) / (mu - part / N1);
    __syncthreads();
}

__device__ __forceinline__ double clenshaw(const double* c, int nc, double t) {
    t = fmin(1.0, fmax(-1.0, t));
    double b1 = 0.0, b2 = 0.0;
    const double t2 = 2.0 * t;
    for (int k = nc - 1; k > 0; k--) {
        double b0 = fma(t2, b1, c[k] - b2);
        b2 = b1; b1 = b0;
    }
    return fma(t, b1, c[0] - b2);
}

// node values (one per thread < 64, in f.val) -> Chebyshev coefficients of level L
__device__ void cheb_fit(const EmArgs& a, FixedPointSmem& f, int L) {
    const int k = threadIdx.x;
    if (k == 0) f.imax = 0;
    __syncthreads();
    double c = 0.0;
    if (k < CH_M) {
        for (int i = 0; i < CH_M; i++) c = fma(f.val[i], a.costab[k * CH_M + i], c);
        c *= (k == 0 ? 1.0 : 2.0) / CH_M;
        f.coef[L][k] = c;
        if (fabs(c) > 1e-18) atomicMax(&f.imax, k);
    }
    __syncthreads();
    if (k == 0) f.ncoef[L] = f.imax + 1;
    __syncthreads();
}

// Returns true and sets *alpha_k / *k_done when the jump was made; false => caller runs the plain loop.
__device__ bool fixed_point_jump(const EmArgs& a, int jmax, double mu, double N1, double alpha0, double a1, FixedPointSmem& f,
                                 double* alpha_k, int* k_done) {
    const int tid = threadIdx.x;
    const double s = a1 > alpha0 ? 1.0 : -1.0;
    PH_BEGIN();
    // ---- bracket the fixed point: geometric candidates, then a linear refinement
    if (tid < CH_M) f.pts[tid] = alpha0 * pow(1.25, s * (double)(tid + 1));
    __syncthreads();
    eval_G_points(a, jmax, mu, N1, f);
    int hit = -1;
    for (int i = 0; i < CH_M; i++)
        if (s * (f.val[i] - f.pts[i]) <= 0.0) { hit = i; break; }
    if (hit < 0) return false;
    const double lo_c = hit == 0 ? alpha0 : f.pts[hit - 1], hi_c = f.pts[hit];
    __syncthreads();
    if (tid < CH_M) f.pts[tid] = lo_c + (hi_c - lo_c) * (double)(tid + 1) / (double)CH_M;
    __syncthreads();
    eval_G_points(a, jmax, mu, N1, f);
    double b = hi_c;
    for (int i = 0; i < CH_M; i++)
        if (s * (f.val[i] - f.pts[i]) <= 0.0) { b = f.pts[i]; break; }
    const double xa = log(fmin(alpha0, b)), xb = log(fmax(alpha0, b));
    const double half = 0.5 * (xb - xa), mid = 0.5 * (xb + xa);
    if (!(half > 0.0) || !isfinite(half)) return false;
    __syncthreads();
    PH_END(PH_BRACKET);
    // ---- interpolate F(x) = log G(exp x) at the Chebyshev-Gauss nodes t_i = cos(theta_i)
    if (tid < CH_M) f.pts[tid] = exp(mid + half * a.costab[CH_M + tid]);
    __syncthreads();
    eval_G_points(a, jmax, mu, N1, f);
    // slopes between neighbouring nodes must lie in (0, 1): monotone, contracting iteration
    if (tid == 0) f.flag = 1;
    __syncthreads();
    double fx = 0.0;
    if (tid < CH_M) {
        fx = log(f.val[tid]);
        if (!isfinite(fx)) f.flag = 0;
    }
    __syncthreads();
    if (tid < CH_M) f.val[tid] = fx;
    __syncthreads();
    if (tid < CH_M - 1) {
        double dx = half * (a.costab[CH_M + tid] - a.costab[CH_M + tid + 1]);     // > 0 (nodes descend)
        double df = f.val[tid] - f.val[tid + 1];
        if (!(df > 0.0) || !(df < dx)) f.flag = 0;
    }
    __syncthreads();
    if (!f.flag) return false;
    cheb_fit(a, f, 0);
    PH_END(PH_NODES);
    // ---- accuracy check at the Chebyshev extrema against direct evaluation
    if (tid < CH_M) f.pts[tid] = exp(mid + half * cos(3.14159265358979323846 * (double)tid / (double)CH_M));
    __syncthreads();
    eval_G_points(a, jmax, mu, N1, f);
    if (tid < CH_M) {
        double t = cos(3.14159265358979323846 * (double)tid / (double)CH_M);
        double err = fabs(clenshaw(f.coef[0], f.ncoef[0], t) - log(f.val[tid]));
        if (!(err < 2e-12)) f.flag = 0;
    }
    __syncthreads();
    if (!f.flag) return false;
    PH_END(PH_VERIFY);
    // ---- powers of two of the map, built until one jump from x0 already overshoots the stopping point
    const double x0 = log(alpha0);
    const double epsx = 1e-6 * (1.0 + 1e-3);
    const double inv_half = 1.0 / half;
    int n_levels = 1;
    while (n_levels < CH_LEVELS) {
        const double* cl = f.coef[n_levels - 1];
        const int nc = f.ncoef[n_levels - 1];
        double y0 = clenshaw(cl, nc, (x0 - mid) * inv_half);
        double step = fabs(clenshaw(f.coef[0], f.ncoef[0], (y0 - mid) * inv_half) - y0);
        if (step < epsx) break;
        double v = 0.0;
        if (tid < CH_M) {
            double y = clenshaw(cl, nc, a.costab[CH_M + tid]);
            v = clenshaw(cl, nc, (y - mid) * inv_half);
        }
        __syncthreads();
        if (tid < CH_M) f.val[tid] = v;
        __syncthreads();
        cheb_fit(a, f, n_levels);
        n_levels++;
    }
    PH_END(PH_DOUBLING);
    // ---- binary descent: keep every jump that still lands before the stopping point
    double x = x0;
    int k = 0;
    for (int L = n_levels - 1; L >= 0; L--) {
        double y = clenshaw(f.coef[L], f.ncoef[L], (x - mid) * inv_half);
        double step = fabs(clenshaw(f.coef[0], f.ncoef[0], (y - mid) * inv_half) - y);
        if (step >= epsx && k + (1 << L) <= 9998) { x = y; k += 1 << L; }
    }
    PH_END(PH_DESCENT);
    *alpha_k = exp(x);
    *k_done = k;
    return true;
}

// M-step (peaks.py:90-128).  Returns false when the reference would raise ZeroDivisionError.
__device__ double dispersion_mle(const EmArgs& a, int n, double N1, double mu, double vmax1, double* sm, FixedPointSmem& fps, long long& inner_total,
                                 long long& direct_total);

__device__ bool m_step(const EmArgs& a, int n, Theta& th, double* sm, FixedPointSmem& fps, long long& inner_total, long long& direct_total) {
    const int tid = threadIdx.x;
    PH_BEGIN();
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};   // N0 K0 N1 M1 N2 K2 vmax1
    double vmax1 = 0.0;
    for (int i = tid; i < n; i += EM_THREADS) {
        double v = a.vd[i], w = a.wd[i];
        double r0 = a.R0[i], r1 = a.R1[i], r2 = a.R2[i];
        double o0 = w * r0, o1 = w * r1, o2 = w * r2;
        acc[0] += o0; acc[1] += v * o0;
        acc[2] += o1; acc[3] += v * o1;
        acc[4] += o2; acc[5] += v * o2;
        if (r1 > 0.0) vmax1 = fmax(vmax1, v);
    }
    double s6[6] = {acc[0], acc[1], acc[2], acc[3], acc[4], acc[5]};
    block_sum<6>(s6, sm);
    vmax1 = block_max(vmax1, sm);
    double N0 = s6[0], K0 = s6[1], N1 = s6[2], M1 = s6[3], N2 = s6[4], K2 = s6[5];
    if (N0 == 0.0 || N2 == 0.0) return false;   // MLE_geometric_parameter on an empty mask: 0 / (0 + 0)
    th.p_zero = N0 / (N0 + (K0 - N0));          // em_fitting.py:86-87
    th.p_signal = N2 / (N2 + (K2 - N2));
    double mu = M1 / N1;                        // em_fitting.py:14 (NaN when the mask is empty, like numpy)
    th.mean_bg = mu;
    double FT = N0 + N1 + N2;
    th.f_zero = N0 / FT;
    th.f_bg = N1 / FT;
    th.alpha_bg = dispersion_mle(a, n, N1, mu, vmax1, sm, fps, inner_total, direct_total);
    return true;
}

// MLE_dispersion (em_fitting.py:35-74) of the component whose responsibilities are a.R1: N1 = sum of w * R1, mu its weighted
// mean, vmax1 the largest value with a positive responsibility.  Shared by the peak fit (m_step above) and the cell-caller
// fit (k_cell_fit below, which points a.R1 at the component in turn).
__device__ double dispersion_mle(const EmArgs& a, int n, double N1, double mu, double vmax1, double* sm, FixedPointSmem& fps, long long& inner_total,
                                 long long& direct_total) {
    const int tid = threadIdx.x;
    PH_BEGIN();
    if (N1 == 0.0) return 1e-6;                               // (counts <= 0).all() on an empty array
    double sq[1] = {0.0};
    for (int i = tid; i < n; i += EM_THREADS) {
        double d = mu - a.vd[i];
        sq[0] += d * d * (a.wd[i] * a.R1[i]);
    }
    block_sum<1>(sq, sm);
    double var = sq[0] / N1;
    double alpha = fmax(1e-6, (var - mu) / (mu * mu));        // MOM_dispersion
    if (vmax1 > 1e8 || !(var > mu)) return alpha;
    PH_END(PH_MSUMS);
    // tail weights T(j) = sum_{i: v_i > j} w_i R1_i  for j = 0 .. vmax1-1
    suffix_sums(a, n, sm);
    const int jmax = (int)vmax1;               // terms j = 0 .. jmax-1  (np.arange(max(counts)))
    for (int j = tid; j < jmax; j += EM_THREADS) a.T[j] = a.Q[a.first_gt[j]];
    __syncthreads();
    PH_END(PH_SUFFIX);
    double newv = alpha;
    int it = 0;
    long long direct = 0;
    auto G_block = [&](double al) {
        double part[1] = {0.0};
        for (int j = tid + 1; j < jmax; j += EM_THREADS) {    // j = 0 contributes 0
            double aj = al * (double)j;
            part[0] += aj / (1.0 + aj) * a.T[j];
        }
        block_sum<1>(part, sm);
        direct++;
        return log(1.0 + al * mu) / (mu - part[0] / N1);
    };
    bool done = false;
    if (a.fast_fixed_point) {
        // first iteration exactly as the reference; it also gives the direction of travel
        newv = G_block(alpha);
        PH_END(PH_DIRECT);
        it = 1;
        if (2.0 * fabs(newv - alpha) / (newv + alpha) < 1e-6) done = true;
        else {
            double ak = alpha;
            int kd = 0;
            if (fixed_point_jump(a, jmax, mu, N1, alpha, newv, fps, &ak, &kd) && kd > 0) { alpha = ak; it = kd; }
            else { alpha = newv; }    // no jump: carry on from iterate 1
        }
    }
    if (!done) {
        _ph_t0 = clock64();
        for (; it < 10000;) {
            newv = G_block(alpha);
            it++;
            if (2.0 * fabs(newv - alpha) / (newv + alpha) < 1e-6) break;
            alpha = newv;
        }
    }
    PH_END(PH_DIRECT);
    direct_total += direct;
    inner_total += it;
    return newv;
}

// E-step (peaks.py:66-87)
__device__ void e_step(const EmArgs& a, int n, const Theta& th) {
    PH_BEGIN();
    const double f_sig = 1.0 - th.f_zero - th.f_bg;
    const double nn = 1.0 / th.alpha_bg;
    const double pp = 1.0 / (1.0 + th.alpha_bg * th.mean_bg);
    for (int i = threadIdx.x; i < n; i += EM_THREADS) {
        double v = a.vd[i];
        double z = pow(1.0 - th.p_zero, v - 1.0) * th.p_zero * th.f_zero;
        double b = nb_pmf(v, nn, pp) * th.f_bg;
        double s = pow(1.0 - th.p_signal, v - 1.0) * th.p_signal * f_sig;
        double tot = z + b + s;
        if (tot == 0.0) { a.R0[i] = 0.0; a.R1[i] = 0.0; a.R2[i] = 0.0; }
        else { a.R0[i] = z / tot; a.R1[i] = b / tot; a.R2[i] = s / tot; }
    }
    __syncthreads();
    PH_END(PH_ESTEP);
}

__device__ __forceinline__ void normalize(Theta& t) {   // peaks.py:131-141
    if (t.p_zero < t.p_signal) {
        double pz = t.p_signal, ps = t.p_zero, fz = 1.0 - (t.f_zero + t.f_bg);
        t.p_zero = pz; t.p_signal = ps; t.f_zero = fz;
    }
}

__device__ __forceinline__ bool moved(const Theta& last, const Theta& cur, double eps) {
    // any(abs(last[k] - params[k]) / params[k] > eps)
    return (fabs(last.p_zero - cur.p_zero) / cur.p_zero > eps) || (fabs(last.mean_bg - cur.mean_bg) / cur.mean_bg > eps) ||
           (fabs(last.alpha_bg - cur.alpha_bg) / cur.alpha_bg > eps) || (fabs(last.p_signal - cur.p_signal) / cur.p_signal > eps) ||
           (fabs(last.f_zero - cur.f_zero) / cur.f_zero > eps) || (fabs(last.f_bg - cur.f_bg) / cur.f_bg > eps);
}

// estimate_parameters (peaks.py:144-170).  ok=false => ZeroDivisionError in the reference.
__device__ bool estimate_parameters(const EmArgs& a, int n, int t0, bool seeded, Theta& th, double* sm, FixedPointSmem& fps, long long& iters, long long& inner, long long& direct) {
    const int tid = threadIdx.x;
    if (!seeded) {
        for (int i = tid; i < n; i += EM_THREADS) {
            double v = a.vd[i];
            a.R0[i] = (v <= (double)DEFAULT_THRESHOLD) ? 1.0 : 0.0;
            a.R1[i] = (v <= (double)t0 && v > (double)DEFAULT_THRESHOLD) ? 1.0 : 0.0;
            a.R2[i] = (v > (double)t0) ? 1.0 : 0.0;
        }
        __syncthreads();
        if (!m_step(a, n, th, sm, fps, inner, direct)) return false;
    }
    Theta last = th;
    bool have_last = false;
    int i = 0;
    while (i < 500 && (!have_last || moved(last, th, 1e-6))) {
        i++;
        e_step(a, n, th);
        last = th;
        have_last = true;
        Theta nt = th;
        if (!m_step(a, n, nt, sm, fps, inner, direct)) return false;
        normalize(nt);
        th = nt;
    }
    iters += i;
    return true;
}

__global__ void __launch_bounds__(EM_THREADS) k_em_fit(EmArgs a) {
    __shared__ double sm[EM_WARPS * 8];
    __shared__ int s_int;
    __shared__ FixedPointSmem fps;
    const int tid = threadIdx.x;
    if (a.status[ST_EM_MODE] == 2) return;   // the register-resident kernel (em_fast.cu) already did the fit
    int64_t n64 = a.n_bins_host >= 0 ? a.n_bins_host : a.status[ST_N_BINS];
    const int n = (int)n64;
    long long iters = 0, inner = 0, direct = 0;
    if (tid == 0) for (int q = 0; q < PH_COUNT; q++) g_em_cycles[q] = 0;
    for (int q = tid; q < CH_M * CH_M; q += EM_THREADS) {
        int k = q / CH_M, i = q % CH_M;
        a.costab[q] = cos((double)k * 3.14159265358979323846 * ((double)i + 0.5) / (double)CH_M);
    }
    __syncthreads();
    // ---- gate (peaks.py:176-177)
    double tot[1] = {0.0};
    double vmax = 0.0;
    for (int i = tid; i < n; i += EM_THREADS) {
        double v = (double)a.values[i], w = (double)a.weights[i];
        a.vd[i] = v; a.wd[i] = w;
        tot[0] += w;
        vmax = fmax(vmax, v);
    }
    block_sum<1>(tot, sm);
    vmax = block_max(vmax, sm);
    if (n < 30 || tot[0] < 1e5) {
        if (tid == 0) {
            a.status[ST_THRESHOLD] = DEFAULT_THRESHOLD; a.status[ST_HAS_PARAMS] = 0;
            a.status[ST_EM_ITERS] = 0; a.status[ST_EM_INNER] = 0;
            for (int k = 0; k < 6; k++) a.params_out[k] = 0.0;
        }
        return;
    }
    if (n > a.work_n || (int64_t)vmax + 2 > a.work_v) { if (tid == 0) raise_error(a.status, CRATAC_ERR_CAPACITY, n); return; }
    // first_gt[j] = first bin index i with v_i > j   (values ascending)
    {
        const int jn = (int)vmax + 1;
        for (int j = tid; j < jn; j += EM_THREADS) {
            int lo = 0, hi = n;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (a.vd[mid] > (double)j) hi = mid; else lo = mid + 1; }
            a.first_gt[j] = lo;
        }
    }
    // ---- simple_threshold_estimate (peaks.py:28-40); thread 0, sequential exact integer cumsum
    if (tid == 0) {
        long long total = 0;
        int first_above = -1, n_above = 0;
        for (int i = 0; i < n; i++)
            if (a.values[i] > DEFAULT_THRESHOLD) { if (first_above < 0) first_above = i; n_above++; total += a.values[i] * a.weights[i]; }
        int t0 = -1;
        long long cum = 0;
        bool any = false;
        for (int i = 0; i < n; i++) {
            if (a.values[i] <= DEFAULT_THRESHOLD) continue;
            cum += a.values[i] * a.weights[i];
            if ((double)cum / (double)total <= 0.25) { t0 = (int)a.values[i]; any = true; }
        }
        if (!any) t0 = (n_above >= 2) ? (int)a.values[first_above + 1] : -1;   // count_values[1] (IndexError if missing)
        s_int = t0;
    }
    __syncthreads();
    const int t0 = s_int;
    if (t0 < 0) { if (tid == 0) raise_error(a.status, CRATAC_ERR_EM_INDEX, n); return; }

    Theta th;
    bool ok = estimate_parameters(a, n, t0, false, th, sm, fps, iters, inner, direct);
    int tries = 0;
    while (ok && (1.0 / th.p_zero < th.mean_bg) && tries < 3) {   // peaks.py:183-191
        Theta seed;
        seed.p_zero = th.p_signal; seed.mean_bg = 1.0 / th.p_zero; seed.alpha_bg = th.alpha_bg; seed.p_signal = 1.0 / th.mean_bg;
        seed.f_zero = 1.0 - (th.f_zero + th.f_bg); seed.f_bg = th.f_bg;
        th = seed;
        ok = estimate_parameters(a, n, t0, true, th, sm, fps, iters, inner, direct);
        tries++;
    }
    if (!ok) { if (tid == 0) raise_error(a.status, CRATAC_ERR_EM_ZERODIV, n); return; }

    // ---- estimate_threshold (peaks.py:43-63): largest x in [0, 1000) with sig/bg < odds
    const double f_sig = 1.0 - th.f_zero - th.f_bg;
    const double nn = 1.0 / th.alpha_bg, pp = 1.0 / (1.0 + th.alpha_bg * th.mean_bg);
    double best = -1.0;
    for (int x = tid; x < 1000; x += EM_THREADS) {
        double xv = (double)x;
        double ysig = f_sig * (pow(1.0 - th.p_signal, xv) * th.p_signal);
        double ybg = th.f_zero * (pow(1.0 - th.p_zero, xv) * th.p_zero) + th.f_bg * nb_pmf(xv, nn, pp);
        if (ysig / ybg < a.odds_ratio) best = xv;
    }
    best = block_max(best, sm);
    if (tid == 0) {
        a.status[ST_THRESHOLD] = best < 0.0 ? 0 : (int64_t)best;   // None in the reference filters nothing == threshold 0
        a.status[ST_HAS_PARAMS] = best < 0.0 ? 2 : 1;
        a.status[ST_EM_ITERS] = iters; a.status[ST_EM_INNER] = inner; a.status[ST_EM_DIRECT] = direct;
        a.params_out[0] = th.p_zero; a.params_out[1] = th.mean_bg; a.params_out[2] = th.alpha_bg;
        a.params_out[3] = th.p_signal; a.params_out[4] = th.f_zero; a.params_out[5] = th.f_bg;
        for (int q = 0; q < PH_COUNT; q++) a.params_out[8 + q] = (double)g_em_cycles[q];
    }
}

// ------------------------------------------------------------------------------------------------
// The cell caller's mixture: two negative binomials -- background barcodes and cell-containing barcodes --
// fitted to the fragments-per-barcode histogram, mro/atac/stages/processing/cell_calling/detect_cell_barcodes/__init__.py:84-201.
// Same weighted mean / MLE_dispersion as the peak fit (dispersion_mle above); E-step with the saddle-point NB pmf
// (nb_pmf.cuh: values reach tens of thousands here).  One CTA; results: params_out[0..4] = mu_noise, alpha_noise, mu_signal,
// alpha_signal, frac_noise; params_out[5] = initial threshold (simple_threshold_estimate), params_out[6] = threshold
// (estimate_threshold; -1: None), params_out[7] = EM iterations.
struct CellTheta { double mu_n, a_n, mu_s, a_s, f_n; };

__device__ __forceinline__ double cell_nb(double k, double mu, double alpha) {
    const double n = 1.0 / alpha, p = 1.0 / (1.0 + alpha * mu);
    return nb_pmf_saddle(k, n, p, 1.0 - p);
}

// weighted_parameter_estimates (:105-122) from the responsibilities in a.R0 (noise) and a.R2 (signal)
__device__ void cell_m_step(const EmArgs& a, int n, CellTheta& th, double* sm, FixedPointSmem& fps, long long& inner, long long& direct) {
    const int tid = threadIdx.x;
    double acc[4] = {0, 0, 0, 0};
    double vmax_n = 0.0, vmax_s = 0.0;
    for (int i = tid; i < n; i += EM_THREADS) {
        const double v = a.vd[i], w = a.wd[i], rn = a.R0[i], rs = a.R2[i];
        acc[0] += w * rn; acc[1] += v * (w * rn); acc[2] += w * rs; acc[3] += v * (w * rs);
        if (rn > 0.0) vmax_n = fmax(vmax_n, v);
        if (rs > 0.0) vmax_s = fmax(vmax_s, v);
    }
    block_sum<4>(acc, sm);
    vmax_n = block_max(vmax_n, sm);
    vmax_s = block_max(vmax_s, sm);
    const double Nn = acc[0], Ns = acc[2];
    CellTheta t;
    t.mu_n = acc[1] / Nn; t.mu_s = acc[3] / Ns;
    t.f_n = Nn / (Nn + Ns);
    EmArgs b = a;
    b.R1 = a.R0;
    t.a_n = dispersion_mle(b, n, Nn, t.mu_n, vmax_n, sm, fps, inner, direct);
    __syncthreads();
    b.R1 = a.R2;
    t.a_s = dispersion_mle(b, n, Ns, t.mu_s, vmax_s, sm, fps, inner, direct);
    __syncthreads();
    if (t.mu_n > t.mu_s) { CellTheta u; u.mu_n = t.mu_s; u.a_n = t.a_s; u.mu_s = t.mu_n; u.a_s = t.a_n; u.f_n = 1.0 - t.f_n; t = u; }   // normalize_parameters (:167-173)
    th = t;
}

__global__ void __launch_bounds__(EM_THREADS) k_cell_fit(EmArgs a, double odds_ratio) {
    __shared__ double sm[EM_WARPS * 8];
    __shared__ double s_d[2];
    __shared__ FixedPointSmem fps;
    const int tid = threadIdx.x;
    const int n = (int)a.n_bins_host;
    long long inner = 0, direct = 0;
    for (int q = tid; q < CH_M * CH_M; q += EM_THREADS) {
        int k = q / CH_M, i = q % CH_M;
        a.costab[q] = cos((double)k * 3.14159265358979323846 * ((double)i + 0.5) / (double)CH_M);
    }
    double vmax = 0.0;
    for (int i = tid; i < n; i += EM_THREADS) {
        a.vd[i] = (double)a.values[i]; a.wd[i] = (double)a.weights[i];
        vmax = fmax(vmax, a.vd[i]);
    }
    vmax = block_max(vmax, sm);
    if (n < 3 || n > a.work_n || (int64_t)vmax + 2 > a.work_v) { if (tid == 0) raise_error(a.status, n < 3 ? CRATAC_ERR_EM_INDEX : CRATAC_ERR_CAPACITY, n); return; }
    {
        const int jn = (int)vmax + 1;
        for (int j = tid; j < jn; j += EM_THREADS) {
            int lo = 0, hi = n;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (a.vd[mid] > (double)j) hi = mid; else lo = mid + 1; }
            a.first_gt[j] = lo;
        }
    }
    // ---- simple_threshold_estimate (:150-164), thread 0, exact integer sums
    if (tid == 0) {
        long long total = 0, cum = 0;
        for (int i = 0; i < n; i++) total += a.values[i] * a.weights[i];
        long long thr = -1;
        for (int i = 0; i < n; i++) {
            cum += a.values[i] * a.weights[i];
            if ((double)cum / (double)total <= 0.01) thr = a.values[i];
        }
        if (thr < 0) thr = a.values[1];
        if (thr < a.values[1]) thr = a.values[1];
        if (thr > a.values[n - 2]) thr = a.values[n - 2];
        s_d[0] = (double)thr;
    }
    __syncthreads();
    const double thr0 = s_d[0];
    // ---- estimate_parameters (:176-201)
    for (int i = tid; i < n; i += EM_THREADS) { a.R0[i] = a.vd[i] <= thr0 ? 1.0 : 0.0; a.R2[i] = a.vd[i] > thr0 ? 1.0 : 0.0; }
    __syncthreads();
    CellTheta th;
    cell_m_step(a, n, th, sm, fps, inner, direct);
    CellTheta last = th;
    bool have_last = false;
    int it = 0;
    auto moved = [](const CellTheta& l, const CellTheta& c) {
        const double eps = 1e-6;
        return fabs(l.mu_n - c.mu_n) / c.mu_n > eps || fabs(l.a_n - c.a_n) / c.a_n > eps || fabs(l.mu_s - c.mu_s) / c.mu_s > eps ||
               fabs(l.a_s - c.a_s) / c.a_s > eps || fabs(l.f_n - c.f_n) / c.f_n > eps;
    };
    while (it < 250 && (!have_last || moved(last, th))) {
        it++;
        for (int i = tid; i < n; i += EM_THREADS) {       // calculate_probability_distribution (:84-102)
            const double v = a.vd[i];
            const double pn = th.f_n * cell_nb(v, th.mu_n, th.a_n), ps = (1.0 - th.f_n) * cell_nb(v, th.mu_s, th.a_s);
            const double tot = pn + ps;
            if (tot == 0.0) { a.R0[i] = 0.0; a.R2[i] = 0.0; } else { a.R0[i] = pn / tot; a.R2[i] = ps / tot; }
        }
        __syncthreads();
        last = th;
        have_last = true;
        cell_m_step(a, n, th, sm, fps, inner, direct);
    }
    // ---- estimate_threshold (:125-147)
    // low = max(int(mu_noise), nbinom.ppf(0.001, n_signal, p_signal)): smallest k with cdf(k) >= 0.001, by chunks of the CTA
    double ppf = -1.0;
    {
        double carry = 0.0;
        for (int base = 0; base < (1 << 22) && ppf < 0.0; base += EM_THREADS) {
            const double pm = cell_nb((double)(base + tid), th.mu_s, th.a_s);
            double incl = pm;
            const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { double y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
            __syncthreads();
            if (lane == 31) sm[warp] = incl;
            __syncthreads();
            double off = carry, tot = 0.0;
            for (int w = 0; w < EM_WARPS; w++) { if (w < warp) off += sm[w]; tot += sm[w]; }
            const double cdf = off + incl;
            double hit = cdf >= 0.001 ? (double)(EM_THREADS - tid) : 0.0;   // block_max finds the smallest index
            hit = block_max(hit, sm);
            if (hit > 0.0) ppf = (double)(base + EM_THREADS - (int)hit);
            carry += tot;
        }
    }
    const long long low = (long long)fmax(floor(th.mu_n), ppf);
    const long long top = (long long)floor(th.mu_s);          // test counts low .. top
    double found = -1.0;
    for (long long base = low; base <= top && found < 0.0; base += EM_THREADS) {
        const long long k = base + tid;
        double hit = 0.0;
        if (k <= top) {
            const double r = ((1.0 - th.f_n) * cell_nb((double)k, th.mu_s, th.a_s)) / (th.f_n * cell_nb((double)k, th.mu_n, th.a_n));
            if (r >= odds_ratio) hit = (double)(EM_THREADS - tid);
        }
        hit = block_max(hit, sm);
        if (hit > 0.0) found = (double)(base + EM_THREADS - (long long)hit);
    }
    if (tid == 0) {
        a.params_out[0] = th.mu_n; a.params_out[1] = th.a_n; a.params_out[2] = th.mu_s; a.params_out[3] = th.a_s; a.params_out[4] = th.f_n;
        a.params_out[5] = thr0;
        a.params_out[6] = found >= 0.0 ? found : (top >= low ? (double)top : -1.0);   // the last test count when none passes; None on an empty range
        a.params_out[7] = (double)it;
    }
    (void)inner; (void)direct;
}

int cell_fit(Ctx* c, const int64_t* h_values, const int64_t* h_weights, int64_t n, double odds_ratio, double out[8]) {
    if (n < 3 || n > (1 << 16)) { set_error("cratac_fit_cell_mixture: between 3 and 65536 distinct counts are needed"); return CRATAC_ERR_ARG; }
    const int64_t work_n = 1 << 16, work_v = HIST_CAP + 2;
    size_t doubles = (size_t)work_n * 6 + 2 + (size_t)work_v + 64 * 64;
    CRATAC_TRY(c->d_em_work.reserve(doubles * 8 + (size_t)work_v * 4));
    CRATAC_TRY(c->d_em_params.reserve(8 * 32));
    CRATAC_TRY(c->d_hist_values.reserve(sizeof(int64_t) * HIST_CAP));
    CRATAC_TRY(c->d_hist_weights.reserve(sizeof(int64_t) * HIST_CAP));
    CRATAC_CUDA(cudaMemcpyAsync(c->d_hist_values.p, h_values, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CRATAC_CUDA(cudaMemcpyAsync(c->d_hist_weights.p, h_weights, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    EmArgs a;
    a.values = c->d_hist_values.as<int64_t>(); a.weights = c->d_hist_weights.as<int64_t>(); a.n_bins_host = n; a.status = c->d_status.as<int64_t>();
    a.params_out = c->d_em_params.as<double>(); a.odds_ratio = odds_ratio;
    double* w = c->d_em_work.as<double>();
    a.vd = w; w += work_n; a.wd = w; w += work_n; a.R0 = w; w += work_n; a.R1 = w; w += work_n; a.R2 = w; w += work_n;
    a.Q = w; w += work_n + 2; a.T = w; w += work_v;
    a.costab = w; w += 64 * 64;
    a.fast_fixed_point = c->em_fast ? 1 : 0;
    a.first_gt = reinterpret_cast<int*>(w);
    a.work_n = work_n; a.work_v = work_v;
    int64_t* st = c->d_status.as<int64_t>();
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_ERROR], 0, sizeof(int64_t) * 2, c->stream));
    k_cell_fit<<<1, EM_THREADS, 0, c->stream>>>(a, odds_ratio);
    c->launches++;
    CRATAC_CUDA(cudaGetLastError());
    CRATAC_CUDA(cudaMemcpyAsync(out, c->d_em_params.p, sizeof(double) * 8, cudaMemcpyDeviceToHost, c->stream));
    return sync_status(c);
}

int em_fast_launch(Ctx* c, const int64_t* d_values, const int64_t* d_weights, int64_t n_bins_host, double odds_ratio);   // em_fast.cu

static int em_launch(Ctx* c, const int64_t* d_values, const int64_t* d_weights, int64_t n_bins_host, double odds_ratio) {
    const int64_t work_n = 1 << 16, work_v = HIST_CAP + 2;
    size_t doubles = (size_t)work_n * 6 + 2 + (size_t)work_v + 64 * 64;
    CRATAC_TRY(c->d_em_work.reserve(doubles * 8 + (size_t)work_v * 4));
    CRATAC_TRY(c->d_em_params.reserve(8 * 32));
    EmArgs a;
    a.values = d_values; a.weights = d_weights; a.n_bins_host = n_bins_host; a.status = c->d_status.as<int64_t>();
    a.params_out = c->d_em_params.as<double>(); a.odds_ratio = odds_ratio;
    double* w = c->d_em_work.as<double>();
    a.vd = w; w += work_n; a.wd = w; w += work_n; a.R0 = w; w += work_n; a.R1 = w; w += work_n; a.R2 = w; w += work_n;
    a.Q = w; w += work_n + 2; a.T = w; w += work_v;
    a.costab = w; w += 64 * 64;
    a.fast_fixed_point = c->em_fast ? 1 : 0;
    int64_t* st = c->d_status.as<int64_t>();
    CRATAC_CUDA(cudaMemsetAsync(&st[ST_EM_MODE], 0, sizeof(int64_t), c->stream));
    a.first_gt = reinterpret_cast<int*>(w);
    a.work_n = work_n; a.work_v = work_v;
    c->stage_begin(SG_EM_FIT);
    if (c->em_fast >= 2) CRATAC_TRY(em_fast_launch(c, d_values, d_weights, n_bins_host, odds_ratio));
    k_em_fit<<<1, EM_THREADS, 0, c->stream>>>(a);
    c->stage_end(SG_EM_FIT);
    c->launches++;
    CRATAC_CUDA(cudaGetLastError());
    return CRATAC_OK;
}

int fit_threshold_device(Ctx* c, double odds_ratio) {
    return em_launch(c, c->d_hist_values.as<int64_t>(), c->d_hist_weights.as<int64_t>(), -1, odds_ratio);
}

int fit_threshold_arrays(Ctx* c, const int64_t* d_values, const int64_t* d_weights, int64_t n, double odds_ratio) {
    return em_launch(c, d_values, d_weights, n, odds_ratio);
}

}  // namespace cratac
