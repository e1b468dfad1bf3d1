"""Restatement of the peak-caller mixture fit (TEST INFRASTRUCTURE ONLY).

Follows lib/python/analysis/peaks.py and lib/python/analysis/em_fitting.py of
the reference operation by operation, including which
sums are numpy pairwise ``ndarray.sum()`` and which are Python's left-to-right
``sum()``.  scipy.stats supplies geom/nbinom pmfs exactly as in the reference.
``tests/golden/make_golden.py`` runs the unmodified reference module next to
this one in the build container and the committed vectors pin both.

Histogram convention: ``values`` = sorted distinct window counts (ints >= 1),
``weights`` = number of bases with that count.
"""
from collections import namedtuple

import numpy as np
from scipy import stats

DEFAULT_THRESHOLD = 15                      # peaks.py:17

Params = namedtuple("Params", ["p_zero_noise", "mean_bg_noise", "alpha_bg_noise",
                               "p_signal", "frac_zero_noise", "frac_bg_noise"])  # peaks.py:19-25


def _hist_arrays(count_dict):
    values = np.array(sorted(k for k in count_dict))
    weights = np.array([count_dict[k] for k in values])
    return values, weights


# ---- em_fitting.py ---------------------------------------------------------
def wmean(x, w):                            # em_fitting.py:10-14
    return (x * w).sum() / w.sum()


def wvar(x, w):                             # em_fitting.py:17-22
    d2 = (wmean(x, w) - x) ** 2
    return (d2 * w).sum() / w.sum()


def mom_dispersion(x, w):                   # em_fitting.py:25-32
    mu = wmean(x, w)
    return max(1e-6, (wvar(x, w) - mu) / np.power(mu, 2))


def mle_dispersion(x, w, maxiter=10000, epsilon=1e-6, stats_out=None):   # em_fitting.py:35-74
    if (x <= 0).all():
        return 1e-6
    alpha = mom_dispersion(x, w)
    if max(x) > 1e8:
        return alpha
    mu = wmean(x, w)
    if wvar(x, w) <= mu:
        return alpha
    idx = x - 1
    idx[idx < 0] = 0
    j = np.arange(max(x))
    new = alpha
    for _ in range(maxiter):
        if stats_out is not None:
            stats_out["inner"] = stats_out.get("inner", 0) + 1
        csum = np.cumsum((alpha * j) / (1 + alpha * j))
        new = np.log(1 + alpha * mu) / (mu - wmean(csum[idx], w))
        if 2 * abs(new - alpha) / (new + alpha) < epsilon:
            break
        alpha = new
    return new


def mle_geometric(x, w):                    # em_fitting.py:77-87 (exclude_zeros=True; python sum())
    n = sum(w)
    k_total = sum(x * w) - n * 1
    return n / (n + k_total)


# ---- peaks.py ----------------------------------------------------------------
def simple_threshold_estimate(count_dict, cutoff=0.25):         # peaks.py:28-40
    v = np.array(sorted(k for k in count_dict if k > DEFAULT_THRESHOLD))
    w = np.array([count_dict[k] for k in v])
    reads = v * w
    cum_frac = reads.cumsum() / reads.sum()
    return v[cum_frac <= cutoff][-1] if any(cum_frac <= cutoff) else v[1]


def e_step(values, params):                                     # peaks.py:66-87
    z = stats.geom.pmf(values, params.p_zero_noise, loc=0)
    b = stats.nbinom.pmf(values, 1 / params.alpha_bg_noise,
                         1 / (1 + params.alpha_bg_noise * params.mean_bg_noise))
    s = stats.geom.pmf(values, params.p_signal, loc=0)
    z *= params.frac_zero_noise
    b *= params.frac_bg_noise
    s *= (1 - params.frac_zero_noise - params.frac_bg_noise)
    out = np.vstack((z, b, s))
    dead = out.sum(axis=0) == 0
    out[:, dead] = 0.5
    out /= np.sum(out, axis=0)
    out[:, dead] = 0.0
    return out


def m_step(values, weights, resp, stats_out=None):              # peaks.py:90-128
    r0, r1, r2 = resp[0, :], resp[1, :], resp[2, :]
    m0, m1, m2 = r0 > 0, r1 > 0, r2 > 0
    p_zero = mle_geometric(values[m0], weights[m0] * r0[m0])
    mean_bg = wmean(values[m1], weights[m1] * r1[m1])
    alpha_bg = mle_dispersion(values[m1], weights[m1] * r1[m1], stats_out=stats_out)
    p_sig = mle_geometric(values[m2], weights[m2] * r2[m2])
    f_zero = (r0 * weights).sum() / (resp * weights).sum()
    f_bg = (r1 * weights).sum() / (resp * weights).sum()
    return Params(p_zero, mean_bg, alpha_bg, p_sig, f_zero, f_bg)


def normalize(p):                                               # peaks.py:131-141
    if p.p_zero_noise < p.p_signal:
        return Params(p.p_signal, p.mean_bg_noise, p.alpha_bg_noise, p.p_zero_noise,
                      1 - (p.frac_zero_noise + p.frac_bg_noise), p.frac_bg_noise)
    return p


def estimate_parameters(count_dict, epsilon=1e-6, maxiter=500, seed_params=None, stats_out=None):
    """peaks.py:144-170."""
    t0 = simple_threshold_estimate(count_dict)
    values, weights = _hist_arrays(count_dict)
    resp = np.empty((3, len(values)), dtype=float)
    resp[0, :] = values <= DEFAULT_THRESHOLD
    resp[1, :] = (values <= t0) & (values > DEFAULT_THRESHOLD)
    resp[2, :] = values > t0
    params = m_step(values, weights, resp, stats_out) if seed_params is None else seed_params
    last = None
    i = 0
    while i < maxiter and (last is None or any(
            abs(last[k] - params[k]) / params[k] > epsilon for k in range(len(params)))):
        i += 1
        resp = e_step(values, params)
        last = params
        params = normalize(m_step(values, weights, resp, stats_out))
        if stats_out is not None:
            stats_out["em_iters"] = stats_out.get("em_iters", 0) + 1
    return params


def estimate_threshold(params, odds_ratio=20, maxval=1000):     # peaks.py:43-63
    x = np.arange(maxval)
    f_sig = 1 - params.frac_zero_noise - params.frac_bg_noise
    y_zero = params.frac_zero_noise * stats.geom.pmf(x, params.p_zero_noise, loc=-1)
    y_bg = params.frac_bg_noise * stats.nbinom.pmf(
        x, 1 / params.alpha_bg_noise, 1 / (1 + params.alpha_bg_noise * params.mean_bg_noise))
    y_sig = f_sig * stats.geom.pmf(x, params.p_signal, loc=-1)
    total_bg = y_zero + y_bg
    with np.errstate(divide="ignore", invalid="ignore"):
        for val in x[::-1]:
            if y_sig[val] / total_bg[val] < odds_ratio:
                return val
    return None


def estimate_final_threshold(count_dict, odds_ratio=20, stats_out=None):   # peaks.py:173-193
    if len(count_dict) < 30 or sum(count_dict.values()) < 1e5:
        return DEFAULT_THRESHOLD, None
    tries = 0
    params = estimate_parameters(count_dict, stats_out=stats_out)
    while 1 / params.p_zero_noise < params.mean_bg_noise and tries < 3:
        seed = Params(params.p_signal, 1 / params.p_zero_noise, params.alpha_bg_noise,
                      1 / params.mean_bg_noise,
                      1 - (params.frac_zero_noise + params.frac_bg_noise), params.frac_bg_noise)
        params = estimate_parameters(count_dict, seed_params=seed, stats_out=stats_out)
        tries += 1
    return estimate_threshold(params, odds_ratio), params
