"""Restatement of the cell-caller mixture fit (TEST INFRASTRUCTURE ONLY; SURVEY 8f rank 4, oracle ahead of the kernel).

DETECT_CELL_BARCODES fits two negative binomials -- background barcodes and cell-containing barcodes -- to the
fragments-per-barcode histogram (mro/atac/stages/processing/cell_calling/detect_cell_barcodes/__init__.py:84-201) with
the same weighted_mean / MLE_dispersion as the peak caller (lib/python/analysis/em_fitting.py, restated in oracle/em.py).
``tests/golden/cell_em.json`` holds the results of the reference's own functions on seeded histograms.

Histogram convention: ``count_dict`` = {fragments per barcode: number of barcodes}.
"""
from collections import namedtuple

import numpy as np
from scipy import stats

from . import em as _em

Params = namedtuple("Params", ["mu_noise", "alpha_noise", "mu_signal", "alpha_signal", "frac_noise"])   # :50-51


def _nb(values, mu, alpha):
    return stats.nbinom.pmf(values, 1 / alpha, 1 / (1 + alpha * mu))


def e_step(values, params):                                    # calculate_probability_distribution, :84-102
    out = np.vstack((params.frac_noise * _nb(values, params.mu_noise, params.alpha_noise),
                     (1 - params.frac_noise) * _nb(values, params.mu_signal, params.alpha_signal)))
    dead = out.sum(axis=0) == 0                                 # both pmfs underflow: weight 0 for either component
    out[:, dead] = 0.5
    out /= np.sum(out, axis=0)
    out[:, dead] = 0.0
    return out


def normalize(p):                                              # normalize_parameters, :167-173
    if p.mu_noise > p.mu_signal:
        return Params(p.mu_signal, p.alpha_signal, p.mu_noise, p.alpha_noise, 1 - p.frac_noise)
    return p


def m_step(values, weights, resp):                             # weighted_parameter_estimates, :105-122
    r_noise, r_signal = resp[0, :], resp[1, :]
    m_noise, m_signal = r_noise > 0, r_signal > 0
    mu_noise = _em.wmean(values[m_noise], weights[m_noise] * r_noise[m_noise])
    mu_signal = _em.wmean(values[m_signal], weights[m_signal] * r_signal[m_signal])
    alpha_noise = _em.mle_dispersion(values[m_noise], weights[m_noise] * r_noise[m_noise])
    alpha_signal = _em.mle_dispersion(values[m_signal], weights[m_signal] * r_signal[m_signal])
    frac_noise = (r_noise * weights).sum() / (resp * weights).sum()
    return normalize(Params(mu_noise, alpha_noise, mu_signal, alpha_signal, frac_noise))


def simple_threshold_estimate(count_dict, cutoff=0.01):        # :150-164
    values, weights = _em._hist_arrays(count_dict)
    reads = values * weights
    cum = reads.cumsum() / reads.sum()
    threshold = values[cum <= cutoff][-1] if any(cum <= cutoff) else values[1]
    threshold = max(threshold, values[1])
    return min(threshold, values[-2])


def estimate_parameters(count_dict, epsilon=1e-6, maxiter=250, threshold=None):    # :176-201
    if threshold is None:
        threshold = simple_threshold_estimate(count_dict)
    values, weights = _em._hist_arrays(count_dict)
    resp = np.empty((2, len(values)), dtype=float)
    resp[0, :] = values <= threshold
    resp[1, :] = values > threshold
    params = m_step(values, weights, resp)
    last = None
    i = 0
    while i < maxiter and (last is None or any(abs(last[j] - params[j]) / params[j] > epsilon for j in range(len(params)))):
        i += 1
        resp = e_step(values, params)
        last = params
        params = m_step(values, weights, resp)
    return params


def estimate_threshold(params, odds_ratio=20):                  # :125-147
    def ratio(k):
        return ((1 - params.frac_noise) * _nb(k, params.mu_signal, params.alpha_signal)) / \
               (params.frac_noise * _nb(k, params.mu_noise, params.alpha_noise))

    low = max([int(params.mu_noise), stats.nbinom.ppf(0.001, 1 / params.alpha_signal, 1 / (1 + params.alpha_signal * params.mu_signal))])
    test_count = None
    for test_count in np.arange(low, int(params.mu_signal) + 1):
        if ratio(test_count) >= odds_ratio:
            return test_count
    return test_count
