"""
DETECT_PEAKS on the GPU: mixture-model threshold on the window-count histogram, then threshold + 500 bp merge.

Drop-in for mro/atac/stages/processing/detect_peaks/__init__.py of cellranger-atac 1.2.0 (same __MRO__, same
peaks.bed and peak_metrics JSON).  split = the fit (cratac_fit_threshold replaces analysis/peaks.py:173-193);
one chunk calls peaks for all contigs from the bedgraph rows (rows with count >= threshold are the signal
intervals; cratac_peaks_from_rows merges them, which is what the per-base loop of :40-62 computes);
join = sort by genome.fa.fai order + metrics.
"""
from __future__ import division

import json
import pickle

import numpy as np

from cratac import _ffi, stage_util

__MRO__ = """
stage DETECT_PEAKS(
    in  bedgraph   cut_sites,
    in  path       reference_path,
    in  pickle     count_dict,
    out bed        peaks,
    out json       peak_metrics,
    src py         "stages/processing/detect_peaks",
) split (
    in  string     contig,
    in  float[]    params,
    in  float      threshold,
) using (
    mem_gb = 6,
)
"""

PARAM_NAMES = ["p_zero_noise", "mean_bg_noise", "alpha_bg_noise", "p_signal", "frac_zero_noise", "frac_bg_noise"]


def split(args):
    if args.cut_sites is None:
        return {'chunks': [], 'join': {}}

    with open(args.count_dict, 'rb') as f:
        count_dict = pickle.load(f)
    mgr, ctx = stage_util.context_for_reference(args.reference_path)
    values = np.array(sorted(count_dict), dtype=np.int64)
    weights = np.array([count_dict[v] for v in values.tolist()], dtype=np.int64)
    threshold, params, _ = ctx.fit_threshold(values, weights, stage_util.PEAK_ODDS_RATIO)
    ctx.close()
    chunks = [{'params': None if params is None else list(params),
               'contig': None,              # one chunk for all contigs
               'threshold': threshold,
               '__mem_gb': 6}]
    return {'chunks': chunks, 'join': {'__mem_gb': 5}}


def main(args, outs):
    '''Call peaks on all chromosomes from the bedgraph rows'''
    if args.cut_sites is None:
        outs.peak_metrics = None
        outs.peaks = None
        return

    mgr, ctx = stage_util.context_for_reference(args.reference_path)
    # the reference compares the count STRING with the float threshold (utils.py:112), which never filters in Python 2;
    # the integer test of detect_peaks/__init__.py:49 is the one that counts.  The rows are parsed natively on all cores
    # (cratac_read_bedgraph): genome-wide the bedgraph has hundreds of millions of them.
    chrom, start, end = _ffi.read_bedgraph(args.cut_sites, list(mgr.all_contigs), args.threshold)
    order = np.lexsort((start, chrom))
    pc, ps, pe = ctx.peaks_from_rows(chrom[order], start[order], end[order])
    ctx.close()
    with open(outs.peaks, 'w') as out:
        for c, s, e in zip(pc.tolist(), ps.tolist(), pe.tolist()):
            out.write('{}\t{}\t{}\n'.format(mgr.all_contigs[c], s, e))
    with open(outs.peak_metrics, 'w') as f:
        json.dump({'total_peaks_detected': int(len(ps)), 'bases_in_peaks': int((pe.astype(np.int64) - ps).sum())}, f)


def join(args, outs, chunk_defs, chunk_outs):
    '''Join peaks'''
    if args.cut_sites is None:
        outs.peak_metrics = None
        outs.peaks = None
        return

    # the single chunk already wrote the peaks in genome.fa.fai order (= `bedtools sort -g`, tools/regions.py:406-409);
    # distinct intervals cannot repeat, so `uniq` has nothing to drop
    with open(chunk_outs[0].peaks, 'r') as src, open(outs.peaks, 'w') as dst:
        for line in src:
            dst.write(line)
    with open(chunk_outs[0].peak_metrics, 'r') as f:
        peak_metrics = json.load(f)

    # add fit data to summary (detect_peaks/__init__.py:121-129)
    fit_threshold = chunk_defs[0].threshold
    if chunk_defs[0].params is not None:
        for name, value in zip(PARAM_NAMES, chunk_defs[0].params):
            peak_metrics["peakcaller_{}".format(name)] = value
    peak_metrics["peakcaller_threshold"] = fit_threshold
    with open(outs.peak_metrics, 'w') as f:
        json.dump(peak_metrics, f)
