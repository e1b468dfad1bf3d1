"""
SETUP_AGGR_SAMPLES on the GPU: per-library depth and signal for `aggr` normalisation.

Drop-in for mro/atac/stages/aggr/setup_aggr_samples/__init__.py of cellranger-atac 1.2.0 (same __MRO__).  For the
"signal_mean" / "signal_noise_threshold" normalisations the reference piles every fragment of the library up over
every contig of the reference (+-200 bp windows, :96-108), histograms the per-base counts and runs
estimate_final_threshold on the histogram (:119-122) -- the same pileup + fit as COUNT_CUT_SITES / DETECT_PEAKS, here
over ALL contigs rather than the primary ones.  That part is cratac_decode + cratac_count_cut_sites +
cratac_fit_threshold_device on a context whose contigs are all primary; the csv bookkeeping stays pandas.
"""
from __future__ import division

import json
import pickle

import numpy as np
import pandas as pd

from cratac import _ffi, stage_util
from cratac.refmgr import ReferenceManager

__MRO__ = """
stage SETUP_AGGR_SAMPLES(
    in  csv        aggr_csv,
    in  string     normalization,
    in  string     reference_path,
    out pickle     library_info,
    out map        gem_group_index,
    out json       gem_group_index_json,
    src py         "stages/aggr/setup_aggr_samples",
) split (
    in  int        n,
) using (
    volatile = strict,
)
"""

SIGNAL_NORMALIZATIONS = ("signal_mean", "signal_noise_threshold")
KEY_AND_KIND = {"unique": ("unique_fragments_per_cell", "unique"), "total": ("total_fragments_per_cell", "total"),
                "signal_noise_threshold": ("original_threshold", "unique"), "signal_mean": ("signal_mean", "unique")}


def split(args):
    """one chunk per library of the aggr csv (the GPU needs no per-contig count array on the host)"""
    n_libraries = len(pd.read_csv(args.aggr_csv, sep=','))
    return {'chunks': [{'n': g, '__mem_gb': 8} for g in range(n_libraries)], 'join': {'__mem_gb': 12}}


def library_signal(reference_path, fragments_path, device=0):
    """(threshold, signal_mean) of one library: window-count histogram over every contig, then the mixture fit."""
    mgr = ReferenceManager(reference_path)
    contigs = mgr.contig_table()
    ctx = _ffi.Context(contigs, primary=[c[0] for c in contigs], device=device)     # every contig is piled up (:98-108)
    try:
        stage_util.load_fragments(ctx, fragments_path)
        ctx.count_cut_sites()
        threshold, params, _ = ctx.fit_threshold_device(stage_util.PEAK_ODDS_RATIO)
    finally:
        ctx.close()
    if params is None:
        raise RuntimeError("the threshold fit returned no parameters (fewer than 30 distinct counts or 1e5 bases)")
    return threshold, 1.0 / params[3]                                               # 1 / p_signal (:122)


def main(args, outs):
    lib_id = args.n + 1
    aggr = pd.read_csv(args.aggr_csv, sep=',')
    row = aggr.iloc[args.n]
    info = {label: str(row[label]) for label in aggr.columns.values.tolist()}
    library_info = {lib_id: info}
    if args.normalization is not None:
        cells = pd.read_csv(info['cells'], sep=',')
        mask = np.zeros(len(cells), dtype=bool)
        for species in ReferenceManager(args.reference_path).list_species():
            mask |= (cells['is_{}_cell_barcode'.format(species)] == 1).values
        sel = cells[mask]
        total = sel['total'] if 'total' in sel.columns else sel['passed_filters'] + sel['duplicate']
        info['total_fragments_per_cell'] = float(np.median(total))
        info['unique_fragments_per_cell'] = float(np.median(sel['passed_filters']))
        if args.normalization in SIGNAL_NORMALIZATIONS:
            threshold, signal_mean = library_signal(args.reference_path, row['fragments'])
            info['original_threshold'] = threshold
            info['signal_mean'] = signal_mean
    with open(outs.library_info, 'wb') as f:
        pickle.dump(library_info, f, protocol=2)      # py2 consumers (the rest of the 1.2.0 pipeline) read protocol <= 2


def join(args, outs, chunk_defs, chunk_outs):
    library_info = {}
    for chunk in chunk_outs:
        if chunk.library_info is not None:
            with open(chunk.library_info, 'rb') as f:
                library_info.update(pickle.load(f))
    if args.normalization is None:
        for n in library_info:
            library_info[n]['kind'] = "unique"
            library_info[n]['rate'] = 1.0
    else:
        key, kind = KEY_AND_KIND.get(args.normalization, ("", ""))
        lowest = min([1e20] + [library_info[n][key] for n in library_info])
        for n in library_info:
            library_info[n]['kind'] = kind
            library_info[n]['rate'] = lowest / library_info[n][key]
    with open(outs.library_info, 'wb') as f:
        pickle.dump(library_info, f, protocol=2)      # py2 consumers (the rest of the 1.2.0 pipeline) read protocol <= 2
    outs.gem_group_index = {"gem_group_index": {n: (library_info[n]['library_id'], 1) for n in library_info}}
    with open(outs.gem_group_index_json, 'w') as f:
        json.dump(outs.gem_group_index, f)
