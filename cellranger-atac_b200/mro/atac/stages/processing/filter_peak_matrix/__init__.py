"""
FILTER_PEAK_MATRIX on the GPU: the raw peak-barcode matrix pruned to the cell barcodes and to non-empty columns.

Drop-in for mro/atac/stages/processing/filter_peak_matrix/__init__.py of cellranger-atac 1.2.0 (same __MRO__, same
filtered_peak_bc_matrix.h5 layout and MEX directory).  The scipy column slicing and the (m > 0).sum(axis=0) of the
reference's join/prune (:46-70, :84-96) become one column gather in HBM (cratac_select_columns).
"""
from __future__ import division

import os

from cratac import _ffi, h5, mex, stage_util

try:
    import martian
    _version = martian.get_pipelines_version
    _log = martian.log_info
except ImportError:            # outside Martian (tests, notebooks)
    martian = None

    def _version():
        return "cratac-b200"

    def _log(msg):
        pass

MAX_N_CLUSTERS_DEFAULT = 10    # cellranger/analysis/constants.py

__MRO__ = """
stage FILTER_PEAK_MATRIX(
    in  h5     raw_matrix,
    in  int    num_analysis_bcs,
    in  int    random_seed,
    in  csv    cell_barcodes,
    out h5     filtered_matrix,
    out path   filtered_matrix_mex,
    src py     "stages/processing/filter_peak_matrix",
)
"""


def split(args):
    if args.raw_matrix is None:
        return {'chunks': []}
    return {'chunks': [], 'join': {'__mem_gb': 8}}


def main(args, outs):
    raise RuntimeError('No chunks defined.')


def join(args, outs, chunk_defs, chunk_outs):
    if args.raw_matrix is None:
        outs.filtered_matrix = None
        return
    if args.cell_barcodes is None:
        raise SystemExit("cell barcodes not provided")
    cell_barcodes = stage_util.load_cell_barcodes(args.cell_barcodes)

    raw = h5.read_matrix_h5(args.raw_matrix)

    def _strs(a):
        return [x.decode() if isinstance(x, bytes) else str(x) for x in a.tolist()]
    feature_ids, genome, barcodes = _strs(raw["features"]["id"]), _strs(raw["features"]["genome"]), _strs(raw["barcodes"])
    n_features = len(feature_ids)
    if n_features == 0:
        _log("data has no peaks, skipping the clustering analysis")
        outs.filtered_matrix = None
        outs.filtered_matrix_mex = None
        return
    ctx = _ffi.Context([("matrix", 1)])
    ctx.set_matrix(raw["indptr"], raw["indices"], raw["data"])
    bcs, indptr, indices, data = stage_util.filter_peak_matrix(ctx, barcodes, cell_barcodes, args.num_analysis_bcs, args.random_seed)
    ctx.close()

    if len(bcs) <= MAX_N_CLUSTERS_DEFAULT:
        _log("Insufficient number of cell barcodes present after processing")
        outs.filtered_matrix = None
        outs.filtered_matrix_mex = None
        return
    if n_features < MAX_N_CLUSTERS_DEFAULT:
        _log("Insufficient number of peaks present after processing")
        outs.filtered_matrix = None
        outs.filtered_matrix_mex = None
        return

    h5.write_matrix_h5(outs.filtered_matrix, indptr, indices, data, n_features, bcs, feature_ids, genome, sw_version=_version())
    if not os.path.exists(outs.filtered_matrix_mex):
        os.mkdir(outs.filtered_matrix_mex)
    mex.save_mex(outs.filtered_matrix_mex, indptr, indices, data, n_features, bcs, feature_ids, _version())
