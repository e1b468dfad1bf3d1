"""
GENERATE_PEAK_MATRIX on the GPU: fragment ends x peaks counted per barcode into the raw peak-barcode matrix.

Drop-in for mro/atac/stages/processing/generate_peak_matrix/__init__.py of cellranger-atac 1.2.0 (same __MRO__,
same raw_peak_bc_matrix.h5 layout and MEX directory).  The 30 barcode chunks of the reference, each re-reading
the whole fragments file (:47,107), become one GPU pass over all barcodes done in join.
"""
from __future__ import division

import os

import numpy as np

from cratac import h5, mex, stage_util
from cratac.refmgr import generate_genome_tag, genome_of_contig

try:
    import martian
    _version = martian.get_pipelines_version
except ImportError:            # outside Martian (tests, notebooks)
    martian = None

    def _version():
        return "cratac-b200"

__MRO__ = """
stage GENERATE_PEAK_MATRIX(
    in  string reference_path,
    in  tsv.gz fragments,
    in  bed    peaks,
    out path   raw_matrix_mex,
    out h5     raw_matrix,
    src py     "stages/processing/generate_peak_matrix",
) split (
    in  file   barcodes,
) using (
    mem_gb = 4,
)
"""


def split(args):
    if args.fragments is None:
        return {'chunks': [], 'join': {}}
    return {'chunks': [], 'join': {'__mem_gb': 16, '__threads': 4}}


def main(args, outs):
    outs.raw_matrix_mex = None
    outs.raw_matrix = None


def join(args, outs, chunk_defs, chunk_outs):
    if args.fragments is None:
        outs.raw_matrix = None
        outs.raw_matrix_mex = None
        return

    mgr, ctx = stage_util.context_for_reference(args.reference_path)
    stage_util.load_fragments(ctx, args.fragments)
    names, starts, ends = stage_util.read_bed3(args.peaks)
    cid = {name: i for i, name in enumerate(mgr.all_contigs)}
    ctx.set_peaks([cid[c] for c in names], starts, ends)
    indptr, indices, data = ctx.peak_matrix()
    barcodes = ctx.get_barcodes()
    ctx.close()

    # FeatureDef(id = name = "chr:start-end", 'Peaks', tags genome/derivation): cellranger/atac/feature_ref.py:45-71
    feature_ids = ["{}:{}-{}".format(c, s, e) for c, s, e in zip(names, starts.tolist(), ends.tolist())]
    genomes = generate_genome_tag(args.reference_path)
    genome = [genome_of_contig(c, genomes) for c in names]
    h5.write_matrix_h5(outs.raw_matrix, indptr, indices, data, len(feature_ids), barcodes, feature_ids, genome, sw_version=_version())
    if not os.path.exists(outs.raw_matrix_mex):
        os.mkdir(outs.raw_matrix_mex)
    mex.save_mex(outs.raw_matrix_mex, indptr, indices, data, len(feature_ids), barcodes, feature_ids, _version())
