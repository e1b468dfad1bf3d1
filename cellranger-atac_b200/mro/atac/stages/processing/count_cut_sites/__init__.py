"""
COUNT_CUT_SITES on the GPU: windowed Tn5 cut-site counts, their histogram and the bedgraph.

Drop-in for mro/atac/stages/processing/count_cut_sites/__init__.py of cellranger-atac 1.2.0: same __MRO__,
same outs (pickle of a collections.Counter, 4-column bedgraph incl. the zero-gap rows of the reference writer).
The per-contig chunks of the reference (one process per contig, :59-62) become one GPU pass done in join
(join-only stages: precedent filter_peak_matrix/__init__.py:30-43).
"""
from __future__ import division

import pickle
from collections import Counter

from cratac import _ffi, stage_util

__MRO__ = """
stage COUNT_CUT_SITES(
    in  path       reference_path,
    in  tsv.gz     fragments,
    in  tsv.gz.tbi fragments_index,
    out bedgraph   cut_sites,
    out pickle     count_dict,
    src py         "stages/processing/count_cut_sites",
) split (
    in  string     contig,
)
"""


def split(args):
    if args.fragments is None:
        return {'chunks': [], 'join': {}}
    return {'chunks': [], 'join': {'__mem_gb': 16, '__threads': 4}}


def main(args, outs):
    # no chunks are issued; kept for Martian's stage interface
    outs.count_dict = None
    outs.cut_sites = None


def join(args, outs, chunk_defs, chunk_outs):
    '''Pile up cut sites, histogram the windowed counts, write the bedgraph'''
    if args.fragments is None:
        outs.count_dict = None
        outs.cut_sites = None
        return

    mgr, ctx = stage_util.context_for_reference(args.reference_path)
    stage_util.load_fragments(ctx, args.fragments)
    values, weights = ctx.count_cut_sites()
    count_dict = Counter(dict(zip(values.tolist(), weights.tolist())))
    with open(outs.count_dict, 'wb') as f:
        pickle.dump(count_dict, f, protocol=2)

    # bedgraph of * windowed cutsites *, contigs in the chunk order of the reference (primary_contigs)
    # (the rows are formatted natively on all cores, cratac_write_bedgraph: genome-wide there are hundreds of millions)
    wrote = False
    open(outs.cut_sites, 'w').close()
    for contig in mgr.primary_contigs(allow_sex_chromosomes=True):
        starts, ends, counts = ctx.bedgraph_rows(contig)
        if starts.size:
            _ffi.write_bedgraph(outs.cut_sites, contig, starts, ends, counts, append=True)
            wrote = True
    ctx.close()
    if not wrote:
        outs.cut_sites = None
