"""Glue shared by the three Martian stage modules (mro/atac/stages/processing/*)."""
import numpy as np

from . import _ffi
from .refmgr import ReferenceManager

WINDOW_SIZE = 400            # lib/python/constants.py:68
PEAK_MERGE_DISTANCE = 500    # lib/python/constants.py:70
PEAK_ODDS_RATIO = 1 / 5.0    # lib/python/constants.py:71


def context_for_reference(reference_path, device=0):
    """ReferenceManager -> cratac context: contigs in genome.fa.fai order, peaks called on
    primary_contigs(allow_sex_chromosomes=True) (count_cut_sites/__init__.py:54-55)."""
    mgr = ReferenceManager(reference_path)
    ctx = _ffi.Context(mgr.contig_table(), primary=mgr.primary_contigs(allow_sex_chromosomes=True), device=device)
    return mgr, ctx


def load_fragments(ctx, fragments_path):
    """fragments.tsv.gz -> SoA in HBM: multi-threaded BGZF inflate (replaces the `gzip -c -d` pipe of
    lib/python/tools/io.py:197-217) + the GPU decoder (replaces io.py:75-79)."""
    text = _ffi.inflate_file(fragments_path)
    return ctx.decode_host(text)


def read_bed3(path):
    """peaks.bed rows in file order -> (contig names, starts, ends)."""
    chrom, start, end = [], [], []
    with open(path, "r") as f:
        for line in f:
            if not line.strip():
                continue
            c, s, e = line.rstrip("\n").split("\t")[:3]
            chrom.append(c)
            start.append(int(s))
            end.append(int(e))
    return chrom, np.array(start, dtype=np.int64), np.array(end, dtype=np.int64)
