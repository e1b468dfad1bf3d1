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
    """fragments.tsv.gz -> SoA in HBM (cratac_decode_file): BGZF blocks inflated on all host cores into a ring of pinned
    chunks (replaces the `gzip -c -d` pipe of lib/python/tools/io.py:197-217), each chunk copied and parsed by the GPU
    decoder (replaces io.py:75-79) while the next ones inflate."""
    return ctx.decode_file(fragments_path)


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


def load_cell_barcodes(filename):
    """cell_barcodes.csv -> {species: set(barcodes)} (lib/python/utils.py:40-53, with_species=True)."""
    cell_barcodes = {}
    with open(filename, "r") as f:
        for line in f:
            items = line.strip("\n").rstrip(",").split(",")
            cell_barcodes[items[0]] = {item for item in items[1:] if item != "null" and item != ""} if len(items) > 1 else set()
    return cell_barcodes


def filter_peak_matrix(ctx, bcs, cell_barcodes, num_analysis_bcs=None, random_seed=None):
    """FILTER_PEAK_MATRIX on the matrix held by `ctx` (after peak_matrix() / set_matrix()): columns of the cell barcodes
    present in the matrix, sorted (filter_peak_matrix/__init__.py:84-96, cellranger/matrix.py:874-883), an optional random
    subset drawn exactly as prune does (:52-56, numpy's legacy generator seeded with random_seed or 0), then the columns
    with at least one positive entry (:60-63).  The column gather and the pruning run on the GPU (cratac_select_columns).
    -> (barcodes, indptr, indices, data)."""
    index = {}
    for j, b in enumerate(bcs):
        index.setdefault(b, j)
    present = set()
    for species in cell_barcodes:
        for bc in cell_barcodes[species]:
            if bc in index:
                present.add(bc)
    sel = sorted(present)
    cols = np.array([index[b] for b in sel], dtype=np.int64)
    if num_analysis_bcs:
        np.random.seed(0 if random_seed is None else random_seed)
        n = cols.size
        pick = np.sort(np.random.choice(np.arange(n), size=min(num_analysis_bcs, n), replace=False))
        cols = cols[pick]
        sel = [sel[i] for i in pick]
    indptr, indices, data, kept = ctx.select_columns(cols, drop_empty=True)
    return [sel[i] for i in kept.tolist()], indptr, indices, data
