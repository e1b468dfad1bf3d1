"""The reference's three stages executed the way Martian runs them (TEST INFRASTRUCTURE ONLY).

Used as the timed CPU baseline (bench.py ``cpu_baseline`` and ``--impl reference``) and as an
end-to-end checker.  The reference itself cannot run here (Python 2.7, martian/pysam/pybedtools/
h5py absent), so this drives the literal restatement in oracle/stages.py
with the reference's own chunking:

  COUNT_CUT_SITES       one task per primary contig   (count_cut_sites/__init__.py:59-62)
  DETECT_PEAKS.split    one mixture fit               (detect_peaks/__init__.py:72-74)
  DETECT_PEAKS.main     one task per primary contig, each scanning the whole-genome bedgraph
                        (lib/python/utils.py:105-115)
  GENERATE_PEAK_MATRIX  split = one full parse for the barcode set (:44), <= 30 barcode chunks each
                        parsing the whole file (:47,107), join = hstack (:62-72)

Tasks run on a multiprocessing pool of ``n_procs`` workers (Martian local mode with that many cores).
"""
import multiprocessing as mp
import os
import time
from collections import Counter

import numpy as np

from . import em, py2set, stages

_G = {}


def _lines_of(contig):
    lo, hi = _G["contig_span"].get(contig, (0, 0))
    return _G["lines"][lo:hi]


def _task_count_cut_sites(contig):
    """count_cut_sites/__init__.py:85-112 (tabix fetch of one contig = that contig's line span)."""
    L = _G["contig_len"][contig]
    frags = ((s, e) for _, s, e, _, _ in stages.parse_fragments(_lines_of(contig)))
    Cuts = stages.pileup(L, frags)
    cd = stages.count_dict_from_cuts(Cuts)
    rows = stages.bedgraph_rows(contig, L, Cuts) if len(cd) else []
    return contig, cd, rows


def _task_detect_peaks(args):
    """detect_peaks/__init__.py:85-97: every chunk walks ALL bedgraph rows and keeps its contig."""
    contig, threshold = args
    text_rows = _G["bedgraph_text_rows"]
    peaks, n = stages.identify_peaks(stages.cut_site_counter_from_bedgraph(text_rows, threshold, contig), threshold)
    return contig, peaks


def _task_matrix_chunk(chunk):
    """generate_peak_matrix/__init__.py:84-125 for one barcode chunk (a full pass over the file)."""
    s, e = chunk
    bcs = _G["barcodes"][s:e]
    return stages.peak_matrix_chunk(_G["peak_lines"], bcs, stages.parse_fragments(_G["lines"]))


def run(text, contigs, primary, n_procs=None, odds_ratio=stages.PEAK_ODDS_RATIO):
    """text: decompressed fragments.tsv bytes.  Returns dict(peaks, threshold, params, barcodes,
    indptr, indices, data, seconds={stage: s}, n_fragments)."""
    n_procs = n_procs or os.cpu_count() or 1
    t_all = time.time()
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    names = [c[0] for c in contigs]
    # contig line spans (what the tabix index provides to the reference)
    first = [ln.split(b"\t", 1)[0].decode() for ln in lines] if lines else []
    span = {}
    i = 0
    while i < len(first):
        j = i
        while j < len(first) and first[j] == first[i]:
            j += 1
        span[first[i]] = (i, j)
        i = j
    _G.clear()
    _G.update(lines=lines, contig_span=span, contig_len=dict(contigs))
    secs = {}
    ctx = mp.get_context("fork")

    t0 = time.time()
    prim = [c for c in names if c in set(primary)]
    with ctx.Pool(min(n_procs, max(len(prim), 1))) as pool:
        res = pool.map(_task_count_cut_sites, prim, chunksize=1)
    count_dict = stages.merge_count_dicts([r[1] for r in res])
    bedgraph = [row for r in res for row in r[2]]
    secs["count_cut_sites"] = time.time() - t0

    t0 = time.time()
    threshold, params = em.estimate_final_threshold(count_dict, odds_ratio)
    secs["detect_peaks_split_fit"] = time.time() - t0

    t0 = time.time()
    _G["bedgraph_text_rows"] = [(c, str(s), str(e), str(v)) for c, s, e, v in bedgraph]
    thr_arg = float(threshold) if threshold is not None else float("-inf")
    with ctx.Pool(min(n_procs, max(len(prim), 1))) as pool:
        res = pool.map(_task_detect_peaks, [(c, thr_arg) for c in prim], chunksize=1)
    peaks = stages.sort_and_uniq_peaks([p for r in res for p in r[1]], names)
    secs["detect_peaks_main_join"] = time.time() - t0

    t0 = time.time()
    seen = dict.fromkeys(bc for _, _, _, bc, _ in stages.parse_fragments(lines))
    barcodes = py2set.py2_set_order(list(seen))
    chunks = stages.get_chunks(len(barcodes), 30)
    _G.update(barcodes=barcodes, peak_lines=["%s\t%d\t%d\n" % p for p in peaks])
    secs["matrix_split"] = time.time() - t0

    t0 = time.time()
    if chunks:
        with ctx.Pool(min(n_procs, len(chunks))) as pool:
            parts = pool.map(_task_matrix_chunk, chunks, chunksize=1)
        indptr, indices, data, _shape = stages.hstack_csc(parts)
    else:
        indptr, indices, data = np.zeros(1, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64)
    secs["matrix_main_join"] = time.time() - t0
    secs["total"] = time.time() - t_all
    return dict(peaks=peaks, threshold=threshold, params=params, barcodes=barcodes, indptr=indptr, indices=indices,
                data=data, seconds=secs, n_fragments=len(lines), count_dict=count_dict)
