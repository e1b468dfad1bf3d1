"""Shared helpers for the parity tests (tests only: may import oracle/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cellranger-atac_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import cfast, em as oracle_em, py2set, stages  # noqa: E402


def oracle_decode(text, contig_names):
    """Oracle view of a fragments text: SoA + barcode strings in first-appearance order."""
    p = cfast.parse(text)
    cid = {c: i for i, c in enumerate(contig_names)}
    chrom = np.array([cid[text[o:o + l].decode()] for o, l in zip(p["contig_off"], p["contig_len"])], dtype=np.int32)
    bcs = [text[o:o + l].decode() for o, l in zip(p["bc_off"], p["bc_len"])]
    first_seen = list(dict.fromkeys(bcs))
    return dict(chrom=chrom, start=p["start"], end=p["stop"], dup=p["dup"], barcodes=bcs, first_seen=first_seen)


def oracle_columns(first_seen):
    """Matrix column order of the reference: CPython-2.7 list(set(...)) (generate_peak_matrix/__init__.py:44)."""
    return py2set.py2_set_order(first_seen)


def oracle_peaks(contigs, primary, chrom, start, end, threshold):
    """COUNT_CUT_SITES + DETECT_PEAKS through the literal C loops, per primary contig, .fai order."""
    out = []
    for ci, (name, L) in enumerate(contigs):
        if name not in primary:
            continue
        m = chrom == ci
        ps, pe = cfast.call_peaks_contig(L, start[m], end[m], threshold)
        out += [(ci, int(s), int(e)) for s, e in zip(ps, pe)]
    return out


def oracle_hist(contigs, primary, chrom, start, end):
    total = {}
    for ci, (name, L) in enumerate(contigs):
        if name not in primary:
            continue
        m = chrom == ci
        W = cfast.window_counts(L, start[m], end[m])
        h = cfast.count_hist(W)
        for v in np.nonzero(h)[0]:
            total[int(v)] = total.get(int(v), 0) + int(h[v])
    return total


def oracle_matrix(n_contigs, peaks, chrom, start, end, col_of_fragment, n_cols):
    pc = np.array([p[0] for p in peaks], dtype=np.int64)
    ps = np.array([p[1] for p in peaks], dtype=np.int64)
    pe = np.array([p[2] for p in peaks], dtype=np.int64)
    if len(peaks) == 0:
        return np.zeros(n_cols + 1, dtype=np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64)
    return cfast.peak_matrix(pc, ps, pe, n_contigs, chrom, start, end, col_of_fragment, n_cols)
