"""Vectorised numpy equivalents of oracle/stages.py (TEST INFRASTRUCTURE ONLY).

Closed forms verified against the literal loops (tests/test_oracle.py, SURVEY.md
section 8c "GV5"):
  c[s] += 1 for s in {start, stop}            (length L+1: stop may equal L)
  W[p] = S[min(p+201, L+1)] - S[max(p-200, 0)],  S = exclusive prefix of c
  bedgraph rows = run-length encoding of the NON-ZERO bases of W, zeros skipped
  peaks = rows with count >= thr, merged while next.start - 500 <= cur.end
"""
from collections import Counter

import numpy as np

HALF = 200
MERGE = 500


def window_counts(L, starts, stops):
    """count_cut_sites/__init__.py:94-101 in closed form -> int32[L]."""
    starts = np.asarray(starts, dtype=np.int64)
    stops = np.asarray(stops, dtype=np.int64)
    c = np.bincount(starts, minlength=L + 1) + np.bincount(stops, minlength=L + 1)
    S = np.zeros(L + 2, dtype=np.int64)
    np.cumsum(c[:L + 1], out=S[1:])
    p = np.arange(L, dtype=np.int64)
    hi = np.minimum(p + HALF + 1, L + 1)
    lo = np.maximum(p - HALF, 0)
    return (S[hi] - S[lo]).astype(np.int32)


def count_dict(W):
    """count_cut_sites/__init__.py:104."""
    bc = np.bincount(W[W > 0])
    nz = np.nonzero(bc)[0]
    return Counter({int(k): int(bc[k]) for k in nz})


def bedgraph_rows(W):
    """count_cut_sites/__init__.py:27-48 -> (start[], end[], count[]) incl. the zero-gap quirk."""
    idx = np.nonzero(W)[0]
    if idx.size == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z, z
    v = W[idx].astype(np.int64)
    brk = np.nonzero(v[1:] != v[:-1])[0] + 1
    first = np.concatenate(([0], brk))
    last = np.concatenate((brk - 1, [idx.size - 1]))
    return idx[first].astype(np.int64), idx[last].astype(np.int64) + 1, v[first]


def peaks_from_rows(row_start, row_end, row_count, threshold):
    """utils.py:105-115 + detect_peaks/__init__.py:40-62 on bedgraph rows of ONE contig."""
    keep = row_count >= threshold
    s, e = row_start[keep], row_end[keep]
    if s.size == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z
    new = np.concatenate(([True], s[1:] - MERGE > e[:-1]))
    heads = np.nonzero(new)[0]
    tails = np.concatenate((heads[1:] - 1, [s.size - 1]))
    return s[heads], e[tails]


def call_peaks_contig(L, starts, stops, threshold):
    W = window_counts(L, starts, stops)
    rs, re_, rc = bedgraph_rows(W)
    return peaks_from_rows(rs, re_, rc, threshold)


def peak_matrix(peak_chrom, peak_start, peak_end, frag_chrom, frag_start, frag_stop, frag_bc, n_bc):
    """generate_peak_matrix/__init__.py:107-119 for integer-coded inputs.

    peaks: arrays in peaks.bed row order, assumed non-overlapping (reference
    raises KeyError otherwise).  Returns canonical CSC (indptr, indices, data)."""
    peak_chrom = np.asarray(peak_chrom, dtype=np.int64)
    peak_start = np.asarray(peak_start, dtype=np.int64)
    peak_end = np.asarray(peak_end, dtype=np.int64)
    n_peaks = peak_chrom.size
    # per-chrom sorted view (Regions sorts on construction: tenkit/regions.py:35)
    order = np.lexsort((peak_end, peak_start, peak_chrom))
    pc, ps, pe = peak_chrom[order], peak_start[order], peak_end[order]
    BIG = np.int64(1) << 40
    pkey = pc * BIG + ps
    fc = np.asarray(frag_chrom, dtype=np.int64)
    bc = np.asarray(frag_bc, dtype=np.int64)
    keys = []
    for pos in (np.asarray(frag_start, dtype=np.int64), np.asarray(frag_stop, dtype=np.int64)):
        q = fc * BIG + pos
        i = np.searchsorted(pkey, q, side="right") - 1
        ok = i >= 0
        ii = np.where(ok, i, 0)
        ok &= (pc[ii] == fc) & (ps[ii] <= pos) & (pos < pe[ii])
        keys.append(bc[ok] * n_peaks + order[ii[ok]])
    k = np.concatenate(keys) if keys else np.zeros(0, np.int64)
    uk, cnt = np.unique(k, return_counts=True)
    col = uk // max(n_peaks, 1)
    row = uk % max(n_peaks, 1)
    indptr = np.zeros(n_bc + 1, dtype=np.int64)
    np.cumsum(np.bincount(col, minlength=n_bc), out=indptr[1:])
    return indptr, row.astype(np.int64), cnt.astype(np.int64)
