"""ctypes wrapper over oracle/fast.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i64 = ctypes.c_int64


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "fast.c")):
            build()
        L = ctypes.CDLL(path)
        L.orc_window_counts.argtypes = [_i64, _i32p, _i32p, _i64, _i32p]
        L.orc_window_counts.restype = ctypes.c_int
        L.orc_count_hist.argtypes = [_i32p, _i64, _i64p, _i64]
        L.orc_count_hist.restype = _i64
        L.orc_bedgraph_rows.argtypes = [_i32p, _i64, _i64p, _i64p, _i64p, _i64]
        L.orc_bedgraph_rows.restype = _i64
        L.orc_identify_peaks.argtypes = [_i64p, _i64p, _i64p, _i64, _i64, _i64p, _i64p, _i64]
        L.orc_identify_peaks.restype = _i64
        L.orc_peak_matrix.argtypes = [_i64p, _i32p, _i32p, _i32p, ctypes.c_int32,
                                      _i32p, _i32p, _i32p, _i32p, _i64, _i64,
                                      _i64p, _i64p, _i64p, _i64]
        L.orc_peak_matrix.restype = _i64
        L.orc_parse.argtypes = [ctypes.c_char_p, _i64, _i64, _i64p, _i32p, _i32p, _i32p, _i64p, _i32p, _i32p]
        L.orc_parse.restype = _i64
        _LIB = L
    return _LIB


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def window_counts(L, starts, stops):
    W = np.zeros(L, dtype=np.int32)
    rc = lib().orc_window_counts(L, _c32(starts), _c32(stops), len(starts), W)
    assert rc == 0
    return W


def count_hist(W, cap=1 << 20):
    hist = np.zeros(cap, dtype=np.int64)
    mx = lib().orc_count_hist(W, W.size, hist, cap)
    assert mx >= 0
    return hist[: mx + 1]


def bedgraph_rows(W):
    cap = 1 << 16
    while True:
        rs, re_, rc = (np.zeros(cap, dtype=np.int64) for _ in range(3))
        n = lib().orc_bedgraph_rows(W, W.size, rs, re_, rc, cap)
        if n <= cap:
            return rs[:n], re_[:n], rc[:n]
        cap = int(n)


def identify_peaks(rs, re_, rc, threshold):
    cap = max(16, rs.size)
    ps, pe = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64)
    n = lib().orc_identify_peaks(np.ascontiguousarray(rs), np.ascontiguousarray(re_),
                                 np.ascontiguousarray(rc), rs.size, int(threshold), ps, pe, cap)
    assert n <= cap
    return ps[:n], pe[:n]


def call_peaks_contig(L, starts, stops, threshold):
    W = window_counts(L, starts, stops)
    rs, re_, rc = bedgraph_rows(W)
    return identify_peaks(rs, re_, rc, threshold)


def peak_matrix(peak_chrom, peak_start, peak_end, n_contigs, f_chrom, f_start, f_stop, f_bc, n_bc):
    """Peaks in peaks.bed row order (any order); returns canonical CSC (indptr, indices, data)."""
    peak_chrom = np.asarray(peak_chrom, dtype=np.int64)
    order = np.lexsort((np.asarray(peak_end), np.asarray(peak_start), peak_chrom))
    pc = peak_chrom[order]
    off = np.zeros(n_contigs + 1, dtype=np.int64)
    np.cumsum(np.bincount(pc, minlength=n_contigs), out=off[1:])
    cap = 2 * len(f_start) + 1
    indptr = np.zeros(n_bc + 1, dtype=np.int64)
    indices = np.zeros(cap, dtype=np.int64)
    data = np.zeros(cap, dtype=np.int64)
    nnz = lib().orc_peak_matrix(off, _c32(np.asarray(peak_start)[order]), _c32(np.asarray(peak_end)[order]),
                                _c32(order), n_contigs, _c32(f_chrom), _c32(f_start), _c32(f_stop),
                                _c32(f_bc), len(f_start), n_bc, indptr, indices, data, cap)
    assert 0 <= nnz <= cap
    return indptr, indices[:nnz].copy(), data[:nnz].copy()


def parse(text):
    """bytes -> dict of arrays; contig/barcode returned as (offset, length) pairs."""
    cap = text.count(b"\n") + 1
    co, bo = np.zeros(cap, np.int64), np.zeros(cap, np.int64)
    cl, st, sp, bl, du = (np.zeros(cap, np.int32) for _ in range(5))
    n = lib().orc_parse(text, len(text), cap, co, cl, st, sp, bo, bl, du)
    if n < 0:
        raise ValueError("malformed fragment line %d" % (-n - 1))
    return dict(n=n, contig_off=co[:n], contig_len=cl[:n], start=st[:n], stop=sp[:n],
                bc_off=bo[:n], bc_len=bl[:n], dup=du[:n])
