#!/usr/bin/env python
"""Host-side file throughput of the drop-in (no GPU needed): BGZF inflate with the library's own DEFLATE decoder against
the zlib route, the bedgraph writer / reader and the Matrix Market writer, at several thread counts.  These are the
numbers quoted in DESIGN.md section 8 (host ingest, stage files).

    python tools/ingest_bench.py [--fragments 4000000] [--rows 20000000]
"""
from __future__ import print_function

import argparse
import ctypes
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cellranger-atac_b200"))

import numpy as np  # noqa: E402

from cratac import _ffi, bgzf, synth  # noqa: E402


def best_of(fn, reps=4):
    best = float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fragments", type=int, default=4000000)
    ap.add_argument("--rows", type=int, default=20000000)
    args = ap.parse_args()
    cores = os.cpu_count() or 1
    threads = sorted({1, 2, 4, cores})
    lib = _ffi.load()
    with tempfile.TemporaryDirectory() as tmp:
        # ---- inflate
        text = synth.to_text(synth.generate(synth.GRCH38, args.fragments, 2000, 20000, seed=3))
        path = os.path.join(tmp, "fragments.tsv.gz")
        bgzf.write(path, text, level=6)
        n = ctypes.c_int64()
        lib.cratac_inflate_file(path.encode(), 1, None, 0, ctypes.byref(n))
        out = np.zeros(n.value, dtype=np.uint8)
        print("inflate: %.1f MB BGZF -> %.1f MB text" % (os.path.getsize(path) / 1e6, n.value / 1e6))
        for mode, name in ((0, "zlib"), (1, "own decoder")):
            _ffi.inflate_mode(mode)
            for th in threads:
                s = best_of(lambda: lib.cratac_inflate_file(path.encode(), th, out.ctypes.data_as(ctypes.c_void_p), out.size, ctypes.byref(n)))
                print("  %-12s %2d threads  %.3f s  %6.0f MB/s" % (name, th, s, n.value / 1e6 / s))
        _ffi.inflate_mode(1)
        assert out.tobytes() == text
        # ---- bedgraph
        rng = np.random.default_rng(1)
        rows = args.rows
        st = np.cumsum(rng.integers(1, 12, size=rows)).astype(np.int64)
        en = st + rng.integers(1, 10, size=rows)
        cnt = rng.integers(1, 80, size=rows).astype(np.int64)
        bg = os.path.join(tmp, "cut_sites.bedgraph")
        for th in threads:
            s = best_of(lambda: _ffi.write_bedgraph(bg, "chr1", st, en, cnt, append=False, threads=th), reps=2)
            print("bedgraph write %d M rows  %2d threads  %.2f s (%.0f MB)" % (rows // 1000000, th, s, os.path.getsize(bg) / 1e6))
        for th in threads:
            s = best_of(lambda: _ffi.read_bedgraph(bg, ["chr1"], 30.0, threads=th), reps=2)
            print("bedgraph read + threshold  %2d threads  %.2f s" % (th, s))
        # ---- Matrix Market
        n_cols = 200000
        per = np.bincount(rng.integers(0, n_cols, size=rows), minlength=n_cols)
        indptr = np.concatenate(([0], np.cumsum(per))).astype(np.int64)
        indices = rng.integers(0, 124000, size=rows).astype(np.int64)
        data = rng.integers(1, 5, size=rows).astype(np.int32)
        mtx = os.path.join(tmp, "matrix.mtx")
        for th in threads:
            s = best_of(lambda: _ffi.write_mtx(mtx, "%%MatrixMarket matrix coordinate integer general\n", indptr, indices, data, threads=th), reps=2)
            print("mtx write %d M entries  %2d threads  %.2f s (%.0f MB)" % (rows // 1000000, th, s, os.path.getsize(mtx) / 1e6))


if __name__ == "__main__":
    main()
