"""ctypes binding of libcratac.so (include/cratac.h).

Modelled on the reference's one ctypes shim,
tenkit/lib/python/striped_smith_waterman/ssw_wrap.py:54-72: ``cdll.LoadLibrary``,
explicit ``argtypes``/``restype`` for every symbol, an opaque handle and an
explicit destroy.  There is no CPU fallback: a missing library or a missing
CUDA device raises.
"""
import ctypes
import os

import numpy as np

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("CRATAC_LIB", os.path.join(_PKG_ROOT, "lib", "libcratac.so"))

c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_vp = ctypes.c_void_p
c_cp = ctypes.c_char_p
P = ctypes.POINTER


class CratacError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "libcratac error %d: %s" % (code, message))
        self.code = code


ERR_CUDA, ERR_ARG, ERR_PARSE, ERR_UNSORTED, ERR_CONTIG, ERR_CAPACITY = -1, -2, -3, -4, -5, -6
ERR_EM_ZERODIV, ERR_EM_INDEX, ERR_PEAKS_OVERLAP, ERR_INTERNAL, ERR_IO = -7, -8, -9, -10, -11
BC_ORDER_PY2SET, BC_ORDER_FIRST_SEEN, BC_ORDER_SORTED = 0, 1, 2


class RunResult(ctypes.Structure):
    _fields_ = [("n_fragments", c_i64), ("n_barcodes", c_i64), ("n_bins", c_i64), ("threshold", c_i64),
                ("has_params", c_i64), ("n_peaks", c_i64), ("nnz", c_i64), ("em_iterations", c_i64),
                ("em_inner_iterations", c_i64), ("params", c_dbl * 6), ("n_hits", c_i64)]


# symbol -> (restype, argtypes); every entry of include/cratac.h
SIGNATURES = {
    "cratac_last_error": (c_cp, []),
    "cratac_abi_version": (c_int, []),
    "cratac_create": (c_int, [c_int, c_int, P(c_cp), P(c_i64), P(c_i32), P(c_vp)]),
    "cratac_destroy": (None, [c_vp]),
    "cratac_set_option": (c_int, [c_vp, c_cp, c_i64]),
    "cratac_launch_count": (c_i64, [c_vp, c_int]),
    "cratac_decode_declined": (c_i64, [c_vp, c_int]),
    "cratac_decode_profile": (c_int, [c_vp, P(c_i64), c_int]),
    "cratac_synchronize": (c_int, [c_vp]),
    "cratac_stream": (c_int, [c_vp, P(ctypes.c_uint64)]),
    "cratac_num_stages": (c_int, []),
    "cratac_stage_name": (c_cp, [c_int]),
    "cratac_stage_times": (c_int, [c_vp, P(ctypes.c_float)]),
    "cratac_host_times": (c_int, [c_vp, c_vp, c_i64]),
    "cratac_em_debug": (c_int, [c_vp, P(c_dbl)]),
    "cratac_decode_host": (c_int, [c_vp, c_vp, c_i64]),
    "cratac_decode_device": (c_int, [c_vp, c_vp, c_i64]),
    "cratac_set_fragments": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, P(c_cp)]),
    "cratac_num_fragments": (c_int, [c_vp, P(c_i64), P(c_i64)]),
    "cratac_get_fragments": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cratac_get_barcodes": (c_int, [c_vp, c_vp, c_i64]),
    "cratac_max_barcode_len": (c_int, [c_vp, P(c_i64)]),
    "cratac_count_cut_sites": (c_int, [c_vp, c_vp, c_vp, c_i64, P(c_i64)]),
    "cratac_window_counts": (c_int, [c_vp, c_int, c_vp]),
    "cratac_bedgraph_rows": (c_int, [c_vp, c_int, c_vp, c_vp, c_vp, c_i64, P(c_i64)]),
    "cratac_fit_threshold": (c_int, [c_vp, c_vp, c_vp, c_i64, c_dbl, P(c_dbl), P(c_int), P(c_i64), P(c_i64), P(c_i64)]),
    "cratac_fit_threshold_device": (c_int, [c_vp, c_dbl, P(c_dbl), P(c_int), P(c_i64)]),
    "cratac_fit_threshold_result": (c_int, [c_vp, P(c_dbl), P(c_int), P(c_i64)]),
    "cratac_call_peaks": (c_int, [c_vp, c_i64, P(c_i64)]),
    "cratac_peaks_from_rows": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, P(c_i64)]),
    "cratac_get_peaks": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "cratac_set_peaks": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "cratac_peak_matrix": (c_int, [c_vp, P(c_i64)]),
    "cratac_get_matrix": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "cratac_matrix_device": (c_int, [c_vp, P(c_vp), P(c_vp), P(c_vp)]),
    "cratac_set_matrix": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "cratac_select_columns": (c_int, [c_vp, c_i64, c_vp, c_int, P(c_i64), P(c_i64)]),
    "cratac_get_selected": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cratac_hist_device": (c_int, [c_vp, P(c_vp), P(c_i64)]),
    "cratac_run": (c_int, [c_vp, c_vp, c_i64, c_int, c_dbl, P(RunResult)]),
    "cratac_run_from_fragments": (c_int, [c_vp, c_dbl, P(RunResult)]),
    "cratac_speculation_stats": (c_int, [c_vp, P(c_i64), P(c_i64)]),
    "cratac_speculation_second_guesses": (c_i64, [c_vp]),
    "cratac_pileup_tile_bases": (c_i64, []),
    "cratac_fit_revised": (c_int, [c_vp, P(c_int), P(c_i64)]),
    "cratac_fit_speculative_launch": (c_int, [c_vp, c_dbl]),
    "cratac_fit_tentative": (c_int, [c_vp, P(c_int), P(c_i64)]),
    "cratac_speculation_scope": (c_int, [c_vp, c_int, P(ctypes.c_uint64)]),
    "cratac_fit_cell_mixture": (c_int, [c_vp, c_vp, c_vp, c_i64, c_dbl, P(c_dbl)]),
    "cratac_low_targeting": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
    "cratac_insert_sizes": (c_int, [c_vp, c_i64, c_vp, c_int, c_vp]),
    "cratac_counts_by_barcode": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_int, c_vp, c_vp]),
    "cratac_decode_file": (c_int, [c_vp, c_cp, c_int]),
    "cratac_write_bgzf": (c_int, [c_cp, c_vp, c_i64, c_int, c_int]),
    "cratac_h5_filter_chunks": (c_int, [c_vp, c_i64, c_int, c_i64, c_int, c_int, c_vp, c_i64, c_vp]),
    "cratac_run_file": (c_int, [c_vp, c_cp, c_int, c_dbl, P(RunResult)]),
    "cratac_py2_set_order": (c_int, [c_i64, P(c_cp), c_vp]),
    "cratac_status_device": (c_int, [c_vp, P(c_vp), P(c_i64)]),
    "cratac_window_histogram": (c_int, [c_vp]),
    "cratac_compact_histogram": (c_int, [c_vp]),
    "cratac_barcode_table_device": (c_int, [c_vp, P(c_vp), P(c_vp), P(c_vp), P(c_i64)]),
    "cratac_set_column_map": (c_int, [c_vp, c_i64, c_vp]),
    "cratac_set_row_offset": (c_int, [c_vp, c_i64, c_i64]),
    "cratac_union_columns": (c_int, [c_vp, c_int, c_i64, c_vp, P(c_i64), P(c_i64), ctypes.c_uint64, P(c_i64)]),
    "cratac_column_keys_device": (c_int, [c_vp, P(c_vp), P(c_i64)]),
    "cratac_peak_matrix_begin": (c_int, [c_vp]),
    "cratac_peak_matrix_finish": (c_int, [c_vp, P(c_i64)]),
    "cratac_merge_rank_csc": (c_int, [c_vp, c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64]),
    "cratac_py2_set_order_hashes": (c_int, [c_i64, c_vp, c_vp]),
    "cratac_py2_set_order_hashes_device": (c_int, [c_vp, c_i64, c_vp, c_vp, ctypes.c_uint64]),
    "cratac_inflate_file": (c_int, [c_cp, c_int, c_vp, c_i64, P(c_i64)]),
    "cratac_inflate_range": (c_int, [c_cp, c_int, ctypes.c_uint64, ctypes.c_uint64, c_vp, c_i64, P(c_i64)]),
    "cratac_write_mtx": (c_int, [c_cp, c_cp, c_i64, c_vp, c_vp, c_vp, c_int]),
    "cratac_write_bedgraph": (c_int, [c_cp, c_int, c_cp, c_i64, c_vp, c_vp, c_vp, c_int]),
    "cratac_read_bedgraph": (c_int, [c_cp, c_int, P(c_cp), c_int, c_dbl, c_int, P(c_vp), P(c_vp), P(c_vp), P(c_i64)]),
    "cratac_free_host": (None, [c_vp]),
    "cratac_inflate_stats": (c_int, [P(c_i64), P(c_i64), c_int]),
    "cratac_inflate_mode": (c_int, [c_int]),
    "cratac_encode_text_device": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, P(c_i64)]),
}

_lib = None


def load():
    """Load libcratac.so and declare every signature.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libcratac.so not found at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C cellranger-atac_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.cdll.LoadLibrary(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise CratacError(rc, load().cratac_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(c_vp)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Context(object):
    """One GPU context.  contigs: list of (name, length) in genome.fa.fai order."""

    def __init__(self, contigs, primary=None, device=0):
        lib = load()
        self._lib = lib
        self.contig_names = [c[0] for c in contigs]
        self.contig_lens = [int(c[1]) for c in contigs]
        n = len(contigs)
        names = (c_cp * n)(*[s.encode("ascii") for s in self.contig_names])
        lens = (c_i64 * n)(*self.contig_lens)
        if primary is None:
            prim = None
        else:
            pset = set(primary)
            prim = (c_i32 * n)(*[1 if s in pset else 0 for s in self.contig_names])
        self.primary = list(primary) if primary is not None else list(self.contig_names)
        h = c_vp()
        check(lib.cratac_create(int(device), n, names, lens, prim, ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cratac_destroy(self._h)
            self._h = None

    __del__ = close

    def set_option(self, key, value):
        check(self._lib.cratac_set_option(self._h, key.encode(), int(value)))

    def launch_count(self, reset=False):
        return int(self._lib.cratac_launch_count(self._h, 1 if reset else 0))

    def fit_cell_mixture(self, values, weights, odds_ratio=20):
        """DETECT_CELL_BARCODES mixture fit -> dict(params=(mu_noise, alpha_noise, mu_signal, alpha_signal, frac_noise),
        simple_threshold, threshold (None when the reference returns None), em_iterations)."""
        values = np.ascontiguousarray(values, dtype=np.int64)
        weights = np.ascontiguousarray(weights, dtype=np.int64)
        out = (c_dbl * 8)()
        check(self._lib.cratac_fit_cell_mixture(self._h, _ptr(values), _ptr(weights), values.size, float(odds_ratio), out))
        return dict(params=tuple(float(out[k]) for k in range(5)), simple_threshold=int(out[5]),
                    threshold=None if out[6] < 0 else int(out[6]), em_iterations=int(out[7]))

    # ---- sibling scans
    def _regions(self, rows):
        """rows: iterable of (contig name, start, end[, minus]) -> arrays grouped by contig in reference order, tools/regions.py
        order (sorted (start, end) pairs) inside a contig."""
        cid = {nm: i for i, nm in enumerate(self.contig_names)}
        rows = sorted(((cid[r[0]], int(r[1]), int(r[2]), int(r[3]) if len(r) > 3 else 0) for r in rows if r[0] in cid),
                      key=lambda r: r[:3])           # stable: rows equal in (start, end) keep the caller's order (name, strand in the reference)
        a = np.array(rows, dtype=np.int64).reshape(-1, 4)
        return _i32(a[:, 0]), _i32(a[:, 1]), _i32(a[:, 2]), np.ascontiguousarray(a[:, 3], dtype=np.int8)

    def low_targeting(self, regions, paddings=(0, 50000)):
        """REMOVE_LOW_TARGETING_BARCODES.main over all contigs -> dict of int64 arrays in matrix column order:
        fragments, targeted, lengths{padding}, covered{padding}.  regions: (contig, start, end) rows."""
        c, s, e, _ = self._regions(regions)
        _, nb = self.num_fragments()
        pads = np.ascontiguousarray(paddings, dtype=np.int64)
        out = np.zeros((2 + 2 * pads.size, max(nb, 1)), dtype=np.int64)
        check(self._lib.cratac_low_targeting(self._h, c.size, _ptr(c), _ptr(s), _ptr(e), pads.size, _ptr(pads), _ptr(out)))
        out = out[:, :nb]
        return dict(fragments=out[0], targeted=out[1], lengths={int(p): out[2 + k] for k, p in enumerate(pads)},
                    covered={int(p): out[2 + pads.size + k] for k, p in enumerate(pads)})

    def insert_sizes(self, cols, exclude_non_primary=False):
        """REPORT_INSERT_SIZES.main for the barcodes of the matrix columns `cols` -> int64[len(cols), 1001]."""
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        table = np.zeros((max(cols.size, 1), 1001), dtype=np.int64)
        check(self._lib.cratac_insert_sizes(self._h, cols.size, _ptr(cols), int(bool(exclude_non_primary)), _ptr(table)))
        return table[:cols.size]

    COUNT_ROWS = ("passed_filters", "total", "duplicate", "peak_region_cutsites", "TSS_fragments", "DNase_sensitive_region_fragments",
                  "enhancer_region_fragments", "promoter_region_fragments", "on_target_fragments", "blacklist_region_fragments",
                  "peak_region_fragments")

    def counts_by_barcode(self, tracks, only_contig=None, species_list=(), species_of_contig=None, rel_half=1 << 16):
        """get_counts_by_barcode (tools/io.py:429-555).  tracks: dict with the keys tss, ctcf, dnase, enhancer, promoter,
        blacklist, peaks -> rows (contig, start, end[, minus]) already padded as the reference pads them (TSS 2000, CTCF 250),
        or None.  -> ({row name: int64[n_barcodes] in matrix column order}, tss profile {position: count}, ctcf profile)."""
        order = ("tss", "ctcf", "dnase", "enhancer", "promoter", "blacklist", "peaks")
        parts = [self._regions(tracks.get(k) or []) for k in order]
        off = np.concatenate(([0], np.cumsum([p[0].size for p in parts]))).astype(np.int64)
        rc, rs, re_, rm = (np.ascontiguousarray(np.concatenate([p[k] for p in parts])) for k in range(4))
        n_sp = len(species_list) if species_of_contig is not None and len(species_list) > 1 else 0
        sp = None
        if n_sp:
            sp = _i32([species_list.index(species_of_contig[nm]) if nm in species_of_contig else -1 for nm in self.contig_names])
        _, nb = self.num_fragments()
        rows = 11 + 2 * n_sp
        out = np.zeros((rows, max(nb, 1)), dtype=np.int64)
        rel = np.zeros((2, 2 * rel_half + 1), dtype=np.int64)
        oc = -1 if only_contig is None else self.contig_names.index(only_contig)
        check(self._lib.cratac_counts_by_barcode(self._h, _ptr(off), _ptr(rc), _ptr(rs), _ptr(re_), _ptr(rm), oc, n_sp, _ptr(sp), rel_half,
                                                 _ptr(out), _ptr(rel)))
        res = {name: out[k, :nb] for k, name in enumerate(self.COUNT_ROWS)}
        for k, spn in enumerate(species_list if n_sp else ()):
            res["passed_filters_" + spn] = out[11 + k, :nb]
            res["peak_region_fragments_" + spn] = out[11 + n_sp + k, :nb]
        prof = [{int(i) - rel_half: int(v) for i, v in zip(np.nonzero(rel[t])[0], rel[t][np.nonzero(rel[t])[0]])} for t in range(2)]
        return res, prof[0], prof[1]

    def fit_speculative_launch(self, odds_ratio=0.2):
        check(self._lib.cratac_fit_speculative_launch(self._h, float(odds_ratio)))

    def fit_tentative(self):
        """-> the tentative threshold the running fit published, or None when it ended without publishing one."""
        pub, thr = c_int(), c_i64()
        check(self._lib.cratac_fit_tentative(self._h, ctypes.byref(pub), ctypes.byref(thr)))
        return thr.value if pub.value else None

    def fit_revised(self):
        """-> the fit's second guess (it extrapolated where its threshold is heading and that is not the tentative one), or
        None when it ended without correcting itself."""
        pub, thr = c_int(), c_i64()
        check(self._lib.cratac_fit_revised(self._h, ctypes.byref(pub), ctypes.byref(thr)))
        return thr.value if pub.value else None

    def speculation_scope(self, on):
        """on = 1: tentative threshold, 2: the fit's second guess, 0: off -> the stream the library's calls are queued on from now on."""
        s = ctypes.c_uint64()
        check(self._lib.cratac_speculation_scope(self._h, int(on), ctypes.byref(s)))
        return s.value

    def speculation_second_guesses(self):
        """Runs in which the fit corrected its tentative threshold while it was still running."""
        return int(self._lib.cratac_speculation_second_guesses(self._h))

    def speculation_stats(self):
        """-> (runs whose speculative peaks + matrix were kept, runs that redid them)."""
        a, b = c_i64(), c_i64()
        check(self._lib.cratac_speculation_stats(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def decode_declined(self, reset=False):
        """Text ranges the single-pass decoder handed to the general parse kernel (0 on a clean file)."""
        return int(self._lib.cratac_decode_declined(self._h, 1 if reset else 0))

    def decode_profile(self):
        """SM cycles per phase of the single-pass decoder's tile loop (option decode_profile = 1) -> dict."""
        out = (c_i64 * 10)()
        check(self._lib.cratac_decode_profile(self._h, out, 10))
        names = ["wait_tile", "separator_masks", "separator_list", "structure_check", "look_back", "field_parse", "commit", "tiles",
                 "look_back_rounds", "look_back_retries"]
        return {k: int(v) for k, v in zip(names, out)}

    def synchronize(self):
        check(self._lib.cratac_synchronize(self._h))

    def stream(self):
        s = ctypes.c_uint64()
        check(self._lib.cratac_stream(self._h, ctypes.byref(s)))
        return s.value

    def stage_times(self):
        """-> {stage name: ms} for the kernels of the last run (needs set_option('profile', 1))."""
        n = self._lib.cratac_num_stages()
        ms = (ctypes.c_float * n)()
        check(self._lib.cratac_stage_times(self._h, ms))
        return {self._lib.cratac_stage_name(i).decode(): float(ms[i]) for i in range(n) if ms[i] >= 0}

    def em_debug(self):
        out = (c_dbl * 11)()
        check(self._lib.cratac_em_debug(self._h, out))
        names = ["e_step", "m_sums", "suffix", "bracket", "nodes", "verify_or_xc", "doubling", "descent", "direct", "retries_with_64_nodes", "levels_sum"]
        return {k: float(v) for k, v in zip(names, out)}

    def host_times(self):
        buf = ctypes.create_string_buffer(4096)
        check(self._lib.cratac_host_times(self._h, buf, 4096))
        out = []
        for kv in buf.value.decode().split(";"):
            if kv:
                k, v = kv.split("=")
                out.append((k, float(v)))
        return out

    # ---- decode
    def decode_host(self, text):
        """text: bytes / bytearray / numpy uint8 array / (address, nbytes) of decompressed fragments.tsv."""
        addr, n, keep = _buffer(text)
        check(self._lib.cratac_decode_host(self._h, addr, n))
        return self.num_fragments()

    def decode_file(self, path, threads=0):
        """fragments.tsv.gz -> fragments in HBM: BGZF inflate on host threads, PCIe copy and parse overlapped."""
        check(self._lib.cratac_decode_file(self._h, path.encode(), int(threads)))
        return self.num_fragments()

    def run_file(self, path, odds_ratio=0.2, threads=0):
        """decode_file + the rest of the path -> RunResult."""
        res = RunResult()
        check(self._lib.cratac_run_file(self._h, path.encode(), int(threads), float(odds_ratio), ctypes.byref(res)))
        return res

    def decode_device(self, dev_ptr, nbytes):
        check(self._lib.cratac_decode_device(self._h, c_vp(int(dev_ptr)), int(nbytes)))
        return self.num_fragments()

    def set_fragments(self, chrom, start, end, bc, dup, barcodes):
        chrom, start, end, bc = _i32(chrom), _i32(start), _i32(end), _i32(bc)
        dup = None if dup is None else _i32(dup)
        nb = len(barcodes)
        arr = (c_cp * max(nb, 1))(*[b.encode("ascii") for b in barcodes])
        check(self._lib.cratac_set_fragments(self._h, chrom.size, _ptr(chrom), _ptr(start), _ptr(end), _ptr(bc), _ptr(dup), nb, arr))

    def num_fragments(self):
        n, b = c_i64(), c_i64()
        check(self._lib.cratac_num_fragments(self._h, ctypes.byref(n), ctypes.byref(b)))
        return n.value, b.value

    def get_fragments(self):
        n, _ = self.num_fragments()
        out = [np.zeros(n, dtype=np.int32) for _ in range(5)]
        check(self._lib.cratac_get_fragments(self._h, *[_ptr(a) for a in out]))
        return dict(chrom=out[0], start=out[1], end=out[2], bc=out[3], dup=out[4])

    def get_barcodes(self):
        ln = c_i64()
        check(self._lib.cratac_max_barcode_len(self._h, ctypes.byref(ln)))
        _, nb = self.num_fragments()
        stride = ln.value + 1
        buf = np.zeros(max(nb, 1) * stride, dtype=np.uint8)
        check(self._lib.cratac_get_barcodes(self._h, _ptr(buf), stride))
        raw = buf.tobytes()
        return [raw[j * stride:(j + 1) * stride].split(b"\0", 1)[0].decode("ascii") for j in range(nb)]

    # ---- COUNT_CUT_SITES
    def count_cut_sites(self):
        """-> (values int64[], weights int64[]) of Counter(v for v in Cuts if v > 0), ascending."""
        cap = 1 << 20
        v = np.zeros(cap, dtype=np.int64)
        w = np.zeros(cap, dtype=np.int64)
        n = c_i64()
        check(self._lib.cratac_count_cut_sites(self._h, _ptr(v), _ptr(w), cap, ctypes.byref(n)))
        return v[:n.value].copy(), w[:n.value].copy()

    def window_counts(self, contig):
        ci = self.contig_names.index(contig) if isinstance(contig, str) else int(contig)
        W = np.zeros(self.contig_lens[ci], dtype=np.int32)
        check(self._lib.cratac_window_counts(self._h, ci, _ptr(W)))
        return W

    def bedgraph_rows(self, contig):
        ci = self.contig_names.index(contig) if isinstance(contig, str) else int(contig)
        n = c_i64()
        z = np.zeros(1, dtype=np.int64)
        check(self._lib.cratac_bedgraph_rows(self._h, ci, _ptr(z), _ptr(z), _ptr(z), 0, ctypes.byref(n)))
        s, e, c = (np.zeros(max(n.value, 1), dtype=np.int64) for _ in range(3))
        check(self._lib.cratac_bedgraph_rows(self._h, ci, _ptr(s), _ptr(e), _ptr(c), n.value, ctypes.byref(n)))
        return s[:n.value], e[:n.value], c[:n.value]

    # ---- threshold
    def fit_threshold(self, values, weights, odds_ratio=0.2):
        """-> (threshold or None, params tuple or None, stats dict) like estimate_final_threshold."""
        values = np.ascontiguousarray(values, dtype=np.int64)
        weights = np.ascontiguousarray(weights, dtype=np.int64)
        params = (c_dbl * 6)()
        has = c_int()
        thr, it, inner = c_i64(), c_i64(), c_i64()
        check(self._lib.cratac_fit_threshold(self._h, _ptr(values), _ptr(weights), values.size, float(odds_ratio), params,
                                             ctypes.byref(has), ctypes.byref(thr), ctypes.byref(it), ctypes.byref(inner)))
        return _fit_result(thr.value, has.value, params, it.value, inner.value)

    def fit_threshold_device(self, odds_ratio=0.2):
        params = (c_dbl * 6)()
        has = c_int()
        thr = c_i64()
        check(self._lib.cratac_fit_threshold_device(self._h, float(odds_ratio), params, ctypes.byref(has), ctypes.byref(thr)))
        return _fit_result(thr.value, has.value, params, None, None)

    # ---- FILTER_PEAK_MATRIX
    def set_matrix(self, indptr, indices, data):
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int64)
        data = np.ascontiguousarray(data, dtype=np.int32)
        check(self._lib.cratac_set_matrix(self._h, indptr.size - 1, _ptr(indptr), _ptr(indices), _ptr(data)))

    def select_columns_device(self, cols, drop_empty=True):
        """Same, result left in HBM -> (n_cols, nnz)."""
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        nc, nz = c_i64(), c_i64()
        check(self._lib.cratac_select_columns(self._h, cols.size, _ptr(cols) if cols.size else None, int(bool(drop_empty)),
                                              ctypes.byref(nc), ctypes.byref(nz)))
        return nc.value, nz.value

    def select_columns(self, cols, drop_empty=True):
        """m[:, cols] of the current matrix, without the columns left empty -> (indptr, indices, data, kept)."""
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        nc, nz = c_i64(), c_i64()
        check(self._lib.cratac_select_columns(self._h, cols.size, _ptr(cols) if cols.size else None, int(bool(drop_empty)),
                                              ctypes.byref(nc), ctypes.byref(nz)))
        indptr = np.zeros(nc.value + 1, dtype=np.int64)
        indices = np.zeros(nz.value, dtype=np.int64)
        data = np.zeros(nz.value, dtype=np.int32)
        kept = np.zeros(nc.value, dtype=np.int64)
        check(self._lib.cratac_get_selected(self._h, _ptr(indptr), _ptr(indices) if nz.value else None, _ptr(data) if nz.value else None,
                                            _ptr(kept) if nc.value else None))
        return indptr, indices, data, kept

    def py2_set_order_hashes_device(self, n, d_hashes, d_order_out, stream=0):
        """CPython-2.7 list(set) order from the hashes of n distinct keys, device pointers in and out."""
        check(self._lib.cratac_py2_set_order_hashes_device(self._h, int(n), c_vp(int(d_hashes)), c_vp(int(d_order_out)), int(stream)))

    def fit_threshold_result(self):
        """Numbers of a fit queued under option defer_sync = 1."""
        params = (c_dbl * 6)()
        has = c_int()
        thr = c_i64()
        check(self._lib.cratac_fit_threshold_result(self._h, params, ctypes.byref(has), ctypes.byref(thr)))
        return _fit_result(thr.value, has.value, params, None, None)

    # ---- peaks
    def call_peaks(self, threshold=None):
        n = c_i64()
        check(self._lib.cratac_call_peaks(self._h, -1 if threshold is None else int(threshold), ctypes.byref(n)))
        return self.get_peaks(n.value)

    def peaks_from_rows(self, chrom, start, end):
        """Signal intervals (bedgraph rows with count >= threshold, sorted by contig, start) -> merged peaks."""
        chrom = _i32(chrom)
        start = np.ascontiguousarray(start, dtype=np.int64)
        end = np.ascontiguousarray(end, dtype=np.int64)
        n = c_i64()
        check(self._lib.cratac_peaks_from_rows(self._h, chrom.size, _ptr(chrom), _ptr(start), _ptr(end), ctypes.byref(n)))
        return self.get_peaks(n.value)

    def get_peaks(self, n):
        c, s, e = (np.zeros(n, dtype=np.int32) for _ in range(3))
        if n:
            check(self._lib.cratac_get_peaks(self._h, _ptr(c), _ptr(s), _ptr(e)))
        return c, s, e

    def set_peaks(self, chrom, start, end):
        chrom, start, end = _i32(chrom), _i32(start), _i32(end)
        check(self._lib.cratac_set_peaks(self._h, chrom.size, _ptr(chrom), _ptr(start), _ptr(end)))

    # ---- matrix
    def peak_matrix(self):
        """-> (indptr int64[B+1], indices int64[nnz], data int32[nnz]) canonical CSC, rows = peaks."""
        nnz = c_i64()
        check(self._lib.cratac_peak_matrix(self._h, ctypes.byref(nnz)))
        return self.get_matrix(nnz.value)

    def get_matrix(self, nnz, out=None):
        """out: optional (indptr, indices, data) numpy views of (pinned) host buffers to fill."""
        _, nb = self.num_fragments()
        if out is not None:
            indptr, indices, data = out[0][:nb + 1], out[1][:nnz], out[2][:nnz]
        else:
            indptr = np.zeros(nb + 1, dtype=np.int64)
            indices = np.zeros(nnz, dtype=np.int64)
            data = np.zeros(nnz, dtype=np.int32)
        check(self._lib.cratac_get_matrix(self._h, _ptr(indptr), _ptr(indices), _ptr(data)))
        return indptr, indices, data

    # ---- whole path
    def run(self, text=None, odds_ratio=0.2, device_ptr=None, nbytes=None):
        res = RunResult()
        if device_ptr is not None:
            check(self._lib.cratac_run(self._h, c_vp(int(device_ptr)), int(nbytes), 1, float(odds_ratio), ctypes.byref(res)))
        elif text is not None:
            addr, n, keep = _buffer(text)
            check(self._lib.cratac_run(self._h, addr, n, 0, float(odds_ratio), ctypes.byref(res)))
        else:
            check(self._lib.cratac_run_from_fragments(self._h, float(odds_ratio), ctypes.byref(res)))
        return res

    # ---- multi-GPU building blocks (device pointers; the host code wraps them as torch tensors)
    def status_device(self):
        p, n = c_vp(), c_i64()
        check(self._lib.cratac_status_device(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def hist_device(self):
        p, n = c_vp(), c_i64()
        check(self._lib.cratac_hist_device(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def window_histogram(self):
        check(self._lib.cratac_window_histogram(self._h))

    def compact_histogram(self):
        check(self._lib.cratac_compact_histogram(self._h))

    def barcode_table_device(self):
        k, f, hsh, n = c_vp(), c_vp(), c_vp(), c_i64()
        check(self._lib.cratac_barcode_table_device(self._h, ctypes.byref(k), ctypes.byref(f), ctypes.byref(hsh), ctypes.byref(n)))
        return k.value, f.value, hsh.value, n.value

    def set_column_map(self, n_cols, d_col_to_dense):
        check(self._lib.cratac_set_column_map(self._h, int(n_cols), c_vp(int(d_col_to_dense)) if d_col_to_dense else None))

    def set_row_offset(self, row_base, total_rows):
        check(self._lib.cratac_set_row_offset(self._h, int(row_base), int(total_rows)))

    def union_columns(self, world, max_b, d_gathered, n_fragments, n_barcodes, stream=0):
        """Global columns from the all-gathered barcode tables of all shards -> n_cols (sets the column map)."""
        nf = (c_i64 * world)(*[int(x) for x in n_fragments])
        nb = (c_i64 * world)(*[int(x) for x in n_barcodes])
        n = c_i64()
        check(self._lib.cratac_union_columns(self._h, int(world), int(max_b), c_vp(int(d_gathered)), nf, nb, int(stream), ctypes.byref(n)))
        return n.value

    def column_keys_device(self):
        p, n = c_vp(), c_i64()
        check(self._lib.cratac_column_keys_device(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def peak_matrix_begin(self):
        check(self._lib.cratac_peak_matrix_begin(self._h))

    def peak_matrix_finish(self):
        """-> (nnz, d_indptr, d_indices, d_data) of the CSC left on the device."""
        nnz = c_i64()
        check(self._lib.cratac_peak_matrix_finish(self._h, ctypes.byref(nnz)))
        a, b, c = c_vp(), c_vp(), c_vp()
        check(self._lib.cratac_matrix_device(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return nnz.value, a.value, b.value, c.value

    def peak_matrix_device(self):
        """Build the CSC on the device only -> (nnz, d_indptr, d_indices, d_data)."""
        nnz = c_i64()
        check(self._lib.cratac_peak_matrix(self._h, ctypes.byref(nnz)))
        a, b, c = c_vp(), c_vp(), c_vp()
        check(self._lib.cratac_matrix_device(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return nnz.value, a.value, b.value, c.value

    def matrix_device(self, nnz):
        """Device pointers of the current CSC -> (nnz, d_indptr, d_indices, d_data)."""
        a, b, c = c_vp(), c_vp(), c_vp()
        check(self._lib.cratac_matrix_device(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return int(nnz), a.value, b.value, c.value

    def merge_rank_csc(self, world, n_cols, d_indptr_all, d_rank_off, d_indices_all, d_data_all, total_nnz):
        check(self._lib.cratac_merge_rank_csc(self._h, int(world), int(n_cols), c_vp(int(d_indptr_all)), c_vp(int(d_rank_off)),
                                              c_vp(int(d_indices_all)), c_vp(int(d_data_all)), int(total_nnz)))

    def get_matrix_indptr(self, n_cols):
        """indptr of the current CSC only (8 * (n_cols + 1) bytes)."""
        indptr = np.zeros(n_cols + 1, dtype=np.int64)
        check(self._lib.cratac_get_matrix(self._h, _ptr(indptr), None, None))
        return indptr

    def get_matrix_cols(self, n_cols, nnz, out=None):
        """D2H of the current CSC when its column count is not the shard's own barcode count."""
        if out is not None:
            indptr, indices, data = out[0][:n_cols + 1], out[1][:nnz], out[2][:nnz]
        else:
            indptr = np.zeros(n_cols + 1, dtype=np.int64)
            indices = np.zeros(nnz, dtype=np.int64)
            data = np.zeros(nnz, dtype=np.int32)
        check(self._lib.cratac_get_matrix(self._h, _ptr(indptr), _ptr(indices), _ptr(data)))
        return indptr, indices, data

    def encode_text_device(self, n, d_chrom, d_start, d_end, d_seq, d_dup, d_out, cap):
        nb = c_i64()
        check(self._lib.cratac_encode_text_device(self._h, int(n), c_vp(int(d_chrom)), c_vp(int(d_start)), c_vp(int(d_end)),
                                                  c_vp(int(d_seq)), c_vp(int(d_dup)), c_vp(int(d_out)) if d_out else None,
                                                  int(cap), ctypes.byref(nb)))
        return nb.value


def _fit_result(thr, has, params, it, inner):
    stats = dict(em_iterations=it, inner_iterations=inner)
    if has == 0:
        return thr, None, stats
    p = tuple(float(x) for x in params)
    return (None if has == 2 else thr), p, stats


def _buffer(text):
    """-> (c_void_p address, nbytes, keepalive)."""
    if isinstance(text, tuple):
        return c_vp(int(text[0])), int(text[1]), None
    if isinstance(text, np.ndarray):
        a = np.ascontiguousarray(text).view(np.uint8)
        return a.ctypes.data_as(c_vp), a.size, a
    if isinstance(text, (bytes, bytearray, memoryview)):
        a = np.frombuffer(text, dtype=np.uint8)
        return a.ctypes.data_as(c_vp), a.size, a
    raise TypeError("unsupported text buffer type %r" % type(text))


def py2_set_order(keys):
    """CPython-2.7 list(set(keys)) order; keys given in insertion order -> list of indices."""
    lib = load()
    n = len(keys)
    arr = (c_cp * max(n, 1))(*[k.encode("latin-1") if isinstance(k, str) else k for k in keys])
    out = np.zeros(max(n, 1), dtype=np.int64)
    check(lib.cratac_py2_set_order(n, arr, _ptr(out)))
    return out[:n].tolist()


def py2_set_order_hashes(hashes):
    """CPython-2.7 list(set) order of distinct str keys from their hashes (insertion order) -> int32 index array."""
    lib = load()
    h = np.ascontiguousarray(hashes, dtype=np.int64)
    out = np.zeros(max(h.size, 1), dtype=np.int32)
    check(lib.cratac_py2_set_order_hashes(h.size, _ptr(h), _ptr(out)))
    return out[:h.size]


def inflate_file(path, threads=None):
    """Multi-threaded BGZF / gzip inflate -> numpy uint8 array."""
    lib = load()
    threads = threads or (os.cpu_count() or 1)
    n = c_i64()
    check(lib.cratac_inflate_file(path.encode(), threads, None, 0, ctypes.byref(n)))
    out = np.empty(max(n.value, 1), dtype=np.uint8)
    check(lib.cratac_inflate_file(path.encode(), threads, _ptr(out), out.size, ctypes.byref(n)))
    return out[:n.value]


def write_bgzf(path, text, level=1, threads=0):
    """text (bytes / numpy uint8 / (address, nbytes)) -> BGZF file, deflated on `threads` host threads (0: all)."""
    addr, n, keep = _buffer(text)
    check(load().cratac_write_bgzf(path.encode(), addr, n, int(level), int(threads)))


def h5_filter_chunks(arr, chunk_elems, level=4, threads=0):
    """HDF5 shuffle + deflate of every chunk of a 1-D array on host threads -> (buffer uint8, stride, sizes int64[n_chunks])."""
    import zlib  # noqa: F401  (compressBound is not exposed; this bound is zlib's own formula)
    arr = np.ascontiguousarray(arr).reshape(-1)
    chunk_bytes = chunk_elems * arr.dtype.itemsize
    stride = chunk_bytes + (chunk_bytes >> 12) + (chunk_bytes >> 14) + (chunk_bytes >> 25) + 13 + 64
    n_chunks = (arr.size + chunk_elems - 1) // chunk_elems
    out = np.empty(max(n_chunks, 1) * stride, dtype=np.uint8)
    sizes = np.zeros(max(n_chunks, 1), dtype=np.int64)
    check(load().cratac_h5_filter_chunks(_ptr(arr) if arr.size else None, arr.size, arr.dtype.itemsize, int(chunk_elems), int(level), int(threads),
                                         _ptr(out), stride, _ptr(sizes)))
    return out, stride, sizes[:n_chunks]


def write_mtx(path, header, indptr, indices, data, threads=None):
    """Matrix Market file: `header` text, then one "row col value" line (1-based) per stored entry, column by column."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    nnz = int(indptr[-1]) if indptr.size else 0
    indices = np.ascontiguousarray(indices[:nnz], dtype=np.int64)
    data = np.ascontiguousarray(data[:nnz], dtype=np.int32)
    threads = threads or (os.cpu_count() or 1)
    check(load().cratac_write_mtx(path.encode(), header.encode("ascii"), max(indptr.size - 1, 0), _ptr(indptr), _ptr(indices), _ptr(data), threads))


def write_bedgraph(path, contig, starts, ends, counts, append=True, threads=None):
    """Append the rows "contig\\tstart\\tend\\tcount" of one contig to a bedgraph file."""
    starts, ends, counts = (np.ascontiguousarray(a, dtype=np.int64) for a in (starts, ends, counts))
    threads = threads or (os.cpu_count() or 1)
    check(load().cratac_write_bedgraph(path.encode(), 1 if append else 0, contig.encode(), starts.size, _ptr(starts), _ptr(ends), _ptr(counts), threads))


def read_bedgraph(path, contig_names, threshold=None, threads=None):
    """Rows of a bedgraph with count >= threshold (all rows when threshold is None), in file order
    -> (chrom int32 = index into contig_names, start int64, end int64)."""
    lib = load()
    threads = threads or (os.cpu_count() or 1)
    names = (c_cp * len(contig_names))(*[n.encode() for n in contig_names])
    pc, ps, pe, n = c_vp(), c_vp(), c_vp(), c_i64()
    check(lib.cratac_read_bedgraph(path.encode(), len(contig_names), names, 0 if threshold is None else 1,
                                   0.0 if threshold is None else float(threshold), threads,
                                   ctypes.byref(pc), ctypes.byref(ps), ctypes.byref(pe), ctypes.byref(n)))
    try:
        k = n.value
        chrom = np.ctypeslib.as_array(ctypes.cast(pc, P(c_i32)), shape=(max(k, 1),))[:k].copy()
        start = np.ctypeslib.as_array(ctypes.cast(ps, P(c_i64)), shape=(max(k, 1),))[:k].copy()
        end = np.ctypeslib.as_array(ctypes.cast(pe, P(c_i64)), shape=(max(k, 1),))[:k].copy()
    finally:
        for q in (pc, ps, pe):
            lib.cratac_free_host(q)
    return chrom, start, end


def inflate_stats(reset=False):
    """-> (BGZF blocks decoded by the library's own DEFLATE decoder, blocks handed to zlib) since the last reset."""
    a, b = c_i64(), c_i64()
    check(load().cratac_inflate_stats(ctypes.byref(a), ctypes.byref(b), 1 if reset else 0))
    return a.value, b.value


def inflate_mode(mode):
    """0: zlib only; 1 (default): own decoder first."""
    check(load().cratac_inflate_mode(int(mode)))


def inflate_range(path, voff_begin, voff_end, threads=None):
    """Inflate the BGZF virtual-offset range [voff_begin, voff_end) of a file -> numpy uint8 array."""
    lib = load()
    threads = threads or (os.cpu_count() or 1)
    n = c_i64()
    check(lib.cratac_inflate_range(path.encode(), threads, int(voff_begin), int(voff_end), None, 0, ctypes.byref(n)))
    out = np.empty(max(n.value, 1), dtype=np.uint8)
    check(lib.cratac_inflate_range(path.encode(), threads, int(voff_begin), int(voff_end), _ptr(out), out.size, ctypes.byref(n)))
    return out[:n.value]


def inflate_contigs(path, contigs, index=None, threads=None):
    """Text of the records of `contigs`, in the order given, from fragments.tsv.gz + its .tbi: what a rank that owns this
    contig range reads (the reference fetches one contig at a time, lib/python/tools/io.py:82-93).  Contigs that follow
    one another in the file are inflated as one span."""
    from . import tabix
    idx = tabix.read_index(index or path + ".tbi") if not isinstance(index, dict) else index
    spans = tabix.contig_spans(idx)
    runs = []
    for c in contigs:
        if c not in spans:
            continue
        if runs and runs[-1][1] == spans[c][0]:
            runs[-1][1] = spans[c][1]
        else:
            runs.append(list(spans[c]))
    parts = [inflate_range(path, b, e, threads) for b, e in runs]
    if not parts:
        return np.empty(0, dtype=np.uint8)
    return parts[0] if len(parts) == 1 else np.concatenate(parts)
