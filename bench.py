#!/usr/bin/env python
"""Benchmark of the fragment -> peaks -> peak x barcode matrix path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cratac|reference] [--fragments F]

One "step" = one pass of the whole hot path (decode, pileup/window/histogram, threshold fit, peak
calling, peak x barcode CSC) over one synthetic GRCh38-shaped input (BASELINE.json configs[1]:
10k cells, 250M fragments, ~100k latent peaks; `--fragments` scales it down for quick runs and is
reported in config).  `value` times the step with the decompressed text already resident in HBM;
`e2e` times the same step through the C ABI from a pinned HOST text buffer to peaks + CSC arrays
back in pinned host memory (H2D and D2H inside the timed region).  Per-kernel times come from CUDA
events on the library's own stream (cratac_stage_times).

N > 1 (torchrun, one rank per GPU): the genome is sharded by contig (LPT), each rank runs the path
on its contigs; the window-count histogram is all-reduced over NCCL before the threshold fit so
every rank calls peaks with the global threshold; "scaling" is strong (total work fixed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "cellranger-atac_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "fragments/sec (DETECT_PEAKS+peak-bc matrix)"
UNIT = "fragments/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region, in-process through NVML: one light query per GPU every
    50 ms, from ONE process (rank 0 watches all GPUs of the run).  Spawning nvidia-smi per rank stalled the driver for the
    first timed steps of an 8-GPU run, and eight in-process NVML clients starting together still cost the first step 10-30 ms."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_indices):
        self.samples, self.reasons = [], set()
        self.smax = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.handles = []
        self.err = "not sampled on this rank"
        if not gpu_indices:
            self.nv = None
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in gpu_indices]
            self.smax = float(min(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM) for h in self.handles))
        except Exception as e:                                # noqa: BLE001 - reported in the JSON line
            self.nv = None
            self.err = "NVML unavailable: %s" % (e,)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            for h in self.handles:
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    for bit, name in self.REASONS:
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:                             # noqa: BLE001
                    pass
            self.stop_flag.wait(0.05)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err], "samples": 0}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "reasons": sorted(self.reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- workload
def make_workload(torch, ctx, contigs, n_frag, cells, n_peaks, seed, device, chrom_offset=0, keep=None, in_peak_frac=0.5):
    """Seeded synthetic fragments generated ON the device (torch is plumbing here), sorted by
    (contig, start, stop) and rendered to fragments.tsv text by the library's encoder kernel.
    Same distributions as cratac/synth.py.
    keep = (first, last_exclusive) contig range: the WHOLE library is generated (identical on every rank: one seed, no
    rank in it) and only the fragments of those contigs are rendered -- the sharded runs at N = 1, 2, 4, 8 therefore work
    on the same library and their results (bench line `result.checksum`) must agree."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lens = torch.tensor([c[1] for c in contigs], dtype=torch.int64, device=device)
    cum = torch.cat([torch.zeros(1, dtype=torch.int64, device=device), torch.cumsum(lens, 0)])
    total = int(cum[-1].item())
    n = int(n_frag)
    centres = torch.randint(0, total, (max(int(n_peaks), 1),), generator=g, device=device, dtype=torch.int64)
    in_peak = torch.rand(n, generator=g, device=device) < in_peak_frac
    pick = torch.randint(0, centres.numel(), (n,), generator=g, device=device, dtype=torch.int64)
    gpos = centres[pick] + (torch.randn(n, generator=g, device=device) * 150.0).to(torch.int64)
    del pick
    bg = torch.randint(0, total, (n,), generator=g, device=device, dtype=torch.int64)
    gpos = torch.where(in_peak, gpos, bg).clamp_(0, total - 1)
    del bg, in_peak
    chrom = torch.searchsorted(cum[1:].contiguous(), gpos, right=True)
    start = gpos - cum[chrom]
    del gpos
    comp = torch.rand(n, generator=g, device=device)
    z = torch.randn(n, generator=g, device=device)
    length = torch.where(comp < 0.5, 120.0 + 35.0 * z, torch.where(comp < 0.85, 320.0 + 45.0 * z, 520.0 + 60.0 * z))
    del comp, z
    length = length.to(torch.int64).clamp_(20, 1200)
    L = lens[chrom]
    start = torch.minimum(start, L - 1)
    end = torch.minimum(start + length, L)
    start = torch.minimum(start, end - 1).clamp_(min=0)
    del length, L
    # sort by (contig, start, end)
    key = ((cum[chrom] + start) << 11) | (end - start)
    order = torch.argsort(key)
    del key
    chrom = (chrom[order] + chrom_offset).to(torch.int32)   # index into the context's (global) contig list
    start = start[order].to(torch.int32)
    end = end[order].to(torch.int32)
    del order
    dup = torch.empty(n, device=device).geometric_(0.5, generator=g).to(torch.int32)
    n_bg = 20 * cells
    gb = torch.Generator(device=device)
    gb.manual_seed(977)                      # the barcode universe is the same on every rank
    depth = torch.empty(cells, device=device).log_normal_(0.0, 0.5, generator=gb)
    p = torch.cat([0.8 * depth / depth.sum(), torch.full((n_bg,), 0.2 / n_bg, device=device)])
    cdf = torch.cumsum(p.double(), 0)
    cdf = cdf / cdf[-1]
    bc_idx = torch.searchsorted(cdf, torch.rand(n, generator=g, device=device, dtype=torch.float64)).clamp_(max=cells + n_bg - 1)
    seqs = torch.randint(0, 1 << 32, (cells + n_bg,), generator=gb, device=device, dtype=torch.int64)
    seqs = torch.where(seqs >= (1 << 31), seqs - (1 << 32), seqs).to(torch.int32)
    seq = seqs[bc_idx].contiguous()
    if keep is not None:
        lo = int((chrom < keep[0] + chrom_offset).sum().item())
        hi = int((chrom < keep[1] + chrom_offset).sum().item())
        chrom, start, end, dup, seq = (x[lo:hi].contiguous() for x in (chrom, start, end, dup, seq))
        n = hi - lo
    torch.cuda.synchronize()   # the encoder runs on the library's stream
    nbytes = ctx.encode_text_device(n, chrom.data_ptr(), start.data_ptr(), end.data_ptr(), seq.data_ptr(), dup.data_ptr(), 0, 0)
    del bc_idx
    text = torch.empty(nbytes + 16, dtype=torch.uint8, device=device)
    got = ctx.encode_text_device(n, chrom.data_ptr(), start.data_ptr(), end.data_ptr(), seq.data_ptr(), dup.data_ptr(), text.data_ptr(), nbytes)
    assert got == nbytes
    torch.cuda.synchronize()
    return text, nbytes, n


def algorithmic_bytes(T, N, G, P, K, nnz, B):
    """SURVEY.md section 8d per-kernel algorithmic bytes, summed over the stages each kernel replaces."""
    return {
        "decode": T + 20 * N,                         # K1 (newline count + parse)
        "pileup_window_hist": 16 * N + 4 * G,         # K2 + K3 (fused; the dense count array never reaches HBM)
        "pileup_call_peaks": 16 * N + 4 * G + 12 * P,  # K2 (recomputed) + K5
        "overlap": 12 * N + 8 * K,                    # K6 (count pass + scatter pass)
        "sort_reduce": 8 * K + 12 * nnz,              # K7
        "csc_build": 12 * nnz + 8 * (B + 1),          # K8
    }


KERNEL_GROUPS = {
    "decode": ("decode_count_newlines", "decode_parse"),
    "pileup_window_hist": ("pileup_window_hist",),
    "pileup_call_peaks": ("pileup_call_peaks",),
    "overlap": ("overlap_count", "overlap_scatter"),
    "sort_reduce": ("reduce_small", "reduce_large"),
    "csc_build": ("csc_build",),
}


# --------------------------------------------------------------------------------------------- reference arm
CPU_SLICE = 1000000          # bases per sample contig


def cpu_sample(args):
    """Bounded CPU sample of the configured workload: one 1 Mb contig per host core (at most 24) at the configuration's
    fragment density, so that the reference's per-contig tasks keep every core busy as they do on the whole genome."""
    from cratac import synth
    if args.config == 1:
        # BASELINE.json configs[0] is the reference's own CPU-runnable case: run in FULL, no sampling, no extrapolation
        contigs = synth.GRCH38[:1]
        frags = synth.generate(contigs, args.config_fragments, args.cells, args.peaks, seed=1000, background_factor=20)
        desc = ("the whole configuration: %d fragments on chr1 (%d bases), %d cells; reference-shaped flow (oracle/reference_flow.py) "
                "on a multiprocessing pool" % (args.config_fragments, contigs[0][1], args.cells))
        return (contigs, synth.to_text(frags)), desc
    k = max(1, min(os.cpu_count() or 1, 24))      # GRCh38 has 24 primary contigs: no more per-contig tasks than that
    contigs = [("chr%d" % (i + 1), CPU_SLICE) for i in range(k)]
    density = args.config_fragments / float(args.genome_bases)
    n = int(density * CPU_SLICE * k)
    peaks = max(int(args.peaks * CPU_SLICE * k / args.genome_bases), 4)
    frags = synth.generate(contigs, n, max(args.cells // 16, 10), peaks, seed=1000, background_factor=20)
    desc = ("%d fragments on %d contigs of %d bases at the config's fragment density (%.4f/base); the whole reference-shaped "
            "flow (oracle/reference_flow.py: per-contig pileup + per-base Counter/bedgraph loops, mixture fit, <= 30 barcode "
            "chunks each parsing the file) on a multiprocessing pool" % (n, k, CPU_SLICE, density))
    return (contigs, synth.to_text(frags)), desc


def run_reference_flow(sample):
    from oracle import reference_flow
    contigs, text = sample
    t0 = time.time()
    res = reference_flow.run(text, contigs, [c[0] for c in contigs], n_procs=os.cpu_count())
    dt = time.time() - t0
    return res["n_fragments"], dt, res["seconds"]


_FULL_FIT = {}


def full_size_fit_seconds(args):
    """The mixture fit does not scale with the sample: it runs once per library on the whole-genome histogram.  For the
    250 M-fragment configuration that histogram is a committed fixture (tests/golden/em.json, case bench_250m); the
    oracle's fit on it is timed once.  Other sizes: None (the sample's own fit time is used)."""
    if "t" in _FULL_FIT:
        return _FULL_FIT["t"]
    t = None
    path = os.path.join(ROOT, "tests", "golden", "em.json")
    if args.config_fragments == 250000000 and os.path.exists(path):
        from oracle import em as oracle_em
        for case in json.load(open(path)):
            if case.get("name") == "bench_250m":
                cd = {int(v): int(w) for v, w in zip(case["values"], case["weights"])}
                t0 = time.time()
                oracle_em.estimate_final_threshold(cd, 0.2)
                t = time.time() - t0
    _FULL_FIT["t"] = t
    return t


def cpu_whole_job_estimate(args, n_sample, secs, fit_full):
    if args.config == 1:          # measured in full
        return n_sample / secs["total"], {"whole_job_seconds_measured": secs["total"]}
    return _cpu_whole_job_estimate(args, n_sample, secs, fit_full)


def _cpu_whole_job_estimate(args, n_sample, secs, fit_full):
    """fragments/s of the whole configured job on this host: the per-contig / per-chunk stages scale with the number of
    fragments (linear scaling: slightly generous to the CPU, DETECT_PEAKS.main re-reads the whole bedgraph per contig),
    the fit is paid once."""
    fit_sample = secs["detect_peaks_split_fit"]
    scalable = secs["total"] - fit_sample
    t_full = scalable * args.config_fragments / float(n_sample) + (fit_full if fit_full is not None else fit_sample)
    return args.config_fragments / t_full, {"scalable_seconds_sample": scalable, "fit_seconds_sample": fit_sample,
                                            "fit_seconds_full_histogram": fit_full, "whole_job_seconds_estimate": t_full}


def cpu_baseline_block(args, n_sample, secs, fit_full, desc):
    value, model = cpu_whole_job_estimate(args, n_sample, secs, fit_full)
    return {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": desc + ("; value = fragments / measured wall time" if args.config == 1 else
                              "; value = configured fragments / (sample's scalable stage time x fragments ratio + one fit on the "
                              "full-size histogram)"), "seconds_by_stage": secs, "model": model}


def reference_arm(args, world, rank):
    """--impl reference: the reference-shaped CPU flow (oracle port) on all host cores, bounded sample per step."""
    if rank != 0:
        return
    sample, desc = cpu_sample(args)
    fit_full = full_size_fit_seconds(args)
    for _ in range(args.warmup):
        run_reference_flow(sample)
    t0 = time.time()
    values = []
    for _ in range(args.steps):
        nf, dt, secs = run_reference_flow(sample)
        values.append(cpu_whole_job_estimate(args, nf, secs, fit_full)[0])
    total = time.time() - t0
    block = cpu_baseline_block(args, nf, secs, fit_full, desc)
    value = sum(values) / len(values)
    block["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": block,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args):
    genome = "GRCh38 chr1 only" if args.config == 1 else "GRCh38 chr1-22,X,Y"
    return {"workload": "BASELINE.json configs[%d]: synthetic %d-cell run, %d fragments, %s, %d latent peak centres"
                        % (args.config - 1, args.cells, args.config_fragments, genome, args.peaks),
            "fragments": args.config_fragments, "cells": args.cells, "latent_peaks": args.peaks, "genome": genome,
            "l2": "inputs larger than L2 (text %.1f GB per step)" % (46.0 * args.config_fragments / 1e9),
            "library": "one seeded library, identical for every --gpus N, cut by contig range" if not args.per_rank_library
                       else "one seeded library per rank (its own contig range)",
            "in_peak_frac": args.in_peak_frac,
            "parallelism": "genome-range sharding by contig, %d GPU(s)" % args.gpus}


def apply_config(args, synth):
    """BASELINE.json configs: 1 = configs[0] (chr1 only, 500 cells, 5 M fragments: the reference's own CPU-runnable case),
    2 = configs[1] (the metric's configuration: 250 M fragments, 10 k cells, whole GRCh38), 4 = configs[3] (aggr shape:
    2 B fragments, 80 k cells, 8 GPUs).  Explicit --fragments / --cells / --peaks win."""
    defaults = {1: (5000000, 500, 8000), 2: (250000000, 10000, 100000), 4: (2000000000, 80000, 100000)}[args.config]
    if args.fragments is None:
        args.fragments = defaults[0]
    if args.cells is None:
        args.cells = defaults[1]
    if args.peaks is None:
        args.peaks = defaults[2]
    args.config_fragments = args.fragments
    contigs = synth.GRCH38[:1] if args.config == 1 else synth.GRCH38
    args.genome_bases = sum(c[1] for c in contigs)
    if args.config == 4:
        args.per_rank_library = True      # 2 B fragments are generated shard by shard: no single GPU holds the generator's temporaries
    return contigs


# 64-bit mixing constants (as signed int64) of the result checksum
_K1, _K2, _K3, _K4 = -7046029254386353131, -4417276706812531889, 1609587929392839161, -4658895280553007687


def _mix_sum(torch, a, b, c):
    """Order-independent 64-bit checksum of the triples (a, b, c): sum of a mixed word per triple, wrapping."""
    if a.numel() == 0:
        return 0
    h = a.to(torch.int64) * _K1 + b.to(torch.int64) * _K2 + c.to(torch.int64) * _K3
    h = (h ^ (h >> 29)) * _K4
    h = h ^ (h >> 32)
    return int(h.sum().item())


def result_checksum(torch, dist, world, device, threshold, peaks, row_base, indptr, indices, data, col_keys):
    """One 64-bit word over (threshold, peaks.bed rows, column order, indptr / indices / data): every rank adds the part
    it holds -- its peaks with their global row numbers, its matrix entries with global (column, row) -- so that the sum is
    the same for every way of sharding the same library.  Outside the timed region; torch is only the calculator here."""
    c, s_, e = peaks
    n = int(c.numel())
    rows = torch.arange(n, dtype=torch.int64, device=device) + int(row_base)
    part = _mix_sum(torch, rows * 64 + c.to(torch.int64), s_, e)
    ncols = int(indptr.numel()) - 1
    if int(indices.numel()) > 0:
        cols = torch.repeat_interleave(torch.arange(ncols, dtype=torch.int64, device=device), (indptr[1:] - indptr[:-1]))
        part += _mix_sum(torch, cols, indices, data)
    t = torch.tensor([((part + (1 << 63)) % (1 << 64)) - (1 << 63)], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(t)                   # int64 sum wraps
    total = int(t.item())
    total += _mix_sum(torch, torch.arange(int(col_keys.numel()), dtype=torch.int64, device=device), col_keys, torch.zeros_like(col_keys))
    total += int(threshold) * 1000003
    return "%016x" % (total % (1 << 64))


def file_to_file(args, torch, ctx, text, nbytes, contigs, nnz_hint, nb_hint):
    """`e2e_file`: the same metric from fragments.tsv.gz ON DISK to peaks.bed + raw_peak_bc_matrix.h5 ON DISK,
    through the calls the stage modules make: cratac_run_file (BGZF inflate on all host cores pipelined with the PCIe copy
    and the decode, then the path), the D2H of peaks and CSC, the peaks.bed text and the HDF5 file (cratac/h5.py, the
    layout of cellranger/matrix.py:415-447).  Host wall clock; the input file is written once, outside the timed region."""
    import shutil
    import tempfile
    import numpy as np
    from cratac import _ffi, h5
    tmpd = tempfile.mkdtemp(prefix="cratac_bench_")
    try:
        host_text = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host_text.copy_(text[:nbytes])
        torch.cuda.synchronize()
        frag_path = os.path.join(tmpd, "fragments.tsv.gz")
        t0 = time.time()
        _ffi.write_bgzf(frag_path, (host_text.data_ptr(), nbytes), level=1)
        t_make = time.time() - t0
        file_bytes = os.path.getsize(frag_path)
        # inflate alone (no GPU): a 1 GiB prefix of the text through cratac_inflate_file, as the bound to compare with
        probe = min(nbytes, 1 << 30)
        probe_path = os.path.join(tmpd, "probe.gz")
        _ffi.write_bgzf(probe_path, (host_text.data_ptr(), probe), level=1)
        _ffi.inflate_file(probe_path)
        t0 = time.time()
        _ffi.inflate_file(probe_path)
        inflate_gbs = probe / 1e9 / (time.time() - t0)
        del host_text
        cap = int(nnz_hint * 1.05) + 1024
        out_indptr = torch.empty(nb_hint + 1, dtype=torch.int64, pin_memory=True)
        out_indices = torch.empty(cap, dtype=torch.int64, pin_memory=True)
        out_data = torch.empty(cap, dtype=torch.int32, pin_memory=True)
        outs = (out_indptr.numpy(), out_indices.numpy(), out_data.numpy())
        names = np.array([c[0] for c in contigs])
        phases = {}
        best = None
        for it in range(args.e2e_file_steps + 1):             # first pass: warm-up (ring allocation, page cache)
            t = [time.time()]

            def lap(name):
                now = time.time()
                phases[name] = 1e3 * (now - t[0])
                t[0] = now

            t_start = t[0]
            r = ctx.run_file(frag_path, odds_ratio=0.2)
            lap("run_file")
            pc, ps, pe = ctx.get_peaks(int(r.n_peaks))
            indptr, indices, data = ctx.get_matrix(int(r.nnz), out=outs)
            barcodes = ctx.get_barcodes()
            lap("d2h")
            chrom_names = names[pc]
            with open(os.path.join(tmpd, "peaks.bed"), "w") as f:
                f.write("".join("%s\t%d\t%d\n" % x for x in zip(chrom_names.tolist(), ps.tolist(), pe.tolist())))
            lap("write_peaks_bed")
            feature_ids = ["%s:%d-%d" % x for x in zip(chrom_names.tolist(), ps.tolist(), pe.tolist())]
            h5.write_matrix_h5(os.path.join(tmpd, "raw_peak_bc_matrix.h5"), indptr, indices, data, len(feature_ids), barcodes, feature_ids,
                               "GRCh38", sw_version="cratac-b200")
            lap("write_h5")
            total_ms = 1e3 * (time.time() - t_start)
            if it > 0 and (best is None or total_ms < best[0]):
                best = (total_ms, dict(phases), [kv for kv in ctx.host_times() if kv[0].startswith("decode")][-3:])
        out_bytes = os.path.getsize(os.path.join(tmpd, "peaks.bed")) + os.path.getsize(os.path.join(tmpd, "raw_peak_bc_matrix.h5"))
        pcie_ms = 1e3 * nbytes / 55e9
        inflate_ms = 1e3 * nbytes / 1e9 / inflate_gbs
        return {"value": args.fragments / (best[0] / 1e3), "unit": UNIT, "ms_per_step": best[0], "phases_ms": best[1], "decode_host_phases_ms": best[2],
                "input_file_bytes": file_bytes, "text_bytes": int(nbytes), "output_file_bytes": out_bytes, "host_threads": os.cpu_count(),
                "inflate_only_gbs_text": inflate_gbs, "bound_ms": {"inflate_alone": inflate_ms, "pcie_text_at_55gbs": pcie_ms},
                "run_file_over_bound": best[1]["run_file"] / max(inflate_ms, pcie_ms), "input_written_in_s": t_make, "tmpdir": tempfile.gettempdir(),
                "steps": args.e2e_file_steps, "timing": "host wall clock, best of the steps after one warm-up"}
    finally:
        shutil.rmtree(tmpd, ignore_errors=True)


# --------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cratac", choices=["cratac", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 4], help="BASELINE.json configuration (1-based): 1 chr1 / 5 M / 500 cells, "
                    "2 (default, the metric's) whole GRCh38 / 250 M / 10 k cells, 4 aggr shape 2 B / 80 k cells")
    ap.add_argument("--fragments", type=int, default=None, help="total fragments of the workload (default: the configuration's)")
    ap.add_argument("--cells", type=int, default=None)
    ap.add_argument("--peaks", type=int, default=None)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--in-peak-frac", type=float, default=0.5, help="share of the fragments placed around latent peak centres; the rest is uniform background")
    ap.add_argument("--per-rank-library", action="store_true", help="every rank generates only its own contig range (different libraries for different N)")
    ap.add_argument("--ncu-csv", default=None, help="per-kernel CSV (tools/ncu_summary.py) of an `ncu --set full` capture of THIS command line: "
                    "source of roofline.traffic; without it traffic is null")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=VALUE", help="cratac_set_option on the context (experiments; recorded in config)")
    ap.add_argument("--no-checksum", action="store_true")
    ap.add_argument("--no-e2e-file", action="store_true")
    ap.add_argument("--e2e-file-steps", type=int, default=2)
    args = ap.parse_args()

    from cratac import synth
    contigs_all = apply_config(args, synth)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, world, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from cratac import _ffi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    # ---- shard the genome into contiguous contig ranges (rank order = genome order)
    from cratac import multi
    parts = multi.partition_contigs([c[1] for c in contigs_all], world)
    first_c, last_c = parts[rank]
    contigs = contigs_all[first_c:last_c]
    share = sum(c[1] for c in contigs) / float(args.genome_bases)
    n_local = int(round(args.fragments * share))
    peaks_local = max(int(round(args.peaks * share)), 1)
    G = sum(c[1] for c in contigs)

    # every rank knows all contigs (global ids); it piles up / calls peaks on its own range only
    ctx = _ffi.Context(contigs_all, primary=[c[0] for c in contigs], device=local_rank)
    ctx.set_option("profile", 1)
    for kv in args.option:
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    if args.per_rank_library:
        text, nbytes, n_local = make_workload(torch, ctx, contigs, n_local, args.cells, peaks_local, args.seed + 17 * rank, device, chrom_offset=first_c, in_peak_frac=args.in_peak_frac)
    else:
        text, nbytes, n_local = make_workload(torch, ctx, contigs_all, args.fragments, args.cells, args.peaks, args.seed, device, keep=(first_c, last_c), in_peak_frac=args.in_peak_frac)
    torch.cuda.empty_cache()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    trace = []
    last_shard = [None]

    def step_device():
        if world == 1:
            return ctx.run(device_ptr=text.data_ptr(), nbytes=nbytes, odds_ratio=0.2)
        with torch.cuda.stream(stream):
            shard = multi.CudaShard(ctx, torch, device, text_dev_ptr=text.data_ptr(), nbytes=nbytes)
            last_shard[0] = shard
            trace.clear()
            return multi.step(shard, torch, dist, device, trace=trace)

    # ---- device-resident timing
    res = None
    for _ in range(args.warmup):
        res = step_device()
    stage_ms = {}
    sampler = ClockSampler(list(range(world)) if rank == 0 else [])   # single-node runs: local GPU index = rank
    # host hygiene: the objects of the set-up (workload, contexts) leave the collector's generations, so that a collection
    # inside the timed region has little to walk; what the collector still costs there is measured and reported
    import gc
    gc.collect()
    gc.freeze()
    gc_pause = {"t": 0.0, "ms": 0.0, "n": 0}

    def on_gc(phase, info):
        if phase == "start":
            gc_pause["t"] = time.time()
        else:
            gc_pause["ms"] += 1e3 * (time.time() - gc_pause["t"])
            gc_pause["n"] += 1
    gc.callbacks.append(on_gc)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count(reset=True)
    ev0.record(stream)
    wall0 = time.time()
    loop_ms = []
    slow_trace = None
    for _ in range(args.steps):
        t_a = time.time()
        res = step_device()
        t_b = time.time()
        if world > 1 and (slow_trace is None or t_b - t_a > slow_trace[0]):
            slow_trace = (t_b - t_a, [(k, round(v, 2)) for k, v in trace])
        for k, v in ctx.stage_times().items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v
        loop_ms.append((round(1e3 * (t_b - t_a), 2), round(1e3 * (time.time() - t_b), 2)))
    ev1.record(stream)
    barrier()
    gc.callbacks.remove(on_gc)
    wall = time.time() - wall0
    clocks = sampler.stop()
    launches = ctx.launch_count(reset=True)
    dev_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    traces_by_rank = None
    if world > 1:
        traces_by_rank = [None] * world
        dist.all_gather_object(traces_by_rank, {"rank": rank, "dev_ms_per_step": ev0.elapsed_time(ev1) / args.steps, "phases": [(k, round(v, 2)) for k, v in trace], "loop_ms": loop_ms, "gc_collections": gc_pause["n"], "gc_ms": round(gc_pause["ms"], 2),
                                                "slowest_step_phases": slow_trace[1] if slow_trace else None})
    total_frags = args.steps * args.fragments
    value = total_frags / (dev_ms / 1000.0)
    free_b, total_b = torch.cuda.mem_get_info(device)
    hbm_used = torch.tensor([(total_b - free_b) / 1e9], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(hbm_used, op=dist.ReduceOp.MAX)
    hbm_used_gb = float(hbm_used.item())      # per GPU, largest over the ranks: text + SoA + matrix buffers + the generator's leftovers

    # ---- checksum of the result (outside the timed region): equal for every --gpus N on the same library
    checksum = None
    if not args.no_checksum:
        if world == 1:
            pk = tuple(torch.from_numpy(x).to(device) for x in ctx.get_peaks(int(res.n_peaks)))
            nnz1, ipp, ixp, dtp = ctx.matrix_device(int(res.nnz))
            ip_t = multi.dev_tensor(torch, ipp, int(res.n_barcodes) + 1, "<i8", device)
            ix_t = multi.dev_tensor(torch, ixp, nnz1, "<i8", device)
            dt_t = multi.dev_tensor(torch, dtp, nnz1, "<i4", device)
            keys = torch.from_numpy(np.array([multi.packed_barcode_key(b) for b in ctx.get_barcodes()], dtype=np.int64)).to(device)
            checksum = result_checksum(torch, dist, 1, device, res.threshold, pk, 0, ip_t, ix_t, dt_t, keys)
        else:
            pk = tuple(torch.from_numpy(np.ascontiguousarray(x)).to(device) for x in last_shard[0].peaks)
            ip_t, ix_t, dt_t = last_shard[0].held_matrix()
            checksum = result_checksum(torch, dist, world, device, res.threshold, pk, res.row_base, ip_t, ix_t, dt_t, res.col_keys)

    # ---- end to end from pinned host text to pinned host results
    e2e = None
    if not args.no_e2e:
        host_text = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host_text.copy_(text[:nbytes])
        torch.cuda.synchronize()
        nb = res.n_barcodes
        out_indptr = torch.empty(nb + 1, dtype=torch.int64, pin_memory=True)
        cap_nnz = int(getattr(res, "nnz_local", res.nnz) * 1.05) + 1024      # N > 1: every rank keeps and copies its own row block
        out_indices = torch.empty(cap_nnz, dtype=torch.int64, pin_memory=True)
        out_data = torch.empty(cap_nnz, dtype=torch.int32, pin_memory=True)
        outs = (out_indptr.numpy(), out_indices.numpy(), out_data.numpy())
        out_indptr_global = torch.empty(nb + 1, dtype=torch.int64, pin_memory=True)
        e2e_trace = []

        def step_e2e():
            if world == 1:
                r = ctx.run(text=(host_text.data_ptr(), nbytes), odds_ratio=0.2)
                ctx.get_peaks(r.n_peaks)
                ctx.get_matrix(r.nnz, out=outs)
                return r
            with torch.cuda.stream(stream):
                shard = multi.CudaShard(ctx, torch, device, text_host=(host_text.data_ptr(), nbytes), nbytes=nbytes)
                e2e_trace.clear()
                r = multi.step(shard, torch, dist, device, trace=e2e_trace)
            # every rank copies ITS row block (CSC over the global columns) and its peaks to pinned host memory over its own
            # PCIe link; the all-reduced indptr of the whole matrix goes with rank 0's
            ctx.get_matrix_cols(r.n_barcodes, r.nnz_local, out=outs)
            if rank == 0:
                out_indptr_global.copy_(r.indptr_global)
            e2e_trace.append(("d2h_block", 0.0))
            return r

        for _ in range(min(args.warmup, 2)):
            step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            r = step_e2e()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        d2h = 12 * int(r.n_peaks) + 8 * (int(r.n_barcodes) + 1) * (world + 1 if world > 1 else 1) + 12 * int(r.nnz)
        e2e_traces = None
        if world > 1:
            e2e_traces = [None] * world
            dist.all_gather_object(e2e_traces, {"rank": rank, "phases_last_step_ms": [(k, round(v, 2)) for k, v in e2e_trace]})
        nbytes_all = torch.tensor([nbytes], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(nbytes_all)
        h2d_total = int(nbytes_all.item())
        e2e = {"value": total_frags / (ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms / args.steps, "phases_by_rank": e2e_traces,
               "ends_with": "peaks + CSC in pinned host memory" if world == 1 else
                            "every rank's peaks + CSC row block (global columns) in its own pinned host memory, global indptr on rank 0"}

    # ---- file to file: fragments.tsv.gz on disk -> peaks.bed + raw_peak_bc_matrix.h5 on disk
    e2e_file = None
    if world == 1 and not args.no_e2e_file:
        e2e_file = file_to_file(args, torch, ctx, text, nbytes, contigs_all, int(res.nnz), int(res.n_barcodes))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant HBM-bound kernel (rank 0's shard)
    peak_gbs, peak_src = measured_peaks()
    K_hits = int(res.n_hits)
    n_frag_r0 = getattr(res, "n_fragments_local", res.n_fragments)
    n_peaks_r0 = getattr(res, "n_peaks_local", res.n_peaks)
    nnz_r0 = getattr(res, "nnz_local", res.nnz)
    alg = algorithmic_bytes(nbytes, n_frag_r0, G, n_peaks_r0, K_hits, nnz_r0, res.n_barcodes)
    kernels = {}
    for name, members in KERNEL_GROUPS.items():
        ms = sum(stage_ms.get(m, 0.0) for m in members) / args.steps
        if ms > 0:
            gbs = alg[name] / 1e9 / (ms / 1000.0)
            kernels[name] = {"ms": ms, "algorithmic_gb": alg[name] / 1e9, "gbs": gbs, "frac": gbs / peak_gbs}
    em_ms = stage_ms.get("em_fit", 0.0) / args.steps
    dom = max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None
    roofline = None
    traffic = None
    ncu_csv = args.ncu_csv
    if dom and ncu_csv and os.path.exists(ncu_csv):
        # DRAM bytes of the dominant group's kernels from the `ncu --set full` capture named on the command line (a capture
        # of this same command; tools/ncu_summary.py writes the CSV): one launch of each kernel of the group
        import csv
        name_of = {"decode": ("k_parse_stream", "k_count_newlines", "k_parse_tiles"), "pileup_window_hist": ("k_pileup<0>",), "pileup_call_peaks": ("k_pileup<1>",),
                   "overlap": ("k_overlap(",), "sort_reduce": ("k_reduce_warp", "k_reduce_small", "k_reduce_large"),
                   "csc_build": ("k_csc_copy",)}[dom]
        rows = list(csv.reader(open(ncu_csv)))
        hdr = rows[0]
        seen, total = set(), 0.0
        for r in rows[2:]:
            for nm in name_of:
                if nm in r[0] and nm not in seen:               # one launch of each kernel
                    seen.add(nm)
                    total += float(r[hdr.index("dram__bytes_read.sum")]) + float(r[hdr.index("dram__bytes_write.sum")])
        if seen:
            traffic = total * 1e9                               # the capture reports Gbyte
    if dom:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak_gbs, "unit": "GB/s",
                    "frac": kernels[dom]["frac"], "traffic": traffic,
                    "traffic_note": ("bytes per step: ncu dram__bytes_read.sum + dram__bytes_write.sum of the group's kernels, " + str(ncu_csv)) if traffic else "no --ncu-csv given: not measured by this run",
                    "peak_source": peak_src,
                    "note": "achieved = SURVEY section 8d algorithmic bytes of the stages the kernel replaces / CUDA-event time"}

    # ---- the stage right after the path (FILTER_PEAK_MATRIX) on the matrix just built: the 10 000
    #      columns with most entries stand in for the cell barcodes; outside the timed region, reported on its own
    filter_info = None
    if world == 1 and int(res.nnz) > 0:
        ip = ctx.get_matrix_indptr(int(res.n_barcodes))
        per_col = np.diff(ip)
        cols = np.sort(np.argsort(-per_col, kind="stable")[: min(args.cells, per_col.size)]).astype(np.int64)
        ctx.select_columns_device(cols)
        torch.cuda.synchronize()
        t0 = time.time()
        reps = 3
        for _ in range(reps):
            n_sel, nnz_sel = ctx.select_columns_device(cols)
        torch.cuda.synchronize()
        ms = 1e3 * (time.time() - t0) / reps
        gb = 2 * 12 * nnz_sel / 1e9
        filter_info = {"ms_host_wall": ms, "columns_in": int(cols.size), "columns_out": int(n_sel), "nnz_out": int(nnz_sel),
                       "algorithmic_gb": gb, "gbs": gb / (ms / 1e3), "frac": gb / (ms / 1e3) / peak_gbs,
                       "note": "cratac_select_columns incl. its host round trips; 12 B read + 12 B written per kept entry"}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        sample, desc = cpu_sample(args)
        nf, dt, secs = run_reference_flow(sample)
        cpu_baseline = cpu_baseline_block(args, nf, secs, full_size_fit_seconds(args), desc)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": workload_config(args),
            "clocks": clocks, "e2e": e2e, "e2e_file": e2e_file, "hbm_used_gb_per_gpu_max": hbm_used_gb, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "kernels": kernels, "filter_peak_matrix": filter_info, "em_fit_ms": em_ms, "em_iterations": int(res.em_iterations), "em_inner_iterations": int(res.em_inner_iterations),
            "result": {"n_fragments": int(res.n_fragments), "n_fragments_rank0": int(n_frag_r0), "n_barcodes": int(res.n_barcodes),
                       "threshold": int(res.threshold), "n_peaks": int(res.n_peaks), "n_peaks_rank0": int(n_peaks_r0), "nnz": int(res.nnz),
                       "nnz_rank0": int(nnz_r0), "hits_rank0": K_hits, "text_bytes_rank0": int(nbytes), "checksum": checksum},
            "wall_ms_per_step": 1000.0 * wall / args.steps, "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
            "host_phases_last_step_ms": ctx.host_times()[-12:], "multi_phases_last_step_ms": traces_by_rank,
            "em_cycles_per_iteration": {k: round(v / max(int(res.em_iterations), 1)) for k, v in ctx.em_debug().items()},
            "n_bins": int(getattr(res, "n_bins", 0))}
    line["host_gc_in_timed_region"] = {"collections": gc_pause["n"], "ms": round(gc_pause["ms"], 2)}
    hits, misses = ctx.speculation_stats()
    line["speculation"] = {"kept": int(hits), "redone": int(misses), "second_guesses": ctx.speculation_second_guesses(),
                           "note": "peaks + matrix computed from an early iterate's threshold while the fit finishes; redone when the final "
                                   "threshold differs (cumulative over warm-up, timed and e2e steps; at N > 1 the multi-GPU step keeps its own count)"}
    if args.option:
        line["config"]["options"] = list(args.option)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
