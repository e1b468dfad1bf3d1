#!/usr/bin/env python
"""BASELINE.json configs[4]: the mm10 sweep -- peak density 20k..300k x barcodes 1k..100k -- on ONE GPU.

Each point generates a seeded synthetic mm10-shaped fragment text on the device (bench.make_workload), runs the
whole path (decode -> pileup/histogram -> fit -> peaks -> matrix) from device-resident text and prints one JSON line:
fragments/s, ms/step, the per-stage CUDA-event times and each kernel group's algorithmic GB/s as a fraction of the
measured HBM peak (the same accounting as bench.py).  Not a bench line of the driver contract: a table for DESIGN.md.

    python tools/mm10_sweep.py [--fragments 100000000] [--steps 3] [--warmup 2]
"""
from __future__ import print_function

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cellranger-atac_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fragments", type=int, default=100000000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--peaks", default="20000,100000,300000")
    ap.add_argument("--cells", default="1000,10000,100000")
    ap.add_argument("--py2-device-min", default="", help="comma list of values of the library option of that name: every point is timed "
                                                         "once per value on the same text (empty: library default only)")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=VALUE", help="cratac_set_option on every point's context")
    args = ap.parse_args()

    import torch
    import bench
    from cratac import _ffi, synth

    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device")
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    contigs = synth.MM10
    G = sum(c[1] for c in contigs)
    peak_gbs, _ = bench.measured_peaks()
    for n_peaks in [int(x) for x in args.peaks.split(",")]:
        for cells in [int(x) for x in args.cells.split(",")]:
            point = {"genome": "mm10", "fragments": args.fragments, "latent_peaks": n_peaks, "cells": cells}
            try:
                ctx = _ffi.Context(contigs, primary=[c[0] for c in contigs], device=0)
                ctx.set_option("profile", 1)
                for kv in args.option:
                    ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
                    point[kv.split("=")[0]] = int(kv.split("=")[1])
                text, nbytes, _n = bench.make_workload(torch, ctx, contigs, args.fragments, cells, n_peaks, 5, device)
                stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
                settings = [int(x) for x in args.py2_device_min.split(",") if x] or [None]
                for si, setting in enumerate(settings):
                    if setting is not None:
                        ctx.set_option("py2_device_min", setting)
                        point["py2_device_min"] = setting
                    for _ in range(args.warmup):
                        res = ctx.run(device_ptr=text.data_ptr(), nbytes=nbytes, odds_ratio=0.2)
                    torch.cuda.synchronize()
                    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    stage_ms = {}
                    ev0.record(stream)
                    for _ in range(args.steps):
                        res = ctx.run(device_ptr=text.data_ptr(), nbytes=nbytes, odds_ratio=0.2)
                        for k, v in ctx.stage_times().items():
                            stage_ms[k] = stage_ms.get(k, 0.0) + v
                    ev1.record(stream)
                    torch.cuda.synchronize()
                    ms = ev0.elapsed_time(ev1) / args.steps
                    alg = bench.algorithmic_bytes(nbytes, res.n_fragments, G, res.n_peaks, int(res.n_hits), res.nnz, res.n_barcodes)
                    groups = {}
                    for name, members in bench.KERNEL_GROUPS.items():
                        gms = sum(stage_ms.get(m, 0.0) for m in members) / args.steps
                        if gms > 0:
                            groups[name] = {"ms": round(gms, 3), "frac": round(alg[name] / 1e9 / (gms / 1e3) / peak_gbs, 3)}
                    point.update({"ms_per_step": round(ms, 3), "fragments_per_s": args.fragments / (ms / 1e3), "threshold": int(res.threshold),
                                  "n_peaks": int(res.n_peaks), "n_barcodes": int(res.n_barcodes), "nnz": int(res.nnz),
                                  "em_fit_ms": round(stage_ms.get("em_fit", 0.0) / args.steps, 3), "em_iterations": int(res.em_iterations),
                                  "speculation_kept_redone": list(ctx.speculation_stats()), "second_guesses": ctx.speculation_second_guesses(),
                                  "host_phases_last_step_ms": ctx.host_times()[-10:], "stage_ms": {k: round(v / args.steps, 3) for k, v in stage_ms.items()}, "kernels": groups})
                    if si + 1 < len(settings):
                        print(json.dumps(point))
                        sys.stdout.flush()
                del text
                ctx.close()
                torch.cuda.empty_cache()
            except Exception as exc:   # a failing point is reported, the sweep goes on
                point["error"] = "%s: %s" % (type(exc).__name__, exc)
            print(json.dumps(point))
            sys.stdout.flush()


if __name__ == "__main__":
    main()
