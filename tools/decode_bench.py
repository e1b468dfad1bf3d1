#!/usr/bin/env python
"""Decode kernel variants on one synthetic text resident in HBM (one B200): ms per decode and GB/s of algorithmic bytes.

  python tools/decode_bench.py [--fragments N] [--reps R]

Prints one JSON line per variant: option "decode_kernel" 0 (count pass + general kernel), 1 (single pass, 16 KiB tiles),
2 (single pass, 32 KiB tiles), each at a few caps on resident CTAs per SM."""
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
    ap.add_argument("--cells", type=int, default=10000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import bench
    from cratac import _ffi, synth
    device = torch.device("cuda", 0)
    contigs = synth.GRCH38
    ctx = _ffi.Context(contigs, primary=[c[0] for c in contigs], device=0)
    ctx.set_option("profile", 1)
    text, nbytes, _n = bench.make_workload(torch, ctx, contigs, args.fragments, args.cells, 100000, 2, device)
    peak = bench.measured_peaks()[0]
    # (decode_kernel, decode_mode): mode bit 0 = ticket taken at tile start (no prefetch), bit 1 = line numbers from a count pass, bit 2 = one tile per CTA
    variants = [(0, 0), (1, 1), (1, 2), (1, 6), (2, 6)]
    for kernel, mode in variants:
        ctx.set_option("decode_kernel", kernel)
        ctx.set_option("decode_mode", mode)
        tot = {}
        n = nb = 0
        for rep in range(args.reps + 2):
            n, nb = ctx.decode_device(text.data_ptr(), nbytes)
            if rep >= 2:
                for k, v in ctx.stage_times().items():
                    if k.startswith("decode"):
                        tot[k] = tot.get(k, 0.0) + v
        ms = sum(tot.values()) / args.reps
        alg = nbytes + 20 * n
        print(json.dumps({"decode_kernel": kernel, "decode_mode": mode, "fragments": n, "barcodes": nb, "text_bytes": nbytes,
                          "ms": round(ms, 3), "stages_ms": {k: round(v / args.reps, 3) for k, v in tot.items()},
                          "algorithmic_gbs": round(alg / 1e9 / (ms / 1e3), 1), "frac_of_measured_hbm": round(alg / 1e9 / (ms / 1e3) / peak, 3),
                          "declined_ranges": ctx.decode_declined(reset=True)}), flush=True)
    # where the single-pass kernel's time goes: cycles per phase and tile (one thread's clock per CTA)
    for kernel, mode in ((1, 2),):
        ctx.set_option("decode_kernel", kernel)
        ctx.set_option("decode_mode", mode)
        ctx.set_option("decode_profile", 1)
        ctx.decode_device(text.data_ptr(), nbytes)
        prof = ctx.decode_profile()
        tiles = max(prof.pop("tiles"), 1)
        print(json.dumps({"decode_kernel": kernel, "decode_mode": mode, "cycles_per_tile": {k: round(v / tiles) for k, v in prof.items()}, "tiles": tiles}), flush=True)
        ctx.set_option("decode_profile", 0)
    ctx.close()


if __name__ == "__main__":
    main()
