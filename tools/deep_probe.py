"""Phase by phase over a chr1 shard of a library as deep as BASELINE.json configs[3] (160 M fragments on chr1, 80 k cells):
wall time per call with a device synchronize after each, every line flushed so that a hang names its phase."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cellranger-atac_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fragments", type=int, default=160000000)
    ap.add_argument("--cells", type=int, default=80000)
    ap.add_argument("--peaks", type=int, default=8000)
    ap.add_argument("--in-peak-frac", type=float, default=0.5)
    args = ap.parse_args()
    import torch
    import bench
    from cratac import _ffi, synth
    device = torch.device("cuda", 0)
    contigs = synth.GRCH38[:1]
    t = [time.time()]

    def mark(what):
        torch.cuda.synchronize()
        now = time.time()
        print("%-28s %9.3f s" % (what, now - t[0]))
        sys.stdout.flush()
        t[0] = now

    ctx = _ffi.Context(contigs, primary=[contigs[0][0]], device=0)
    ctx.set_option("profile", 1)
    text, nbytes, n = bench.make_workload(torch, ctx, contigs, args.fragments, args.cells, args.peaks, 2, device, in_peak_frac=args.in_peak_frac)
    mark("make_workload (%d fragments, %.1f GB)" % (n, nbytes / 1e9))
    for rep in range(2):
        ctx.decode_device(text.data_ptr(), nbytes)
        mark("decode")
        ctx.window_histogram()
        mark("window_histogram")
        try:
            res = ctx.fit_threshold_device(0.2)
            mark("fit -> %s" % (res[0],))
            thr = res[0]
        except _ffi.CratacError as exc:
            mark("fit raised: %s" % exc)
            thr = 300
        pk = ctx.call_peaks(thr if thr is not None else 300)
        mark("call_peaks(%s) -> %d peaks" % (thr, len(pk[0])))
        out = ctx.peak_matrix_device()
        mark("peak_matrix_device -> nnz %d" % out[0])
        print({k: round(v, 3) for k, v in ctx.stage_times().items()})
        print("HBM in use: %.1f GB" % ((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
