#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Every rank runs cratac.multi.step on its contiguous contig range; rank 0 also runs the whole input through a
single-GPU context (cratac_run) and requires identical threshold, peaks, barcode column order and CSC arrays."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cellranger-atac_b200"))
from cratac import _ffi, multi, synth  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    contigs = [("chr1", 5000000), ("chr2", 3000000), ("chr3", 4000000), ("chr4", 2500000), ("chr5", 1500000), ("chrX", 2000000)]
    n_frag = int(os.environ.get("CHECK_FRAGMENTS", "600000"))
    frags = synth.generate(contigs, n_frag, 80, 1500, seed=41, background_factor=6)
    parts = multi.partition_contigs([c[1] for c in contigs], world)
    a, b = parts[rank]
    m = (frags["chrom"] >= a) & (frags["chrom"] < b)
    sub = {k: (v[m] if isinstance(v, np.ndarray) else v) for k, v in frags.items()}
    text = synth.to_text(sub)
    ctx = _ffi.Context(contigs, primary=[c[0] for c in contigs[a:b]], device=local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
    with torch.cuda.stream(stream):
        # every rank keeps its row block ...
        shard = multi.CudaShard(ctx, torch, device, text_host=text, nbytes=len(text))
        res0 = multi.step(shard, torch, dist, device)
        block = ctx.get_matrix_cols(res0.n_barcodes, res0.nnz_local)
        indptr_global = res0.indptr_global.cpu().numpy()
        # ... or the blocks are funnelled to rank 0 and interleaved on the device
        shard = multi.CudaShard(ctx, torch, device, text_host=text, nbytes=len(text))
        res = multi.step(shard, torch, dist, device, gather=True)
    peaks_local = np.stack(shard.peaks, axis=1) if res.n_peaks_local else np.zeros((0, 3), np.int32)
    gathered = [None] * world
    dist.all_gather_object(gathered, peaks_local)
    blocks = [None] * world
    dist.all_gather_object(blocks, block)
    ok = True
    if rank == 0:
        indptr, indices, data = ctx.get_matrix_cols(res.n_barcodes, res.nnz)
        sip, six, sdt = multi.stack_row_blocks(res.n_barcodes, blocks)
        peaks = np.concatenate(gathered)
        ref = _ffi.Context(contigs, device=local)
        full = synth.to_text(frags)
        r1 = ref.run(text=full)
        c, s, e = ref.get_peaks(r1.n_peaks)
        rip, rix, rdt = ref.get_matrix(r1.nnz)
        keys = [multi.packed_barcode_key(x) for x in ref.get_barcodes()]
        checks = {
            "threshold": res.threshold == r1.threshold,
            "n_peaks": res.n_peaks == r1.n_peaks,
            "peaks": np.array_equal(peaks, np.stack([c, s, e], axis=1)),
            "barcode_order": res.col_keys.cpu().numpy().tolist() == keys,
            "indptr": np.array_equal(indptr, rip),
            "indices": np.array_equal(indices, rix),
            "data": np.array_equal(data, rdt),
            "row_blocks_stacked": np.array_equal(sip, rip) and np.array_equal(six, rix) and np.array_equal(sdt, rdt),
            "column_totals_all_reduced": np.array_equal(indptr_global, rip),
            "same_without_funnel": (res0.threshold, res0.n_peaks, res0.nnz) == (res.threshold, res.n_peaks, res.nnz),
        }
        ok = all(checks.values())
        print("MULTI_GPU_CHECK world=%d fragments=%d peaks=%d nnz=%d %s %s" % (world, res.n_fragments, res.n_peaks, res.nnz,
                                                                              "OK" if ok else "FAIL", checks), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
