"""Genome-range sharding of the path over the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink; gloo on CPU for the tests).  The genome is cut into
`world` CONTIGUOUS contig ranges (rank order = genome order), so no halo and no peak stitching is needed and
matrix rows of different ranks never interleave.  Exchanges, all small except the last:

  1. window-count histogram: all-reduce(max) of the largest count, all-reduce(sum) of the bins  -- replaces the
     `Counter +=` of count_cut_sites/__init__.py:77-81; the mixture fit then runs redundantly on every rank.
  2. peak counts and fragment counts: all-gather -> matrix row base and global line base of every rank.
  3. barcode dictionary: one all-gather of (64-bit key, first line in the shard, CPython-2.7 hash); the union, the
     insertion order and the set order are computed by the library on every rank (cratac_union_columns,
     csrc/multi.cu) -> the global column order of generate_peak_matrix/__init__.py:44 and each rank's column ->
     local barcode map.
  4. per-column entry counts: all-reduce(sum) -> the indptr of the whole matrix.  The CSC row blocks themselves stay
     on their ranks (every rank copies its own block to the host over its own PCIe link; `stack_row_blocks` is the
     writer's interleave); `gather=True` funnels them to rank 0 and interleaves them on the device instead -- the
     hstack of generate_peak_matrix/__init__.py:62-72.

`step()` is written against a small backend interface so that the exchange logic runs unchanged on CPU tensors
with gloo (tests/test_multi_cpu.py, where the per-shard compute is played by the oracle) and on the GPU with the
CUDA library (`CudaShard`).
"""
import numpy as np


# --------------------------------------------------------------------------------------------- partition
def partition_contigs(contig_lens, world):
    """Contiguous partition of the contigs into `world` ranges minimising the largest range (DP).
    -> list of (first, last_exclusive) index pairs, one per rank (ranges may be empty when world > n)."""
    n = len(contig_lens)
    pre = [0]
    for L in contig_lens:
        pre.append(pre[-1] + int(L))
    INF = float("inf")
    best = [[INF] * (n + 1) for _ in range(world + 1)]
    cut = [[0] * (n + 1) for _ in range(world + 1)]
    best[0][0] = 0
    for k in range(1, world + 1):
        for j in range(0, n + 1):
            for i in range(0, j + 1):
                if best[k - 1][i] == INF:
                    continue
                cost = max(best[k - 1][i], pre[j] - pre[i])
                if cost < best[k][j]:
                    best[k][j] = cost
                    cut[k][j] = i
    bounds = []
    j = n
    for k in range(world, 0, -1):
        i = cut[k][j]
        bounds.append((i, j))
        j = i
    return bounds[::-1]


def rank_text(fragments_path, contigs, rank, world, index=None, threads=None):
    """Host text of rank `rank`'s contig range of fragments.tsv.gz: only the BGZF blocks the tabix index assigns to
    those contigs are read and inflated (the reference's per-contig chunks go through
    pysam.TabixFile.fetch, lib/python/tools/io.py:82-93).  contigs = [(name, length), ...] in reference order.
    -> (numpy uint8 text, (first, last_exclusive) contig range)."""
    from . import _ffi
    first, last = partition_contigs([c[1] for c in contigs], world)[rank]
    return _ffi.inflate_contigs(fragments_path, [c[0] for c in contigs[first:last]], index=index, threads=threads), (first, last)


def packed_barcode_key(bc):
    """The library's 64-bit key of a 10x barcode "[ACGT]{16}-<group>": 2 bits per base (A 0, C 1, T 2, G 3; first base in
    the top bits) | group << 32.  (Other strings are keyed by a 63-bit hash on the device; not reproduced here.)"""
    seq, grp = bc.split("-")
    v = 0
    for ch in seq:
        v = (v << 2) | ((ord(ch) >> 1) & 3)
    return (int(grp) << 32) | v


# --------------------------------------------------------------------------------------------- the step
class StepResult(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _all_gather_i64(torch, dist, values, device):
    t = torch.tensor(values, dtype=torch.int64, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()


def step(shard, torch, dist, device, odds_ratio=0.2, trace=None, gather=False):
    """One pass of the path on this rank's shard + the cross-rank exchanges.  `shard` implements the backend interface
    (see CudaShard).  Every rank keeps its own CSC row block (global columns, global row numbers; rank order = row
    order, so the whole matrix is the vertical stack of the blocks) together with the all-reduced per-column totals
    (`StepResult.indptr_global`).  gather=True also funnels the blocks to rank 0 and interleaves them per column there
    (shard.merged) -- the hstack of generate_peak_matrix/__init__.py:62-72 as ONE device array, used by the parity check."""
    import time
    t_last = [time.time()]

    def mark(name):      # host wall clock between phase boundaries (no extra synchronisation)
        if trace is not None:
            now = time.time()
            trace.append((name, 1e3 * (now - t_last[0])))
            t_last[0] = now

    world, rank = dist.get_world_size(), dist.get_rank()
    shard.check_stream()         # the collectives below must be ordered after the library's kernels
    n_frag, n_bc = shard.decode()
    mark("decode")
    # 1. histogram: queued, all-reduced in place (largest count and bins), fit queued behind it -- no host wait
    hist, mx = shard.window_histogram()                      # int64 tensors: bins [cap], largest count [1]
    # one small all-gather carries everything the ranks need to know of each other at this point: the largest window count
    # (only the used bins are all-reduced: an 8 MB all-reduce of the whole table stalled NCCL for 50-150 ms now and then) and
    # the fragment / barcode counts of the dictionary exchange below -- every host-synchronised collective less is one
    # exposure less to a late peer
    mine = torch.cat([mx.to(torch.int64).reshape(1), torch.tensor([n_frag, n_bc], dtype=torch.int64, device=device)])
    info = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(info, mine)
    info = torch.stack(info).cpu().numpy()                   # (waits for the histogram pass)
    gmax = int(info[:, 0].max())
    counts = info[:, 1:3]
    mx.fill_(gmax)                                           # the fit reads the largest count from here
    dist.all_reduce(hist[: gmax + 1], op=dist.ReduceOp.SUM)
    shard.fit_launch(odds_ratio)
    mark("hist_fit_launch")
    # 2. barcode dictionary, exchanged and united while the fit runs (side stream on a GPU: the fit occupies one SM).
    #    One all-gather of (key, first line, CPython-2.7 hash) per barcode; the union, the insertion order and the
    #    set order are the library's (cratac_union_columns): same columns on every rank, no host-side tensor algebra.
    with shard.side_stream():
        keys, first, hsh = shard.barcode_table()             # int64 tensors [n_bc]
        max_b = max(int(counts[:, 1].max()), 1)
        pad = torch.zeros((3, max_b), dtype=torch.int64, device=device)
        pad[0, :n_bc] = keys
        pad[1, :n_bc] = first
        pad[2, :n_bc] = hsh
        gathered = torch.empty((world, 3, max_b), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(gathered.view(-1), pad.view(-1))
        mark("d_all_gather")
        n_cols = shard.union_columns(gathered, counts[:, 0], counts[:, 1], max_b)
        mark("d_union_columns")
    # 3. peaks, row bases, local CSC over the global columns, per-column totals combined with one all-reduce (the file's
    #    indptr).  Every rank fits the same histogram, so every rank sees the same tentative threshold a few EM iterations
    #    into the fit (and the same answer to "was one published at all"): peaks + matrix start from it on the speculation
    #    stream while the fit -- one CTA -- runs on, and are kept iff the fit ends on that threshold.
    def peaks_and_matrix(tag):
        n_peaks = shard.call_peaks()
        mark(tag + "call_peaks")
        pk = _all_gather_i64(torch, dist, [n_peaks], device)[:, 0]
        row_base = int(pk[:rank].sum())
        total_rows = int(pk.sum())
        shard.set_row_offset(row_base, total_rows)
        mark(tag + "row_bases")
        indptr, indices, data = shard.peak_matrix(n_cols)    # tensors (views) on `device`
        mark(tag + "peak_matrix")
        nnz = int(indices.numel())
        per_col = (indptr[1:] - indptr[:-1]).contiguous()
        dist.all_reduce(per_col, op=dist.ReduceOp.SUM)
        indptr_global = torch.zeros(n_cols + 1, dtype=torch.int64, device=device)
        torch.cumsum(per_col, 0, out=indptr_global[1:])
        nnz_all = _all_gather_i64(torch, dist, [nnz], device)[:, 0]
        mark(tag + "column_totals")
        return n_peaks, row_base, total_rows, indptr, indices, data, nnz, indptr_global, nnz_all

    tentative = shard.fit_tentative()                        # None: the fit published nothing (same on every rank)
    out = None
    if tentative is not None:
        with shard.speculating():
            out = peaks_and_matrix("spec_")
        # the fit extrapolates where its threshold is heading and says so when that is not the tentative one (same on every
        # rank, at most once): the pass is queued again behind the first, still beside the fit
        revised = shard.fit_revised()
        if revised is not None and revised != tentative:
            tentative = revised
            with shard.speculating(second=True):
                out = peaks_and_matrix("spec2_")
    threshold, params = shard.fit_result()
    mark("fit_wait")
    if out is None or threshold != tentative:
        out = peaks_and_matrix("")
    n_peaks, row_base, total_rows, indptr, indices, data, nnz, indptr_global, nnz_all = out
    total_nnz = int(nnz_all.sum())
    if gather:
        if rank == 0:
            indptr_all = torch.empty(world * (n_cols + 1), dtype=torch.int64, device=device)
            indices_all = torch.empty(max(total_nnz, 1), dtype=torch.int64, device=device)
            data_all = torch.empty(max(total_nnz, 1), dtype=torch.int32, device=device)
            offs = np.concatenate(([0], np.cumsum(nnz_all)))
            indptr_all[: n_cols + 1] = indptr
            indices_all[: nnz] = indices
            data_all[: nnz] = data
            ops = []
            for r in range(1, world):                             # all receives posted at once: the transfers of the
                ops.append(dist.P2POp(dist.irecv, indptr_all[r * (n_cols + 1):(r + 1) * (n_cols + 1)], r))   # ranks overlap
                if nnz_all[r] > 0:
                    ops.append(dist.P2POp(dist.irecv, indices_all[offs[r]:offs[r + 1]], r))
                    ops.append(dist.P2POp(dist.irecv, data_all[offs[r]:offs[r + 1]], r))
            for w in (dist.batch_isend_irecv(ops) if ops else []):
                w.wait()
            rank_off = torch.as_tensor(offs[:-1].astype(np.int64), device=device)
            shard.merge(world, n_cols, indptr_all, rank_off, indices_all, data_all, total_nnz)
        else:
            ops = [dist.P2POp(dist.isend, indptr.contiguous(), 0)]
            if nnz > 0:
                ops.append(dist.P2POp(dist.isend, indices.contiguous(), 0))
                ops.append(dist.P2POp(dist.isend, data.contiguous(), 0))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        mark("gather_merge")
    return StepResult(n_fragments=int(counts[:, 0].sum()), n_fragments_local=n_frag, n_barcodes=n_cols, threshold=threshold,
                      params=params, n_peaks=total_rows, n_peaks_local=n_peaks, row_base=row_base, nnz=total_nnz, nnz_local=nnz,
                      indptr_global=indptr_global, col_keys=shard.col_keys(n_cols), em_iterations=getattr(shard, "em_iterations", 0),
                      em_inner_iterations=getattr(shard, "em_inner_iterations", 0), n_hits=getattr(shard, "n_hits", 0))


def stack_row_blocks(n_cols, blocks):
    """The whole CSC from the ranks' row blocks [(indptr, indices, data), ...] in rank order (numpy, host): column j of
    the result is column j of block 0, then of block 1, ... -- rows of different ranks never interleave, so the row
    indices stay ascending.  What a writer does with the blocks every rank copied to the host over its own PCIe link."""
    counts = sum(np.diff(b[0]) for b in blocks)
    indptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
    indices = np.empty(int(indptr[-1]), dtype=np.int64)
    data = np.empty(int(indptr[-1]), dtype=np.int32)
    fill = indptr[:-1].copy()
    for ip, ix, dt in blocks:
        lens = np.diff(ip)
        if ix.size == 0:
            continue
        # destination of every entry of this block: its column's running offset + position inside the column
        dst = np.repeat(fill - ip[:-1], lens) + np.arange(ix.size, dtype=np.int64)
        indices[dst] = ix
        data[dst] = dt
        fill += lens
    return indptr, indices, data


# --------------------------------------------------------------------------------------------- CUDA backend
class _DevArray(object):
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def dev_tensor(torch, ptr, n, typestr, device):
    if n == 0:
        dt = {"<i8": torch.int64, "<i4": torch.int32}[typestr]
        return torch.empty(0, dtype=dt, device=device)
    return torch.as_tensor(_DevArray(ptr, n, typestr), device=device)


class CudaShard(object):
    """Backend of step() on one GPU: every compute call goes through libcratac (cratac._ffi.Context)."""

    def __init__(self, ctx, torch, device, text_dev_ptr=None, text_host=None, nbytes=0):
        self.ctx, self.torch, self.device = ctx, torch, device
        self.text_dev_ptr, self.text_host, self.nbytes = text_dev_ptr, text_host, nbytes
        self.merged = None
        self._side = None
        self.speculate = True

    def check_stream(self):
        """step() queues NCCL collectives on torch's CURRENT stream between the library's kernels: that stream must be the
        context's own (enter `torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream()))` first), else the all-reduce of the
        histogram would not be ordered after the kernel that fills it and the fit would run on un-reduced bins."""
        cur = self.torch.cuda.current_stream(self.device).cuda_stream
        if int(cur) != int(self.ctx.stream()):
            raise RuntimeError("cratac.multi.step: torch's current stream (0x%x) is not the context's stream (0x%x); wrap the call in "
                               "torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream()))" % (int(cur), int(self.ctx.stream())))

    def decode(self):
        if self.text_dev_ptr is not None:
            return self.ctx.decode_device(self.text_dev_ptr, self.nbytes)
        return self.ctx.decode_host(self.text_host)

    def window_histogram(self):
        self.ctx.set_option("defer_sync", 1)
        self.ctx.window_histogram()
        sp, ns = self.ctx.status_device()
        self.status = dev_tensor(self.torch, sp, ns, "<i8", self.device)
        hp, cap = self.ctx.hist_device()
        return dev_tensor(self.torch, hp, cap, "<i8", self.device), self.status[6:7]

    def fit_launch(self, odds_ratio):
        # the all-reduces were queued by torch on the current stream; the library's stream must be that stream
        self.ctx.compact_histogram()
        self.ctx.fit_speculative_launch(odds_ratio)          # queued only; publishes a tentative threshold on its way
        self.ctx.set_option("defer_sync", 0)

    def fit_tentative(self):
        return self.ctx.fit_tentative() if self.speculate else None

    def fit_revised(self):
        return self.ctx.fit_revised() if self.speculate else None

    def speculating(self, second=False):
        """Library calls and torch collectives go to the speculation stream, the peak kernels read the tentative threshold
        (second: the fit's second guess)."""
        import contextlib

        @contextlib.contextmanager
        def scope():
            s = self.ctx.speculation_scope(2 if second else 1)
            try:
                with self.torch.cuda.stream(self.torch.cuda.ExternalStream(s, device=self.device)):
                    yield
            finally:
                self.ctx.speculation_scope(False)
        return scope()

    def side_stream(self):
        if self._side is None:
            self._side = self.torch.cuda.Stream(device=self.device)
        # no wait on the main stream: the dictionary inputs were final when decode() returned (it synchronises)
        return self.torch.cuda.stream(self._side)

    def fit_result(self):
        thr, params, _ = self.ctx.fit_threshold_result()
        self.em_iterations = int(self.status[15].item())
        self.em_inner_iterations = int(self.status[16].item())
        return thr, params

    def call_peaks(self):
        c, s, e = self.ctx.call_peaks(None)
        self.peaks = (c, s, e)
        return len(s)

    def set_row_offset(self, base, total):
        self.ctx.set_row_offset(base, total)

    def barcode_table(self):
        k, f, h, n = self.ctx.barcode_table_device()
        t = self.torch
        return dev_tensor(t, k, n, "<i8", self.device), dev_tensor(t, f, n, "<i8", self.device), dev_tensor(t, h, n, "<i8", self.device)

    def union_columns(self, gathered, n_fragments, n_barcodes, max_b):
        cur = self.torch.cuda.current_stream(self.device)    # the side stream of step(): the fit has the context's own
        self._gathered = gathered                            # kept alive until the library has read it
        return self.ctx.union_columns(len(n_fragments), max_b, gathered.data_ptr(), n_fragments, n_barcodes, stream=cur.cuda_stream)

    def col_keys(self, n_cols):
        p, n = self.ctx.column_keys_device()
        return dev_tensor(self.torch, p, n, "<i8", self.device)

    def peak_matrix(self, n_cols):
        self.ctx.peak_matrix_begin()
        nnz, ip, ix, dt = self.ctx.peak_matrix_finish()
        t = self.torch
        data = dev_tensor(t, dt, nnz, "<i4", self.device)
        self.n_hits = int(data.sum(dtype=t.int64).item()) if nnz else 0
        self._held = (dev_tensor(t, ip, n_cols + 1, "<i8", self.device), dev_tensor(t, ix, nnz, "<i8", self.device), data)
        return self._held

    def held_matrix(self):
        """This rank's CSC row block (device tensors; overwritten on rank 0 by a gather=True step)."""
        return self._held

    def merge(self, world, n_cols, indptr_all, rank_off, indices_all, data_all, total_nnz):
        self.torch.cuda.current_stream(self.device).synchronize()
        self.ctx.merge_rank_csc(world, n_cols, indptr_all.data_ptr(), rank_off.data_ptr(), indices_all.data_ptr(), data_all.data_ptr(), total_nnz)
        self.merged = (n_cols, total_nnz)
