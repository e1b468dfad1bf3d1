"""Genome-range sharding of the path over the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink; gloo on CPU for the tests).  The genome is cut into
`world` CONTIGUOUS contig ranges (rank order = genome order), so no halo and no peak stitching is needed and
matrix rows of different ranks never interleave.  Exchanges, all small except the last:

  1. window-count histogram: all-reduce(max) of the largest count, all-reduce(sum) of the bins  -- replaces the
     `Counter +=` of count_cut_sites/__init__.py:77-81; the mixture fit then runs redundantly on every rank.
  2. peak counts and fragment counts: all-gather -> matrix row base and global line base of every rank.
  3. barcode dictionary: all-gather of (64-bit key, first global line, CPython-2.7 hash); union -> the global
     column order of generate_peak_matrix/__init__.py:44 and each rank's column -> local barcode map.
  4. CSC column blocks: gathered to rank 0 and interleaved per column (rank order keeps rows ascending) --
     replaces the hstack of generate_peak_matrix/__init__.py:62-72.

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


# --------------------------------------------------------------------------------------------- barcode union
def global_columns(torch, keys_all, first_all, hash_all, valid_all, py2_order_fn, order="py2set"):
    """keys/first/hash: int64 tensors [world, maxB] (padded), valid_all: bool [world, maxB].
    -> (n_cols, col_of[world, maxB] int64 with -1 on padding, uniq_keys in column order)."""
    flat_valid = valid_all.reshape(-1)
    keys = keys_all.reshape(-1)[flat_valid]
    first = first_all.reshape(-1)[flat_valid]
    hsh = hash_all.reshape(-1)[flat_valid]
    uniq, inv = torch.unique(keys, return_inverse=True)
    n = uniq.numel()
    first_min = torch.full((n,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=keys.device)
    first_min = first_min.scatter_reduce(0, inv, first, reduce="amin", include_self=True)
    hash_u = torch.zeros(n, dtype=torch.int64, device=keys.device).scatter(0, inv, hsh)
    by_first = torch.argsort(first_min)                      # insertion order of the reference's set
    if order == "py2set":
        perm = py2_order_fn(hash_u[by_first].contiguous())   # slot order of the CPython-2.7 set (tensor in, tensor out)
        col_to_uniq = by_first[perm.to(torch.int64)]
    else:
        col_to_uniq = by_first
    col_of_uniq = torch.empty(n, dtype=torch.int64, device=keys.device)
    col_of_uniq[col_to_uniq] = torch.arange(n, dtype=torch.int64, device=keys.device)
    col_flat = torch.full(flat_valid.shape, -1, dtype=torch.int64, device=keys.device)
    col_flat[flat_valid] = col_of_uniq[inv]
    return n, col_flat.reshape(valid_all.shape), uniq[col_to_uniq]


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


def step(shard, torch, dist, device, odds_ratio=0.2, order="py2set", trace=None):
    """One pass of the path on this rank's shard + the cross-rank exchanges.  `shard` implements the backend
    interface (see CudaShard).  Returns a StepResult; the merged matrix lives on rank 0 (shard.merged)."""
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
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    gmax = int(mx.item())                                    # (waits for the histogram pass; only the used bins are exchanged:
    dist.all_reduce(hist[: gmax + 1], op=dist.ReduceOp.SUM)  #  an 8 MB all-reduce of the whole table stalled NCCL for 50-150 ms now and then)
    shard.fit_launch(odds_ratio)
    mark("hist_fit_launch")
    # 2. barcode dictionary, exchanged while the fit runs (side stream on a GPU: the fit occupies one SM)
    with shard.side_stream():
        counts = _all_gather_i64(torch, dist, [n_frag, n_bc], device)
        mark("d_gather_counts")
        line_base = int(counts[:rank, 0].sum())
        keys, first, hsh = shard.barcode_table()             # int64 tensors [n_bc]
        max_b = int(counts[:, 1].max())
        pad = torch.zeros((3, max(max_b, 1)), dtype=torch.int64, device=device)
        pad[0, :n_bc] = keys
        pad[1, :n_bc] = first + line_base
        pad[2, :n_bc] = hsh
        gathered = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(gathered, pad)
        if trace is not None: torch.cuda.current_stream().synchronize()
        mark("d_all_gather")
        g = torch.stack(gathered)                            # [world, 3, maxB]
        valid = torch.arange(max(max_b, 1), device=device)[None, :] < torch.as_tensor(counts[:, 1], device=device)[:, None]
        n_cols, col_of, col_keys = global_columns(torch, g[:, 0], g[:, 1], g[:, 2], valid, shard.py2_order, order)
        mark("d_global_columns")
        c2d = torch.full((max(n_cols, 1),), -1, dtype=torch.int32, device=device)
        mine = col_of[rank, :n_bc]
        c2d[mine] = torch.arange(n_bc, dtype=torch.int32, device=device)
    mark("dictionary")
    # 3. threshold, peaks, row bases
    threshold, params = shard.fit_result()
    mark("fit_wait")
    n_peaks = shard.call_peaks()
    mark("call_peaks")
    pk = _all_gather_i64(torch, dist, [n_peaks], device)[:, 0]
    row_base = int(pk[:rank].sum())
    total_rows = int(pk.sum())
    shard.set_row_offset(row_base, total_rows)
    shard.set_column_map(n_cols, c2d)
    # 4. local CSC over the global columns, gathered to rank 0
    mark("row_col_maps")
    indptr, indices, data = shard.peak_matrix(n_cols)        # tensors (views) on `device`
    mark("peak_matrix")
    nnz = int(indices.numel())
    nnz_all = _all_gather_i64(torch, dist, [nnz], device)[:, 0]
    total_nnz = int(nnz_all.sum())
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
                      col_keys=col_keys, em_iterations=getattr(shard, "em_iterations", 0),
                      em_inner_iterations=getattr(shard, "em_inner_iterations", 0), n_hits=getattr(shard, "n_hits", 0))


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
        self.ctx.fit_threshold_device(odds_ratio)
        self.ctx.set_option("defer_sync", 0)

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
        return (dev_tensor(t, k, n, "<i8", self.device).clone(), dev_tensor(t, f, n, "<i8", self.device).clone(),
                dev_tensor(t, h, n, "<i8", self.device).clone())

    def py2_order(self, hashes):
        t = self.torch
        out = t.empty(hashes.numel(), dtype=t.int32, device=self.device)
        cur = t.cuda.current_stream(self.device)             # the side stream of step(): the fit has the context's own
        cur.synchronize()                                    # `hashes` is complete before the library reads it
        self.ctx.py2_set_order_hashes_device(hashes.numel(), hashes.data_ptr(), out.data_ptr(), stream=cur.cuda_stream)
        return out

    def set_column_map(self, n_cols, c2d):
        if self._side is not None:
            self._side.synchronize()                          # c2d was computed there
        self.torch.cuda.current_stream(self.device).synchronize()
        self.ctx.set_column_map(n_cols, c2d.data_ptr())

    def peak_matrix(self, n_cols):
        nnz, ip, ix, dt = self.ctx.peak_matrix_device()
        t = self.torch
        data = dev_tensor(t, dt, nnz, "<i4", self.device)
        self.n_hits = int(data.sum(dtype=t.int64).item()) if nnz else 0
        return dev_tensor(t, ip, n_cols + 1, "<i8", self.device), dev_tensor(t, ix, nnz, "<i8", self.device), data

    def merge(self, world, n_cols, indptr_all, rank_off, indices_all, data_all, total_nnz):
        self.torch.cuda.current_stream(self.device).synchronize()
        self.ctx.merge_rank_csc(world, n_cols, indptr_all.data_ptr(), rank_off.data_ptr(), indices_all.data_ptr(), data_all.data_ptr(), total_nnz)
        self.merged = (n_cols, total_nnz)
