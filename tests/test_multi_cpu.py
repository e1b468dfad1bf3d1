"""The N>1 exchange logic (cratac/multi.py) on CPU with gloo, world_size 2: the per-shard compute is played by
the oracle, the exchanges (histogram all-reduce, row/line bases, barcode union, CSC gather + merge) are the
product code.  The merged result must equal the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from cratac import _ffi, bgzf, multi, synth, tabix
from oracle import cfast, em as oracle_em, py2set


def test_partition_contigs():
    lens = [c[1] for c in synth.GRCH38]
    for world in (1, 2, 4, 8):
        parts = multi.partition_contigs(lens, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(lens)
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        loads = [sum(lens[a:b]) for a, b in parts]
        assert max(loads) <= 1.25 * sum(lens) / world        # contiguous ranges stay within 25 % of perfect balance
    assert multi.partition_contigs([5, 5], 4)[-1][1] == 2


class OracleShard(object):
    """multi.step backend whose compute is the CPU oracle (tests only)."""

    def __init__(self, contigs, primary, text, first_contig):
        self.contigs, self.primary, self.text, self.first_contig = contigs, primary, text, first_contig
        self.names = [c[0] for c in contigs]

    def check_stream(self):
        pass

    def decode(self):
        o = util.oracle_decode(self.text, self.names)
        self.o = o
        self.bc_list = o["first_seen"]
        self.bc_id = {b: i for i, b in enumerate(self.bc_list)}
        self.frag_bc = np.array([self.bc_id[b] for b in o["barcodes"]], dtype=np.int64)
        first = {}
        for i, b in enumerate(o["barcodes"]):
            first.setdefault(b, i)
        self.first = np.array([first[b] for b in self.bc_list], dtype=np.int64)
        return o["start"].size, len(self.bc_list)

    def window_histogram(self):
        h = util.oracle_hist(self.contigs, self.primary, self.o["chrom"], self.o["start"], self.o["end"])
        cap = 1 << 14
        self.hist = torch.zeros(cap, dtype=torch.int64)
        for v, w in h.items():
            self.hist[v] = w
        self.mx = torch.tensor([max(h) if h else 0], dtype=torch.int64)
        return self.hist, self.mx

    # guesses = (d1, d2): what a speculating fit would publish, as offsets from the fitted threshold (None: nothing)
    guesses = (None, None)

    def fit_launch(self, odds):
        self.odds = odds

    def _final(self):
        gmax = int(self.mx.item())
        cd = {int(v): int(self.hist[v]) for v in range(1, gmax + 1) if int(self.hist[v]) > 0}
        return oracle_em.estimate_final_threshold(cd, self.odds)

    def fit_tentative(self):
        if self.guesses[0] is None:
            return None             # no speculation
        self._tent = int(self._final()[0]) + self.guesses[0]
        return self._tent

    def fit_revised(self):
        if self.guesses[1] is None:
            return None
        self._rev = int(self._final()[0]) + self.guesses[1]
        return self._rev

    def speculating(self, second=False):
        import contextlib

        @contextlib.contextmanager
        def scope():
            self.thr = self._rev if second else self._tent
            yield
        return scope()

    def side_stream(self):
        import contextlib
        return contextlib.nullcontext()

    def fit_result(self):
        thr, params = self._final()
        self.thr = int(thr)
        return self.thr, params

    def call_peaks(self):
        self.peaks = util.oracle_peaks(self.contigs, self.primary, self.o["chrom"], self.o["start"], self.o["end"], self.thr)
        return len(self.peaks)

    def set_row_offset(self, base, total):
        self.row_base, self.total_rows = base, total

    def barcode_table(self):
        # any injective 64-bit key works for the union; use the low 63 bits of the CPython-2.7 hash pair
        keys = np.array([(py2set.py2_str_hash(b) ^ (len(b) << 50)) & ((1 << 62) - 1) for b in self.bc_list], dtype=np.int64)
        assert len(set(keys.tolist())) == len(keys)
        hashes = np.array([py2set.py2_str_hash(b) for b in self.bc_list], dtype=np.int64)
        return torch.from_numpy(keys), torch.from_numpy(self.first.copy()), torch.from_numpy(hashes)

    def union_columns(self, gathered, n_fragments, n_barcodes, max_b):
        """What cratac_union_columns (csrc/multi.cu) does on the device, restated with numpy for the CPU backend: union
        of the keys, smallest global first line per key, insertion order, CPython-2.7 set order of the hashes."""
        g = gathered.numpy()
        line_base = np.concatenate(([0], np.cumsum(n_fragments)[:-1]))
        first_of, hash_of = {}, {}
        for r in range(g.shape[0]):
            for i in range(int(n_barcodes[r])):
                k, f, h = int(g[r, 0, i]), int(g[r, 1, i]) + int(line_base[r]), int(g[r, 2, i])
                first_of[k] = min(first_of.get(k, f), f)
                hash_of[k] = h
        by_first = sorted(first_of, key=first_of.get)
        perm = _ffi.py2_set_order_hashes(np.array([hash_of[k] for k in by_first], dtype=np.int64))   # host-only entry point of libcratac
        self.col_key_list = [by_first[j] for j in perm]
        col_of = {k: j for j, k in enumerate(self.col_key_list)}
        self.n_cols = len(self.col_key_list)
        self.c2d = np.full(max(self.n_cols, 1), -1, dtype=np.int64)
        my_keys = self.barcode_table()[0].numpy()
        for d, k in enumerate(my_keys.tolist()):
            self.c2d[col_of[k]] = d
        return self.n_cols

    def col_keys(self, n_cols):
        return torch.tensor(self.col_key_list, dtype=torch.int64)

    def peak_matrix(self, n_cols):
        d2c = np.full(len(self.bc_list), -1, dtype=np.int64)
        for j, d in enumerate(self.c2d[:n_cols]):
            if d >= 0:
                d2c[d] = j
        indptr, indices, data = util.oracle_matrix(len(self.contigs), self.peaks, self.o["chrom"], self.o["start"], self.o["end"],
                                                   d2c[self.frag_bc].astype(np.int32), n_cols)
        self.last_block = (indptr, indices + self.row_base, data.astype(np.int32))
        return torch.from_numpy(indptr), torch.from_numpy(indices + self.row_base), torch.from_numpy(data.astype(np.int32))

    def merge(self, world, n_cols, indptr_all, rank_off, indices_all, data_all, total_nnz):
        ip = indptr_all.numpy().reshape(world, n_cols + 1)
        cnt = (ip[:, 1:] - ip[:, :-1]).sum(axis=0)
        indptr = np.concatenate(([0], np.cumsum(cnt)))
        indices = np.zeros(total_nnz, dtype=np.int64)
        data = np.zeros(total_nnz, dtype=np.int32)
        for j in range(n_cols):
            dst = indptr[j]
            for r in range(world):
                s, e = int(rank_off[r]) + ip[r, j], int(rank_off[r]) + ip[r, j + 1]
                indices[dst:dst + e - s] = indices_all.numpy()[s:e]
                data[dst:dst + e - s] = data_all.numpy()[s:e]
                dst += e - s
        self.merged = (indptr, indices, data)


def _dataset():
    contigs = [("chr1", 400000), ("chr2", 250000), ("chr3", 300000), ("chr4", 120000)]
    frags = synth.generate(contigs, 60000, 30, 120, seed=31, background_factor=4)
    return contigs, frags


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    contigs, frags = _dataset()
    # every rank reads its own contig range of fragments.tsv.gz through the tabix index (host ingest of the product)
    text, (a, b) = multi.rank_text(os.path.join(out_dir, "fragments.tsv.gz"), contigs, rank, world, threads=2)
    m = (frags["chrom"] >= a) & (frags["chrom"] < b)
    sub = {k: (v[m] if isinstance(v, np.ndarray) else v) for k, v in frags.items()}
    text = text.tobytes()
    assert text == synth.to_text(sub)
    shard = OracleShard(contigs, [c[0] for c in contigs[a:b]], text, a)
    # without the funnel: every rank keeps its row block; the blocks stacked on the host are the whole matrix
    res0 = multi.step(shard, torch, dist, torch.device("cpu"))
    block = tuple(np.asarray(x) for x in shard.last_block)
    np.savez(os.path.join(out_dir, "block_%d.npz" % rank), indptr=block[0], indices=block[1], data=block[2],
             indptr_global=res0.indptr_global.numpy())
    # speculation: a right first guess is kept; a wrong one corrected by the fit's second guess costs a second pass beside
    # the fit; an uncorrected wrong one (or a wrong correction) is redone after the fit.  Same result every time.
    for guesses, want in (((0, None), ["spec_"]), ((1, 0), ["spec_", "spec2_"]), ((1, None), ["spec_", ""]), ((0, 2), ["spec_", "spec2_", ""]),
                          ((3, 3), ["spec_", ""])):
        shard.guesses = guesses
        trace = []
        r = multi.step(shard, torch, dist, torch.device("cpu"), trace=trace)
        assert (r.threshold, r.n_peaks, r.nnz) == (res0.threshold, res0.n_peaks, res0.nnz)
        assert all(np.array_equal(x, y) for x, y in zip(shard.last_block, block))
        assert [name[:-len("call_peaks")] for name, _ in trace if name.endswith("call_peaks")] == want
    shard.guesses = (None, None)
    res = multi.step(shard, torch, dist, torch.device("cpu"), gather=True)
    assert (res.threshold, res.n_peaks, res.nnz) == (res0.threshold, res0.n_peaks, res0.nnz)
    if rank == 0:
        indptr, indices, data = shard.merged
        np.savez(os.path.join(out_dir, "merged.npz"), indptr=indptr, indices=indices, data=data, threshold=res.threshold,
                 n_peaks=res.n_peaks, n_cols=res.n_barcodes, col_keys=res.col_keys.numpy())
    # every rank reports its peaks for the global peak list
    np.save(os.path.join(out_dir, "peaks_%d.npy" % rank), np.array(shard.peaks, dtype=np.int64).reshape(-1, 3))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_matches_single_process_oracle(tmp_path):
    world = 2
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    contigs, frags = _dataset()
    bgzf.write(str(tmp_path / "fragments.tsv.gz"), synth.to_text(frags))
    tabix.build_index(str(tmp_path / "fragments.tsv.gz"))
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(str(tmp_path / "merged.npz"))
    # single-process oracle on the whole dataset
    contigs, frags = _dataset()
    names = [c[0] for c in contigs]
    text = synth.to_text(frags)
    o = util.oracle_decode(text, names)
    hist = util.oracle_hist(contigs, names, o["chrom"], o["start"], o["end"])
    thr, _ = oracle_em.estimate_final_threshold(hist, 0.2)
    assert int(got["threshold"]) == int(thr)
    peaks = util.oracle_peaks(contigs, names, o["chrom"], o["start"], o["end"], int(thr))
    shard_peaks = np.concatenate([np.load(str(tmp_path / ("peaks_%d.npy" % r))) for r in range(world)])
    assert [tuple(p) for p in shard_peaks.tolist()] == peaks and int(got["n_peaks"]) == len(peaks)
    cols = util.oracle_columns(o["first_seen"])              # CPython-2.7 order over the WHOLE file
    assert int(got["n_cols"]) == len(cols)
    want_keys = [(py2set.py2_str_hash(b) ^ (len(b) << 50)) & ((1 << 62) - 1) for b in cols]
    assert got["col_keys"].tolist() == want_keys
    col_id = {b: j for j, b in enumerate(cols)}
    fcol = np.array([col_id[b] for b in o["barcodes"]], dtype=np.int32)
    oip, oix, od = util.oracle_matrix(len(contigs), peaks, o["chrom"], o["start"], o["end"], fcol, len(cols))
    assert np.array_equal(got["indptr"], oip) and np.array_equal(got["indices"], oix) and np.array_equal(got["data"], od)
    # the row blocks every rank kept (gather=False), interleaved by the host helper, are the same matrix
    blocks = [np.load(str(tmp_path / ("block_%d.npz" % r))) for r in range(world)]
    sip, six, sdt = multi.stack_row_blocks(len(cols), [(b["indptr"], b["indices"], b["data"]) for b in blocks])
    assert np.array_equal(sip, oip) and np.array_equal(six, oix) and np.array_equal(sdt, od)
    assert all(np.array_equal(b["indptr_global"], oip) for b in blocks)      # all-reduced per-column totals
