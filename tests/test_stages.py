"""The three Martian stage modules (cellranger-atac_b200/mro/atac/stages/processing/*) driven the way Martian
drives them (split / main / join with Record objects), plus the hand-written HDF5 / MEX writers."""
import importlib.util
import json
import os
import pickle
from collections import Counter

import numpy as np
import pytest

import util
from cratac import bgzf, h5, mex, refmgr, synth
from oracle import reference_flow

STAGES = os.path.join(util.PKG, "mro", "atac", "stages", "processing")


class Record(object):
    """stand-in for martian.Record"""
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def coerce_strings(self):
        pass


def load_stage(name):
    spec = importlib.util.spec_from_file_location("stage_" + name, os.path.join(STAGES, name, "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_mro_declarations_match_the_reference_signatures():
    ccs, dp, gpm = load_stage("count_cut_sites"), load_stage("detect_peaks"), load_stage("generate_peak_matrix")
    assert "stage COUNT_CUT_SITES(" in ccs.__MRO__ and "out pickle     count_dict" in ccs.__MRO__ and "in  string     contig" in ccs.__MRO__
    assert "stage DETECT_PEAKS(" in dp.__MRO__ and "in  float[]    params" in dp.__MRO__ and "out json       peak_metrics" in dp.__MRO__
    assert "stage GENERATE_PEAK_MATRIX(" in gpm.__MRO__ and "out h5     raw_matrix" in gpm.__MRO__ and "in  file   barcodes" in gpm.__MRO__
    for m in (ccs, dp, gpm):
        assert callable(m.split) and callable(m.main) and callable(m.join)


def test_null_inputs_propagate_as_null_outputs():
    ccs, dp, gpm = load_stage("count_cut_sites"), load_stage("detect_peaks"), load_stage("generate_peak_matrix")
    assert ccs.split(Record(fragments=None)) == {'chunks': [], 'join': {}}
    assert dp.split(Record(cut_sites=None)) == {'chunks': [], 'join': {}}
    assert gpm.split(Record(fragments=None)) == {'chunks': [], 'join': {}}
    o = Record(count_dict="x", cut_sites="y")
    ccs.join(Record(fragments=None), o, [], [])
    assert o.count_dict is None and o.cut_sites is None
    o = Record(peaks="x", peak_metrics="y")
    dp.join(Record(cut_sites=None), o, [], [])
    assert o.peaks is None and o.peak_metrics is None
    o = Record(raw_matrix="x", raw_matrix_mex="y")
    gpm.join(Record(fragments=None), o, [], [])
    assert o.raw_matrix is None and o.raw_matrix_mex is None


def test_h5_round_trip(tmp_path):
    rng = np.random.default_rng(8)
    n_feat, n_bc = 57, 23
    dense = (rng.random((n_feat, n_bc)) < 0.1) * rng.integers(1, 9, size=(n_feat, n_bc))
    indptr, indices, data = [0], [], []
    for j in range(n_bc):
        rows = np.nonzero(dense[:, j])[0]
        indices += rows.tolist()
        data += dense[rows, j].tolist()
        indptr.append(len(indices))
    barcodes = ["".join(rng.choice(list("ACGT"), size=16)) + "-1" for _ in range(n_bc)]
    feats = ["chr%d:%d-%d" % (i % 3 + 1, 100 * i, 100 * i + 50) for i in range(n_feat)]
    p = str(tmp_path / "m.h5")
    h5.write_matrix_h5(p, indptr, indices, data, n_feat, barcodes, feats, "GRCh38", sw_version="1.2.0")
    blob = open(p, "rb").read()
    assert blob[:8] == b"\x89HDF\r\n\x1a\n"
    m = h5.read_matrix_h5(p)
    assert m["attrs"] == {"filetype": "matrix", "version": 2, "software_version": "1.2.0"}
    assert m["indptr"].tolist() == indptr and m["indices"].tolist() == indices and m["data"].tolist() == data
    assert m["indptr"].dtype == np.int64 and m["indices"].dtype == np.int64 and m["data"].dtype == np.int32
    assert m["shape"].tolist() == [n_feat, n_bc] and m["shape"].dtype == np.int32
    assert [b.decode() for b in m["barcodes"]] == barcodes
    f = m["features"]
    assert sorted(f) == ["_all_tag_keys", "derivation", "feature_type", "genome", "id", "name"]
    assert [x.decode() for x in f["id"]] == feats and [x.decode() for x in f["name"]] == feats
    assert set(f["feature_type"].tolist()) == {b"Peaks"} and set(f["genome"].tolist()) == {b"GRCh38"}
    assert [x.decode() for x in f["_all_tag_keys"]] == ["genome", "derivation"]
    # empty matrix
    h5.write_matrix_h5(p, [0], [], [], 0, [], [], "GRCh38")
    m = h5.read_matrix_h5(p)
    assert m["indptr"].tolist() == [0] and m["indices"].size == 0 and m["shape"].tolist() == [0, 0]


def test_h5_matrix_arrays_are_stored_like_the_reference_stores_them(tmp_path):
    """cellranger/matrix.py:441-447: chunks of 80000, gzip, shuffle, unlimited maximum dimension.  The bytes are checked
    against the HDF5 file format specification by hand (no libhdf5 in the image): version-3 chunked layout message,
    version-1 filter pipeline (shuffle, then deflate level 4), version-1 chunk B-tree whose leaves hold at most 64 chunks
    -- 200 chunks here, so the tree has two levels -- and every chunk is a zlib stream of the byte-shuffled, zero-padded
    80000 elements.  Contiguous storage (compression=None) reads back the same."""
    import struct
    import zlib
    rng = np.random.default_rng(9)
    nnz, n_bc = 200 * 80000 - 12345, 3000
    counts = rng.multinomial(nnz, np.ones(n_bc) / n_bc)
    indptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
    indices = rng.integers(0, 124000, size=nnz).astype(np.int64)
    data = rng.integers(1, 4, size=nnz).astype(np.int32)
    p = str(tmp_path / "m.h5")
    for comp in ("gzip", None):
        h5.write_matrix_h5(p, indptr, indices, data, 124000, ["B%d-1" % i for i in range(n_bc)], [], "GRCh38", compression=comp)
        m = h5.read_matrix_h5(p)
        assert np.array_equal(m["indptr"], indptr) and np.array_equal(m["indices"], indices) and np.array_equal(m["data"], data)
    h5.write_matrix_h5(p, indptr, indices, data, 124000, ["B%d-1" % i for i in range(n_bc)], [], "GRCh38")
    blob = open(p, "rb").read()
    r = h5._Reader(blob)
    mat = r.children(r.children(r.root_hdr)["matrix"])
    msgs = dict(r.messages(mat["indices"]))
    layout = msgs[0x0008]
    assert layout[:3] == bytes([3, 2, 2]) and struct.unpack_from("<II", layout, 11) == (80000, 8)          # v3, chunked, rank + 1, chunk dims
    assert msgs[0x0001][:3] == bytes([1, 1, 1]) and struct.unpack_from("<QQ", msgs[0x0001], 8) == (nnz, 0xFFFFFFFFFFFFFFFF)
    pipe = msgs[0x000B]
    assert pipe[:2] == bytes([1, 2]) and struct.unpack_from("<HHHHI", pipe, 8) == (2, 0, 1, 1, 8) and struct.unpack_from("<HHHHI", pipe, 24) == (1, 0, 1, 1, 4)
    root = struct.unpack_from("<Q", layout, 3)[0]
    assert blob[root:root + 4] == b"TREE" and blob[root + 4] == 1 and blob[root + 5] == 1                  # chunk tree, level 1
    n_leaves = struct.unpack_from("<H", blob, root + 6)[0]
    assert n_leaves == 4                                                                                    # ceil(200 / 64)
    leaf0 = struct.unpack_from("<Q", blob, root + 24 + 24)[0]
    assert blob[leaf0 + 5] == 0 and struct.unpack_from("<H", blob, leaf0 + 6)[0] == 64
    assert struct.unpack_from("<QQ", blob, leaf0 + 8) == (0xFFFFFFFFFFFFFFFF, struct.unpack_from("<Q", blob, root + 24 + 32 + 24)[0])   # siblings
    size, mask, off, zero, addr = struct.unpack_from("<IIQQQ", blob, leaf0 + 24 + 32)                       # second chunk of the first leaf
    assert (mask, off, zero) == (0, 80000, 0)
    raw = zlib.decompress(blob[addr:addr + size])
    assert np.array_equal(np.frombuffer(raw, dtype=np.uint8).reshape(8, 80000).T.copy().view("<i8").ravel(), indices[80000:160000])


def test_mex_writer(tmp_path):
    d = str(tmp_path / "mex")
    mex.save_mex(d, [0, 2, 2, 3], [0, 2, 1], [4, 1, 7], 3, ["A-1", "B-1", "C-1"], ["chr1:1-5", "chr1:9-20", "chr2:0-3"], "1.2.0")
    lines = open(os.path.join(d, "matrix.mtx")).read().splitlines()
    assert lines[0] == "%%MatrixMarket matrix coordinate integer general"
    assert lines[1].startswith("%metadata_json: ") and json.loads(lines[1][len("%metadata_json: "):]) == {"software_version": "1.2.0", "format_version": 2}
    assert lines[2:] == ["3 3 3", "1 1 4", "3 1 1", "2 3 7"]
    assert open(os.path.join(d, "peaks.bed")).read() == "chr1\t1\t5\nchr1\t9\t20\nchr2\t0\t3\n"
    assert open(os.path.join(d, "barcodes.tsv")).read() == "A-1\nB-1\nC-1\n"


def test_native_mtx_writer_equals_per_entry_formatting(tmp_path):
    """cratac_write_mtx against the per-entry "%i %i %i" formatting of CountMatrix.save_mex (cellranger/matrix.py:812-814):
    random CSC matrices incl. empty columns, more chunks than threads, one thread."""
    import time
    from cratac import _ffi
    rng = np.random.default_rng(6)
    for n_rows, n_cols, nnz, threads in ((50, 40, 300, 3), (100000, 3000, 2_500_000, 4), (10, 0, 0, 2), (7, 5, 0, 1)):
        per_col = np.bincount(rng.integers(0, max(n_cols, 1), size=nnz), minlength=n_cols)[:n_cols] if n_cols else np.zeros(0, dtype=np.int64)
        indptr = np.concatenate(([0], np.cumsum(per_col))).astype(np.int64)
        indices = rng.integers(0, n_rows, size=int(indptr[-1])).astype(np.int64)
        data = rng.integers(1, 70000, size=int(indptr[-1])).astype(np.int32)
        p = str(tmp_path / "m.mtx")
        t0 = time.time()
        _ffi.write_mtx(p, "%%header\n3 4 5\n", indptr, indices, data, threads=threads)
        t_native = time.time() - t0
        got = open(p, "rb").read()
        assert got.startswith(b"%%header\n3 4 5\n")
        body = got[len(b"%%header\n3 4 5\n"):]
        if indptr[-1] <= 1000:
            cols = np.repeat(np.arange(1, n_cols + 1), np.diff(indptr))
            want = "".join("%i %i %i\n" % (r + 1, c, d) for r, c, d in zip(indices.tolist(), cols.tolist(), data.tolist()))
            assert body == want.encode()
        else:
            arr = np.array(body.split(), dtype=np.int64).reshape(-1, 3)         # parse back, compare as numbers
            cols = np.repeat(np.arange(1, n_cols + 1), np.diff(indptr))
            assert np.array_equal(arr[:, 0], indices + 1) and np.array_equal(arr[:, 1], cols) and np.array_equal(arr[:, 2], data)
            assert body.count(b"\n") == int(indptr[-1]) and t_native < 20


def test_count_cut_sites_and_detect_peaks_stage_files_on_cpu(tmp_path, monkeypatch):
    """The two stage modules around the cut_sites bedgraph with the GPU context replaced by an oracle-backed fake: what
    is exercised is the host side of the stages -- the Counter pickle, the natively written bedgraph (row for row the
    reference writer's), the natively parsed and thresholded rows, peaks.bed and the metrics."""
    from cratac import stage_util
    from oracle import em as oracle_em, stages
    contigs = [("chr1", 60000), ("chr2", 30000), ("chrM", 5000)]
    primary = ["chr1", "chr2"]
    ref = str(tmp_path / "ref")
    refmgr.write_reference(ref, contigs)
    defs_path = os.path.join(ref, "fasta", "contig-defs.json")
    defs = json.load(open(defs_path))
    defs["primary_contigs"] = primary
    defs["non_nuclear_contigs"] = ["chrM"]
    json.dump(defs, open(defs_path, "w"))
    frags = synth.generate(contigs[:2], 9000, 25, 12, seed=21, background_factor=4)
    rows = stages.parse_fragments(synth.to_text(frags).decode().splitlines())
    by_contig = {c: [(s, e) for cc, s, e, _b, _d in rows if cc == c] for c in primary}
    cuts = {c: stages.pileup(dict(contigs)[c], by_contig[c]) for c in primary}
    count_dict = stages.merge_count_dicts([stages.count_dict_from_cuts(cuts[c]) for c in primary])
    bed_rows = {c: stages.bedgraph_rows(c, dict(contigs)[c], cuts[c]) for c in primary}

    class FakeCtx(object):
        def count_cut_sites(self):
            v = np.array(sorted(count_dict), dtype=np.int64)
            return v, np.array([count_dict[k] for k in v.tolist()], dtype=np.int64)

        def bedgraph_rows(self, contig):
            r = bed_rows[contig]
            return tuple(np.array([x[k] for x in r], dtype=np.int64) for k in (1, 2, 3))

        def fit_threshold(self, values, weights, odds):
            thr, params = oracle_em.estimate_final_threshold(dict(zip(values.tolist(), weights.tolist())), odds)
            return (None if thr is None else int(thr)), params, {}

        def peaks_from_rows(self, chrom, start, end):
            out = []
            for c, s, e in zip(chrom.tolist(), start.tolist(), end.tolist()):
                if out and out[-1][0] == c and s - stage_util.PEAK_MERGE_DISTANCE <= out[-1][2]:
                    out[-1][2] = max(out[-1][2], e)
                else:
                    out.append([c, s, e])
            a = np.array(out, dtype=np.int64).reshape(-1, 3)
            return a[:, 0].astype(np.int32), a[:, 1], a[:, 2]

        def close(self):
            pass

    monkeypatch.setattr(stage_util, "context_for_reference", lambda path, device=0: (refmgr.ReferenceManager(path), FakeCtx()))
    monkeypatch.setattr(stage_util, "load_fragments", lambda ctx, path: None)
    ccs, dp = load_stage("count_cut_sites"), load_stage("detect_peaks")
    args = Record(reference_path=ref, fragments="fragments.tsv.gz", fragments_index="fragments.tsv.gz.tbi")
    outs = Record(cut_sites=str(tmp_path / "cut_sites.bedgraph"), count_dict=str(tmp_path / "count_dict.pickle"))
    ccs.join(args, outs, [], [])
    assert dict(pickle.load(open(outs.count_dict, "rb"))) == dict(count_dict)
    assert open(outs.cut_sites).read() == "".join(stages.format_bedgraph(bed_rows[c]) for c in primary)
    args = Record(cut_sites=outs.cut_sites, reference_path=ref, count_dict=outs.count_dict)
    sp = dp.split(args)
    thr, _params = oracle_em.estimate_final_threshold(count_dict, stage_util.PEAK_ODDS_RATIO)
    assert sp["chunks"][0]["threshold"] == int(thr)
    chunk = Record(**sp["chunks"][0])
    mouts = Record(peaks=str(tmp_path / "chunk_peaks.bed"), peak_metrics=str(tmp_path / "chunk_metrics.json"))
    dp.main(Record(cut_sites=outs.cut_sites, reference_path=ref, count_dict=outs.count_dict, contig=None, params=chunk.params,
                   threshold=float(chunk.threshold)), mouts)
    want = [p for c in primary for p in stages.detect_peaks_contig(c, dict(contigs)[c], by_contig[c], int(thr))[3]]
    assert len(want) > 0
    assert open(mouts.peaks).read() == "".join("%s\t%d\t%d\n" % tuple(p) for p in want)
    jouts = Record(peaks=str(tmp_path / "peaks.bed"), peak_metrics=str(tmp_path / "peak_metrics.json"))
    dp.join(args, jouts, [chunk], [mouts])
    metrics = json.load(open(jouts.peak_metrics))
    assert metrics["total_peaks_detected"] == len(want) and metrics["bases_in_peaks"] == sum(p[2] - p[1] for p in want)


def test_native_bedgraph_io(tmp_path):
    """cratac_write_bedgraph / cratac_read_bedgraph against the per-row Python they replace (count_cut_sites/__init__.py:38-48,
    utils.py:105-115 + detect_peaks/__init__.py:49): many rows, several contigs, every threshold form, error cases."""
    from cratac import _ffi
    rng = np.random.default_rng(10)
    names = ["chr1", "chr2", "GL000219.1", "chrX"]
    p = str(tmp_path / "cs.bedgraph")
    open(p, "w").close()
    want_lines = []
    for c in ("chr1", "GL000219.1", "chrX"):
        n = 1_300_000 if c == "chr1" else 700
        s = np.cumsum(rng.integers(1, 40, size=n)).astype(np.int64)
        e = s + rng.integers(1, 30, size=n)
        v = rng.integers(1, 60, size=n).astype(np.int64)
        _ffi.write_bedgraph(p, c, s, e, v, append=True, threads=3)
        want_lines.append((c, s, e, v))
    text = open(p).read()
    assert text.count("\n") == sum(x[1].size for x in want_lines)
    c, s, e, v = want_lines[1]
    assert "".join("{}\t{}\t{}\t{}\n".format(c, a, b, d) for a, b, d in zip(s.tolist(), e.tolist(), v.tolist())) in text
    assert text.startswith("chr1\t%d\t%d\t%d\n" % (want_lines[0][1][0], want_lines[0][2][0], want_lines[0][3][0]))
    for thr in (None, 1.0, 17.0, 17.5, 1000.0):
        gc, gs, ge = _ffi.read_bedgraph(p, names, thr, threads=4)
        wc, ws, we = [], [], []
        for c, s, e, v in want_lines:
            keep = np.ones(s.size, dtype=bool) if thr is None else ~(v < thr)
            wc.append(np.full(int(keep.sum()), names.index(c), dtype=np.int32)); ws.append(s[keep]); we.append(e[keep])
        assert np.array_equal(gc, np.concatenate(wc)) and np.array_equal(gs, np.concatenate(ws)) and np.array_equal(ge, np.concatenate(we))
    with pytest.raises(_ffi.CratacError):
        _ffi.read_bedgraph(p, ["chr1", "chr2"], None)                    # contig not in the reference (KeyError in the Python loop)
    open(p, "a").write("chr1\t5\t6\n")
    with pytest.raises(_ffi.CratacError):
        _ffi.read_bedgraph(p, names, None)                               # three fields: the unpacking of the Python loop raises
    open(p, "w").write("chr1\t5\t9\t3")                                 # last line without a newline
    gc, gs, ge = _ffi.read_bedgraph(p, names, 2.0)
    assert gc.tolist() == [0] and gs.tolist() == [5] and ge.tolist() == [9]
    open(p, "w").close()
    assert _ffi.read_bedgraph(p, names, None)[0].size == 0


def test_reference_manager(tmp_path):
    contigs = [("chr1", 1000), ("chr2", 500), ("chrX", 300), ("chrM", 16)]
    ref = str(tmp_path / "ref")
    refmgr.write_reference(ref, contigs)
    defs = json.load(open(os.path.join(ref, "fasta", "contig-defs.json")))
    defs["primary_contigs"] = ["chr1", "chr2", "chrX"]
    defs["non_nuclear_contigs"] = ["chrM"]
    json.dump(defs, open(os.path.join(ref, "fasta", "contig-defs.json"), "w"))
    m = refmgr.ReferenceManager(ref)
    assert m.contig_table() == contigs and m.genome == "GRCh38"
    assert m.primary_contigs(allow_sex_chromosomes=True) == ["chr1", "chr2", "chrX"]
    assert m.primary_contigs() == ["chr1", "chr2"]
    assert refmgr.generate_genome_tag(ref) == ["GRCh38"]


@pytest.mark.gpu
def test_three_stages_end_to_end_match_reference_flow(tmp_path):
    contigs = [("chr1", 900000), ("chr2", 500000), ("chrX", 300000)]
    ref = str(tmp_path / "ref")
    refmgr.write_reference(ref, contigs)
    frags = synth.generate(contigs, 110000, 40, 400, seed=77, background_factor=5)
    text = synth.to_text(frags)
    fpath = str(tmp_path / "fragments.tsv.gz")
    bgzf.write(fpath, text)
    want = reference_flow.run(text, contigs, [c[0] for c in contigs], n_procs=4)

    ccs, dp, gpm = load_stage("count_cut_sites"), load_stage("detect_peaks"), load_stage("generate_peak_matrix")
    # ---- COUNT_CUT_SITES
    args = Record(reference_path=ref, fragments=fpath, fragments_index=fpath + ".tbi")
    s = ccs.split(args)
    assert s["chunks"] == []
    outs = Record(cut_sites=str(tmp_path / "cut_sites.bedgraph"), count_dict=str(tmp_path / "count_dict.pickle"))
    ccs.join(args, outs, [], [])
    cd = pickle.load(open(outs.count_dict, "rb"))
    assert isinstance(cd, Counter) and dict(cd) == dict(want["count_dict"])
    # ---- DETECT_PEAKS
    args = Record(cut_sites=outs.cut_sites, reference_path=ref, count_dict=outs.count_dict)
    s = dp.split(args)
    assert len(s["chunks"]) == 1 and s["chunks"][0]["threshold"] == int(want["threshold"])
    chunk = Record(**s["chunks"][0])
    margs = Record(cut_sites=outs.cut_sites, reference_path=ref, count_dict=outs.count_dict, contig=chunk.contig, params=chunk.params,
                   threshold=chunk.threshold)
    mouts = Record(peaks=str(tmp_path / "chunk_peaks.bed"), peak_metrics=str(tmp_path / "chunk_metrics.json"))
    dp.main(margs, mouts)
    jouts = Record(peaks=str(tmp_path / "peaks.bed"), peak_metrics=str(tmp_path / "peak_metrics.json"))
    dp.join(args, jouts, [chunk], [mouts])
    peaks_text = open(jouts.peaks).read()
    assert peaks_text == "".join("%s\t%d\t%d\n" % p for p in want["peaks"])
    metrics = json.load(open(jouts.peak_metrics))
    assert metrics["total_peaks_detected"] == len(want["peaks"]) and metrics["peakcaller_threshold"] == int(want["threshold"])
    assert metrics["bases_in_peaks"] == sum(e - s for _, s, e in want["peaks"])
    if want["params"] is not None:
        assert np.allclose([metrics["peakcaller_" + k] for k in dp.PARAM_NAMES], np.array(want["params"], dtype=float), rtol=1e-6)
    # ---- GENERATE_PEAK_MATRIX
    args = Record(reference_path=ref, fragments=fpath, peaks=jouts.peaks)
    assert gpm.split(args)["chunks"] == []
    gouts = Record(raw_matrix=str(tmp_path / "raw_peak_bc_matrix.h5"), raw_matrix_mex=str(tmp_path / "raw_peak_bc_matrix_mex"))
    gpm.join(args, gouts, [], [])
    m = h5.read_matrix_h5(gouts.raw_matrix)
    assert [b.decode() for b in m["barcodes"]] == want["barcodes"]
    assert np.array_equal(m["indptr"], want["indptr"]) and np.array_equal(m["indices"], want["indices"])
    assert np.array_equal(m["data"], want["data"])
    assert m["shape"].tolist() == [len(want["peaks"]), len(want["barcodes"])]
    assert [x.decode() for x in m["features"]["id"]] == ["%s:%d-%d" % p for p in want["peaks"]]
    assert open(os.path.join(gouts.raw_matrix_mex, "peaks.bed")).read() == peaks_text
    # the bedgraph the stage wrote is the reference writer's, row for row
    assert os.path.getsize(outs.cut_sites) > 0


def test_filter_peak_matrix_stage_null_and_declaration():
    fpm = load_stage("filter_peak_matrix")
    assert "stage FILTER_PEAK_MATRIX(" in fpm.__MRO__ and "in  csv    cell_barcodes" in fpm.__MRO__ and "out path   filtered_matrix_mex" in fpm.__MRO__
    assert fpm.split(Record(raw_matrix=None)) == {'chunks': []}
    o = Record(filtered_matrix="x", filtered_matrix_mex="y")
    fpm.join(Record(raw_matrix=None), o, [], [])
    assert o.filtered_matrix is None


@pytest.mark.gpu
def test_filter_peak_matrix_stage_end_to_end(tmp_path):
    """raw_peak_bc_matrix.h5 + cell_barcodes.csv -> filtered matrix h5 / MEX, against the oracle restatement of the stage."""
    from oracle import stages as ostages
    fpm = load_stage("filter_peak_matrix")
    rng = np.random.default_rng(12)
    n_feat, n_bc = 500, 300
    sizes = np.where(rng.random(n_bc) < 0.25, 0, rng.integers(1, 60, size=n_bc))
    indptr = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
    indices = np.concatenate([np.sort(rng.choice(n_feat, size=k, replace=False)) for k in sizes] + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
    data = rng.integers(1, 7, size=indices.size).astype(np.int32)
    bcs = ["".join("ACGT"[i] for i in rng.integers(0, 4, size=16)) + "-1" for _ in range(n_bc)]
    feats = ["chr1:%d-%d" % (1000 * i, 1000 * i + 500) for i in range(n_feat)]
    raw = str(tmp_path / "raw.h5")
    h5.write_matrix_h5(raw, indptr, indices, data, n_feat, bcs, feats, "GRCh38", sw_version="t")
    cells = {"GRCh38": [bcs[i] for i in rng.choice(n_bc, size=120, replace=False)], "mm10": [bcs[i] for i in rng.choice(n_bc, size=30, replace=False)] + ["AAAA-7"]}
    csv = str(tmp_path / "cell_barcodes.csv")
    with open(csv, "w") as f:
        for sp, v in cells.items():
            f.write(",".join([sp] + v) + "\n")
    outs = Record(filtered_matrix=str(tmp_path / "f.h5"), filtered_matrix_mex=str(tmp_path / "mex"))
    fpm.join(Record(raw_matrix=raw, num_analysis_bcs=80, random_seed=4, cell_barcodes=csv), outs, [], [])
    want = ostages.filter_peak_matrix(n_feat, bcs, indptr.tolist(), indices.tolist(), data.tolist(), {k: set(v) for k, v in cells.items()}, 80, 4)
    got = h5.read_matrix_h5(outs.filtered_matrix)
    assert [b.decode() for b in got["barcodes"].tolist()] == want[0]
    assert got["indptr"].tolist() == want[1] and got["indices"].tolist() == want[2] and got["data"].tolist() == want[3]
    assert got["shape"].tolist() == [n_feat, len(want[0])]
    assert os.path.exists(os.path.join(outs.filtered_matrix_mex, "matrix.mtx.gz")) or os.listdir(outs.filtered_matrix_mex)


@pytest.mark.gpu
def test_setup_aggr_samples_signal_normalization(tmp_path):
    """SETUP_AGGR_SAMPLES (another caller of the same pileup + fit): per-library threshold and signal
    mean over ALL contigs of the reference, rates in join -- against the oracle's histogram + fit."""
    from oracle import em as oracle_em
    spec = importlib.util.spec_from_file_location("stage_setup_aggr", os.path.join(util.PKG, "mro", "atac", "stages", "aggr", "setup_aggr_samples", "__init__.py"))
    sas = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sas)
    assert "stage SETUP_AGGR_SAMPLES(" in sas.__MRO__ and "out pickle     library_info" in sas.__MRO__
    contigs = [("chr1", 700000), ("chr2", 500000), ("chrX", 250000)]
    ref = str(tmp_path / "ref")
    refmgr.write_reference(ref, contigs)
    rows = []
    want = {}
    for lib, (n_frag, seed) in enumerate(((25000, 5), (40000, 6))):
        frags = synth.generate(contigs, n_frag, 30, 300, seed=seed, background_factor=5)
        text = synth.to_text(frags)
        fpath = str(tmp_path / ("fragments_%d.tsv.gz" % lib))
        bgzf.write(fpath, text)
        cells = str(tmp_path / ("singlecell_%d.csv" % lib))
        with open(cells, "w") as f:
            f.write("barcode,total,duplicate,passed_filters,is__cell_barcode\n")
            for k in range(50):
                f.write("BC%d-1,%d,%d,%d,%d\n" % (k, 1000 + 10 * k + lib, 100, 700 + 5 * k, 1 if k % 2 == 0 else 0))
        rows.append("lib%d,%s,%s" % (lib, fpath, cells))
        o = util.oracle_decode(text, [c[0] for c in contigs])
        hist = util.oracle_hist(contigs, [c[0] for c in contigs], o["chrom"], o["start"], o["end"])
        thr, params = oracle_em.estimate_final_threshold(hist, 1 / 5.0)
        want[lib + 1] = (int(thr), 1.0 / params[3])
    csv = str(tmp_path / "aggr.csv")
    with open(csv, "w") as f:
        f.write("library_id,fragments,cells\n" + "\n".join(rows) + "\n")
    s = sas.split(Record(aggr_csv=csv, normalization="signal_mean", reference_path=ref))
    assert [c['n'] for c in s['chunks']] == [0, 1]
    chunk_outs = []
    for c in s['chunks']:
        outs = Record(library_info=str(tmp_path / ("lib_%d.pickle" % c['n'])))
        sas.main(Record(aggr_csv=csv, normalization="signal_mean", reference_path=ref, n=c['n']), outs)
        chunk_outs.append(outs)
    outs = Record(library_info=str(tmp_path / "library_info.pickle"), gem_group_index=None, gem_group_index_json=str(tmp_path / "ggi.json"))
    sas.join(Record(aggr_csv=csv, normalization="signal_mean", reference_path=ref), outs, s['chunks'], chunk_outs)
    info = pickle.load(open(outs.library_info, "rb"))
    for lib_id, (thr, sm) in want.items():
        assert info[lib_id]['original_threshold'] == thr
        assert abs(info[lib_id]['signal_mean'] - sm) <= 1e-6 * sm
        assert info[lib_id]['kind'] == "unique"
    lowest = min(v[1] for v in want.values())
    for lib_id, (thr, sm) in want.items():
        assert abs(info[lib_id]['rate'] - lowest / sm) < 1e-6
    assert info[1]['total_fragments_per_cell'] == float(np.median([1000 + 10 * k for k in range(0, 50, 2)]))
    assert outs.gem_group_index == {"gem_group_index": {1: ("lib0", 1), 2: ("lib1", 1)}}

# ----------------------------------------------------------------------------- SETUP_AGGR_SAMPLES against the reference's own run
SETUP_AGGR = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "setup_aggr.json")))


def _aggr_library_text(lib):
    names = [c[0] for c in SETUP_AGGR["contigs"]]
    return "".join("%s\t%d\t%d\tAAACGAAAGAAACGAA-1\t1\n" % (names[c], a, b) for c, a, b in lib["fragments"]).encode()


def test_setup_aggr_golden_pins_the_oracle():
    """tests/golden/setup_aggr.json holds what the reference's SETUP_AGGR_SAMPLES main + join (executed as they stand by
    make_golden.py) make of two small libraries in every normalisation mode.  The oracle's histogram over ALL contigs + fit
    gives the same thresholds and signal means; the rates follow setup_aggr_samples/__init__.py:150-154."""
    from oracle import em as oracle_em
    contigs = [tuple(c) for c in SETUP_AGGR["contigs"]]
    names = [c[0] for c in contigs]
    want = {c["normalization"]: c for c in SETUP_AGGR["cases"]}
    fitted = {}
    for i, lib in enumerate(SETUP_AGGR["libraries"]):
        o = util.oracle_decode(_aggr_library_text(lib), names)
        hist = util.oracle_hist(contigs, names, o["chrom"], o["start"], o["end"])
        thr, params = oracle_em.estimate_final_threshold(hist, 1 / 5.0)
        fitted[str(i + 1)] = (int(thr), 1.0 / params[3])
    for mode in ("signal_mean", "signal_noise_threshold"):
        info = want[mode]["library_info"]
        for lib_id, (thr, sm) in fitted.items():
            assert info[lib_id]["original_threshold"] == thr
            assert abs(info[lib_id]["signal_mean"] - sm) <= 1e-9 * sm
        key = "signal_mean" if mode == "signal_mean" else "original_threshold"
        lowest = min(info[k][key] for k in info)
        assert all(abs(info[k]["rate"] - lowest / info[k][key]) < 1e-12 and info[k]["kind"] == "unique" for k in info)
    assert all(v["rate"] == 1.0 and v["kind"] == "unique" for v in want[None]["library_info"].values())
    assert {k: v["kind"] for k, v in want["total"]["library_info"].items()} == {"1": "total", "2": "total"}


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [None, "unique", "total", "signal_mean", "signal_noise_threshold"])
def test_setup_aggr_samples_stage_equals_the_reference_run(tmp_path, mode):
    """The drop-in stage (GPU pileup over every contig + fit) on the golden's libraries: library_info after join and the gem
    group index equal what the reference wrote (floats to 1e-8: the fit)."""
    spec = importlib.util.spec_from_file_location("stage_setup_aggr", os.path.join(util.PKG, "mro", "atac", "stages", "aggr", "setup_aggr_samples", "__init__.py"))
    sas = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sas)
    contigs = [tuple(c) for c in SETUP_AGGR["contigs"]]
    ref = str(tmp_path / "ref")
    refmgr.write_reference(ref, contigs)
    rows = []
    for i, lib in enumerate(SETUP_AGGR["libraries"]):
        fpath = str(tmp_path / ("fragments_%d.tsv.gz" % i))
        bgzf.write(fpath, _aggr_library_text(lib))
        cells = str(tmp_path / ("singlecell_%d.csv" % i))
        with open(cells, "w") as f:
            f.write(lib["cells_csv"])
        rows.append("%s,%s,%s" % (lib["library_id"], fpath, cells))
    csv = str(tmp_path / "aggr.csv")
    with open(csv, "w") as f:
        f.write("library_id,fragments,cells\n" + "\n".join(rows) + "\n")
    want = [c for c in SETUP_AGGR["cases"] if c["normalization"] == mode][0]
    s = sas.split(Record(aggr_csv=csv, normalization=mode, reference_path=ref))
    chunk_outs = []
    for c in s['chunks']:
        outs = Record(library_info=str(tmp_path / ("lib_%d.pickle" % c['n'])))
        sas.main(Record(aggr_csv=csv, normalization=mode, reference_path=ref, n=c['n']), outs)
        chunk_outs.append(outs)
    outs = Record(library_info=str(tmp_path / "library_info.pickle"), gem_group_index=None, gem_group_index_json=str(tmp_path / "ggi.json"))
    sas.join(Record(aggr_csv=csv, normalization=mode, reference_path=ref), outs, s['chunks'], chunk_outs)
    info = pickle.load(open(outs.library_info, "rb"))
    assert sorted(str(k) for k in info) == sorted(want["library_info"])
    for lib_id, ref_info in want["library_info"].items():
        got = info[int(lib_id)]
        assert sorted(k for k in got if k not in ("fragments", "cells")) == sorted(ref_info)
        for key, val in ref_info.items():
            if isinstance(val, float):
                assert abs(got[key] - val) <= 1e-8 * abs(val), (lib_id, key, got[key], val)
            else:
                assert got[key] == val, (lib_id, key, got[key], val)
    assert {str(k): list(v) for k, v in outs.gem_group_index["gem_group_index"].items()} == want["gem_group_index"]
    assert json.load(open(outs.gem_group_index_json)) == want["gem_group_index_json"]
