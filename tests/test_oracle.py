"""The oracle against the golden vectors produced by executing the reference's own
code (tests/golden/make_golden.py), and the fast/C restatements against the literal one."""
import json
import os
from collections import Counter

import numpy as np
import pytest

from oracle import stages, fast, cfast, em, py2set

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


CUT_CASES = _load("cut_sites.json")
MATRIX_CASES = _load("matrix.json")


def _bed_to_peaks(text):
    return [(c, int(s), int(e)) for c, s, e in (l.split("\t") for l in text.strip("\n").split("\n") if l)]


@pytest.mark.parametrize("case", CUT_CASES, ids=[c["name"] for c in CUT_CASES])
def test_cut_sites_literal_vs_reference(case):
    L = case["contig_len"]
    frags = [tuple(f) for f in case["fragments"]]
    Cuts = stages.pileup(L, frags)
    cd = stages.count_dict_from_cuts(Cuts)
    assert {str(k): v for k, v in cd.items()} == case["count_dict"]
    rows = stages.bedgraph_rows(case["contig"], L, Cuts) if len(cd) else []
    if case["bedgraph"] is None:
        assert rows == []
    else:
        assert stages.format_bedgraph(rows) == case["bedgraph"]
    for thr, want in case["peaks"].items():
        text_rows = [(c, str(s), str(e), str(v)) for c, s, e, v in rows]
        peaks, n = stages.identify_peaks(
            stages.cut_site_counter_from_bedgraph(text_rows, float(thr), case["contig"]), float(thr))
        assert peaks == _bed_to_peaks(want["bed"])
        assert n == want["total_peaks_detected"]


@pytest.mark.parametrize("case", CUT_CASES, ids=[c["name"] for c in CUT_CASES])
def test_cut_sites_fast_and_c_vs_reference(case):
    L = case["contig_len"]
    starts = np.array([f[0] for f in case["fragments"]], dtype=np.int32)
    stops = np.array([f[1] for f in case["fragments"]], dtype=np.int32)
    for mod in (fast, cfast):
        W = mod.window_counts(L, starts, stops)
        if mod is fast:
            cd = fast.count_dict(W)
        else:
            h = cfast.count_hist(W)
            cd = Counter({int(k): int(h[k]) for k in np.nonzero(h)[0]})
        assert {str(k): v for k, v in cd.items()} == case["count_dict"]
        rs, re_, rc = mod.bedgraph_rows(W)
        text = "".join("%s\t%d\t%d\t%d\n" % (case["contig"], s, e, v) for s, e, v in zip(rs, re_, rc))
        assert text == (case["bedgraph"] or "")
        for thr, want in case["peaks"].items():
            if mod is fast:
                ps, pe = fast.peaks_from_rows(rs, re_, rc, int(thr))
            else:
                ps, pe = cfast.identify_peaks(rs, re_, rc, int(thr))
            got = [(case["contig"], int(s), int(e)) for s, e in zip(ps, pe)]
            assert got == _bed_to_peaks(want["bed"])


@pytest.mark.parametrize("case", MATRIX_CASES, ids=[c["name"] for c in MATRIX_CASES])
def test_matrix_literal_vs_reference(case):
    peak_lines = case["peaks"].splitlines(True)
    frags = [tuple(r) for r in case["fragments"]]
    indptr, indices, data, shape = stages.peak_matrix_chunk(peak_lines, case["barcodes"], frags)
    want = case["result"]
    assert list(shape) == want["shape"]
    assert indptr.tolist() == want["indptr"]
    assert indices.tolist() == want["indices"]
    assert data.tolist() == want["data"]
    assert stages.feature_names(peak_lines) == want["features"]


@pytest.mark.parametrize("case", MATRIX_CASES, ids=[c["name"] for c in MATRIX_CASES])
def test_matrix_fast_and_c_vs_reference(case):
    peak_rows = [l.split("\t") for l in case["peaks"].splitlines() if l]
    contigs = sorted({r[0] for r in peak_rows} | {f[0] for f in case["fragments"]})
    cid = {c: i for i, c in enumerate(contigs)}
    bcid = {b: i for i, b in enumerate(case["barcodes"])}
    pc = np.array([cid[r[0]] for r in peak_rows])
    ps = np.array([int(r[1]) for r in peak_rows])
    pe = np.array([int(r[2]) for r in peak_rows])
    fr = [f for f in case["fragments"] if f[3] in bcid]
    fc = np.array([cid[f[0]] for f in fr], dtype=np.int32)
    fs = np.array([f[1] for f in fr], dtype=np.int32)
    fe = np.array([f[2] for f in fr], dtype=np.int32)
    fb = np.array([bcid[f[3]] for f in fr], dtype=np.int32)
    want = case["result"]
    n_bc = len(case["barcodes"])
    a = fast.peak_matrix(pc, ps, pe, fc, fs, fe, fb, n_bc)
    b = cfast.peak_matrix(pc, ps, pe, len(contigs), fc, fs, fe, fb, n_bc)
    for indptr, indices, data in (a, b):
        assert indptr.tolist() == want["indptr"]
        assert indices.tolist() == want["indices"]
        assert data.tolist() == want["data"]


def test_coo_to_csc_matches_scipy():
    from scipy.sparse import csc_matrix, coo_matrix
    rng = np.random.default_rng(5)
    counter = Counter()
    for _ in range(500):
        counter[int(rng.integers(0, 7)), int(rng.integers(0, 40))] += int(rng.integers(1, 4))
    indptr, indices, data, shape = stages.coo_to_csc(counter, 40, 7)
    d, col, row = zip(*[(v, k[0], k[1]) for k, v in counter.items()])
    m = csc_matrix(coo_matrix((d, (row, col)), shape=(40, 7), dtype=int))
    m.sort_indices()
    assert m.indptr.tolist() == indptr.tolist()
    assert m.indices.tolist() == indices.tolist()
    assert m.data.tolist() == data.tolist()


def test_random_fast_vs_literal():
    """vectorised == literal on random contigs with injected piles."""
    rng = np.random.default_rng(99)
    for _ in range(12):
        L = int(rng.integers(3000, 12000))
        n = int(rng.integers(5, 300))
        starts = rng.integers(0, L, size=n)
        stops = np.minimum(starts + rng.integers(0, 900, size=n), L)
        for _ in range(int(rng.integers(0, 4))):
            s = int(rng.integers(0, L))
            k = int(rng.integers(2, 20))
            starts = np.concatenate([starts, [s] * k])
            stops = np.concatenate([stops, [s] * k])
        thr = int(rng.integers(2, 20))
        Cuts, cd, rows, peaks = stages.detect_peaks_contig("c", L, list(zip(starts.tolist(), stops.tolist())), thr)
        W = fast.window_counts(L, starts, stops)
        assert np.array_equal(W, Cuts)
        assert np.array_equal(cfast.window_counts(L, starts, stops), Cuts)
        ps, pe = fast.call_peaks_contig(L, starts, stops, thr)
        assert [(int(s), int(e)) for s, e in zip(ps, pe)] == [(s, e) for _, s, e in peaks]
        ps, pe = cfast.call_peaks_contig(L, starts, stops, thr)
        assert [(int(s), int(e)) for s, e in zip(ps, pe)] == [(s, e) for _, s, e in peaks]


def test_c_parser():
    text = b"chr1\t10\t250\tAAACGAAAGACCATAA-1\t2\nchr1\t11\t99\tTTTGGTTTCACCTTAG-12\t1\r\n\nchrX\t0\t5\tfoo\t7"
    p = cfast.parse(text)
    assert p["n"] == 3
    rows = list(stages.parse_fragments([l for l in text.split(b"\n") if l.strip()]))
    for i, (c, s, e, b, d) in enumerate(rows):
        assert text[p["contig_off"][i]: p["contig_off"][i] + p["contig_len"][i]].decode() == c
        assert text[p["bc_off"][i]: p["bc_off"][i] + p["bc_len"][i]].decode() == b
        assert (p["start"][i], p["stop"][i], p["dup"][i]) == (s, e, d)


def test_get_chunks():
    assert stages.get_chunks(10, 30) == [(i, i + 1) for i in range(10)]
    ch = stages.get_chunks(100, 30)
    assert ch[0][0] == 0 and ch[-1][1] == 100 and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))


def test_py2_hash_known_values():
    # CPython 2.7 64-bit, hash randomisation off: published values
    assert py2set.py2_str_hash("") == 0
    assert py2set.py2_str_hash("a") == 12416037344
    assert py2set.py2_str_hash("abc") == 1453079729188098211
    assert py2set.py2_str_hash("hello") == 840651671246116861


def test_py2_set_order_small():
    # set(['a','b','c','d']) iterates as a, c, b, d in CPython 2.7 (64-bit)
    assert py2set.py2_set_order(["a", "b", "c", "d"]) == ["a", "c", "b", "d"]
    # resize keeps all keys, no duplicates, and is insertion-order dependent only through collisions
    keys = ["BC%05d-1" % i for i in range(5000)]
    out = py2set.py2_set_order(keys + keys[:100])
    assert sorted(out) == sorted(keys)


_EMP = os.path.join(GOLDEN, "em.json")
EM_CASES = _load("em.json") if os.path.exists(_EMP) and os.path.getsize(_EMP) > 0 else []
# every tenth of the 60 seeded histograms of em_random.json (the GPU suite checks all of them against the CUDA fit)
EM_CASES = EM_CASES + _load("em_random.json")[::10]


@pytest.mark.parametrize("case", EM_CASES, ids=[c["name"] for c in EM_CASES])
def test_em_oracle_vs_reference(case):
    cd = {int(v): int(w) for v, w in zip(case["values"], case["weights"])}
    if case.get("raises"):
        with pytest.raises(ZeroDivisionError):
            em.estimate_final_threshold(cd, 1 / 5)
        return
    thr, params = em.estimate_final_threshold(cd, 1 / 5)
    if case["params"] is None:
        assert params is None and thr == case["threshold"]
        return
    assert (None if thr is None else int(thr)) == case["threshold"]
    assert np.allclose(np.array(params, dtype=float), np.array(case["params"]), rtol=1e-12, atol=0)


FILTER_CASES = json.load(open(os.path.join(GOLDEN, "filter.json")))


@pytest.mark.parametrize("case", FILTER_CASES, ids=[c["name"] for c in FILTER_CASES])
def test_filter_peak_matrix_restatement_matches_reference(case):
    """oracle.stages.filter_peak_matrix against vectors produced by the reference's own prune / CountMatrix code."""
    bcs, indptr, indices, data = stages.filter_peak_matrix(case["n_features"], case["bcs"], case["indptr"], case["indices"], case["data"],
                                                           case["cell_barcodes"], case["num_analysis_bcs"], case["random_seed"])
    want = case["result"]
    assert bcs == want["bcs"] and list(indptr) == want["indptr"] and list(indices) == want["indices"] and list(data) == want["data"]


LOW_TARGETING_CASES = json.load(open(os.path.join(GOLDEN, "low_targeting.json")))


@pytest.mark.parametrize("case", LOW_TARGETING_CASES, ids=[c["name"] for c in LOW_TARGETING_CASES])
def test_remove_low_targeting_restatement_matches_reference(case):
    """oracle.stages.remove_low_targeting_contig (the next sibling stage) against vectors produced by
    the reference's own REMOVE_LOW_TARGETING_BARCODES.main + tools/regions.py: fragment-INTERVAL overlap with the
    bisect-on-ends rule (touching on the left counts; nested rows leave `ends` unsorted), per-barcode padded lengths and
    covered bases, peak coverage within 2 kb."""
    got = stages.remove_low_targeting_contig(case["contig"], case["contig_len"], [tuple(p) for p in case["peaks"]],
                                             [tuple(r) for r in case["fragments"]])
    want = case["expect"]
    assert dict(got["fragment_counts"]) == want["fragment_counts"]
    assert {k: v for k, v in got["targeted_counts"].items() if v} == want["targeted_counts"]
    for p in (0, 50000):
        assert dict(got["fragment_lengths"][p]) == want["fragment_lengths"][str(p)]
        assert dict(got["covered_bases"][p]) == want["covered_bases"][str(p)]
    assert got["peak_coverage"] == want["peak_coverage"]


COUNTS_CASES = json.load(open(os.path.join(GOLDEN, "counts_by_barcode.json")))


@pytest.mark.parametrize("case", COUNTS_CASES, ids=[c["name"] for c in COUNTS_CASES])
def test_get_counts_by_barcode_restatement_matches_reference(case):
    """oracle.stages.get_counts_by_barcode against vectors produced by the reference's own tools/io.py:429-555 and
    tools/regions.py: per-barcode fragment / cut-site counts against seven region tracks, TSS and CTCF relative-position
    profiles (padded, strand-aware).  `cell_id` is not in the vectors (CPython-2.7 dict order, see make_golden.py)."""
    read_count, counts, tss_relpos, ctcf_relpos = stages.get_counts_by_barcode(
        case["tracks"], case["peaks"], [tuple(r) for r in case["fragments"]], case["species_list"], case["species_of_contig"],
        case["known_cells"], case["contig"])
    want = case["expect"]
    assert read_count == want["read_count"]
    got = {bc: {k: int(v) for k, v in c.items() if k != "cell_id"} for bc, c in counts.items()}
    # a Counter keeps keys that were written as 0 (is_<species>_cell_barcode) and those that `duplicate += 0` created
    assert got == want["counts_by_barcode"]
    assert {str(k): v for k, v in tss_relpos.items()} == want["tss_relpos"]
    assert {str(k): v for k, v in ctcf_relpos.items()} == want["ctcf_relpos"]
    if case["known_cells"] is not None:
        ids = {bc: c["cell_id"] for bc, c in counts.items()}
        cells = [b for line in case["known_cells"].splitlines() for b in line.split(",")[1:] if b != "null"]
        assert all((ids[bc] == "None") == (bc not in cells) for bc in ids)


CELL_EM_CASES = json.load(open(os.path.join(GOLDEN, "cell_em.json")))


@pytest.mark.parametrize("case", CELL_EM_CASES, ids=[c["name"] for c in CELL_EM_CASES])
def test_cell_caller_fit_restatement_matches_reference(case):
    """oracle.cell_em (two negative binomials over fragments per barcode, DETECT_CELL_BARCODES :84-201) against the
    reference's own functions run on the same histograms: initial threshold and final threshold equal, parameters
    to 1e-9 relative (same arithmetic, same scipy)."""
    from oracle import cell_em
    cd = dict(zip(case["values"], case["weights"]))
    assert int(cell_em.simple_threshold_estimate(cd)) == case["simple_threshold"]
    params = cell_em.estimate_parameters(cd)
    assert np.allclose(np.array(params), np.array(case["params"]), rtol=1e-9, atol=0)
    assert int(cell_em.estimate_threshold(params, 20)) == case["threshold"]


INSERT_SIZES = json.load(open(os.path.join(GOLDEN, "insert_sizes.json")))


@pytest.mark.parametrize("case", INSERT_SIZES["cases"], ids=[c["name"] for c in INSERT_SIZES["cases"]])
def test_insert_size_table_restatement_matches_reference(case):
    """oracle.stages.insert_size_table against REPORT_INSERT_SIZES.main run by make_golden.py: per-barcode size histogram
    (1..1000, >1000; size 0 dropped), row order = barcode chunk order, totals."""
    rows, total = stages.insert_size_table(case["barcodes"], [tuple(r) for r in INSERT_SIZES["fragments"]], case["primary"],
                                           case["exclude_non_nuclear"])
    want = case["expect"]
    assert want["header_ok"] and list(rows) == want["row_order"]
    for bc, hist in rows.items():
        got = {str(n + 1): int(v) for n, v in enumerate(hist[:1000]) if v}
        if hist[1000]:
            got[">1000"] = int(hist[1000])
        assert got == want["table"][bc]
    assert total.tolist() == want["total"]
