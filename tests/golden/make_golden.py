#!/usr/bin/env python
"""Generate tests/golden/*.json by EXECUTING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference; the GPU box never runs
this).  No reference source is copied into the repository: each stage module is
read from /root/reference at generation time, its import statements are dropped
with ``ast``, three mechanical py2->py3 token shims are applied
(``xrange``->``range``, ``.iteritems()``->``.items()``, open mode ``'rU'``->``'r'``)
and the module body is exec'd in a namespace whose I/O collaborators (martian,
ReferenceManager, pysam-backed fragment iterators, pickle/h5 writers) are
stubs that feed in-memory data and capture results.  The arithmetic that is
executed -- the pileup slice adds, the Counter, write_chrom_bedgraph,
cut_site_counter_from_bedgraph, identify_peaks, estimate_final_threshold,
tenkit Regions bisect, the (barcode, peak) Counter and scipy's COO->CSC -- is
the reference's, unmodified.

Python-2 comparison semantics are supplied by value wrappers, not code edits:
  * ``None >= min_counts``   -> False   (count_cut_sites/__init__.py:42,47)
  * ``"12" < threshold``     -> False   (lib/python/utils.py:112, str vs float)

filter.json (FILTER_PEAK_MATRIX) is produced the same way: the reference's `prune`
(filter_peak_matrix/__init__.py:46-70) and its `CountMatrix` class (cellranger/matrix.py:
__init__, bc_to_int, select_barcodes, filter_barcodes) are exec'd unmodified on scipy
matrices; only cellranger/bisect.pyx (Cython, 12 lines) is restated, and numpy's
``np.array(list, copy=False)`` is given its pre-2.0 meaning.

low_targeting.json and counts_by_barcode.json (the sibling stages of SURVEY 8f rank 3, oracle only so far) come from
REMOVE_LOW_TARGETING_BARCODES.main (cell_calling/remove_low_targeting_barcodes/__init__.py:178-243) and
get_counts_by_barcode (lib/python/tools/io.py:429-555), exec'd with the reference's own Region / Regions
(lib/python/tools/regions.py).

Usage: python tests/golden/make_golden.py [cut_sites|matrix|em|filter|low_targeting|counts_by_barcode|cell_em|insert_sizes ...]   (writes next to this file)
"""
import ast
import io
import re
import json
import os
import sys
import tempfile
import time
import types
from collections import Counter, OrderedDict

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- py2 value shims
class Py2Int(int):
    """int that answers ``None >= self`` with False like CPython 2."""
    def __le__(self, other):
        if other is None:
            return False
        return int.__le__(self, other)


class Py2Float(float):
    """float that answers ``some_str < self`` with False like CPython 2
    (numbers order before every other type)."""
    def __gt__(self, other):
        if isinstance(other, str):
            return False
        return float.__gt__(self, other)


# --------------------------------------------------------------------------- loader
def load_reference_module(relpath, namespace, keep_funcs=None):
    src = open(os.path.join(REF, relpath)).read()
    src = src.replace("xrange(", "range(").replace(".iteritems()", ".items()").replace("'rU'", "'r'")
    if keep_funcs is not None:
        # files with py2-only syntax elsewhere cannot be parsed whole: cut out the
        # wanted top-level def/class blocks by indentation first
        lines = src.split("\n")
        picked, i = [], 0
        while i < len(lines):
            m = re.match(r"(def|class)\s+(\w+)", lines[i])
            if m and m.group(2) in keep_funcs:
                j = i + 1
                while j < len(lines) and (lines[j].strip() == "" or lines[j][0] in " \t#"):
                    j += 1
                picked.append("\n".join(lines[i:j]))
                i = j
            else:
                i += 1
        src = "\n\n".join(picked)
    tree = ast.parse(src)
    body = []
    for node in tree.body:
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            continue
        if keep_funcs is not None:
            if not (isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in keep_funcs):
                continue
        body.append(node)
    tree.body = body
    code = compile(tree, os.path.join(REF, relpath), "exec")
    exec(code, namespace)
    return namespace


class Record(object):
    """stand-in for martian.Record"""
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def coerce_strings(self):
        pass


class FakePickle(object):
    def __init__(self):
        self.store = {}

    def dump(self, obj, f):
        self.store[f.name] = obj

    def load(self, f):
        return self.store[f.name]


class FakeRefMgr(object):
    contig_lengths = {}
    primary = []

    def __init__(self, path):
        pass

    def get_contig_lengths(self):
        return dict(FakeRefMgr.contig_lengths)

    def primary_contigs(self, allow_sex_chromosomes=False):
        return list(FakeRefMgr.primary)


# --------------------------------------------------------------------------- stage runners
def run_count_cut_sites(contig, contig_len, frags, tmpdir):
    """reference COUNT_CUT_SITES.main on in-memory fragments of one contig."""
    fp = FakePickle()
    FakeRefMgr.contig_lengths = {contig: contig_len}

    def parsed_fragments_from_contig(contig=None, filename=None, index=None):
        for s, e in frags:
            yield contig, s, e, "BC", 1

    ns = {"np": np, "Counter": Counter, "pickle": fp, "ReferenceManager": FakeRefMgr,
          "parsed_fragments_from_contig": parsed_fragments_from_contig, "WINDOW_SIZE": 400,
          "combine_csv": None}
    load_reference_module("mro/atac/stages/processing/count_cut_sites/__init__.py", ns)
    # py2: `None >= min_counts` must be False -> rebind the default through a wrapper value
    wcb = ns["write_chrom_bedgraph"]
    ns["write_chrom_bedgraph"] = lambda chrom, L, Cuts, out, min_counts=1: wcb(chrom, L, Cuts, out, Py2Int(min_counts))
    outs = Record(cut_sites=os.path.join(tmpdir, "cs.bedgraph"), count_dict=os.path.join(tmpdir, "cd.pickle"))
    args = Record(fragments="frag.tsv.gz", fragments_index="frag.tsv.gz.tbi", reference_path="ref", contig=contig)
    ns["main"](args, outs)
    cd = fp.store[outs.count_dict]
    bedgraph = open(outs.cut_sites).read() if outs.cut_sites is not None else None
    return {int(k): int(v) for k, v in cd.items()}, bedgraph, fp


def run_detect_peaks_main(bedgraph_text, contig, threshold, tmpdir):
    """reference DETECT_PEAKS.main (utils.cut_site_counter_from_bedgraph + identify_peaks)."""
    ns = {"Counter": Counter, "json": json, "PEAK_MERGE_DISTANCE": 500, "PEAK_ODDS_RATIO": 1 / 5}
    load_reference_module("lib/python/utils.py", ns, keep_funcs={"cut_site_counter_from_bedgraph"})
    load_reference_module("mro/atac/stages/processing/detect_peaks/__init__.py", ns,
                          keep_funcs={"identify_peaks", "main"})
    bg = os.path.join(tmpdir, "all.bedgraph")
    with open(bg, "w") as f:
        f.write(bedgraph_text)
    outs = Record(peaks=os.path.join(tmpdir, "peaks.bed"), peak_metrics=os.path.join(tmpdir, "pm.json"))
    # split passes `threshold` as a float through JSON; wrap it for the py2 str<float rule
    args = Record(cut_sites=bg, threshold=Py2Float(threshold), contig=contig)
    ns["main"](args, outs)
    return open(outs.peaks).read(), json.load(open(outs.peak_metrics))


def reference_em():
    sys.path.insert(0, os.path.join(REF, "lib/python"))
    import analysis.peaks as ref_peaks
    return ref_peaks


def run_generate_peak_matrix_main(peaks_text, barcodes, frag_rows, tmpdir):
    """reference GENERATE_PEAK_MATRIX.main for one barcode chunk; captures the scipy CSC."""
    from scipy.sparse import csc_matrix, coo_matrix, hstack
    captured = {}

    class CountMatrix(object):
        def __init__(self, feature_ref, bcs, matrix):
            captured["bcs"] = list(bcs)
            captured["m"] = matrix
            captured["features"] = feature_ref

        def save_h5_file(self, filename, sw_version=None):
            captured["h5"] = filename

    tk_ns = {}
    load_reference_module("tenkit/lib/python/tenkit/regions.py", tk_ns, keep_funcs={"Regions"})
    import bisect as _bisect
    tk_ns["bisect"] = _bisect
    tk_ns["Region"] = __import__("collections").namedtuple("Region", "start end")
    load_reference_module("tenkit/lib/python/tenkit/bio_io.py", tk_ns,
                          keep_funcs={"get_target_regions_dict", "get_target_regions"})
    tk_bio = types.SimpleNamespace(get_target_regions=tk_ns["get_target_regions"])

    def open_fragment_file(filename):
        for r in frag_rows:
            yield r

    feat_ns = {}
    ns = {"os": os, "csc_matrix": csc_matrix, "coo_matrix": coo_matrix, "hstack": hstack,
          "OrderedDict": OrderedDict, "Counter": Counter,
          "utils": types.SimpleNamespace(generate_genome_tag=lambda p: ["GRCh38"], get_chunks=None),
          "tk_bio": tk_bio, "cr_matrix": types.SimpleNamespace(CountMatrix=CountMatrix),
          "atac_feature_ref": types.SimpleNamespace(
              from_peaks_bed=lambda path, genomes: ["{}:{}-{}".format(*l.strip("\n").split("\t")) for l in open(path)]),
          "atac_matrix": None, "cr_lib_constants": None,
          "open_fragment_file": open_fragment_file,
          "martian": types.SimpleNamespace(get_pipelines_version=lambda: "golden")}
    load_reference_module("mro/atac/stages/processing/generate_peak_matrix/__init__.py", ns)
    pk = os.path.join(tmpdir, "peaks_in.bed")
    with open(pk, "w") as f:
        f.write(peaks_text)
    bcf = os.path.join(tmpdir, "barcodes.txt")
    with open(bcf, "w") as f:
        f.write("\n".join(barcodes))
    args = Record(fragments="frag.tsv.gz", peaks=pk, barcodes=bcf, reference_path="ref")
    outs = Record(raw_matrix=os.path.join(tmpdir, "m.h5"), raw_matrix_mex=None)
    ns["main"](args, outs)
    if outs.raw_matrix is None:
        return None
    m = captured["m"]
    assert m.has_sorted_indices or True
    m.sort_indices()
    return dict(indptr=m.indptr.tolist(), indices=m.indices.tolist(), data=m.data.tolist(),
                shape=list(m.shape), barcodes=captured["bcs"], features=captured["features"])


# --------------------------------------------------------------------------- case builders
def random_contig_case(rng, L, n_frag, piles):
    starts = rng.integers(0, L, size=n_frag)
    lens = np.clip(rng.normal(250, 150, size=n_frag).astype(int), 0, 1200)
    stops = np.minimum(starts + lens, L)
    frags = list(zip(starts.tolist(), stops.tolist()))
    for _ in range(piles):
        s = int(rng.integers(0, L))
        e = min(L, s + int(rng.integers(0, 600)))
        frags += [(s, e)] * int(rng.integers(2, 25))
    frags.sort()
    return frags


def make_cut_site_cases(tmpdir):
    rng = np.random.default_rng(20191)
    cases = []
    fixed = [
        ("GV1_zero_gap", 10000, [(2000, 2000)] * 4 + [(5000, 5000)] * 4, [5, 8, 9]),
        ("GV2_single", 5000, [(1000, 1500)], [1, 2]),
        ("edge_clip", 1000, [(0, 0), (0, 1000), (999, 1000), (1000, 1000), (5, 400), (5, 400)], [1, 2, 3]),
        ("empty", 3000, [], [1]),
        ("zero_gap_unequal", 12000, [(2000, 2000)] * 4 + [(6000, 6000)] * 5, [4, 5]),
        ("zero_gap_chain", 20000, [(2000, 2000)] * 6 + [(5000, 5000)] * 6 + [(9000, 9000)] * 6 + [(9001, 15000)] * 3, [3, 6, 7, 12]),
        ("merge_boundary", 8000, [(1000, 1000)] * 5 + [(1901, 1901)] * 5 + [(2803, 2803)] * 5, [5]),
    ]
    for name, L, frags, thrs in fixed:
        cases.append((name, L, sorted(frags), thrs))
    for k in range(14):
        L = int(rng.integers(3000, 20000))
        frags = random_contig_case(rng, L, int(rng.integers(5, 400)), int(rng.integers(0, 6)))
        thrs = sorted(set(int(t) for t in rng.integers(1, 20, size=3)))
        cases.append(("rand%02d" % k, L, frags, thrs))
    out = []
    for name, L, frags, thrs in cases:
        cd, bedgraph, _ = run_count_cut_sites("chr1", L, frags, tmpdir)
        peaks = {}
        if bedgraph is not None:
            for t in thrs:
                text, metrics = run_detect_peaks_main(bedgraph, "chr1", float(t), tmpdir)
                peaks[str(t)] = dict(bed=text, total_peaks_detected=metrics["total_peaks_detected"])
        out.append(dict(name=name, contig="chr1", contig_len=L, fragments=[list(f) for f in frags],
                        count_dict={str(k): v for k, v in sorted(cd.items())}, bedgraph=bedgraph, peaks=peaks))
    return out


def gv4_histogram():
    """SURVEY.md section 8c GV4."""
    rng = np.random.default_rng(0)
    L = 20000000
    centres = rng.integers(1000, L - 1000, size=2000)
    normals = rng.normal(0, 150, size=(2000, 300))
    uniforms = rng.integers(0, L, size=400000)
    cuts = np.concatenate([uniforms, np.clip((centres[:, None] + normals.astype(int)).ravel(), 0, L - 1)])
    c = np.bincount(cuts, minlength=L + 1)
    S = np.concatenate(([0], np.cumsum(c)))
    p = np.arange(L)
    W = S[np.minimum(p + 201, L + 1)] - S[np.maximum(p - 200, 0)]
    b = np.bincount(W[W > 0])
    return {int(k): int(b[k]) for k in np.nonzero(b)[0]}


def synth_histogram(seed, L, n_uniform, n_peaks, per_peak, sd):
    rng = np.random.default_rng(seed)
    centres = rng.integers(1000, L - 1000, size=n_peaks)
    cuts = np.concatenate([rng.integers(0, L, size=n_uniform),
                           np.clip((centres[:, None] + rng.normal(0, sd, size=(n_peaks, per_peak)).astype(int)).ravel(), 0, L - 1)])
    c = np.bincount(cuts, minlength=L + 1)
    S = np.concatenate(([0], np.cumsum(c)))
    p = np.arange(L)
    W = S[np.minimum(p + 201, L + 1)] - S[np.maximum(p - 200, 0)]
    b = np.bincount(W[W > 0])
    return {int(k): int(b[k]) for k in np.nonzero(b)[0]}


def make_em_cases():
    ref = reference_em()
    cases = []
    hists = [("GV4", gv4_histogram()),
             ("all_above_15", synth_histogram(11, 4000000, 600000, 800, 400, 120)),
             ("medium", synth_histogram(14, 5000000, 250000, 800, 400, 120)),
             ("sparse", synth_histogram(12, 8000000, 90000, 500, 250, 200)),
             ("many_peaks", synth_histogram(13, 6000000, 300000, 3000, 150, 100)),
             ("gate_few_bins", {k: 100000 for k in range(1, 20)}),
             ("gate_few_bases", {k: 10 for k in range(1, 60)})]
    for name, cd in hists:
        t0 = time.time()
        try:
            thr, params = ref.estimate_final_threshold(dict(cd), 1 / 5)
        except Exception as exc:      # the reference itself fails on this histogram: record how
            print("EM case %-14s raises %s" % (name, type(exc).__name__), flush=True)
            cases.append(dict(name=name, values=sorted(cd), weights=[cd[k] for k in sorted(cd)],
                              raises=type(exc).__name__))
            continue
        dt = time.time() - t0
        print("EM case %-14s bins=%4d thr=%s  %.1fs" % (name, len(cd), thr, dt), flush=True)
        cases.append(dict(name=name, values=sorted(cd), weights=[cd[k] for k in sorted(cd)],
                          threshold=None if thr is None else int(thr),
                          params=None if params is None else [float(x) for x in params]))
    return cases


def random_mixture_histogram(seed):
    """Seeded window-count histogram drawn straight from a zero-noise + background-NB + signal mixture with
    randomised sizes, dispersions and signal tails (60 of them pin the fit at 1e-8, tests/golden/em_random.json)."""
    rng = np.random.default_rng(seed)
    n_zero = int(rng.integers(50000, 600000))
    n_bg = int(rng.integers(150000, 1200000))
    n_sig = int(rng.integers(1500, 30000))
    r = float(rng.uniform(2.0, 12.0))
    p = float(rng.uniform(0.15, 0.6))
    tail = float(rng.uniform(0.008, 0.05))
    shift = int(rng.integers(5, 40))
    vals = np.concatenate([rng.geometric(float(rng.uniform(0.3, 0.7)), size=n_zero),
                           rng.negative_binomial(r, p, size=n_bg) + 1,
                           rng.geometric(tail, size=n_sig) + shift])
    b = np.bincount(vals)
    return {int(k): int(b[k]) for k in np.nonzero(b)[0] if k > 0}


def _fit_one_random(seed):
    ref = reference_em()
    cd = random_mixture_histogram(seed)
    case = dict(name="random_%d" % seed, values=sorted(cd), weights=[cd[k] for k in sorted(cd)])
    try:
        thr, params = ref.estimate_final_threshold(dict(cd), 1 / 5)
    except Exception as exc:
        case["raises"] = type(exc).__name__
        return case
    case["threshold"] = None if thr is None else int(thr)
    case["params"] = None if params is None else [float(x) for x in params]
    return case


def make_em_random_cases(n=60):
    """The reference's own estimate_final_threshold (analysis/peaks.py, unmodified) on n seeded histograms."""
    import multiprocessing
    with multiprocessing.Pool(min(8, os.cpu_count() or 1)) as pool:
        return pool.map(_fit_one_random, range(1000, 1000 + n))


def _library_histogram(genome, density, peaks_per_gb, seed):
    """Window-count histogram of a synthetic library (cratac/synth.py, the generator of bench.py) on a small genome at a
    given fragment density and latent-peak density."""
    sys.path.insert(0, os.path.join(HERE, "..", "..", "cellranger-atac_b200"))
    from cratac import synth
    contigs = [("chr1", genome)]
    fr = synth.generate(contigs, int(genome * density), 50, max(int(peaks_per_gb * genome / 1e9), 1), seed)
    cuts = np.concatenate([fr["start"], fr["end"]]).astype(np.int64)
    c = np.bincount(cuts, minlength=genome + 1)
    S = np.concatenate(([0], np.cumsum(c)))
    pos = np.arange(genome)
    W = S[np.minimum(pos + 201, genome + 1)] - S[np.maximum(pos - 200, 0)]
    b = np.bincount(W[W > 0])
    return {int(k): int(b[k]) for k in np.nonzero(b)[0]}


def _one_trajectory(args):
    """The reference's own EM loop (analysis/peaks.py:144-170, replayed iterate by iterate with its own functions) on one
    histogram: the threshold of every iterate's parameters (estimate_threshold, :43-63) and the real-valued crossing behind
    it (log sig/bg interpolated between t and t + 1 against log odds), and the threshold estimate_final_threshold ends on."""
    name, genome, density, peaks_per_gb, seed = args
    ref = reference_em()
    from scipy import stats
    cd = _library_histogram(genome, density, peaks_per_gb, seed)
    odds = 1 / 5
    final, _ = ref.estimate_final_threshold(dict(cd), odds)

    def crossing(params):
        x = np.arange(1000)
        f_sig = 1 - params.frac_zero_noise - params.frac_bg_noise
        y_zero = params.frac_zero_noise * stats.geom.pmf(x, params.p_zero_noise, loc=-1)
        y_bg = params.frac_bg_noise * stats.nbinom.pmf(x, 1 / params.alpha_bg_noise, 1 / (1 + params.alpha_bg_noise * params.mean_bg_noise))
        y_sig = f_sig * stats.geom.pmf(x, params.p_signal, loc=-1)
        with np.errstate(all="ignore"):
            r = y_sig / (y_zero + y_bg)
        idx = np.nonzero(r < odds)[0]
        if len(idx) == 0:
            return -1, -1.0
        t = int(idx[-1])
        if t + 1 >= 1000 or not (r[t] > 0) or not (r[t + 1] > r[t]) or not np.isfinite(r[t + 1]):
            return t, t + 0.5
        return t, float(t + (np.log(odds) - np.log(r[t])) / (np.log(r[t + 1]) - np.log(r[t])))

    threshold = ref.simple_threshold_estimate(cd)
    count_values = np.array(sorted(cd))
    pm = np.empty((3, len(count_values)), dtype=float)
    pm[0, :] = count_values <= ref.DEFAULT_THRESHOLD
    pm[1, :] = (count_values <= threshold) & (count_values > ref.DEFAULT_THRESHOLD)
    pm[2, :] = count_values > threshold
    params = ref.weighted_parameter_estimates(cd, pm)
    last, i, traj = None, 0, []
    while i < 500 and (last is None or any(abs(last[k] - params[k]) / params[k] > 1e-6 for k in range(len(params)))):
        i += 1
        pm = ref.calculate_probability_distribution(cd, params)
        last = params
        params = ref.normalize_params(ref.weighted_parameter_estimates(cd, pm))
        t, c = crossing(params)
        assert t == (-1 if ref.estimate_threshold(params, odds) is None else int(ref.estimate_threshold(params, odds)))
        traj.append([t, round(c, 9)])
    reseeded = 1 / params.p_zero_noise < params.mean_bg_noise
    print("trajectory %-22s iterates=%3d final=%s reseeded=%s" % (name, len(traj), final, reseeded), flush=True)
    return dict(name=name, final=None if final is None else int(final), reseeded=bool(reseeded), trajectory=traj)


def make_spec_trajectories():
    """EM trajectories for the speculation rule (csrc/spec_rule.cuh): libraries at the depth and peak density of the bench
    workloads (BASELINE.json configs[1] and the mm10 sweep of configs[4]) on a 12 Mb genome, four seeds each."""
    import multiprocessing
    g = 12000000
    jobs = []
    for seed in (2, 5, 7, 11):
        for name, dens, ppg in (("grch38_250m_100k", 250e6 / 3.1e9, 100000 / 3.1), ("mm10_100m_20k", 100e6 / 2.73e9, 20000 / 2.73),
                                ("mm10_100m_100k", 100e6 / 2.73e9, 100000 / 2.73), ("mm10_100m_300k", 100e6 / 2.73e9, 300000 / 2.73),
                                ("grch38_120m_30k", 120e6 / 3.1e9, 30000 / 3.1)):
            jobs.append(("%s_seed%d" % (name, seed), g, dens, ppg, seed))
    with multiprocessing.Pool(min(8, os.cpu_count() or 1)) as pool:
        return pool.map(_one_trajectory, jobs)


def make_setup_aggr_cases(tmpdir):
    """reference SETUP_AGGR_SAMPLES main + join (mro/atac/stages/aggr/setup_aggr_samples/__init__.py, executed as it stands)
    on an aggregation of two small libraries, once per normalisation mode."""
    import pandas as pd
    sys.path.insert(0, os.path.join(HERE, "..", "..", "cellranger-atac_b200"))
    from cratac import synth
    ref_peaks = reference_em()
    contigs = [("chr1", 260000), ("chr2", 180000), ("chrM", 16000)]
    libs = []
    for lib, (n_frag, seed) in enumerate(((2600, 5), (4200, 6))):
        fr = synth.generate(contigs, n_frag, 30, 40, seed=seed, background_factor=5)
        rows = [[int(c), int(a), int(b)] for c, a, b in zip(fr["chrom"], fr["start"], fr["end"])]
        cells_csv = "barcode,total,duplicate,passed_filters,is__cell_barcode\n" + "".join(
            "BC%d-1,%d,%d,%d,%d\n" % (k, 1000 + 10 * k + 37 * lib, 100 + k, 700 + 5 * k + 11 * lib, 1 if k % 2 == 0 else 0) for k in range(50))
        libs.append(dict(library_id="lib%d" % lib, fragments=rows, cells_csv=cells_csv))

    class RefMgr(object):
        def __init__(self, path):
            pass

        def get_contig_lengths(self):
            return dict(contigs)

        def list_species(self):
            return [""]

    class Martian(object):
        @staticmethod
        def exit(msg):
            raise SystemExit(msg)

    frag_files, cell_files = {}, {}
    for i, lib in enumerate(libs):
        cells = os.path.join(tmpdir, "singlecell_%d.csv" % i)
        with open(cells, "w") as f:
            f.write(lib["cells_csv"])
        frag_files["fragments_%d.tsv.gz" % i] = lib["fragments"]
        cell_files[i] = cells
    aggr_csv = os.path.join(tmpdir, "aggr.csv")
    with open(aggr_csv, "w") as f:
        f.write("library_id,fragments,cells\n" + "".join("%s,fragments_%d.tsv.gz,%s\n" % (lib["library_id"], i, cell_files[i]) for i, lib in enumerate(libs)))

    def open_fragment_file(filename=None):
        for c, a, b in frag_files[filename]:
            yield contigs[c][0], a, b, "BC", 1

    def plain(v):
        if isinstance(v, (np.floating, float)):
            return float(v)
        if isinstance(v, (np.integer, int)):
            return int(v)
        return v

    cases = []
    for normalization in (None, "unique", "total", "signal_mean", "signal_noise_threshold"):
        fp = FakePickle()
        ns = {"pd": pd, "np": np, "Counter": Counter, "pickle": fp, "json": json, "martian": Martian,
              "estimate_final_threshold": ref_peaks.estimate_final_threshold, "ReferenceManager": RefMgr,
              "open_fragment_file": open_fragment_file, "WINDOW_SIZE": 400, "PEAK_ODDS_RATIO": 1 / 5}
        load_reference_module("mro/atac/stages/aggr/setup_aggr_samples/__init__.py", ns)
        chunk_outs = []
        for n in range(len(libs)):
            outs = Record(library_info=os.path.join(tmpdir, "li_%s_%d.pickle" % (normalization, n)))
            ns["main"](Record(aggr_csv=aggr_csv, normalization=normalization, reference_path="ref", n=n), outs)
            chunk_outs.append(outs)
        outs = Record(library_info=os.path.join(tmpdir, "li_%s.pickle" % normalization), gem_group_index=None,
                      gem_group_index_json=os.path.join(tmpdir, "ggi_%s.json" % normalization))
        ns["join"](Record(aggr_csv=aggr_csv, normalization=normalization, reference_path="ref"), outs, None, chunk_outs)
        info = fp.store[outs.library_info]
        info = {str(k): {kk: plain(vv) for kk, vv in v.items() if kk not in ("fragments", "cells")} for k, v in info.items()}
        ggi = {str(k): list(v) for k, v in outs.gem_group_index["gem_group_index"].items()}
        print("setup_aggr %-24s %s" % (normalization, {k: (v.get("original_threshold"), v.get("rate")) for k, v in info.items()}), flush=True)
        cases.append(dict(normalization=normalization, library_info=info, gem_group_index=ggi,
                          gem_group_index_json=json.load(open(outs.gem_group_index_json))))
    return dict(contigs=[list(c) for c in contigs], libraries=libs, cases=cases)


def make_matrix_cases(tmpdir):
    rng = np.random.default_rng(77)
    cases = []
    gv3 = dict(name="GV3",
               peaks="chr1\t100\t200\nchr1\t300\t400\nchr2\t50\t60\n",
               barcodes=["A", "B", "C", "D"],
               frags=[("chr1", 150, 350, "A", 1), ("chr1", 100, 200, "A", 3), ("chr1", 199, 300, "B", 1),
                      ("chr1", 120, 180, "A", 1), ("chr2", 10, 50, "C", 1), ("chr3", 5, 9, "D", 1),
                      ("chr2", 59, 60, "B", 2)])
    cases.append(gv3)
    touching = dict(name="touching_peaks",
                    peaks="chr1\t100\t200\nchr1\t200\t300\nchr2\t0\t10\n",
                    barcodes=["X-1", "Y-1"],
                    frags=[("chr1", 199, 200, "X-1", 1), ("chr1", 200, 299, "X-1", 1), ("chr1", 300, 301, "Y-1", 1),
                           ("chr2", 0, 9, "Y-1", 1), ("chr2", 9, 10, "Y-1", 1), ("chr1", 99, 100, "Y-1", 1)])
    cases.append(touching)
    for k in range(6):
        contigs = ["chr1", "chr2", "chrX"][: int(rng.integers(1, 4))]
        peak_rows = []
        for c in contigs:
            pos = 0
            for _ in range(int(rng.integers(1, 30))):
                pos += int(rng.integers(1, 800))
                ln = int(rng.integers(1, 900))
                peak_rows.append((c, pos, pos + ln))
                pos += ln + int(rng.integers(0, 3) == 0)   # sometimes touching the next one
        n_bc = int(rng.integers(1, 12))
        universe = ["".join(rng.choice(list("ACGT"), size=16)) + "-1" for _ in range(n_bc + 2)]
        chunk_bcs = universe[:n_bc]
        frags = []
        for _ in range(int(rng.integers(10, 500))):
            c = contigs[int(rng.integers(0, len(contigs)))]
            s = int(rng.integers(0, 30000))
            e = s + int(rng.integers(0, 1200))
            frags.append((c, s, e, universe[int(rng.integers(0, len(universe)))], int(rng.integers(1, 5))))
        frags.sort(key=lambda r: (contigs.index(r[0]), r[1], r[2]))
        cases.append(dict(name="rand%02d" % k, peaks="".join("%s\t%d\t%d\n" % r for r in peak_rows),
                          barcodes=chunk_bcs, frags=frags))
    out = []
    for case in cases:
        res = run_generate_peak_matrix_main(case["peaks"], case["barcodes"], case["frags"], tmpdir)
        out.append(dict(name=case["name"], peaks=case["peaks"], barcodes=case["barcodes"],
                        fragments=[list(r) for r in case["frags"]], result=res))
    return out


# --------------------------------------------------------------------------- FILTER_PEAK_MATRIX
def _bisect_left(a, x, keys):
    """cellranger/bisect.pyx:6-18 is Cython (not importable here): its 12 lines restated -- index into `keys` of x
    given the argsort `a` of keys."""
    lo, hi = 0, a.shape[0]
    while hi - lo > 1:
        i = (lo + hi) // 2
        y = keys[a[i]]
        if keys[a[i - 1]] < x and x <= y:
            return a[i]
        elif y < x:
            lo = i
        elif hi == i + 1:
            hi = i
        else:
            hi = i + 1
    return a[lo]


class _NumpyProxy(object):
    """numpy with the py2-era ``np.array(list, copy=False)`` meaning 'copy only if needed' (numpy 2 raises)."""
    def __getattr__(self, k):
        return getattr(np, k)

    @staticmethod
    def array(a, dtype=None, copy=True, **kw):
        return np.asarray(a, dtype=dtype) if not copy else np.array(a, dtype=dtype, **kw)


def run_filter_peak_matrix(n_features, bcs, indptr, indices, data, cell_barcodes, num_analysis_bcs, random_seed):
    """filter_peak_matrix/__init__.py join (:84-96) + prune (:46-70) on the reference's CountMatrix
    (cellranger/matrix.py: __init__ :187-202, bc_to_int :379-383, select_barcodes :658-662,
    filter_barcodes :874-883), scipy doing the column slicing as it does there."""
    import functools
    import scipy.sparse as sp_sparse
    ns = {"np": _NumpyProxy(), "sp_sparse": sp_sparse, "reduce": functools.reduce, "DEFAULT_DATA_DTYPE": "int32",
          "cr_bisect": types.SimpleNamespace(bisect_left=_bisect_left)}
    load_reference_module("lib/cellranger/lib/python/cellranger/matrix.py", ns, keep_funcs={"CountMatrix"})
    src = open(os.path.join(REF, "lib/cellranger/lib/python/cellranger/matrix.py")).read()
    assert ".itervalues()" in src                       # filter_barcodes: shimmed below like .iteritems()
    ns2 = {"np": np, "martian": types.SimpleNamespace(log_info=lambda *a: None)}
    load_reference_module("mro/atac/stages/processing/filter_peak_matrix/__init__.py", ns2, keep_funcs={"prune"})
    CountMatrix = ns["CountMatrix"]
    fdefs = [types.SimpleNamespace(id="f%d" % i, index=i) for i in range(n_features)]
    m = sp_sparse.csc_matrix((np.array(data, dtype=np.int32), np.array(indices, dtype=np.int64), np.array(indptr, dtype=np.int64)),
                             shape=(n_features, len(bcs)))
    matrix = CountMatrix(types.SimpleNamespace(feature_defs=fdefs, all_tag_keys=[]), [b.encode() for b in bcs], m)
    peak_matrix_bcs = set(matrix.bcs)
    present = {}
    for species in cell_barcodes:                        # the loop of filter_peak_matrix/__init__.py:84-91
        present[species] = set()
        for bc in cell_barcodes[species]:
            if bc.encode() in peak_matrix_bcs:
                present[species].add(bc.encode())

    class _Dict(dict):                                   # py2 dict.itervalues for matrix.py:882
        def itervalues(self):
            return self.values()
    matrix = matrix.filter_barcodes(_Dict(present))
    matrix = ns2["prune"](matrix, num_analysis_bcs=num_analysis_bcs, random_state=random_seed)
    out = matrix.m.tocsc()
    return {"bcs": [b.decode() for b in matrix.bcs.tolist()], "indptr": out.indptr.astype(np.int64).tolist(),
            "indices": out.indices.astype(np.int64).tolist(), "data": out.data.astype(np.int64).tolist()}


def make_filter_cases():
    cases = []
    rng = np.random.default_rng(77)
    specs = [("plain", 40, 60, 0.15, None, None, 0.0), ("subsample", 30, 80, 0.2, 25, 7, 0.0), ("subsample_more_than_present", 30, 50, 0.2, 500, 3, 0.0),
             ("explicit_zeros", 25, 70, 0.1, None, None, 0.5), ("seed_none_means_zero", 20, 90, 0.3, 10, None, 0.0),
             ("nothing_left", 10, 30, 0.0, None, None, 0.0), ("wide", 300, 400, 0.05, 120, 11, 0.1)]
    for name, n_feat, n_bc, dens, nab, seed, zero_frac in specs:
        cols = []
        indptr = [0]
        indices, data = [], []
        for j in range(n_bc):
            k = rng.binomial(n_feat, dens) if rng.random() > 0.2 else 0          # a fifth of the columns empty
            rows = np.sort(rng.choice(n_feat, size=k, replace=False))
            vals = rng.integers(1, 9, size=k)
            if zero_frac and k:
                vals[rng.random(k) < zero_frac] = 0                               # stored zeros: (m > 0) must ignore them
            indices += rows.tolist()
            data += vals.tolist()
            indptr.append(len(indices))
        letters = "ACGT"
        bcs = []
        seen = set()
        while len(bcs) < n_bc:
            b = "".join(letters[i] for i in rng.integers(0, 4, size=16)) + "-1"
            if b not in seen:
                seen.add(b)
                bcs.append(b)
        perm = rng.permutation(n_bc)
        half = [bcs[i] for i in perm[: n_bc * 2 // 3]]
        cell = {"GRCh38": half[: len(half) * 2 // 3] + ["TTTTTTTTTTTTTTTT-9"], "mm10": half[len(half) // 2:]}   # overlap + one absent barcode
        res = run_filter_peak_matrix(n_feat, bcs, indptr, indices, data, cell, nab, seed)
        cases.append({"name": name, "n_features": n_feat, "bcs": bcs, "indptr": indptr, "indices": indices, "data": data,
                      "cell_barcodes": cell, "num_analysis_bcs": nab, "random_seed": seed, "result": res})
    return cases


# --------------------------------------------------------------------------- REMOVE_LOW_TARGETING_BARCODES
def run_remove_low_targeting_main(contig, contig_len, peaks_text, frag_rows, tmpdir):
    """reference REMOVE_LOW_TARGETING_BARCODES.main for one contig chunk (cell_calling/remove_low_targeting_barcodes/
    __init__.py:178-243) with the reference's own Regions / fragment_overlaps_target (lib/python/tools/regions.py)."""
    import bisect as _bisect
    from collections import namedtuple
    fp = FakePickle()
    FakeRefMgr.contig_lengths = {contig: contig_len}
    reg_ns = {"bisect": _bisect, "namedtuple": namedtuple}
    load_reference_module("lib/python/tools/regions.py", reg_ns,
                          keep_funcs={"Region", "Regions", "get_target_regions_dict", "get_target_regions", "fragment_overlaps_target"})

    def parsed_fragments_from_contig(c, filename, index=None):
        for r in frag_rows:
            if r[0] == c:
                yield r

    ns = {"json": json, "pickle": fp, "Counter": Counter, "ReferenceManager": FakeRefMgr,
          "parsed_fragments_from_contig": parsed_fragments_from_contig,
          "get_target_regions": reg_ns["get_target_regions"], "fragment_overlaps_target": reg_ns["fragment_overlaps_target"]}
    load_reference_module("mro/atac/stages/processing/cell_calling/remove_low_targeting_barcodes/__init__.py", ns)
    pk = os.path.join(tmpdir, "peaks_rlt.bed")
    with open(pk, "w") as f:
        f.write(peaks_text)
    outs = Record(fragment_counts=os.path.join(tmpdir, "fc.pickle"), targeted_counts=os.path.join(tmpdir, "tc.pickle"),
                  fragment_lengths=os.path.join(tmpdir, "fl.pickle"), covered_bases=os.path.join(tmpdir, "cb.pickle"),
                  peak_coverage=None)
    args = Record(peaks=pk, fragments="frag.tsv.gz", fragments_index="frag.tsv.gz.tbi", reference_path="ref", contig=contig)
    ns["main"](args, outs)
    as_dict = lambda c: {k: int(v) for k, v in sorted(c.items())}
    return dict(fragment_counts=as_dict(fp.store[outs.fragment_counts]), targeted_counts=as_dict(fp.store[outs.targeted_counts]),
                fragment_lengths={str(p): as_dict(c) for p, c in fp.store[outs.fragment_lengths].items()},
                covered_bases={str(p): as_dict(c) for p, c in fp.store[outs.covered_bases].items()},
                peak_coverage=int(outs.peak_coverage))


def make_low_targeting_cases(tmpdir):
    cases = []

    def add(name, contig_len, peaks, frags):
        peaks_text = "".join("%s\t%d\t%d\n" % p for p in peaks)
        got = run_remove_low_targeting_main("chr1", contig_len, peaks_text, frags, tmpdir)
        cases.append(dict(name=name, contig="chr1", contig_len=contig_len, peaks=[list(p) for p in peaks],
                          fragments=[list(r) for r in frags], expect=got))

    # touching is overlap on the left only: Regions.overlaps_region bisects the ENDS with bisect_left
    add("touching_edges", 100000, [("chr1", 1000, 2000), ("chr1", 5000, 6000)],
        [("chr1", 500, 1000, "A-1", 1), ("chr1", 2000, 2100, "A-1", 1), ("chr1", 1999, 2001, "B-1", 1), ("chr1", 2001, 2100, "B-1", 2),
         ("chr1", 4000, 5000, "C-1", 1), ("chr1", 4000, 5001, "C-1", 1), ("chr1", 6000, 6001, "D-1", 1), ("chr1", 6001, 6002, "D-1", 1)])
    add("no_peaks_on_contig", 50000, [("chr2", 10, 20)], [("chr1", 10, 200, "A-1", 1), ("chr1", 40000, 49999, "B-1", 1)])
    add("no_fragments", 50000, [("chr1", 100, 200), ("chr1", 3000, 3100)], [])
    add("padding_clipped_at_contig_ends", 60000, [("chr1", 100, 900), ("chr1", 59000, 59900)],
        [("chr1", 0, 120, "A-1", 1), ("chr1", 30, 400, "A-1", 1), ("chr1", 59950, 60000, "A-1", 1), ("chr1", 20000, 20300, "B-1", 1)])
    add("overlapping_and_nested_peaks", 300000, [("chr1", 1000, 9000), ("chr1", 2000, 3000), ("chr1", 8500, 12000), ("chr1", 100000, 100500),
                                                   ("chr1", 103000, 104000)],
        [("chr1", 3100, 3200, "A-1", 1), ("chr1", 9500, 9600, "A-1", 1), ("chr1", 12000, 12100, "B-1", 1), ("chr1", 50000, 50400, "B-1", 1),
         ("chr1", 101000, 101200, "B-1", 1), ("chr1", 160000, 160100, "A-1", 1)])
    rng = np.random.default_rng(77)
    for k, (L, n_peaks, n_frag, n_bc) in enumerate([(400000, 40, 600, 12), (3000000, 300, 4000, 60), (150000, 3, 900, 5)]):
        ps = np.sort(rng.integers(0, L - 3000, size=n_peaks))
        peaks, last = [], -1
        for s in ps.tolist():
            s = max(s, last + 1)
            e = min(L, s + int(rng.integers(150, 2500)))
            if e <= s:
                continue
            peaks.append(("chr1", s, e))
            last = e + int(rng.integers(0, 3)) - 1          # touching and 1-base gaps do occur
        st = np.sort(rng.integers(0, L - 1300, size=n_frag))
        ln = rng.integers(20, 1200, size=n_frag)
        bcs = ["".join(rng.choice(list("ACGT"), size=16)) + "-1" for _ in range(n_bc)]
        frags = [("chr1", int(a), int(min(L, a + b)), bcs[int(rng.integers(0, n_bc))], int(rng.integers(1, 4))) for a, b in zip(st, ln)]
        # a few fragments pinned to peak boundaries
        for j in range(0, len(peaks), max(1, len(peaks) // 6)):
            _, s, e = peaks[j]
            frags.append(("chr1", e, min(L, e + 50), bcs[0], 1))
            frags.append(("chr1", max(0, s - 50), s, bcs[1 % n_bc], 1))
        frags.sort(key=lambda r: (r[1], r[2]))
        add("random_%d" % k, L, peaks, frags)
    return cases


# --------------------------------------------------------------------------- get_counts_by_barcode
TRACKS = ("tss", "ctcf", "dnase", "enhancer", "promoter", "blacklist")


def run_get_counts_by_barcode(tracks, peaks_text, frag_rows, species_list, species_of_contig, known_cells_text, contig, tmpdir):
    """reference tools/io.py:429-555 with the reference's own Region / Regions (lib/python/tools/regions.py)."""
    import bisect as _bisect
    from collections import namedtuple
    reg_ns = {"bisect": _bisect, "namedtuple": namedtuple}
    load_reference_module("lib/python/tools/regions.py", reg_ns,
                          keep_funcs={"Region", "Regions", "get_target_regions_dict", "get_target_regions"})
    paths = {}
    for name in TRACKS:
        paths[name] = None
        if tracks.get(name) is not None:
            paths[name] = os.path.join(tmpdir, "track_%s.bed" % name)
            with open(paths[name], "w") as f:
                f.write(tracks[name])
    pk = None
    if peaks_text is not None:
        pk = os.path.join(tmpdir, "peaks_gcb.bed")
        with open(pk, "w") as f:
            f.write(peaks_text)
    kc = None
    if known_cells_text is not None:
        kc = os.path.join(tmpdir, "cells.csv")
        with open(kc, "w") as f:
            f.write(known_cells_text)

    class RefMgr(object):
        def __init__(self, path):
            self.tss_track, self.ctcf_track, self.dnase_track = paths["tss"], paths["ctcf"], paths["dnase"]
            self.enhancer_track, self.promoter_track, self.blacklist_track = paths["enhancer"], paths["promoter"], paths["blacklist"]

        def list_species(self):
            return list(species_list)

        def species_from_contig(self, c):
            return species_of_contig[c]

    def open_fragment_file(filename):
        for r in frag_rows:
            yield r

    def parsed_fragments_from_contig(c, filename, index=None):
        for r in frag_rows:
            if r[0] == c:
                yield r

    ns = {"Counter": Counter, "ReferenceManager": RefMgr, "regtools": types.SimpleNamespace(get_target_regions=reg_ns["get_target_regions"]),
          "open_fragment_file": open_fragment_file, "parsed_fragments_from_contig": parsed_fragments_from_contig}
    load_reference_module("lib/python/tools/io.py", ns, keep_funcs={"get_counts_by_barcode"})
    read_count, counts, tss_relpos, ctcf_relpos = ns["get_counts_by_barcode"]("ref", pk, "frag.tsv.gz", fragments_index="x.tbi",
                                                                              contig=contig, known_cells=kc)
    # "cell_id" numbers the cells in the iteration order of a CPython-2.7 dict (io.py:482-488); a Python-3 run cannot
    # reproduce that order, so the label is left out of the vectors (everything else does not depend on it)
    return dict(read_count=int(read_count),
                counts_by_barcode={bc: {k: int(v) for k, v in sorted(c.items()) if k != "cell_id"} for bc, c in sorted(counts.items())},
                tss_relpos={str(k): int(v) for k, v in sorted(tss_relpos.items())},
                ctcf_relpos={str(k): int(v) for k, v in sorted(ctcf_relpos.items())})


def make_counts_by_barcode_cases(tmpdir):
    cases = []

    def bed(rows):
        return "".join("\t".join(str(x) for x in r) + "\n" for r in rows)

    def add(name, tracks, peaks, frags, species_list=("GRCh38",), species_of_contig=None, known_cells=None, contig=None):
        species_of_contig = species_of_contig or {"chr1": "GRCh38", "chr2": "GRCh38"}
        tr = {k: (bed(v) if v is not None else None) for k, v in tracks.items()}
        pk = bed(peaks) if peaks is not None else None
        got = run_get_counts_by_barcode(tr, pk, frags, species_list, species_of_contig, known_cells, contig, tmpdir)
        cases.append(dict(name=name, tracks=tr, peaks=pk, fragments=[list(r) for r in frags], species_list=list(species_list),
                          species_of_contig=species_of_contig, known_cells=known_cells, contig=contig, expect=got))

    tss = [("chr1", 10000, 10001, "GENE1", ".", "+"), ("chr1", 30000, 30001, "GENE2", ".", "-"), ("chr2", 5000, 5001, "GENE3", ".", "+")]
    ctcf = [("chr1", 12000, 12020, "CTCF1", ".", "+"), ("chr1", 40000, 40019, "CTCF2", ".", "-")]
    small = dict(tss=tss, ctcf=ctcf, dnase=[("chr1", 100, 500), ("chr1", 20000, 21000)], enhancer=[("chr1", 400, 450)],
                 promoter=[("chr1", 9000, 10500), ("chr2", 4000, 5200)], blacklist=[("chr1", 50000, 60000)])
    frags = [("chr1", 90, 100, "A-1", 1), ("chr1", 100, 130, "A-1", 3), ("chr1", 500, 700, "A-1", 1), ("chr1", 8001, 8100, "B-1", 2),
             ("chr1", 9999, 10001, "B-1", 1), ("chr1", 11749, 11751, "B-1", 1), ("chr1", 12269, 12271, "C-1", 5), ("chr1", 28000, 32001, "C-1", 1),
             ("chr1", 39750, 40269, "A-1", 1), ("chr1", 55000, 55100, "C-1", 1), ("chr2", 3000, 3001, "A-1", 1), ("chr2", 6999, 7001, "D-1", 2)]
    peaks = [("chr1", 95, 120), ("chr1", 9990, 10010), ("chr1", 39000, 41000), ("chr2", 2990, 3001)]
    add("all_tracks_single_species", small, peaks, frags)
    add("one_contig_only", small, peaks, frags, contig="chr1")
    add("no_tracks_no_peaks", {k: None for k in TRACKS}, None, frags)
    add("known_cells", small, peaks, frags, known_cells="GRCh38,A-1,C-1,null\n")
    two = {"chr1": "hg19", "chr2": "mm10"}
    add("two_species_known_cells", small, peaks, frags, species_list=("hg19", "mm10"), species_of_contig=two,
        known_cells="hg19,A-1,B-1\nmm10,A-1,D-1,null\n")
    rng = np.random.default_rng(91)
    L = {"chr1": 600000, "chr2": 300000}

    def rand_rows(n, wmin, wmax, named):
        rows = set()
        out = []
        for _ in range(n):
            c = "chr1" if rng.random() < 0.66 else "chr2"
            s0 = int(rng.integers(3000, L[c] - wmax - 3000))
            e0 = s0 + int(rng.integers(wmin, wmax + 1))
            if (c, s0, e0) in rows:
                continue
            rows.add((c, s0, e0))
            out.append((c, s0, e0, "n%d" % len(out), ".", "+-"[int(rng.integers(0, 2))]) if named else (c, s0, e0))
        return out

    tracks = dict(tss=rand_rows(60, 1, 1, True), ctcf=rand_rows(80, 19, 20, True), dnase=rand_rows(150, 150, 2000, False),
                  enhancer=rand_rows(100, 200, 3000, False), promoter=rand_rows(60, 500, 2500, False), blacklist=rand_rows(6, 1000, 20000, False))
    peaks, last = [], {"chr1": -1, "chr2": -1}
    for c, s0, e0 in sorted(rand_rows(250, 150, 2500, False)):
        s0 = max(s0, last[c] + 1)
        if e0 > s0:
            peaks.append((c, s0, e0))
            last[c] = e0 - 1 + int(rng.integers(0, 2))
    bcs = ["".join(rng.choice(list("ACGT"), size=16)) + "-1" for _ in range(40)]
    fr = []
    for _ in range(5000):
        c = "chr1" if rng.random() < 0.66 else "chr2"
        a = int(rng.integers(0, L[c] - 1300))
        fr.append((c, a, a + int(rng.integers(20, 1200)), bcs[int(rng.integers(0, len(bcs)))], int(rng.integers(1, 5))))
    fr.sort(key=lambda r: (r[0], r[1], r[2]))
    add("random_tracks", tracks, peaks, fr, known_cells="GRCh38," + ",".join(bcs[:15]) + "\n")
    return cases


# --------------------------------------------------------------------------- cell-caller mixture fit
def reference_cell_caller():
    """The two-negative-binomial fit of DETECT_CELL_BARCODES (cell_calling/detect_cell_barcodes/__init__.py:84-201),
    exec'd with the reference's own analysis.em_fitting and scipy.stats."""
    from collections import namedtuple
    from scipy import stats
    sys.path.insert(0, os.path.join(REF, "lib/python"))
    from analysis.em_fitting import weighted_mean, MLE_dispersion
    ns = {"np": np, "stats": stats, "weighted_mean": weighted_mean, "MLE_dispersion": MLE_dispersion,
          "MixedModelParams": namedtuple("MixedModelParams", ["mu_noise", "alpha_noise", "mu_signal", "alpha_signal", "frac_noise"]),
          "print": lambda *a, **k: None}
    load_reference_module("mro/atac/stages/processing/cell_calling/detect_cell_barcodes/__init__.py", ns,
                          keep_funcs={"calculate_probability_distribution", "weighted_parameter_estimates", "estimate_threshold",
                                      "simple_threshold_estimate", "normalize_parameters", "estimate_parameters"})
    return ns


def make_cell_em_cases():
    ns = reference_cell_caller()
    cases = []
    rng = np.random.default_rng(123)

    def nb(mu, alpha, n):
        r = 1.0 / alpha
        return rng.negative_binomial(r, r / (r + mu), size=n)

    shapes = [("pbmc_like", (40, 2.0, 60000), (9000, 0.25, 1200)), ("low_depth", (8, 1.5, 30000), (700, 0.4, 400)),
              ("few_cells", (25, 3.0, 80000), (15000, 0.15, 60)), ("overlapping", (120, 1.0, 20000), (900, 0.8, 3000))]
    for name, noise, signal in shapes:
        counts = np.concatenate((nb(*noise), nb(*signal)))
        counts = counts[counts > 0]
        cd = Counter(int(c) for c in counts)
        t0 = time.time()
        simple = ns["simple_threshold_estimate"](cd)
        params = ns["estimate_parameters"](cd)
        thr = ns["estimate_threshold"](params, 20)
        cases.append(dict(name=name, values=sorted(cd), weights=[cd[k] for k in sorted(cd)], simple_threshold=int(simple),
                          params=[float(x) for x in params], threshold=int(thr), seconds=round(time.time() - t0, 2)))
    return cases


# --------------------------------------------------------------------------- REPORT_INSERT_SIZES.main (another scan of the fragments)
def run_report_insert_sizes_main(barcodes, frag_rows, primary, exclude_non_nuclear, tmpdir):
    """reference REPORT_INSERT_SIZES.main for one barcode chunk (stages/metrics/report_insert_sizes/__init__.py:74-113)."""
    FakeRefMgr.primary = list(primary)

    def open_fragment_file(filename):
        for r in frag_rows:
            yield r

    ns = {"np": np, "Counter": Counter, "OrderedDict": OrderedDict, "ReferenceManager": FakeRefMgr, "open_fragment_file": open_fragment_file,
          "json": json, "martian": None, "utils": None, "combine_csv": None, "signal": None, "stats": None}
    src_path = "mro/atac/stages/metrics/report_insert_sizes/__init__.py"
    load_reference_module(src_path, ns, keep_funcs={"main"})
    ns["MAX_INSERT_SIZE"] = 1000
    ns["GT_MAX_INSERT_SIZE"] = ">1000"
    bcf = os.path.join(tmpdir, "barcodes_ris.txt")
    with open(bcf, "w") as f:
        f.write("\n".join(barcodes))
    args = Record(barcodes=bcf, fragments="frag.tsv.gz", exclude_non_nuclear=exclude_non_nuclear, reference_path="ref")
    outs = Record(insert_sizes=os.path.join(tmpdir, "is.csv"), insert_summary=None, total=os.path.join(tmpdir, "total.csv"))
    src = open(os.path.join(REF, src_path)).read()
    assert ".iterkeys()" in src           # one more py2 token on this path, mapped below
    ns["OrderedDict"] = type("OrderedDictPy2", (OrderedDict,), {"iterkeys": lambda self: iter(self.keys())})
    ns["main"](args, outs)
    if outs.insert_sizes is None:
        return None
    rows = open(outs.insert_sizes).read().split("\n")
    table = {}
    for line in rows[1:]:
        if line:
            f = line.split(",")
            nz = {str(i + 1): int(v) for i, v in enumerate(f[1:1001]) if int(v)}
            if int(f[1001]):
                nz[">1000"] = int(f[1001])
            table[f[0]] = nz
    return dict(header_ok=rows[0] == ",".join(["Barcode"] + [str(n) for n in range(1, 1001)] + [">1000"]),
                row_order=[line.split(",")[0] for line in rows[1:] if line], table=table,
                total=[int(x) for x in np.genfromtxt(outs.total, "float64").tolist()])


def make_insert_size_cases(tmpdir):
    rng = np.random.default_rng(55)
    cases = []
    bcs = ["".join(rng.choice(list("ACGT"), size=16)) + "-1" for _ in range(25)]
    frags = []
    for _ in range(4000):
        c = ["chr1", "chr2", "chrM"][int(rng.integers(0, 3))]
        a = int(rng.integers(0, 100000))
        ln = int(rng.choice([0, 1, 37, 147, 999, 1000, 1001, 1500])) if rng.random() < 0.1 else int(rng.integers(1, 1200))
        frags.append((c, a, a + ln, bcs[int(rng.integers(0, len(bcs)))], 1))
    frags.sort()
    for name, chunk, excl in (("all_contigs", bcs[:18], False), ("nuclear_only", bcs[:18], True), ("absent_barcodes", bcs[20:] + ["TTTTTTTTTTTTTTTT-1"], True)):
        got = run_report_insert_sizes_main(chunk, frags, ["chr1", "chr2"], excl, tmpdir)
        cases.append(dict(name=name, barcodes=chunk, primary=["chr1", "chr2"], exclude_non_nuclear=excl, expect=got))
    return dict(fragments=[list(r) for r in frags], cases=cases)       # the three cases share the fragments


def main():
    only = set(sys.argv[1:])
    with tempfile.TemporaryDirectory() as tmpdir:
        if not only or "cut_sites" in only:
            with open(os.path.join(HERE, "cut_sites.json"), "w") as f:
                json.dump(make_cut_site_cases(tmpdir), f, separators=(",", ":"))
        if not only or "matrix" in only:
            with open(os.path.join(HERE, "matrix.json"), "w") as f:
                json.dump(make_matrix_cases(tmpdir), f, separators=(",", ":"))
        if not only or "filter" in only:
            with open(os.path.join(HERE, "filter.json"), "w") as f:
                json.dump(make_filter_cases(), f, separators=(",", ":"))
        if not only or "low_targeting" in only:
            with open(os.path.join(HERE, "low_targeting.json"), "w") as f:
                json.dump(make_low_targeting_cases(tmpdir), f, separators=(",", ":"))
        if not only or "counts_by_barcode" in only:
            with open(os.path.join(HERE, "counts_by_barcode.json"), "w") as f:
                json.dump(make_counts_by_barcode_cases(tmpdir), f, separators=(",", ":"))
        if not only or "cell_em" in only:
            with open(os.path.join(HERE, "cell_em.json"), "w") as f:
                json.dump(make_cell_em_cases(), f, separators=(",", ":"))
        if not only or "insert_sizes" in only:
            with open(os.path.join(HERE, "insert_sizes.json"), "w") as f:
                json.dump(make_insert_size_cases(tmpdir), f, separators=(",", ":"))
        if not only or "setup_aggr" in only:
            with open(os.path.join(HERE, "setup_aggr.json"), "w") as f:
                json.dump(make_setup_aggr_cases(tmpdir), f, separators=(",", ":"))
        if not only or "spec_trajectories" in only:
            with open(os.path.join(HERE, "spec_trajectories.json"), "w") as f:
                json.dump(make_spec_trajectories(), f, separators=(",", ":"))
        if not only or "em_random" in only:
            with open(os.path.join(HERE, "em_random.json"), "w") as f:
                json.dump(make_em_random_cases(), f, separators=(",", ":"))
        if not only or "em" in only:
            with open(os.path.join(HERE, "em.json"), "w") as f:
                json.dump(make_em_cases(), f, separators=(",", ":"))


if __name__ == "__main__":
    main()
