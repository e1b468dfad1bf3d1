"""Literal Python-3 restatement of the reference stage code (TEST INFRASTRUCTURE ONLY).

Every function names the reference lines it follows (paths relative to the
cellranger-atac 1.2.0 checkout).  The per-base / per-fragment Python loops are
kept on purpose: this file is both the bit-exactness checker for small inputs
and the timed "reference-shaped" CPU baseline.  Vectorised equivalents that
finish on large inputs live in ``oracle/fast.py`` and ``oracle/fast.c`` and are
themselves checked against this file.

Python-2 semantics that had to be written out explicitly:
  * ``None >= 1`` is False in py2 (count_cut_sites/__init__.py:42,47)
  * ``str < float`` is always False in py2 (utils.py:112) -> no filtering
  * ``from __future__ import division`` is on in all three stages
"""
import bisect
from collections import Counter, OrderedDict

import numpy as np

WINDOW_SIZE = 400            # lib/python/constants.py:68
PEAK_MERGE_DISTANCE = 500    # lib/python/constants.py:70
PEAK_ODDS_RATIO = 1 / 5      # lib/python/constants.py:71


# --------------------------------------------------------------------------- #
# fragments text                                                              #
# --------------------------------------------------------------------------- #
def parse_fragments(lines):
    """lib/python/tools/io.py:75-79 (open_fragment_file): yield
    (contig, start, stop, barcode, dup_count) from text lines."""
    for line in lines:
        if isinstance(line, bytes):
            line = line.decode("ascii")
        contig, start, stop, barcode, dup_count = line.strip().split("\t")
        yield contig, int(start), int(stop), barcode, int(dup_count)


# --------------------------------------------------------------------------- #
# COUNT_CUT_SITES                                                             #
# --------------------------------------------------------------------------- #
def pileup(chrom_len, fragments):
    """count_cut_sites/__init__.py:94-101: windowed cut-site counts for ONE contig.
    ``fragments`` yields (start, stop)."""
    half_window = WINDOW_SIZE // 2
    Cuts = np.zeros(chrom_len, dtype="int32")
    for start, stop in fragments:
        Cuts[max(0, start - half_window): min(start + half_window + 1, chrom_len)] += 1
        Cuts[max(0, stop - half_window): min(stop + half_window + 1, chrom_len)] += 1
    return Cuts


def count_dict_from_cuts(Cuts):
    """count_cut_sites/__init__.py:104: Counter(v for v in Cuts if v > 0)."""
    return Counter(int(v) for v in Cuts if v > 0)


def bedgraph_rows(chrom, contig_len, Cuts, min_counts=1):
    """count_cut_sites/__init__.py:27-48 (write_chrom_bedgraph) returning the rows
    instead of writing them.  Zeros are skipped WITHOUT closing the open run
    (:36-37), so equal non-zero values either side of a zero gap form one row."""
    rows = []
    last_count = None
    last_start = None
    last_pos = None
    for pos in range(contig_len):
        count = int(Cuts[pos])
        if count == 0:
            continue
        if last_count == count:
            last_pos = pos
        else:
            if last_count is not None and last_count >= min_counts:   # py2: None >= 1 is False
                rows.append((chrom, last_start, last_pos + 1, last_count))
            last_start = pos
            last_pos = pos
            last_count = count
    if last_count is not None and last_count >= min_counts:
        rows.append((chrom, last_start, last_pos + 1, last_count))
    return rows


def format_bedgraph(rows):
    """count_cut_sites/__init__.py:43: '{}\\t{}\\t{}\\t{}\\n'."""
    return "".join("{}\t{}\t{}\t{}\n".format(*r) for r in rows)


def merge_count_dicts(dicts):
    """count_cut_sites/__init__.py:77-81: Counter += per chunk."""
    total = Counter()
    for d in dicts:
        total += d
    return total


# --------------------------------------------------------------------------- #
# DETECT_PEAKS                                                                #
# --------------------------------------------------------------------------- #
def cut_site_counter_from_bedgraph(rows, threshold=1, contig=None):
    """lib/python/utils.py:105-115 on already-split rows.  The reference compares
    the *string* count with the float threshold (:112); in Python 2 a number
    orders before every str, so the test is always False and nothing is filtered."""
    for chrom, start, stop, count in rows:
        if contig is not None and chrom != contig:
            continue
        for pos in range(int(start), int(stop)):
            yield chrom, pos, int(count)


def identify_peaks(site_iter, threshold):
    """detect_peaks/__init__.py:40-62 returning (peaks, total_peaks_detected)."""
    peaks = []
    window = None
    for chrom, pos, windowed_count in site_iter:
        if windowed_count < threshold:
            continue
        if window is None:
            window = [chrom, pos, pos + 1]
        elif pos - PEAK_MERGE_DISTANCE <= window[2] and chrom == window[0]:
            window[2] = pos + 1
        else:
            peaks.append(tuple(window))
            window = [chrom, pos, pos + 1]
    if window is not None:
        peaks.append(tuple(window))
    return peaks, len(peaks)


def sort_and_uniq_peaks(peaks, fai_contig_order):
    """lib/python/tools/regions.py:392-417: ``bedtools sort -g genome.fa.fai`` then
    ``uniq``.  bedtools sort -g orders by the .fai contig order, then start (then
    end for ties, which cannot occur for non-overlapping peaks)."""
    rank = {c: i for i, c in enumerate(fai_contig_order)}
    out = sorted(peaks, key=lambda p: (rank[p[0]], p[1], p[2]))
    uniq = []
    for p in out:
        if not uniq or uniq[-1] != p:
            uniq.append(p)
    return uniq


def detect_peaks_contig(chrom, chrom_len, fragments, threshold):
    """COUNT_CUT_SITES.main + DETECT_PEAKS.main for one contig through the
    bedgraph, exactly as the two stages chain (count_cut_sites/__init__.py:85-112,
    detect_peaks/__init__.py:85-97)."""
    Cuts = pileup(chrom_len, fragments)
    cd = count_dict_from_cuts(Cuts)
    rows = bedgraph_rows(chrom, chrom_len, Cuts) if len(cd) else []
    text_rows = [(c, str(s), str(e), str(v)) for c, s, e, v in rows]
    peaks, _ = identify_peaks(cut_site_counter_from_bedgraph(text_rows, threshold, chrom), threshold)
    return Cuts, cd, rows, peaks


def peak_metrics(peaks, threshold, params):
    """detect_peaks/__init__.py:113-129 (join): metrics dict.  bases_in_peaks is
    sum(len(fasta[chrom][start:end])) = sum(end-start) because end <= contig_len."""
    m = OrderedDict()
    m["total_peaks_detected"] = len(peaks)
    m["bases_in_peaks"] = int(sum(e - s for _, s, e in peaks))
    if params is not None:
        for name, val in zip(("p_zero_noise", "mean_bg_noise", "alpha_bg_noise", "p_signal",
                              "frac_zero_noise", "frac_bg_noise"), params):
            m["peakcaller_{}".format(name)] = float(val)
    m["peakcaller_threshold"] = threshold
    return m


# --------------------------------------------------------------------------- #
# tenkit Regions (only what GENERATE_PEAK_MATRIX touches)                     #
# --------------------------------------------------------------------------- #
class Regions(object):
    """tenkit/lib/python/tenkit/regions.py:28-43,91-110,136-145."""

    def __init__(self, regions=None, merge=True):
        if regions is not None and not (regions == []):
            sorted_regions = sorted(regions)
            self.starts = [r[0] for r in sorted_regions]
            self.ends = [r[1] for r in sorted_regions]
            if merge:
                self.make_regions_non_overlapping()
        else:
            self.starts = []
            self.ends = []

    def make_regions_non_overlapping(self):
        new_starts = [self.starts[0]]
        new_ends = [self.ends[0]]
        last_index = 0
        for n in range(1, len(self.starts)):
            this_start = self.starts[n]
            this_end = self.ends[n]
            last_end = new_ends[last_index]
            if this_start >= last_end:
                new_starts.append(this_start)
                new_ends.append(this_end)
                last_index += 1
            else:
                new_ends[last_index] = max(this_end, last_end)
        self.starts = new_starts
        self.ends = new_ends

    def get_region_containing_point(self, pt):
        check_index = bisect.bisect(self.starts, pt) - 1
        if check_index == -1:
            return None
        if pt >= self.starts[check_index] and pt < self.ends[check_index]:
            return (self.starts[check_index], self.ends[check_index])
        return None


def get_target_regions(peak_lines):
    """tenkit/lib/python/tenkit/bio_io.py:736-766."""
    targets = {}
    for line in peak_lines:
        info = line.strip().split("\t")
        if (line.startswith("browser") or line.startswith("track") or line.startswith("-browser")
                or line.startswith("-track") or line.startswith("#")):
            continue
        if len(line.strip()) == 0:
            continue
        targets.setdefault(info[0], []).append((int(info[1]), int(info[2])))
    return {chrom: Regions(regions=se) for chrom, se in targets.items()}


# --------------------------------------------------------------------------- #
# GENERATE_PEAK_MATRIX                                                        #
# --------------------------------------------------------------------------- #
def get_chunks(num, chunks):
    """lib/python/utils.py:133-136."""
    bounds = np.linspace(0, num, chunks + 1, dtype=int, endpoint=True).tolist()
    return [(s, e) for s, e in zip(bounds, bounds[1:]) if s != e]


def peak_matrix_chunk(peak_lines, barcodes, fragments):
    """generate_peak_matrix/__init__.py:92-119 for one barcode chunk.

    peak_lines: lines of peaks.bed; barcodes: the chunk's barcode list (column
    order); fragments: iterable of (contig, start, stop, barcode, dup).
    Returns canonical CSC (indptr int64[B+1], indices int64[nnz], data int64[nnz],
    shape) exactly as ``csc_matrix(coo_matrix(...))`` yields it: row indices
    sorted within each column, duplicates summed.  A merged region (overlapping
    user peaks) raises KeyError like the reference (:114)."""
    peak_lines = list(peak_lines)
    full_peaks = get_target_regions(peak_lines)
    peaks_dict = OrderedDict(("{}:{}-{}".format(*peak.strip("\n").split("\t")), num)
                             for num, peak in enumerate(peak_lines))
    barcodes_dict = OrderedDict((bc.strip("\n"), num) for num, bc in enumerate(barcodes))
    peak_bc_counts = Counter()
    for contig, start, stop, barcode, _ in fragments:
        if barcode not in barcodes_dict:
            continue
        for pos in (start, stop):
            if contig in full_peaks.keys():
                peak = full_peaks[contig].get_region_containing_point(pos)
                if peak is not None:
                    peak_bc_counts[barcodes_dict[barcode],
                                   peaks_dict["{}:{}-{}".format(contig, peak[0], peak[1])]] += 1
    n_rows, n_cols = len(peaks_dict), len(barcodes_dict)
    return coo_to_csc(peak_bc_counts, n_rows, n_cols)


def coo_to_csc(counter, n_rows, n_cols):
    """scipy ``csc_matrix(coo_matrix((data,(row,col))))`` (generate_peak_matrix/
    __init__.py:119): canonical CSC.  Restated without scipy so the layout is
    explicit; tests compare it with scipy itself."""
    items = sorted(((col, row, val) for (col, row), val in counter.items()))
    indptr = np.zeros(n_cols + 1, dtype=np.int64)
    indices = np.zeros(len(items), dtype=np.int64)
    data = np.zeros(len(items), dtype=np.int64)
    for k, (col, row, val) in enumerate(items):
        indptr[col + 1] += 1
        indices[k] = row
        data[k] = val
    indptr = np.cumsum(indptr)
    return indptr, indices, data, (n_rows, n_cols)


def hstack_csc(chunks):
    """generate_peak_matrix/__init__.py:62-72: scipy hstack of the chunk matrices."""
    indptr = [np.zeros(1, dtype=np.int64)]
    indices, data = [], []
    off = 0
    n_rows = None
    for ip, ix, d, shape in chunks:
        n_rows = shape[0]
        indptr.append(ip[1:] + off)
        off += ip[-1]
        indices.append(ix)
        data.append(d)
    indptr = np.concatenate(indptr)
    return (indptr, np.concatenate(indices) if indices else np.zeros(0, np.int64),
            np.concatenate(data) if data else np.zeros(0, np.int64), (n_rows, len(indptr) - 1))


def feature_names(peak_lines):
    """cellranger/atac/feature_ref.py:58-68: id = name = 'chr:start-end'."""
    return ["{}:{}-{}".format(*line.strip("\n").split("\t")) for line in peak_lines]


# ---------------------------------------------------------------------------------------------- FILTER_PEAK_MATRIX
def filter_peak_matrix(n_features, bcs, indptr, indices, data, cell_barcodes, num_analysis_bcs=None, random_seed=None):
    """filter_peak_matrix/__init__.py:84-96 (cell barcodes present in the matrix, over all species) ->
    cellranger/matrix.py:874-883 (sorted union, m[:, indices] :658-662) -> prune (:46-70: optional random subset with
    numpy's legacy generator, then the columns with at least one positive entry).
    -> (bcs, indptr, indices, data) of the filtered CSC.  Plain loops; test infrastructure only."""
    import numpy as np
    index = {}
    for j, b in enumerate(bcs):
        index.setdefault(b, j)
    present = set()
    for species in cell_barcodes:
        for bc in cell_barcodes[species]:
            if bc in index:
                present.add(bc)
    sel = sorted(present)
    cols = [index[b] for b in sel]
    if num_analysis_bcs:
        np.random.seed(0 if random_seed is None else random_seed)
        n = len(cols)
        pick = np.sort(np.random.choice(np.arange(n), size=min(num_analysis_bcs, n), replace=False))
        cols = [cols[i] for i in pick]
        sel = [sel[i] for i in pick]
    out_bcs, out_indptr, out_indices, out_data = [], [0], [], []
    for b, c in zip(sel, cols):
        lo, hi = indptr[c], indptr[c + 1]
        if not any(v > 0 for v in data[lo:hi]):
            continue
        out_bcs.append(b)
        out_indices.extend(indices[lo:hi])
        out_data.extend(data[lo:hi])
        out_indptr.append(len(out_indices))
    return out_bcs, out_indptr, out_indices, out_data


# --------------------------------------------------------------------------- #
# REMOVE_LOW_TARGETING_BARCODES (oracle ahead of the kernel) #
# --------------------------------------------------------------------------- #
LOW_TARGETING_DISTANCE = 2000              # cell_calling/remove_low_targeting_barcodes/__init__.py:40
LOW_TARGETING_PADDINGS = (0, 50000)        # :41


def tools_regions(rows):
    """lib/python/tools/regions.py:39-53 (NOT tenkit's Regions: no merging of overlapping rows).  rows = (start, end)
    pairs of one contig -> (starts, ends) in the order of the sorted (start, end) pairs.  `ends` is NOT sorted when rows
    nest, and overlaps_region bisects it all the same -- kept."""
    srt = sorted(rows)
    return [r[0] for r in srt], [r[1] for r in srt]


def overlaps_region(starts, ends, start, end):
    """lib/python/tools/regions.py:83-93: the first row whose end is >= start (bisect_left: a row that ENDS where the
    fragment starts is taken) overlaps iff the fragment ends beyond that row's start."""
    k = bisect.bisect_left(ends, start)
    if k == len(starts):
        return False
    return end > starts[k]


def count_covered_bases(starts, stops, contig_len, padding):
    """cell_calling/remove_low_targeting_barcodes/__init__.py:44-73: bases covered by the union of the padded intervals,
    taken in the order given (a later interval joins the running one iff it starts before the running stop)."""
    occupancy = 0
    run_start = run_stop = None
    for start, stop in zip(starts, stops):
        start = max(0, start - padding)
        stop = min(contig_len, stop + padding)
        if run_start is None:
            run_start, run_stop = start, stop
        elif start < run_stop:
            run_stop = max(stop, run_stop)
        else:
            occupancy += run_stop - run_start
            run_start, run_stop = start, stop
    if run_start is not None:
        occupancy += run_stop - run_start
    return occupancy


def remove_low_targeting_contig(contig, contig_len, peak_rows, fragments):
    """REMOVE_LOW_TARGETING_BARCODES.main for one contig (cell_calling/remove_low_targeting_barcodes/__init__.py:178-243).
    peak_rows = (contig, start, end) of the whole peaks file; fragments = (contig, start, stop, barcode, dup) rows of the
    file in file order (only those of `contig` are used, as tabix fetch would return them).
    -> dict(fragment_counts, targeted_counts, fragment_lengths{padding}, covered_bases{padding}, peak_coverage)."""
    mine = [(s, e) for c, s, e in peak_rows if c == contig]
    starts, ends = tools_regions(mine) if mine else ([], [])
    fragment_counts, targeted_counts = Counter(), Counter()
    lengths = {p: Counter() for p in LOW_TARGETING_PADDINGS}
    running = {p: {} for p in LOW_TARGETING_PADDINGS}          # barcode -> [covered so far, run start, run stop]
    for c, start, stop, barcode, _dup in fragments:
        if c != contig:
            continue
        fragment_counts[barcode] += 1
        if mine and overlaps_region(starts, ends, start, stop):
            targeted_counts[barcode] += 1
        for p in LOW_TARGETING_PADDINGS:
            a, b = max(0, start - p), min(contig_len, stop + p)
            lengths[p][barcode] += b - a
            st = running[p].setdefault(barcode, [0, None, None])
            if st[1] is None:
                st[1], st[2] = a, b
            elif a < st[2]:
                st[2] = max(b, st[2])
            else:
                st[0] += st[2] - st[1]
                st[1], st[2] = a, b
    covered = {p: Counter({bc: st[0] + (st[2] - st[1] if st[1] is not None else 0) for bc, st in running[p].items()})
               for p in LOW_TARGETING_PADDINGS}
    peak_coverage = count_covered_bases(starts, ends, contig_len, LOW_TARGETING_DISTANCE) if mine else 0
    return dict(fragment_counts=fragment_counts, targeted_counts=targeted_counts, fragment_lengths=lengths,
                covered_bases=covered, peak_coverage=peak_coverage)


# --------------------------------------------------------------------------- #
# get_counts_by_barcode (oracle ahead of the kernel)         #
# --------------------------------------------------------------------------- #
def tools_target_regions(lines, padding=0):
    """lib/python/tools/regions.py:96-128: BED lines -> {contig: sorted [(start - padding, end + padding, name, strand)]}.
    Rows are ordered as tuples; None sorts before any string as it does in Python 2."""
    targets = {}
    for line in lines:
        info = line.strip().split("\t")
        if any(line.startswith(c) for c in ("browser", "track", "-browser", "-track", "#")):
            continue
        if len(line.strip()) == 0:
            continue
        name = info[3] if len(info) >= 4 else None
        strand = info[5] if len(info) >= 6 else None
        if strand not in ("+", "-"):
            strand = None
        targets.setdefault(info[0], []).append((int(info[1]) - padding, int(info[2]) + padding, name, strand))
    py2_key = lambda r: (r[0], r[1], r[2] is not None, r[2] or "", r[3] is not None, r[3] or "")
    return {chrom: sorted(rows, key=py2_key) for chrom, rows in targets.items()}


def region_containing_point(rows, starts, point):
    """lib/python/tools/regions.py:66-81 (contains_point / get_region_containing_point)."""
    k = bisect.bisect(starts, point) - 1
    if k >= 0 and starts[k] <= point < rows[k][1]:
        return rows[k]
    return None


def relative_position(region, point):
    """Region.get_relative_position, lib/python/tools/regions.py:24-36 (width = end - start - 1)."""
    start, end, _name, strand = region
    width = end - start - 1
    if strand != "-":
        return (point - start) - width // 2
    return (end - point) - width // 2


def get_counts_by_barcode(tracks, peaks_text, fragments, species_list, species_of_contig, known_cells_text=None, contig=None):
    """lib/python/tools/io.py:429-555.  tracks = {"tss" | "ctcf" | "dnase" | "enhancer" | "promoter" | "blacklist": BED text
    or None}; fragments = (contig, start, stop, barcode, dups) rows in file order.
    -> (read_count, {barcode: Counter}, tss_relpos, ctcf_relpos).  The "cell_id" label (io.py:482-488) numbers cells in
    CPython-2.7 dict order, emulated with the same table walk as the matrix columns (oracle/py2set.py)."""
    from . import py2set

    def load(text, padding=0):
        if text is None:
            return None
        rows = tools_target_regions(text.splitlines(True), padding)
        return {c: (r, [x[0] for x in r], [x[1] for x in r]) for c, r in rows.items()}

    tss, ctcf = load(tracks.get("tss"), 2000), load(tracks.get("ctcf"), 250)
    dnase, enhancer = load(tracks.get("dnase")), load(tracks.get("enhancer"))
    promoter, blacklist = load(tracks.get("promoter")), load(tracks.get("blacklist"))
    peaks = load(peaks_text)

    def point_region(c, position, regions):
        if regions is None or c not in regions:
            return None
        rows, starts, _ends = regions[c]
        return region_containing_point(rows, starts, position)

    def overlaps(c, start, stop, regions):
        if regions is None or c not in regions:
            return False
        _rows, starts, ends = regions[c]
        return overlaps_region(starts, ends, start, stop)

    cell_barcodes, inserted = {}, []
    if known_cells_text is not None:
        for line in known_cells_text.splitlines(True):
            items = line.strip("\n").split(",")
            for barcode in items[1:]:
                if barcode != "null":
                    if barcode not in cell_barcodes:
                        cell_barcodes[barcode] = []
                        inserted.append(barcode)
                    cell_barcodes[barcode] += [items[0]]
    cell_index = {}
    spnum = {species: 0 for species in species_list}
    py2_order = list(py2set.py2_set_order(inserted)) if inserted else []     # returns the keys in table order
    for species in species_list:
        for barcode in py2_order:
            if species in cell_barcodes[barcode]:
                label = "{}_cell_{}".format(species, spnum[species])
                spnum[species] += 1
                cell_index[barcode] = label if barcode not in cell_index else "_".join([cell_index[barcode], label])

    counts, tss_relpos, ctcf_relpos = {}, Counter(), Counter()
    read_count = 0
    for c, start, stop, barcode, dups in fragments:
        if contig is not None and c != contig:
            continue
        read_count += 2
        if barcode not in counts:
            counts[barcode] = Counter()
            if known_cells_text is not None:
                counts[barcode]["cell_id"] = cell_index.get(barcode, "None")
                for species in species_list:
                    counts[barcode]["is_{}_cell_barcode".format(species)] = 1 if species in cell_barcodes.get(barcode, []) else 0
        cb = counts[barcode]
        if known_cells_text is not None and len(species_list) > 1:
            sp = species_of_contig[c]
            cb["passed_filters_{}".format(sp)] += 1
            if overlaps(c, start, stop, peaks):
                cb["peak_region_fragments_{}".format(sp)] += 1
        cb["passed_filters"] += 1
        cb["total"] += dups
        cb["duplicate"] += dups - 1
        for position in (start, stop):
            r = point_region(c, position, tss)
            if r is not None:
                tss_relpos[relative_position(r, position)] += 1
            r = point_region(c, position, ctcf)
            if r is not None:
                ctcf_relpos[relative_position(r, position)] += 1
            if point_region(c, position, peaks) is not None:
                cb["peak_region_cutsites"] += 1
        targeted = False
        for regions, key in ((tss, "TSS_fragments"), (dnase, "DNase_sensitive_region_fragments"),
                             (enhancer, "enhancer_region_fragments"), (promoter, "promoter_region_fragments")):
            if overlaps(c, start, stop, regions):
                cb[key] += 1
                targeted = True
        if targeted:
            cb["on_target_fragments"] += 1
        if overlaps(c, start, stop, blacklist):
            cb["blacklist_region_fragments"] += 1
        if overlaps(c, start, stop, peaks):
            cb["peak_region_fragments"] += 1
    return read_count, counts, tss_relpos, ctcf_relpos


# --------------------------------------------------------------------------- #
# REPORT_INSERT_SIZES (one more scan of the fragments; oracle ahead of the kernel)
# --------------------------------------------------------------------------- #
MAX_INSERT_SIZE = 1000                      # stages/metrics/report_insert_sizes/__init__.py:32


def insert_size_table(barcodes, fragments, primary, exclude_non_nuclear):
    """REPORT_INSERT_SIZES.main for one barcode chunk (stages/metrics/report_insert_sizes/__init__.py:74-113).
    -> (rows in the order of `barcodes`: {barcode: int64[1001]} with column n - 1 = fragments of size n for n = 1..1000
    and column 1000 = fragments longer than 1000, and total int64[1000]).  Fragments of size 0 are counted under "0" by the
    reference and never written: they appear nowhere."""
    index = OrderedDict((bc, k) for k, bc in enumerate(barcodes))
    table = np.zeros((len(index), MAX_INSERT_SIZE + 1), dtype=np.int64)
    primary = set(primary)
    for contig, start, stop, barcode, _dup in fragments:
        if barcode not in index:
            continue
        if exclude_non_nuclear and contig not in primary:
            continue
        size = stop - start
        if size > MAX_INSERT_SIZE:
            table[index[barcode], MAX_INSERT_SIZE] += 1
        elif size >= 1:
            table[index[barcode], size - 1] += 1
    return OrderedDict((bc, table[k]) for bc, k in index.items()), table[:, :MAX_INSERT_SIZE].sum(axis=0)
