"""CPU-side checks of the product package: the C-ABI library loads and exports every symbol the
header declares, and the host-only entry points agree with the oracle.  No GPU compute here."""
import os
import re
import subprocess

import numpy as np
import pytest

import util  # noqa: F401  (sys.path)
from cratac import _ffi, bgzf, synth
from oracle import py2set

ROOT = util.ROOT


def test_header_symbols_are_exported_and_bound():
    header = open(os.path.join(ROOT, "include", "cratac.h")).read()
    declared = set(re.findall(r"\b(cratac_[a-z0-9_]+)\s*\(", header))
    declared -= {"cratac_ctx", "cratac_run_result"}
    assert declared, "no declarations parsed"
    assert declared == set(_ffi.SIGNATURES), (declared ^ set(_ffi.SIGNATURES))
    lib = _ffi.load()
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", _ffi.LIB_PATH]).decode()
    exported = set(re.findall(r"\bT (cratac_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    assert lib.cratac_abi_version() == 1


def test_no_device_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_ffi.CratacError) as e:
        _ffi.Context([("chr1", 1000)])
    assert e.value.code == _ffi.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_py2_set_order_matches_oracle():
    rng = np.random.default_rng(3)
    keys = ["".join(rng.choice(list("ACGT"), size=16)) + "-1" for _ in range(20000)]
    keys = list(dict.fromkeys(keys))
    want = py2set.py2_set_order(keys)
    got = [keys[i] for i in _ffi.py2_set_order(keys)]
    assert got == want
    small = ["a", "b", "c", "d"]
    assert [small[i] for i in _ffi.py2_set_order(small)] == ["a", "c", "b", "d"]


def _parallel_set_rebuild(hashes):
    """numpy model of csrc/py2order.cu: every table of the CPython-2.7 set's life is the result of inserting a known
    sequence of keys one after the other, and that result is the fixed point of `slot <- min(sequence position asking for
    it)` with losers moving on along their probe sequence -- reached here in synchronous rounds, as the kernel does with
    atomicMin."""
    n = len(hashes)
    M64 = (1 << 64) - 1
    h = [int(x) & M64 for x in hashes]
    seq = []                                 # keys in slot order of the previous table
    size, fill, used = 8, 0, 0
    while True:
        room = (2 * size + 2) // 3
        fresh = min(n - used, room - fill if room > fill else (1 if n > used else 0))
        keys = seq + list(range(used, used + fresh))
        m = len(keys)
        mask = size - 1
        i = [h[k] & mask for k in keys]
        pert = [h[k] for k in keys]
        cand = [x & mask for x in i]
        table = [m] * size                   # m = empty
        changed = True
        while changed:
            changed = False
            for t in range(m):               # "parallel": every position asks for its candidate slot
                if table[cand[t]] > t:
                    table[cand[t]] = t
            for t in range(m):               # losers move one probe further
                if table[cand[t]] != t:
                    i[t] = (i[t] * 5 + pert[t] + 1) & M64
                    pert[t] >>= 5
                    cand[t] = i[t] & mask
                    changed = True
        seq = [keys[t] for t in table if t < m]
        fill += fresh
        used += fresh
        grow = fill * 3 >= size * 2
        if used == n and not grow:
            return seq
        if grow:
            minused = used * 2 if used > 50000 else used * 4
            size = 8
            while size <= minused:
                size <<= 1
            fill = used


def test_parallel_set_rebuild_is_the_sequential_insertion():
    """The algorithm behind cratac_py2_set_order_hashes_device, modelled on the host (the CUDA kernels themselves are
    checked on the GPU): the fixed point of the parallel rounds equals one-by-one insertion for random hashes, long
    probe chains and clustered home slots, through several table growths."""
    rng = np.random.default_rng(12)
    cases = [rng.integers(-2 ** 62, 2 ** 62, size=n, dtype=np.int64) for n in (1, 5, 6, 21, 22, 90, 1500)]
    cases.append((rng.integers(0, 8, size=600, dtype=np.int64) << 20) + np.arange(600, dtype=np.int64) * (1 << 40))
    cases.append(np.arange(900, dtype=np.int64) * 8)
    for hs in cases:
        want = _ffi.py2_set_order_hashes(hs).tolist()
        assert _parallel_set_rebuild(hs.tolist()) == want, hs.size


def _warp_upper_bound(arr, lo, hi, key):
    """Lane-by-lane model of warp_upper_bound (csrc/matrix.cu): 32 probes per round, ballot, narrow."""
    while hi - lo > 32:
        step = (hi - lo) // 33 + 1
        probes = [lo + step * (lane + 1) - 1 for lane in range(32)]
        gt = [(key < arr[p]) if p < hi else True for p in probes]
        if not any(gt):
            lo = lo + step * 32
            continue
        j = gt.index(True)
        lo, hi = (lo + step * j if j > 0 else lo), (probes[j] if probes[j] < hi else hi)
    gt = [(key < arr[lo + lane]) if lo + lane < hi else True for lane in range(32)]
    return lo + gt.index(True) if any(gt) else hi


def test_warp_wide_search_model_equals_bisect():
    import bisect
    rng = np.random.default_rng(4)
    for _ in range(4000):
        n = int(rng.choice([0, 1, 5, 31, 32, 33, 34, 64, 65, 100, 1000, 1089, 1090, 5000, 40000]))
        arr = np.sort(rng.integers(0, max(1, n * int(rng.choice([1, 3, 50]))), size=n)).tolist()
        lo = int(rng.integers(0, n + 1))
        hi = int(rng.integers(lo, n + 1))
        key = int(rng.integers(-2, (arr[-1] if n else 0) + 3))
        assert _warp_upper_bound(arr, lo, hi, key) == bisect.bisect_right(arr, key, lo, hi)


def test_inflate_bgzf_and_gzip(tmp_path):
    import gzip
    frags = synth.generate(synth.GRCH38[:2], 30000, 20, 50, seed=7)
    text = synth.to_text(frags)
    p = str(tmp_path / "fragments.tsv.gz")
    bgzf.write(p, text)
    assert gzip.open(p).read() == text                      # any gzip reader can read our BGZF
    got = _ffi.inflate_file(p, threads=4)
    assert got.tobytes() == text
    p2 = str(tmp_path / "plain.tsv.gz")
    with gzip.open(p2, "wb") as f:
        f.write(text)
    assert _ffi.inflate_file(p2, threads=2).tobytes() == text
    p3 = str(tmp_path / "raw.tsv")
    open(p3, "wb").write(text)
    assert _ffi.inflate_file(p3).tobytes() == text


def test_own_deflate_decoder_against_zlib(tmp_path):
    """csrc/inflate.cpp decodes every valid BGZF block itself (no block handed to zlib) and reproduces zlib's output:
    stored, fixed-Huffman and dynamic blocks, long codes (second-level tables), matches at every small distance, runs,
    incompressible data, block sizes from 0 to the BGZF maximum."""
    import zlib
    rng = np.random.default_rng(8)
    text = synth.to_text(synth.generate(synth.GRCH38[:2], 20000, 20, 50, seed=17))
    words = [bytes(rng.integers(97, 123, size=int(rng.integers(2, 12)), dtype=np.uint8)) for _ in range(400)]
    datas = {
        "empty": b"", "one_byte": b"x", "fragments_text": text,
        "random_bytes": rng.integers(0, 256, size=300000, dtype=np.uint8).tobytes(),
        "zeros": bytes(200000), "run_then_noise": b"A" * 70000 + rng.integers(0, 256, size=5000, dtype=np.uint8).tobytes(),
        "small_periods": b"".join((bytes(range(65, 65 + k)) * 3000)[:9000] for k in range(1, 12)),
        "skewed_alphabet": bytes(rng.choice(np.arange(256, dtype=np.uint8), size=250000,
                                             p=(lambda w: w / w.sum())(1.0 / np.arange(1, 257) ** 2.2)).tolist()),
        "dictionary_words": b" ".join(words[int(i)] for i in rng.integers(0, len(words), size=60000)),
        "exact_block": text[:bgzf.BLOCK_IN], "block_plus_one": text[:bgzf.BLOCK_IN + 1],
    }
    settings = [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE), (9, zlib.Z_FILTERED)]
    _ffi.inflate_mode(1)
    p = str(tmp_path / "x.gz")
    for name, data in datas.items():
        for level, strategy in settings:
            for block_in in (bgzf.BLOCK_IN, 700):
                if block_in == 700 and len(data) > 40000:
                    continue
                with open(p, "wb") as f:
                    f.write(bgzf.compress(data, level, strategy, block_in))
                _ffi.inflate_stats(reset=True)
                got = _ffi.inflate_file(p, threads=3).tobytes()
                fast, slow = _ffi.inflate_stats()
                assert got == data, (name, level, strategy, block_in)
                assert slow == 0 and fast > 0, (name, level, strategy, block_in, fast, slow)
    # the zlib route gives the same bytes (measurement switch)
    with open(p, "wb") as f:
        f.write(bgzf.compress(text, 6))
    _ffi.inflate_mode(0)
    _ffi.inflate_stats(reset=True)
    assert _ffi.inflate_file(p, threads=2).tobytes() == text and _ffi.inflate_stats()[0] == 0
    _ffi.inflate_mode(1)


def test_corrupt_bgzf_blocks_fail_cleanly(tmp_path):
    """Damaged payloads: the own decoder declines and zlib confirms, or the block's CRC-32 does not match -> an error;
    never a crash or an out-of-bounds write (damage confined to header bytes gzip ignores may still decode)."""
    rng = np.random.default_rng(21)
    text = synth.to_text(synth.generate(synth.GRCH38[:1], 6000, 10, 20, seed=3))
    blob = bytearray(bgzf.compress(text, 6))
    p = str(tmp_path / "bad.gz")
    outcomes = {"error": 0, "decoded": 0}
    for _ in range(300):
        bad = bytearray(blob)
        for _k in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(18, len(bad) - 40))            # keep the first header and the EOF block intact
            bad[pos] ^= int(rng.integers(1, 256))
        with open(p, "wb") as f:
            f.write(bad)
        try:
            _ffi.inflate_file(p, threads=2)
            outcomes["decoded"] += 1
        except _ffi.CratacError:
            outcomes["error"] += 1
    assert outcomes["error"] > 250


def test_malformed_bgzf_headers_fail_cleanly(tmp_path):
    """a BSIZE smaller than the block's own header + trailer, an extra field running past the block, an ISIZE
    above 64 KiB and a truncated last block must come back as errors (never an out-of-bounds read of the mapping); a
    payload whose CRC-32 does not match the trailer is an error too, as it is for the reference's `gzip -c -d` pipe
    (lib/python/tools/io.py:197-217)."""
    import struct
    text = synth.to_text(synth.generate(synth.GRCH38[:1], 4000, 10, 20, seed=5))
    blob = bytearray(bgzf.compress(text, 6))
    p = str(tmp_path / "bad.gz")

    def expect_error(data, ranged=True, gzip_readable=False):
        with open(p, "wb") as f:
            f.write(data)
        if gzip_readable:        # only the BGZF extra field is damaged: the file is still a valid multi-member gzip, which
            assert _ffi.inflate_file(p, threads=2).tobytes() == text   # the sequential path reads as `gzip -d` would
        else:
            with pytest.raises(_ffi.CratacError):
                _ffi.inflate_file(p, threads=2)
        if ranged:
            with pytest.raises(_ffi.CratacError):
                _ffi.inflate_range(p, 0, (len(data) - 28) << 16, threads=2)

    for bsize in (0, 3, 11, 17, 24):                     # BSIZE field = block size - 1: smaller than header + trailer
        bad = bytearray(blob)
        bad[16:18] = struct.pack("<H", bsize)
        expect_error(bad, gzip_readable=True)
    bad = bytearray(blob)
    bad[10:12] = struct.pack("<H", 60000)                # XLEN beyond the block
    expect_error(bad)
    bad = bytearray(blob)
    bs = struct.unpack("<H", bytes(blob[16:18]))[0] + 1
    bad[bs - 4:bs] = struct.pack("<I", 1 << 20)          # ISIZE above the BGZF limit
    expect_error(bad)
    expect_error(blob[:bs - 3], ranged=False)            # file ends inside the first block
    bad = bytearray(blob)
    bad[bs - 8] ^= 0x5A                                  # CRC-32 of the first block
    expect_error(bad)
    with open(p, "wb") as f:                             # and the undamaged file still reads
        f.write(blob)
    assert _ffi.inflate_file(p, threads=2).tobytes() == text


def test_own_deflate_decoder_under_sanitizers(tmp_path):
    """The decoder on damaged / truncated / mis-sized blocks, compiled with -fsanitize=address,undefined."""
    exe = str(tmp_path / "inflate_fuzz")
    src = [os.path.join(ROOT, "tests", "inflate_fuzz.cpp"), os.path.join(ROOT, "cellranger-atac_b200", "csrc", "inflate.cpp")]
    build = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-o", exe] + src + ["-lpthread"],
                           capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("no sanitizer runtime in this toolchain")
    assert build.returncode == 0, build.stderr
    text = synth.to_text(synth.generate(synth.GRCH38[:2], 30000, 20, 50, seed=23))
    p = str(tmp_path / "f.gz")
    bgzf.write(p, text, level=6)
    run = subprocess.run([exe, p, "6000"], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-2000:]
    ok, declined = [int(x.split("=")[1]) for x in run.stdout.split()]
    assert ok > 1000 and declined > 1000


def test_tabix_index_and_contig_range_inflate(tmp_path):
    """Host ingest: a rank reads only the BGZF blocks of its contig range, located through the
    tabix index (the reference fetches contig by contig through pysam.TabixFile, lib/python/tools/io.py:82-93)."""
    from cratac import tabix
    contigs = [("chr1", 3000000), ("chr2", 2000000), ("chr3", 500000), ("chr4", 900000), ("chrM", 16000)]
    frags = synth.generate(contigs[:4], 40000, 20, 80, seed=9)        # chrM has no record
    text = synth.to_text(frags)
    p = str(tmp_path / "fragments.tsv.gz")
    bgzf.write(p, text)
    tbi = tabix.build_index(p)
    idx = tabix.read_index(tbi)
    assert idx["names"] == ["chr1", "chr2", "chr3", "chr4"] and idx["format"] == tabix.TBX_UCSC
    assert (idx["col_seq"], idx["col_beg"], idx["col_end"]) == (1, 2, 3)
    lines = text.split(b"\n")[:-1]
    by_contig = {}
    for ln in lines:
        by_contig.setdefault(ln.split(b"\t")[0].decode(), []).append(ln)
    spans = tabix.contig_spans(idx)
    assert set(spans) == set(by_contig)
    for name, (v0, v1) in spans.items():                                # every contig on its own
        got = _ffi.inflate_range(p, v0, v1, threads=3).tobytes()
        assert got == b"\n".join(by_contig[name]) + b"\n"
    # contiguous contig ranges, as the multi-GPU step partitions them; a range without records is empty
    for first, last in ((0, 2), (2, 5), (1, 4), (4, 5), (0, 5)):
        names = [c[0] for c in contigs[first:last]]
        want = b"".join(b"\n".join(by_contig[n]) + b"\n" for n in names if n in by_contig)
        assert _ffi.inflate_contigs(p, names, threads=2).tobytes() == want
    assert _ffi.inflate_contigs(p, ["chr3", "chr1"]).tobytes() == b"".join(b"\n".join(by_contig[n]) + b"\n" for n in ("chr3", "chr1"))
    # every rank of a 1..4-GPU step reads its own range; together they read the file once
    from cratac import multi
    for world in (1, 2, 3, 4):
        texts = [multi.rank_text(p, contigs, r, world, threads=2) for r in range(world)]
        assert b"".join(t.tobytes() for t, _ in texts) == text
        assert [rng for _, rng in texts] == multi.partition_contigs([c[1] for c in contigs], world)
    # the spans tile the file: consecutive contigs meet, and the union is the whole text
    order = [spans[n] for n in idx["names"]]
    assert all(order[k][1] == order[k + 1][0] for k in range(len(order) - 1))
    assert _ffi.inflate_range(p, order[0][0], order[-1][1]).tobytes() == text
    # a range inside one block, and ranges that end exactly on a block boundary
    v0 = spans["chr1"][0]
    assert _ffi.inflate_range(p, v0 + 10, v0 + 110).tobytes() == text[10:110]
    assert _ffi.inflate_range(p, v0, v0).size == 0
    assert _ffi.inflate_range(p, v0, v0 + bgzf.BLOCK_IN).tobytes() == text[:bgzf.BLOCK_IN]      # end = isize of the block
    with pytest.raises(_ffi.CratacError):
        _ffi.inflate_range(p, v0 + 7, (v0 + 7) + (1 << 40))                                  # beyond the file


def test_tabix_bins_and_linear_index(tmp_path):
    """reg2bin is the UCSC scheme of the tabix paper (16 kb leaves, 5 levels); the linear index holds, per 16 kb
    window, the offset of the first record that overlaps it."""
    from cratac import tabix
    assert tabix.reg2bin(0, 1) == 4681 and tabix.reg2bin(16384, 16385) == 4682
    assert tabix.reg2bin(0, 16385) == 585 and tabix.reg2bin(0, 1 << 29) == 0
    assert tabix.reg2bin((1 << 17) - 1, (1 << 17) + 1) == 73 and tabix.reg2bin(1 << 26, (1 << 26) + 5) == 4681 + 4096
    rows = [("chrA", 5, 300), ("chrA", 20000, 20100), ("chrA", 20050, 40000), ("chrB", 1, 2)]
    text = "".join("%s\t%d\t%d\tACGTACGTACGTACGT-1\t1\n" % r for r in rows).encode()
    p = str(tmp_path / "f.tsv.gz")
    bgzf.write(p, text)
    idx = tabix.read_index(tabix.build_index(p))
    a, b = idx["refs"]
    first = tabix.contig_spans(idx)["chrA"][0]
    line_len = [len(("%s\t%d\t%d\tACGTACGTACGTACGT-1\t1\n" % r)) for r in rows]
    assert sorted(a["bins"]) == sorted({tabix.reg2bin(5, 300), tabix.reg2bin(20000, 20100), tabix.reg2bin(20050, 40000)})
    assert a["linear"] == [first, first + line_len[0], first + line_len[0] + line_len[1]]
    assert list(b["bins"]) == [4681] and len(b["linear"]) == 1


def test_synth_is_sorted_and_valid():
    contigs = synth.GRCH38[:3]
    f = synth.generate(contigs, 5000, 10, 30, seed=1)
    key = f["chrom"].astype(np.int64) * (1 << 40) + f["start"].astype(np.int64)
    assert np.all(np.diff(key) >= 0)
    L = np.array([c[1] for c in contigs])[f["chrom"]]
    assert np.all(f["start"] >= 0) and np.all(f["start"] < f["end"]) and np.all(f["end"] <= L)
    text = synth.to_text(f)
    dec = util.oracle_decode(text, [c[0] for c in contigs])
    assert np.array_equal(dec["start"], f["start"]) and np.array_equal(dec["end"], f["end"])
    assert dec["barcodes"] == synth.barcode_strings(f).tolist()


def test_saddle_point_negative_binomial_pmf_against_scipy(tmp_path):
    """csrc/nb_pmf.cuh (the NB pmf of the fit kernels when 1/alpha is large, i.e. the dispersion sits near its 1e-6
    floor) compiled for the host: equal to scipy.stats.nbinom.pmf -- what analysis/peaks.py:57,75 calls -- to 5e-12
    relative, where the log-gamma form is off by 1e-9."""
    from scipy.stats import nbinom
    exe = str(tmp_path / "nb_pmf_host_test")
    build = subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "nb_pmf_host_test.cpp"), "-lm"],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    rows = []
    for alpha in (1e-6, 1e-5, 1e-4, 5e-4, 9.9e-4):
        for mu in (0.7, 1.5, 4.6, 12.0, 80.0, 300.0):
            n, p = 1.0 / alpha, 1.0 / (1.0 + alpha * mu)
            rows += [(k, n, p) for k in (0, 1, 2, 3, 5, 8, 13, 15, 16, 17, 20, 40, 100, 300, 999)]
    run = subprocess.run([exe], input="\n".join("%d %.17g %.17g" % r for r in rows), capture_output=True, text=True)
    assert run.returncode == 0
    mine = np.array([float(x) for x in run.stdout.split()])
    ref = np.array([nbinom.pmf(*r) for r in rows])
    big = ref > 1e-290
    assert big.sum() > 300 and np.max(np.abs(mine[big] - ref[big]) / ref[big]) < 5e-12

def test_speculation_rule_on_recorded_em_trajectories(tmp_path):
    """csrc/spec_rule.cuh -- when the fit kernel tells the host which threshold it is heading for -- compiled for the host
    and fed the reference's own EM trajectories (tests/golden/spec_trajectories.json: analysis/peaks.py replayed iterate by
    iterate on 20 libraries at the depth and peak density of the bench workloads).  The last guess it publishes is the
    threshold the fit ends on, in every case; a guess never goes out before iterate spec_iter; a second guess differs from
    the first.  With the fixed iterate of round 2's first version (threshold of iterate 32, nothing else) 9 of the 20
    would have been redone."""
    import json
    exe = str(tmp_path / "spec_rule_host_test")
    build = subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "spec_rule_host_test.cpp"), "-lm"],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "spec_trajectories.json")))
    assert len(cases) == 20 and not any(c["reseeded"] for c in cases)
    spec_iter = 32
    text = "".join("%d %d\n" % (spec_iter, len(c["trajectory"])) + "".join("%d %.9f\n" % (t, x) for t, x in c["trajectory"]) for c in cases)
    run = subprocess.run([exe], input=text, capture_output=True, text=True)
    assert run.returncode == 0
    rows = [[int(v) for v in line.split()] for line in run.stdout.strip().splitlines()]
    assert len(rows) == len(cases)
    fixed_iterate_right = second_guesses = 0
    for c, (g1, i1, g2, i2) in zip(cases, rows):
        assert i1 >= spec_iter and g1 >= 0, c["name"]
        last = g2 if i2 >= 0 else g1
        assert last == c["final"], (c["name"], g1, i1, g2, i2)
        if i2 >= 0:
            assert g2 != g1 and i2 > i1
            second_guesses += 1
        fixed_iterate_right += c["trajectory"][spec_iter - 1][0] == c["final"]
    assert fixed_iterate_right < 12 and 0 < second_guesses < 10
    # a fit that ends before iterate spec_iter publishes nothing; spec_iter = 1 publishes the first iterate's threshold
    run = subprocess.run([exe], input="32 3\n40 40.5\n41 41.5\n41 41.6\n1 2\n40 40.5\n41 41.5\n", capture_output=True, text=True)
    assert [line.split() for line in run.stdout.strip().splitlines()] == [["-1", "-1", "-1", "-1"], ["40", "1", "-1", "-1"]]

def test_tabix_reader_takes_what_htslib_adds(tmp_path):
    """An index as htslib writes it carries, per reference, the pseudo-bin 37450 (two "chunks" that are not chunks: the
    virtual-offset range of the reference, then the record counts) and, after the last reference, the number of records
    without coordinates.  The module's own writer emits neither; here both are added to one of its indexes (serialised by
    hand from the format description of tbx.c / the tabix paper) and the reader must give the same spans and ranges."""
    import struct
    from cratac import tabix
    contigs = [("chr1", 3000000), ("chr2", 2000000), ("chr3", 500000)]
    text = synth.to_text(synth.generate(contigs, 30000, 20, 80, seed=19))
    p = str(tmp_path / "fragments.tsv.gz")
    bgzf.write(p, text)
    plain = tabix.read_index(tabix.build_index(p))
    spans = tabix.contig_spans(plain)
    n_records = {n: sum(1 for ln in text.split(b"\n")[:-1] if ln.split(b"\t")[0].decode() == n) for n in plain["names"]}
    nm = b"".join(x.encode() + b"\x00" for x in plain["names"])
    out = [tabix.MAGIC, struct.pack("<8i", len(plain["names"]), plain["format"], plain["col_seq"], plain["col_beg"], plain["col_end"],
                                     plain["meta"], plain["skip"], len(nm)), nm]
    for name, ref in zip(plain["names"], plain["refs"]):
        bins = dict(ref["bins"])
        bins[tabix.PSEUDO_BIN] = [spans[name], (n_records[name], 0)]          # (ref_beg, ref_end), (n_mapped, n_unmapped)
        out.append(struct.pack("<i", len(bins)))
        for b in sorted(bins, reverse=True):                                  # htslib's hash order is arbitrary: not ascending
            out.append(struct.pack("<Ii", b, len(bins[b])))
            out += [struct.pack("<QQ", c0, c1) for c0, c1 in bins[b]]
        out.append(struct.pack("<i", len(ref["linear"])))
        out.append(struct.pack("<%dQ" % len(ref["linear"]), *ref["linear"]))
    out.append(struct.pack("<Q", 0))                                          # n_no_coor
    p2 = str(tmp_path / "htslib_style.tbi")
    with open(p2, "wb") as f:
        f.write(bgzf.compress(b"".join(out), 6))
    styled = tabix.read_index(p2)
    assert styled["names"] == plain["names"] and tabix.contig_spans(styled) == spans
    assert all(tabix.PSEUDO_BIN in r["bins"] for r in styled["refs"])
    assert tabix.range_span(styled, ["chr2", "chr3"]) == tabix.range_span(plain, ["chr2", "chr3"])
    beg, end = tabix.range_span(styled, ["chr2"])
    got = _ffi.inflate_range(p, beg, end, threads=2).tobytes()
    assert got == b"".join(ln + b"\n" for ln in text.split(b"\n")[:-1] if ln.startswith(b"chr2\t"))
