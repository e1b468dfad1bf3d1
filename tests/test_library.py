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
