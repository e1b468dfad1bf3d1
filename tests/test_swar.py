"""The word-at-a-time field parsers of the decoder (csrc/parse_swar.cuh) are host+device code: exercise them
on the host with randomised inputs against byte-wise references (no GPU needed)."""
import os
import subprocess

import util


def test_swar_helpers_on_host(tmp_path):
    src = os.path.join(os.path.dirname(__file__), "swar_host_test.cpp")
    exe = str(tmp_path / "swar_test")
    subprocess.check_call(["g++", "-O2", "-w", "-I", os.path.join(util.PKG, "csrc"), "-o", exe, src])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and "fails=0" in out.stdout, out.stdout[-2000:]
