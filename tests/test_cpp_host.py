"""Builds and runs tests/cpp/test_rscoding.cpp -- the reference's bitmap.rs / rscoding.rs unit tests restated
against the C++ host mirror (summerset_b200/host/summerset_host.hpp) over the C ABI."""
import os
import subprocess
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "tests" / "cpp" / "test_rscoding"


def _build():
    from summerset_b200 import build as b
    b.build()
    src = ROOT / "tests" / "cpp" / "test_rscoding.cpp"
    hdr = ROOT / "summerset_b200" / "host" / "summerset_host.hpp"
    if BIN.exists() and BIN.stat().st_mtime > max(src.stat().st_mtime, hdr.stat().st_mtime):
        return
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-std=c++17", "-O1", "-Wall", str(src), "-o", str(BIN), f"-L{ROOT / 'summerset_b200'}",
                    "-lsummerset_b200", f"-Wl,-rpath,{ROOT / 'summerset_b200'}"], check=True)


def test_cpp_host_mirror_bookkeeping():
    _build()
    env = dict(os.environ)
    if not torch.cuda.is_available():
        env["SS_EXPECT_NO_GPU"] = "1"
    r = subprocess.run([str(BIN), "host"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_reference_tests_on_gpu():
    _build()
    r = subprocess.run([str(BIN), "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "gpu: 0 failure(s)" in r.stdout
