"""The C++ host mirror (randnla_b200/cpp/randblas.hpp) compiles against include/rnla.h (CPU check) and, on the GPU
box, re-runs the reference's unit tests for this path through it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "randnla_b200", "cpp", "test_randblas.cpp")


def _build(tmp_path):
    exe = str(tmp_path / "test_randblas")
    libdir = os.path.join(ROOT, "randnla_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "randnla_b200", "cpp"),
           SRC, "-o", exe, "-L", libdir, "-l:librnla.so", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True)
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path):
    _build(tmp_path)


@pytest.mark.gpu
def test_cpp_mirror_runs_reference_cases(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout + out.stderr
