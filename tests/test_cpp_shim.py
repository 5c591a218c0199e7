"""The header-only C++ shim (include/clic/trajectory_estimator.hpp) compiles as reference-style host code and runs:
on the CPU box against the oracle library; on a GPU box against libclic_cuda.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "shim_test.cpp")


def _build_and_run(libdir, libname, tmp_path):
    exe = str(tmp_path / "shim_test")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, SRC, "-L" + libdir, "-l" + libname, "-Wl,-rpath," + libdir],
                   check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith("SHIM_OK")]
    assert line, out.stdout
    return [float(x) for x in line[0].split()[1:3]]


def test_shim_on_oracle(oracle, tmp_path):
    init, final = _build_and_run(os.path.join(ROOT, "oracle", "_build"), "clic_oracle", tmp_path)
    assert final < init


@pytest.mark.gpu
def test_shim_on_cuda_matches_oracle(cuda, oracle, tmp_path):
    ig, fg = _build_and_run(os.path.join(ROOT, "clic_b200", "lib"), "clic_cuda", tmp_path)
    io, fo = _build_and_run(os.path.join(ROOT, "oracle", "_build"), "clic_oracle", tmp_path)
    assert abs(ig - io) <= 1e-9 * abs(io) and abs(fg - fo) <= 1e-6 * abs(fo)
