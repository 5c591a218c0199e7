"""The header-only C++ shim (include/clic/trajectory_estimator.hpp) compiles as reference-style host code and runs:
on the CPU box against the oracle library; on a GPU box against libclic_cuda.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "shim_test.cpp")
SRC2 = os.path.join(ROOT, "tests", "cpp", "shim_manager_test.cpp")


def _build_and_run(libdir, libname, tmp_path, src=SRC, tag="SHIM_OK", n_vals=2):
    exe = str(tmp_path / (os.path.basename(src)[:-4] + "_" + libname))
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, src, "-L" + libdir, "-l" + libname, "-Wl,-rpath," + libdir],
                   check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith(tag)]
    assert line, out.stdout
    return [float(x) for x in line[0].split()[1:1 + n_vals]]


def test_shim_on_oracle(oracle, tmp_path):
    init, final = _build_and_run(os.path.join(ROOT, "oracle", "_build"), "clic_oracle", tmp_path)
    assert final < init


@pytest.mark.gpu
def test_shim_on_cuda_matches_oracle(cuda, oracle, tmp_path):
    ig, fg = _build_and_run(os.path.join(ROOT, "clic_b200", "lib"), "clic_cuda", tmp_path)
    io, fo = _build_and_run(os.path.join(ROOT, "oracle", "_build"), "clic_oracle", tmp_path)
    assert abs(ig - io) <= 1e-9 * abs(io) and abs(fg - fo) <= 1e-6 * abs(fo)


def test_manager_style_calls_on_oracle(oracle, tmp_path):
    """InitFactorInfo statics (image weight 200), SetTimeoffsetState, PrepareMarginalizationInfo(RType_Prior, new
    MarginalizationFactor(...)), AddGravityFactor, SaveMarginalizationInfo, next solve on the new prior."""
    v = _build_and_run(os.path.join(ROOT, "oracle", "_build"), "clic_oracle", tmp_path, SRC2, "SHIM2_OK", 8)
    assert v[1] < v[0] and v[3] <= v[2]


@pytest.mark.gpu
def test_manager_style_calls_on_cuda_match_oracle(cuda, oracle, tmp_path):
    g = _build_and_run(os.path.join(ROOT, "clic_b200", "lib"), "clic_cuda", tmp_path, SRC2, "SHIM2_OK", 8)
    o = _build_and_run(os.path.join(ROOT, "oracle", "_build"), "clic_oracle", tmp_path, SRC2, "SHIM2_OK", 8)
    assert g[4:] == o[4:]                                       # prior dimensions n, m of both marginalizations
    assert abs(g[0] - o[0]) <= 1e-9 * abs(o[0]) and abs(g[1] - o[1]) <= 1e-6 * abs(o[1])
    assert abs(g[2] - o[2]) <= 1e-6 * abs(o[2]) and abs(g[3] - o[3]) <= 1e-5 * abs(o[3])
