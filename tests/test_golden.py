"""Committed golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle):
 * CPU: the oracle still reproduces them (regression pin of the oracle itself);
 * GPU: the CUDA path reproduces them through the C ABI (same tolerances as the live oracle comparison)."""
import glob
import json
import os

import numpy as np
import pytest

from clic_b200 import synthetic as S
from golden.make_golden import build

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "c*_small.npz")))


def check(lib, path, tol_eval, tol_pose):
    g = np.load(path, allow_pickle=False)
    sw = S.make_window(**json.loads(str(g["window_kwargs"])))
    ekw = json.loads(str(g["estimator_kwargs"]))
    est = build(lib, sw, **ekw)
    for k in range(4):
        s, u = est.indices(k)
        assert np.array_equal(s, g[f"s{k}"]) and np.array_equal(u, g[f"u{k}"])  # indexing: bit-exact
    H, b, cost, fcost = est.Evaluate()
    d = np.sqrt(np.maximum(np.diag(g["H"]), 1e-300))
    assert (np.abs(H - g["H"]) / np.outer(d, d)).max() < tol_eval
    assert np.abs(b - g["b"]).max() <= tol_eval * np.abs(g["b"]).max()
    assert abs(cost - float(g["cost"])) <= tol_eval * abs(float(g["cost"]))
    summ = est.Solve(8)
    ref = json.loads(str(g["summary"]))
    assert summ["iterations"] == ref["iterations"] and summ["termination"] == ref["termination"]
    x = est.read_state()
    assert np.abs(x["knot_p"] - g["knot_p"]).max() < tol_pose
    dq = S.qmul(S.qconj(x["knot_q"]), g["knot_q"])
    assert (2 * np.arctan2(np.linalg.norm(dq[:, :3], axis=1), np.abs(dq[:, 3]))).max() < tol_pose


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_reproduces_golden(oracle, path):
    check(oracle, path, 1e-13, 1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_cuda_reproduces_golden(cuda, path):
    check(cuda, path, 1e-9, 1e-7)


def test_have_goldens():
    assert len(FILES) >= 4


# ---- single analytic factors (clic_evaluate_factor): tests/golden/factors_extra.npz, written by make_golden_factors.py ----
def check_factors(lib, tol):
    import factor_cases as FC
    from clic_b200 import estimator as E
    from test_oracle_autodiff import reference_test_trajectory
    g = np.load(os.path.join(HERE, "golden", "factors_extra.npz"), allow_pickle=False)
    n = 0
    for seed in (1, 2, 3):
        w = reference_test_trajectory(seed)
        est = E.TrajectoryEstimator(w, E.default_options(lib), lib)
        for c in FC.cases(w, seed):
            r, J = FC.evaluate(est, c)
            rg, Jg = g[f"s{seed}/{c['name']}/r"], g[f"s{seed}/{c['name']}/J"]
            assert r.shape == rg.shape and J.shape == Jg.shape
            assert np.abs(r - rg).max() <= tol * max(1.0, np.abs(rg).max()), (seed, c["name"])
            assert np.abs(J - Jg).max() <= 10 * tol * max(1.0, np.abs(Jg).max()), (seed, c["name"])
            n += 1
        est.close()
    assert n == len(g.files) // 2


def test_oracle_reproduces_factor_goldens(oracle):
    check_factors(oracle, 1e-14)


@pytest.mark.gpu
def test_cuda_reproduces_factor_goldens(cuda):
    check_factors(cuda, 1e-12)
