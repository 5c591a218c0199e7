"""Pins the oracle's restatement of every small analytic factor to the independent torch-f64 auto-diff
lineage through clic_evaluate_factor — the property the reference's only test asserts for the same classes on the same
hard-coded inputs (analytic == auto-diff, src/app/test_analytic_factor.cpp:846-900, 1192-1333)."""
import numpy as np
import pytest

import autodiff_ref as AD
import factor_cases as FC
from clic_b200 import estimator as E
from test_oracle_autodiff import reference_test_trajectory

TOL = 1e-6  # the reference's own bar (test_analytic_factor.cpp:872)
NAMES = [c["name"] for c in FC.cases(reference_test_trajectory(1))]


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("name", NAMES)
def test_factor_matches_autodiff(oracle, seed, name):
    w = reference_test_trajectory(seed)
    est = E.TrajectoryEstimator(w, E.default_options(oracle), oracle)
    c = next(c for c in FC.cases(w, seed) if c["name"] == name)
    r, J = FC.evaluate(est, c)
    K = w.n_knots
    r_ad, Jk, Jx = AD.jac(c["fn"], K, c["extras"])
    assert r.shape == r_ad.shape and J.shape == (r.size, 6 * K + sum(c["x_cols"]))
    scale = max(1.0, np.abs(Jk).max())
    assert np.allclose(r, r_ad, rtol=1e-9, atol=1e-8 * max(1.0, np.abs(r_ad).max())), (r, r_ad)
    assert np.abs(J[:, :6 * K] - Jk.reshape(r.size, -1)).max() < TOL * scale
    col = 6 * K
    for i, wdt in enumerate(c["x_cols"]):
        if i not in c["skip_extra"]:
            assert np.abs(J[:, col:col + wdt] - Jx[i].reshape(r.size, wdt)).max() < TOL * max(scale, np.abs(Jx[i]).max()), (name, i)
        col += wdt
    # residual-only evaluation gives the same residual
    r0, _ = FC.evaluate(est, c, want_jacobian=False)
    assert np.array_equal(r0, r)


def test_time_outside_window_is_refused(oracle):
    w = reference_test_trajectory(1)
    est = E.TrajectoryEstimator(w, E.default_options(oracle), oracle)
    with pytest.raises(E.ClicError):
        est.evaluate_factor(FC._abi.FACTOR_EPIPOLAR, t_a_ns=int(5e9), d=np.ones(14))
    with pytest.raises(E.ClicError):
        est.evaluate_factor(99)


def test_preintegration_in_the_problem(oracle):
    """AddPreIntegrationAnalytic: H = J^T J, b = J^T r of the factor over the knots and the two bias slots."""
    w = reference_test_trajectory(2)
    w.bias = np.array([[0.012, 0.19, -0.008, 0.011, 0.018, -0.012], [0.02, 0.21, -0.02, 0.02, 0.03, -0.02]])
    opt = E.default_options(oracle)
    opt.lock_ab = opt.lock_wb = 0
    est = E.TrajectoryEstimator(w, opt, oracle)
    est.SetFixedIndex(-1)
    pre = FC.reference_preintegration()
    assert est.AddPreIntegrationAnalytic(FC.ns(0.21), FC.ns(0.695), FC.preintegration_struct(pre), 0, 1)
    assert not est.AddPreIntegrationAnalytic(FC.ns(0.21), FC.ns(7.0), FC.preintegration_struct(pre), 0, 1)
    H, b, cost, _ = est.Evaluate()
    L = est.layout()
    x = np.concatenate([w.bias[0, :3], w.bias[1, :3], w.bias[0, 3:], w.bias[1, 3:]])
    r, J = est.evaluate_factor(FC._abi.FACTOR_PREINTEGRATION, FC.ns(0.21), FC.ns(0.695), [], x,
                               pre=FC.preintegration_struct(pre))
    K = w.n_knots
    Jf = np.zeros((15, L.D))
    Jf[:, :6 * K] = J[:, :6 * K]
    for blk, col in enumerate([L.col_bias_gyro, L.col_bias_gyro + 3, L.col_bias_accel, L.col_bias_accel + 3]):
        Jf[:, col:col + 3] = J[:, 6 * K + 3 * blk:6 * K + 3 * blk + 3]
    assert np.allclose(H, Jf.T @ Jf, rtol=1e-11, atol=1e-9 * np.abs(H).max())
    assert np.allclose(b, Jf.T @ r, rtol=1e-11, atol=1e-9 * np.abs(b).max())
    assert np.isclose(cost, 0.5 * r @ r, rtol=1e-12)
