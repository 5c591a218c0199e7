"""CUDA evaluators of the small analytic factors (every class of the reference's
test_analytic_factor.cpp:1192-1333, incl. the six that never enter the shipped pipeline) against the oracle through
clic_evaluate_factor, on the reference test's hard-coded inputs; PreIntegrationFactor additionally inside a solve."""
import numpy as np
import pytest

import factor_cases as FC
from clic_b200 import estimator as E
from test_oracle_autodiff import reference_test_trajectory

pytestmark = pytest.mark.gpu
NAMES = [c["name"] for c in FC.cases(reference_test_trajectory(1))]


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("name", NAMES)
def test_factor_matches_oracle(cuda, oracle, seed, name):
    w = reference_test_trajectory(seed)
    eg = E.TrajectoryEstimator(w, E.default_options(cuda), cuda)
    eo = E.TrajectoryEstimator(w, E.default_options(oracle), oracle)
    c = next(c for c in FC.cases(w, seed) if c["name"] == name)
    rg, Jg = FC.evaluate(eg, c)
    ro, Jo = FC.evaluate(eo, c)
    assert rg.shape == ro.shape and Jg.shape == Jo.shape
    assert np.abs(rg - ro).max() <= 1e-12 * max(1.0, np.abs(ro).max()), np.abs(rg - ro).max()
    assert np.abs(Jg - Jo).max() <= 1e-11 * max(1.0, np.abs(Jo).max()), np.abs(Jg - Jo).max()
    assert np.array_equal(Jg == 0.0, Jo == 0.0) or np.abs(Jg - Jo).max() < 1e-13  # same sparsity (same knots touched)
    r0, _ = FC.evaluate(eg, c, want_jacobian=False)
    assert np.array_equal(r0, rg)
    eg.close()
    eo.close()


def test_time_outside_window_is_refused(cuda):
    w = reference_test_trajectory(1)
    est = E.TrajectoryEstimator(w, E.default_options(cuda), cuda)
    with pytest.raises(E.ClicError):
        est.evaluate_factor(FC._abi.FACTOR_EPIPOLAR, t_a_ns=int(5e9), d=np.ones(14))
    with pytest.raises(E.ClicError):
        est.evaluate_factor(99)
    est.close()


@pytest.mark.parametrize("lock_bias", [False, True])
def test_preintegration_in_a_solve(cuda, oracle, lock_bias):
    """AddPreIntegrationAnalytic next to LiDAR + IMU factors: H, b, cost and the solved state against the oracle."""
    from clic_b200 import synthetic as S
    sw = S.make_window(10, n_lidar=1500, n_imu=150, seed=77)
    w = sw.window
    pre = FC.reference_preintegration()
    res = {}
    for name, lib in (("g", cuda), ("o", oracle)):
        opt = E.default_options(lib)
        opt.lock_ab = opt.lock_wb = int(lock_bias)
        est = E.TrajectoryEstimator(w, opt, lib)
        est.SetFixedIndex(-1)
        S.add_all_factors(est, sw, bias_factor=not lock_bias)
        t0 = w.t0_ns
        assert est.AddPreIntegrationAnalytic(t0 + FC.ns(0.21), t0 + FC.ns(0.695), FC.preintegration_struct(pre), 0, 1)
        assert est.AddPreIntegrationAnalytic(t0 + FC.ns(0.05), t0 + FC.ns(0.09), FC.preintegration_struct(pre), 1, 1)
        assert not est.AddPreIntegrationAnalytic(t0 + FC.ns(0.21), t0 + FC.ns(70.0), FC.preintegration_struct(pre), 0, 1)
        H, b, c, f = est.Evaluate()
        cost, count = est.residual_summary()
        s = est.Solve(6)
        res[name] = (H, b, c, f, s, est.read_state(), cost, count)
        est.close()
    Hg, bg, cg, fg, sg, xg, costg, countg = res["g"]
    Ho, bo, co, fo, so, xo, costo, counto = res["o"]
    assert np.abs(Hg - Ho).max() <= 1e-9 * np.abs(Ho).max()
    assert np.abs(bg - bo).max() <= 1e-9 * np.abs(bo).max()
    assert abs(cg - co) <= 1e-9 * abs(co) and abs(fg - fo) <= 1e-9 * max(1.0, abs(fo))
    assert abs(costg[14] - costo[14]) <= 1e-9 * abs(costo[14]) and countg[14] == counto[14] == 2
    assert sg["iterations"] == so["iterations"] and sg["termination_reason"] == so["termination_reason"]
    assert np.abs(xg["knot_p"] - xo["knot_p"]).max() < 1e-7
    assert np.abs(xg["knot_q"] - xo["knot_q"]).max() < 1e-7
    assert np.abs(xg["bias"] - xo["bias"]).max() < 1e-7
