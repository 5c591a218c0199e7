"""AddGlobalVelocityMeasurement (trajectory_estimator.cpp:609-664) with auto_diff::IMUGlobalVelocityFactor
(factor/auto_diff/trajectory_value_factor.h:436-487): the one extra factor of TrajectoryManager::InitTrajWithPropagation
(trajectory_manager.cpp:233-300: IMU factors + one global velocity + Solve(50)).
CPU: the oracle's written-out Jacobians against central differences of an independent numpy residual (the reference
defines this factor by auto-diff only).  GPU: CUDA vs oracle, evaluation and the init solve."""
import numpy as np
import pytest

from clic_b200 import estimator as E
from clic_b200 import synthetic as S
from test_gpu_parity import rel_scaled
from test_gpu_solve import rot_err


def _window(seed=3, K=10):
    return S.make_window(K, n_lidar=10, n_imu=150, seed=seed, init_noise=(5e-3, 2e-2))


def _estimator(lib, sw, free_toff=False, fixed=2):
    opt = E.default_options(lib)
    if free_toff:
        opt.lock_t_offset[0] = 0
        opt.t_offset_padding_ns = 10_000_000
    est = E.TrajectoryEstimator(sw.window, opt, lib)
    est.SetFixedIndex(fixed)
    return est


def _np_residual(w, knot_p, toff_ns, t_s, v_meas, weight):
    """Independent definition: w * (d/dt of the uniform cubic B-spline at t - v), plain blending matrix."""
    t_ns = float(int(t_s * 1e9)) + toff_ns
    st = (t_ns - w.t0_ns) / w.dt_ns
    s = int(np.floor(st))
    u = st - s
    M = np.array([[1, -3, 3, -1], [4, 0, -6, 3], [1, 3, 3, -3], [0, 0, 0, 1]], dtype=np.float64) / 6.0
    dp = np.array([0.0, 1.0, 2 * u, 3 * u * u])
    c = (1e9 / w.dt_ns) * (M @ dp)
    return weight * (c @ knot_p[s:s + 4] - v_meas), s


def test_oracle_global_velocity_matches_finite_differences(oracle):
    sw = _window()
    w = sw.window
    t_s, v_meas, weight = 0.4321, np.array([1.2, -0.4, 0.3]), 25.0
    est = _estimator(oracle, sw, free_toff=True, fixed=-1)
    assert est.AddGlobalVelocityMeasurement(t_s, v_meas, weight)
    assert not est.AddGlobalVelocityMeasurement(5.0, v_meas, weight)      # outside the window: "return false"
    H, b, cost, _ = est.Evaluate()
    L = est.layout()
    r, s = _np_residual(w, w.knot_p, 0.0, t_s, v_meas, weight)
    sq = r @ r
    a2 = 60.0 ** 2
    assert abs(cost - 0.5 * a2 * np.log1p(sq / a2)) <= 1e-12 * cost       # 1/2 rho(|r|^2), CauchyLoss(60)
    rp = 1.0 / (1.0 + sq / a2)
    # central differences of the numpy residual wrt the 4 touched position knots and the time offset
    cols, J = [], []
    for k in range(s, s + 4):
        for d in range(3):
            kp = w.knot_p.copy(); kp[k, d] += 1e-6
            km = w.knot_p.copy(); km[k, d] -= 1e-6
            J.append((_np_residual(w, kp, 0.0, t_s, v_meas, weight)[0] - _np_residual(w, km, 0.0, t_s, v_meas, weight)[0]) / 2e-6)
            cols.append(6 * (k - L.first_free_knot) + 3 + d)
    J.append((_np_residual(w, w.knot_p, 1e3, t_s, v_meas, weight)[0] - _np_residual(w, w.knot_p, -1e3, t_s, v_meas, weight)[0]) / 2e3)
    cols.append(L.col_t_offset[0])
    J = np.array(J).T                                                       # 3 x 13
    Href, bref = np.zeros_like(H), np.zeros_like(b)
    Href[np.ix_(cols, cols)] = rp * J.T @ J
    bref[cols] = rp * J.T @ r
    assert np.abs(H - Href).max() <= 1e-6 * np.abs(Href).max()              # the reference's own 1e-6 (analytic vs auto-diff)
    assert np.abs(b - bref).max() <= 1e-6 * np.abs(bref).max()
    est.close()


def _add_init_problem(est, sw, v_t, v_meas):
    M = sw.imu
    est.AddIMUMeasurementAnalytic(M["timestamp"], M["gyro"], M["accel"], 1, M["info_vec"])
    assert est.AddGlobalVelocityMeasurement(v_t, v_meas, 25.0)


@pytest.mark.gpu
@pytest.mark.parametrize("free_toff", [False, True])
def test_gpu_global_velocity_evaluation(cuda, oracle, free_toff):
    sw = _window(seed=5)
    v_t = 0.612345678
    v_meas = np.array([0.9, -0.7, 0.25])
    out = {}
    for name, lib in (("g", cuda), ("o", oracle)):
        est = _estimator(lib, sw, free_toff=free_toff)
        _add_init_problem(est, sw, v_t, v_meas)
        assert not est.AddGlobalVelocityMeasurement(-1.0, v_meas, 25.0)
        out[name] = est.Evaluate()
        est.close()
    (Hg, bg, cg, fg), (Ho, bo, co, fo) = out["g"], out["o"]
    d = np.sqrt(np.maximum(np.diag(Ho), 1e-300))
    assert rel_scaled(Hg, Ho) < 1e-9
    assert (np.abs(bg - bo) / d).max() <= 1e-9 * (np.abs(bo) / d).max()
    assert abs(cg - co) <= 1e-9 * abs(co) and abs(fg - fo) <= 1e-9 * max(abs(fo), 1e-300)


@pytest.mark.gpu
def test_gpu_init_traj_with_propagation_solve(cuda, oracle):
    """InitTrajWithPropagation: IMU factors + one global velocity, biases locked, Solve(50)."""
    sw = _window(seed=9, K=12)
    v_t = 0.8765
    gt = S.SplineNP(sw.window.t0_ns, sw.window.dt_ns, sw.gt_knot_q, sw.gt_knot_p)
    t_ns = np.array([int(v_t * 1e9)], dtype=np.int64)
    eps = 1_000_000
    v_meas = ((gt.position(t_ns + eps) - gt.position(t_ns - eps)) / (2 * eps * 1e-9))[0]
    res = {}
    for name, lib in (("g", cuda), ("o", oracle)):
        est = _estimator(lib, sw, fixed=3)
        _add_init_problem(est, sw, v_t, v_meas)
        summ = est.Solve(50)
        res[name] = (summ, est.read_state())
        est.close()
    (sg, xg), (so, xo) = res["g"], res["o"]
    assert sg["iterations"] == so["iterations"] and sg["termination"] == so["termination"]
    assert so["final_cost"] < so["initial_cost"]
    assert abs(sg["final_cost"] - so["final_cost"]) <= 1e-7 * abs(so["final_cost"])
    assert np.abs(xg["knot_p"] - xo["knot_p"]).max() < 1e-7
    assert rot_err(xg["knot_q"], xo["knot_q"]).max() < 1e-7
