"""Pins the oracle with the reference test's own property: analytic Jacobian == auto-diff Jacobian (max-abs 1e-6)
on the hard-coded inputs of src/app/test_analytic_factor.cpp GenerateData (:903-1026), over a random trajectory
with the statistics of Trajectory(0.1)+genRandomTrajectory(8) (:61-62).  Auto-diff = torch-f64 autograd of
residuals written from the reference's auto_diff factors (tests/autodiff_ref.py).  CPU only."""
import numpy as np
import pytest

import autodiff_ref as AD
import oracle_lib
from clic_b200 import estimator as E
from clic_b200 import synthetic as S
from clic_b200._abi import GEO_LINE, GEO_PLANE

TOL = 1e-6  # test_analytic_factor.cpp:865


def reference_test_trajectory(seed):
    """Trajectory(0.1) seeds N=4 identity knots (trajectory.h:52-54), genRandomTrajectory(8) appends 8 random ones."""
    rng = np.random.Generator(np.random.PCG64(seed))
    q = np.concatenate([np.tile([0, 0, 0, 1.0], (4, 1)), S.so3_exp(rng.uniform(-np.pi, np.pi, size=(8, 3)))])
    p = np.concatenate([np.zeros((4, 3)), rng.uniform(-5, 5, size=(8, 3))])
    S_LtoI = S.so3_exp(rng.uniform(-np.pi, np.pi, size=3))
    S_CtoI = S.so3_exp(rng.uniform(-np.pi, np.pi, size=3))
    return E.Window(t0_ns=0, dt_ns=100_000_000, knot_q=q, knot_p=p, S_LtoI=S_LtoI, p_LinI=np.array([0.15, 0.52, 0.35]),
                    S_CtoI=S_CtoI, p_CinI=np.array([0.1, 0.2, 0.3]), t_offset_ns=np.array([1.0e6, 0.0, 1.0e6]),
                    gravity=np.array([0, 0, 9.8]), bias=np.array([[0.1, 0.2, 0.3, 0.1, 0.2, 0.3]]),
                    inv_depth=np.array([0.51]))


def win_dict(w):
    return dict(t0_ns=w.t0_ns, dt_ns=w.dt_ns, knot_q=w.knot_q, knot_p=w.knot_p)


def knot_cols(Jk, layout):
    """(m, K, 6) -> (m, 6*n_free) in canonical order."""
    return Jk[:, layout.first_free_knot:, :].reshape(Jk.shape[0], -1)


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("geo", [GEO_PLANE, GEO_LINE])
def test_loam_feature_factor(oracle, seed, geo):
    w = reference_test_trajectory(seed)
    est = E.TrajectoryEstimator(w, E.default_options(oracle), oracle)
    plane = np.array([1, 2, 3, 4.0]) / np.linalg.norm([1, 2, 3, 4.0])
    normal = np.array([3, 2, 1.0]) / np.linalg.norm([3, 2, 1.0])
    gpoint = np.array([1, 2, 3.0])
    point = np.array([1, 2, 3.0])
    ident, zero = np.array([0, 0, 0, 1.0]), np.zeros(3)
    est.AddLoamMeasurementAnalytic([0.55], [point], [geo], [plane], [normal], [gpoint], ident, zero, w.S_LtoI, w.p_LinI,
                                   1.3)
    L = est.layout()
    r, J = oracle_lib.block_jacobian(oracle, est, 0)
    fn = AD.lidar_residual(win_dict(w), int(0.55 * 1e9), point, geo, plane, normal, gpoint, ident, zero, w.S_LtoI,
                           w.p_LinI, 1.3)
    r_ad, Jk, _ = AD.jac(fn, w.n_knots, [])
    assert np.allclose(r, r_ad, atol=1e-9)
    assert np.abs(J - knot_cols(Jk, L)).max() < TOL
    # and b = J^T r, H = J^T J of the single factor
    H, b, cost, _ = est.Evaluate()
    assert np.allclose(H, J.T @ J, rtol=1e-12, atol=1e-12)
    assert np.allclose(b, J.T @ r, rtol=1e-12, atol=1e-12)
    assert np.isclose(cost, 0.5 * float(r @ r))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_imu_factor(oracle, seed):
    w = reference_test_trajectory(seed)
    opt = E.default_options(oracle)
    opt.lock_ab = opt.lock_wb = 0
    opt.lock_t_offset[0] = 0
    opt.t_offset_padding_ns = 100_000_000  # t_padding_ns_ = 1e8 (:907)
    est = E.TrajectoryEstimator(w, opt, oracle)
    gyro, accel = np.array([0.1, 0.2, 0.3]), np.array([0.1, 0.2, -9.8])
    info = np.full(6, 3.4)
    assert est.AddIMUMeasurementAnalytic([0.55], [gyro], [accel], 0, info) == 0
    L = est.layout()
    r, J = oracle_lib.block_jacobian(oracle, est, 0)
    fn = AD.imu_residual(win_dict(w), int(0.55 * 1e9), gyro, accel, info)
    r_ad, Jk, Jx = AD.jac(fn, w.n_knots, [w.bias[0, :3], w.bias[0, 3:], w.gravity, [w.t_offset_ns[0]]])
    assert np.allclose(r, r_ad, atol=1e-8)
    assert np.abs(J[:, :6 * L.n_free_knots] - knot_cols(Jk, L)).max() < TOL
    assert np.abs(J[:, L.col_bias_gyro:L.col_bias_gyro + 3] - Jx[0]).max() < TOL
    assert np.abs(J[:, L.col_bias_accel:L.col_bias_accel + 3] - Jx[1]).max() < TOL
    assert np.abs(J[:, L.col_t_offset[0]] - Jx[3][:, 0]).max() < TOL


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("case", [0, 1])
def test_image_feature_factor(oracle, seed, case):
    w = reference_test_trajectory(seed)
    ti, tj = [(0.13, 0.34), (0.11, 0.63)][case]  # :936-944
    pi, pj = [(np.array([1, 2, 1.0]), np.array([4, 5, 1.0])), (np.array([4, 5, 1.0]), np.array([1, 2, 1.0]))][case]
    opt = E.default_options(oracle)
    opt.lock_t_offset[2] = 0
    opt.t_offset_padding_ns = 100_000_000
    est = E.TrajectoryEstimator(w, opt, oracle)
    assert est.AddImageFeatureAnalytic([ti], [pi], [tj], [pj], [0], fixed_depth=False) == 0
    L = est.layout()
    r, J = oracle_lib.block_jacobian(oracle, est, 0)
    fn = AD.image_residual(win_dict(w), int(ti * 1e9), pi, int(tj * 1e9), pj, w.S_CtoI, w.p_CinI)
    r_ad, Jk, Jx = AD.jac(fn, w.n_knots, [[w.inv_depth[0]], [w.t_offset_ns[2]]])
    scale = max(1.0, np.abs(Jk).max())
    assert np.allclose(r, r_ad, rtol=1e-9, atol=1e-7)
    assert np.abs(J[:, :6 * L.n_free_knots] - knot_cols(Jk, L)).max() < TOL * scale
    assert np.abs(J[:, L.col_inv_depth] - Jx[0][:, 0]).max() < TOL * scale
    assert np.abs(J[:, L.col_t_offset[2]] - Jx[1][:, 0]).max() < TOL * scale


@pytest.mark.parametrize("weight", [200.0, 450.0])
def test_image_feature_weight_from_config(oracle, weight):
    """ImageFeatureFactor::sqrt_info is not the class default in any shipped run: TrajectoryManager::InitFactorInfo sets
    it to image_weight * I2 (trajectory_manager.cpp:53-61; config/ct_odometry_ntu.yaml:35 = 200, ct_odometry_lvi.yaml:35
    = 450).  The option that carries it is pinned here against the independent auto-diff lineage, which takes the
    weight as a plain argument (round 1 shared one literal between oracle and product, so a wrong weight was invisible)."""
    w = reference_test_trajectory(2)
    ti, tj = 0.13, 0.34
    pi, pj = np.array([1, 2, 1.0]), np.array([4, 5, 1.0])
    opt = E.default_options(oracle)
    assert abs(opt.image_sqrt_info - 450.0 / 1.5) < 1e-12          # the class default (image_feature_factor.h:277-278)
    opt.image_sqrt_info = weight
    est = E.TrajectoryEstimator(w, opt, oracle)
    assert est.AddImageFeatureAnalytic([ti], [pi], [tj], [pj], [0], fixed_depth=True) == 0
    L = est.layout()
    r, J = oracle_lib.block_jacobian(oracle, est, 0)
    fn = AD.image_residual(win_dict(w), int(ti * 1e9), pi, int(tj * 1e9), pj, w.S_CtoI, w.p_CinI, sqrt_info=weight)
    r_ad, Jk, _ = AD.jac(fn, w.n_knots, [[w.inv_depth[0]], [w.t_offset_ns[2]]])
    assert np.allclose(r, r_ad, rtol=1e-9, atol=1e-7)
    assert np.abs(J[:, :6 * L.n_free_knots] - knot_cols(Jk, L)).max() < TOL * max(1.0, np.abs(Jk).max())
    # and it is not the default: the residual scales with the weight
    est0 = E.TrajectoryEstimator(w, E.default_options(oracle), oracle)
    est0.AddImageFeatureAnalytic([ti], [pi], [tj], [pj], [0], fixed_depth=True)
    r0, _ = oracle_lib.block_jacobian(oracle, est0, 0)
    assert np.allclose(r / weight, r0 / (450.0 / 1.5), rtol=1e-12)


def test_synthetic_window_factors(oracle):
    """Same property on a realistic (small-rotation-increment) window, every factor type, a few factors each."""
    sw = S.make_window(10, n_lidar=12, n_imu=6, n_landmarks=3, obs_per_landmark=3, line_fraction=0.5, seed=7)
    w = sw.window
    opt = E.default_options(oracle)
    opt.lock_ab = opt.lock_wb = 0
    est = E.TrajectoryEstimator(w, opt, oracle)
    S.add_all_factors(est, sw, bias_factor=False)
    L = est.layout()
    wd = win_dict(w)
    blk = 0
    Ld = sw.lidar
    for i in range(len(Ld["t_point"])):
        r, J = oracle_lib.block_jacobian(oracle, est, blk)
        fn = AD.lidar_residual(wd, int(Ld["t_point"][i] * 1e9), Ld["point"][i], int(Ld["geo_type"][i]),
                               Ld["geo_plane"][i], Ld["geo_normal"][i], Ld["geo_point"][i], Ld["S_GtoM"], Ld["p_GinM"],
                               Ld["S_LtoI"], Ld["p_LinI"], Ld["weight"])
        r_ad, Jk, _ = AD.jac(fn, w.n_knots, [])
        assert np.allclose(r, r_ad, atol=1e-8)
        assert np.abs(J[:, :6 * L.n_free_knots] - knot_cols(Jk, L)).max() < TOL * max(1, np.abs(J).max())
        blk += 1
    I = sw.image
    for i in range(len(I["t_i"])):
        r, J = oracle_lib.block_jacobian(oracle, est, blk)
        fn = AD.image_residual(wd, int(I["t_i"][i] * 1e9), I["p_i"][i], int(I["t_j"][i] * 1e9), I["p_j"][i], w.S_CtoI,
                               w.p_CinI)
        r_ad, Jk, _ = AD.jac(fn, w.n_knots, [[w.inv_depth[I["landmark"][i]]], [0.0]])
        assert np.allclose(r, r_ad, rtol=1e-8, atol=1e-6)
        assert np.abs(J[:, :6 * L.n_free_knots] - knot_cols(Jk, L)).max() < TOL * max(1, np.abs(J).max())
        blk += 1
    M = sw.imu
    for i in range(len(M["timestamp"])):
        r, J = oracle_lib.block_jacobian(oracle, est, blk)
        fn = AD.imu_residual(wd, int(M["timestamp"][i] * 1e9), M["gyro"][i], M["accel"][i], M["info_vec"])
        slot = M["bias_slot"]
        r_ad, Jk, Jx = AD.jac(fn, w.n_knots, [w.bias[slot, :3], w.bias[slot, 3:], w.gravity, [0.0]])
        sc = max(1, np.abs(J).max())
        assert np.allclose(r, r_ad, rtol=1e-9, atol=1e-6)
        assert np.abs(J[:, :6 * L.n_free_knots] - knot_cols(Jk, L)).max() < TOL * sc
        c = L.col_bias_gyro + 3 * slot
        assert np.abs(J[:, c:c + 3] - Jx[0]).max() < TOL * sc
        blk += 1


# ---- the remaining adders of TrajectoryEstimator (fixtures of test_analytic_factor.cpp:903-1026) ----
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_imu_pose_factor(oracle, seed):
    """IMUPoseFactor (test_analytic_factor.cpp:1325-1333: pose at t = 0.55, w_pose 2.3).  Knot blocks against auto-diff;
    the reference's time-offset column is built from a quaternion made out of a SKEW matrix
    (trajectory_value_factor.h:419-423) and is not the derivative of anything — restated as written, not compared."""
    w = reference_test_trajectory(seed)
    rng = np.random.default_rng(seed)
    q_meas = S.so3_exp(rng.uniform(-np.pi, np.pi, 3))
    p_meas = rng.uniform(-5, 5, 3)
    info = np.array([2.3, 2.3, 2.3, 1.7, 1.7, 1.7])
    est = E.TrajectoryEstimator(w, E.default_options(oracle), oracle)
    assert est.AddPoseMeasurementAnalytic([0.55], [p_meas], [q_meas], info) == 0
    L = est.layout()
    r, J = oracle_lib.block_jacobian(oracle, est, 0)
    fn = AD.pose_residual(win_dict(w), int(0.55 * 1e9), p_meas, q_meas, info)
    r_ad, Jk, _ = AD.jac(fn, w.n_knots, [[w.t_offset_ns[0]]])
    assert np.allclose(r, r_ad, atol=1e-9)
    assert np.abs(J[:, :6 * L.n_free_knots] - knot_cols(Jk, L)).max() < TOL * max(1.0, np.abs(Jk).max())


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_local_velocity_factor(oracle, seed):
    """LocalVelocityFactor (test_analytic_factor.cpp:1224-1232: t = 0.23, v = (0.2, 0.4, 0.5)), time offset unlocked."""
    w = reference_test_trajectory(seed)
    opt = E.default_options(oracle)
    opt.lock_t_offset[0] = 0
    opt.t_offset_padding_ns = 100_000_000
    est = E.TrajectoryEstimator(w, opt, oracle)
    v = np.array([0.2, 0.4, 0.5])
    assert est.AddLocalVelocityMeasurementAnalytic([0.23], [v], 1.9) == 0
    L = est.layout()
    r, J = oracle_lib.block_jacobian(oracle, est, 0)
    fn = AD.local_velocity_residual(win_dict(w), int(0.23 * 1e9), v, 1.9)
    r_ad, Jk, Jx = AD.jac(fn, w.n_knots, [[w.t_offset_ns[0]]])
    sc = max(1.0, np.abs(Jk).max())
    assert np.allclose(r, r_ad, atol=1e-8)
    assert np.abs(J[:, :6 * L.n_free_knots] - knot_cols(Jk, L)).max() < TOL * sc
    assert np.abs(J[:, L.col_t_offset[0]] - Jx[0][:, 0]).max() < TOL * sc


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("case", [0, 1])
def test_relative_orientation_factor(oracle, seed, case):
    """RelativeOrientationFactor (test_analytic_factor.cpp:1192-1202: (ta, tb) = (0.11, 0.33), (0.21, 0.83),
    sqrt_info (0.33, 0.27, 0.11))."""
    w = reference_test_trajectory(seed)
    ta, tb = [(0.11, 0.33), (0.21, 0.83)][case]
    rng = np.random.default_rng(10 + seed)
    S_BtoA = S.so3_exp(rng.uniform(-np.pi, np.pi, 3))
    info = np.array([0.33, 0.27, 0.11])
    est = E.TrajectoryEstimator(w, E.default_options(oracle), oracle)
    assert est.AddRelativeRotationAnalytic(ta, tb, S_BtoA, info)
    L = est.layout()
    r, J = oracle_lib.block_jacobian(oracle, est, 0)
    fn = AD.relative_rotation_residual(win_dict(w), int(ta * 1e9), int(tb * 1e9), S_BtoA, info)
    r_ad, Jk, _ = AD.jac(fn, w.n_knots, [])
    assert np.allclose(r, r_ad, atol=1e-9)
    assert np.abs(J[:, :6 * L.n_free_knots] - knot_cols(Jk, L)).max() < TOL * max(1.0, np.abs(Jk).max())


@pytest.mark.parametrize("case", [0, 1])
def test_image_depth_factor(oracle, case):
    """ImageDepthFactor (test_analytic_factor.cpp:1256-1264: p_i / p_j of the image cases, depth_inv 0.51)."""
    w = reference_test_trajectory(1)
    pi, pj = [(np.array([1, 2, 1.0]), np.array([4, 5, 1.0])), (np.array([4, 5, 1.0]), np.array([1, 2, 1.0]))][case]
    rng = np.random.default_rng(4 + case)
    S_CitoCj = S.so3_exp(rng.uniform(-0.4, 0.4, 3))
    p_CiinCj = rng.uniform(-0.5, 0.5, 3)
    est = E.TrajectoryEstimator(w, E.default_options(oracle), oracle)
    est.AddImageDepthAnalytic(pi, pj, S_CitoCj, p_CiinCj, 0)
    L = est.layout()
    assert L.n_free_landmarks == 1
    r, J = oracle_lib.block_jacobian(oracle, est, 0)
    fn = AD.image_depth_residual(pi, pj, S_CitoCj, p_CiinCj, 450.0 / 1.5)
    r_ad, _, Jx = AD.jac(fn, w.n_knots, [[w.inv_depth[0]]])
    assert np.allclose(r, r_ad, rtol=1e-10, atol=1e-7)
    assert np.abs(J[:, L.col_inv_depth] - Jx[0][:, 0]).max() < TOL * max(1.0, np.abs(Jx[0]).max())
