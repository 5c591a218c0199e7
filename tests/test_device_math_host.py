"""Host emulation of the product's __host__ __device__ factor math (clic_b200/csrc/factor_eval.cuh compiled with
g++) against the oracle's per-factor Jacobians.  Catches restructuring mistakes (hoisting, unit-axis Jr,
vector-first chains) on the CPU-only box; the real parity tests are the -m gpu ones through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from clic_b200 import estimator as E
from clic_b200 import synthetic as S
from clic_b200._abi import GEO_PLANE, dptr

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL_DIR = os.path.join(HERE, "host_emul")


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-C", EMUL_DIR], check=True, capture_output=True)
    lib = C.CDLL(os.path.join(EMUL_DIR, "_build", "libclic_emul.so"))
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.emul_loam.argtypes = [C.c_int, dp, dp, C.c_int64, C.c_int64, C.c_int64, dp, C.c_int, dp, dp, dp, dp, dp,
                              C.c_double, dp, dp, ip, dp]
    lib.emul_imu.argtypes = [C.c_int, dp, dp, C.c_int64, C.c_int64, C.c_int64, dp, dp, dp, dp, dp, dp, C.c_double,
                             C.c_int, dp, dp, ip]
    lib.emul_image.argtypes = [C.c_int, dp, dp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, dp, dp, C.c_double, dp, dp,
                               C.c_double, C.c_int, dp, dp, ip, ip]
    lib.emul_image_packed.argtypes = [C.c_int, dp, dp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, dp, dp, C.c_double, dp,
                                      dp, C.c_double, C.c_int, dp, dp]
    lib.emul_small.argtypes = [C.c_int, dp, dp, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int64, dp, C.c_double, C.c_int,
                               C.c_double, C.c_double, ip, C.c_int, ip, C.c_int, dp, dp, ip]
    return lib


def _windows():
    # realistic small-increment window, the reference test's large random rotations + identical leading knots
    from test_oracle_autodiff import reference_test_trajectory
    sw = S.make_window(10, n_lidar=40, n_imu=25, line_fraction=0.5, seed=11)
    yield "smooth", sw
    sw2 = S.make_window(12, n_lidar=40, n_imu=25, line_fraction=0.5, seed=12)
    w = reference_test_trajectory(5)
    sw2.window.knot_q, sw2.window.knot_p = w.knot_q, w.knot_p
    yield "random+identity", sw2


@pytest.mark.parametrize("name,sw", list(_windows()))
def test_loam_rows(oracle, emul, name, sw):
    w = sw.window
    est = E.TrajectoryEstimator(w, E.default_options(oracle), oracle)
    L = sw.lidar
    est.AddLoamMeasurementAnalytic(L["t_point"], L["point"], L["geo_type"], L["geo_plane"], L["geo_normal"],
                                   L["geo_point"], L["S_GtoM"], L["p_GinM"], L["S_LtoI"], L["p_LinI"], L["weight"])
    s_or, u_or = est.indices(0)
    kq, kp = np.ascontiguousarray(w.knot_q), np.ascontiguousarray(w.knot_p)
    worst = 0.0
    for i in range(len(L["t_point"])):
        r_o, J_o = oracle_lib.block_jacobian(oracle, est, i)
        t_ns = int(np.int64(L["t_point"][i] * 1e9))
        is_plane = int(L["geo_type"][i] == GEO_PLANE)
        geo = np.zeros(6)
        if is_plane:
            geo[:4] = L["geo_plane"][i]
        else:
            geo[:3], geo[3:] = L["geo_point"][i], L["geo_normal"][i]
        r, J, s, u = np.zeros(1), np.zeros(24), C.c_int(0), np.zeros(1)
        pL = np.ascontiguousarray(L["point"][i])
        rc = emul.emul_loam(w.n_knots, dptr(kq), dptr(kp), w.t0_ns, w.dt_ns, t_ns, dptr(pL), is_plane, dptr(geo),
                            dptr(L["S_GtoM"]), dptr(L["p_GinM"]), dptr(L["S_LtoI"]), dptr(L["p_LinI"]),
                            float(L["weight"]), dptr(r), dptr(J), C.byref(s), dptr(u))
        assert rc == 0
        assert s.value == s_or[i] and u[0] == u_or[i]  # indexing bit-exact
        J_ref = J_o[0, 6 * s.value:6 * s.value + 24]
        scale = max(1.0, np.abs(J_ref).max())
        assert abs(r[0] - r_o[0]) <= 1e-10 * max(1.0, abs(r_o[0]))
        worst = max(worst, np.abs(J - J_ref).max() / scale)
        # nothing outside the 4-knot support
        mask = np.ones(J_o.shape[1], bool)
        mask[6 * s.value:6 * s.value + 24] = False
        assert not J_o[0, mask].any()
    assert worst < 1e-11, worst


@pytest.mark.parametrize("marg", [0, 1])
@pytest.mark.parametrize("name,sw", list(_windows()))
def test_imu_rows(oracle, emul, name, sw, marg):
    w = sw.window
    w.t_offset_ns = np.array([250_000.7, 0, 0])
    opt = E.default_options(oracle)
    opt.lock_ab = opt.lock_wb = 0
    opt.lock_t_offset[0] = 0
    opt.t_offset_padding_ns = 1_000_000
    est = E.TrajectoryEstimator(w, opt, oracle)
    M = sw.imu
    # keep samples away from the ends so the padding test keeps them
    keep = (M["timestamp"] > 0.01) & (M["timestamp"] < (w.n_knots - 3) * 0.1 - 0.01)
    ts, gy, ac = M["timestamp"][keep], M["gyro"][keep], M["accel"][keep]
    assert est.AddIMUMeasurementAnalytic(ts, gy, ac, 1, M["info_vec"]) == 0
    Lay = est.layout()
    kq, kp = np.ascontiguousarray(w.knot_q), np.ascontiguousarray(w.knot_p)
    NC = 40 if marg else 32
    worst = 0.0
    for i in range(len(ts)):
        r_o, J_o = oracle_lib.block_jacobian(oracle, est, i, apply_loss=True)
        t_ns = int(np.int64(ts[i] * 1e9)) + int(w.t_offset_ns[0])
        rows, rho, s = np.zeros((6, NC)), np.zeros(1), C.c_int(0)
        g, a = np.ascontiguousarray(gy[i]), np.ascontiguousarray(ac[i])
        bg, ba = np.ascontiguousarray(w.bias[1, :3]), np.ascontiguousarray(w.bias[1, 3:])
        rc = emul.emul_imu(w.n_knots, dptr(kq), dptr(kp), w.t0_ns, w.dt_ns, t_ns, dptr(g), dptr(a),
                           dptr(M["info_vec"]), dptr(bg), dptr(ba), dptr(w.gravity), 10.0, marg, dptr(rows),
                           dptr(rho), C.byref(s))
        assert rc == 0
        sv = s.value
        c_r = 34 if marg else 31
        c_t = 33 if marg else 30
        scale = max(1.0, np.abs(J_o).max())
        d = [np.abs(rows[:, :24] - J_o[:, 6 * sv:6 * sv + 24]).max(),
             np.abs(rows[:, 24:27] - J_o[:, Lay.col_bias_gyro + 3:Lay.col_bias_gyro + 6]).max(),
             np.abs(rows[:, 27:30] - J_o[:, Lay.col_bias_accel + 3:Lay.col_bias_accel + 6]).max(),
             np.abs(rows[:, c_t] - J_o[:, Lay.col_t_offset[0]]).max()]
        worst = max(worst, max(d) / scale)
        assert np.allclose(rows[:, c_r], r_o, rtol=1e-10, atol=1e-9)
        assert not rows[:, c_r + 1:].any()
    assert worst < 1e-11, worst


@pytest.mark.parametrize("free_depth", [False, True])
@pytest.mark.parametrize("unlock_toff", [False, True])
def test_image_rows(oracle, emul, unlock_toff, free_depth):
    sw = S.make_window(10, n_lidar=0, n_imu=0, n_landmarks=12, obs_per_landmark=6, seed=31, time_margin_ns=3_000_000)
    w = sw.window
    if unlock_toff:
        w.t_offset_ns = np.array([0, 0, 120_000.4])
    opt = E.default_options(oracle)
    if unlock_toff:
        opt.lock_t_offset[2] = 0
        opt.t_offset_padding_ns = 1_000_000
    est = E.TrajectoryEstimator(w, opt, oracle)
    I = sw.image
    assert est.AddImageFeatureAnalytic(I["t_i"], I["p_i"], I["t_j"], I["p_j"], I["landmark"], not free_depth) == 0
    Lay = est.layout()
    kq, kp = np.ascontiguousarray(w.knot_q), np.ascontiguousarray(w.knot_p)
    worst = 0.0
    overlaps = 0
    for i in range(len(I["t_i"])):
        r_o, J_o = oracle_lib.block_jacobian(oracle, est, i, apply_loss=True)
        toff = int(w.t_offset_ns[2])
        ti = int(np.int64(I["t_i"][i] * 1e9)) + toff
        tj = int(np.int64(I["t_j"][i] * 1e9)) + toff
        rows, rho, si, sj = np.zeros((2, 56)), np.zeros(1), C.c_int(0), C.c_int(0)
        pi, pj = np.ascontiguousarray(I["p_i"][i]), np.ascontiguousarray(I["p_j"][i])
        rc = emul.emul_image(w.n_knots, dptr(kq), dptr(kp), w.t0_ns, w.dt_ns, ti, tj, dptr(pi), dptr(pj),
                             float(w.inv_depth[I["landmark"][i]]), dptr(w.S_CtoI), dptr(w.p_CinI), 2.0,
                             int(unlock_toff), dptr(rows), dptr(rho), C.byref(si), C.byref(sj))
        assert rc == 0
        J = np.zeros_like(J_o)
        J[:, 6 * si.value:6 * si.value + 24] += rows[:, :24]
        J[:, 6 * sj.value:6 * sj.value + 24] += rows[:, 24:48]   # overlapping windows add up
        if unlock_toff:
            J[:, Lay.col_t_offset[2]] += rows[:, 48]
        if free_depth:   # every landmark of this window is observed, so its rank among the free ones is its index
            J[:, Lay.col_inv_depth + int(I["landmark"][i])] += rows[:, 50]
        overlaps += abs(si.value - sj.value) < 4
        # the per-timestamp pose-pack path (packed visual kernel) emits the same rows
        rows_p, rho_p = np.zeros((2, 56)), np.zeros(1)
        rc = emul.emul_image_packed(w.n_knots, dptr(kq), dptr(kp), w.t0_ns, w.dt_ns, ti, tj, dptr(pi), dptr(pj),
                                    float(w.inv_depth[I["landmark"][i]]), dptr(w.S_CtoI), dptr(w.p_CinI), 2.0,
                                    int(unlock_toff), dptr(rows_p), dptr(rho_p))
        assert rc == 0
        assert np.abs(rows_p - rows).max() <= 1e-12 * max(1.0, np.abs(rows).max())
        assert abs(rho_p[0] - rho[0]) <= 1e-13 * max(1.0, abs(rho[0]))
        scale = max(1.0, np.abs(J_o).max())
        worst = max(worst, np.abs(J - J_o).max() / scale)
        assert np.allclose(rows[:, 49], r_o, rtol=1e-10, atol=1e-9)
    assert overlaps > 0
    assert worst < 1e-11, worst


# ---- the one-thread evaluators of the small factors vs the oracle's factor classes ----
def _small_case(oracle, emul, kind, seed, unlock_toff=False, fixed=1):
    from test_oracle_autodiff import reference_test_trajectory
    w = reference_test_trajectory(seed)
    rng = np.random.default_rng(100 + seed)
    opt = E.default_options(oracle)
    if unlock_toff:
        opt.lock_t_offset[0] = 0
        opt.t_offset_padding_ns = 50_000_000
    est = E.TrajectoryEstimator(w, opt, oracle)
    est.SetFixedIndex(fixed)
    d = np.zeros(16)
    t_a = t_b = 0
    loss_a, lm = 0.0, 0
    if kind == 0:
        t, p, q, info = 0.55, rng.uniform(-5, 5, 3), S.so3_exp(rng.uniform(-np.pi, np.pi, 3)), rng.uniform(0.5, 3, 6)
        assert est.AddPoseMeasurementAnalytic([t], [p], [q], info) == 0
        d[:3], d[3:7], d[7:13], t_a = p, q, info, int(t * 1e9)
    elif kind == 1:
        t, v, wt = 0.23, np.array([0.2, 0.4, 0.5]), 1.9
        assert est.AddLocalVelocityMeasurementAnalytic([t], [v], wt) == 0
        d[:3], d[3], t_a = v, wt, int(t * 1e9)
    elif kind == 2:
        ta, tb = [(0.11, 0.33), (0.21, 0.83), (0.52, 0.58)][seed % 3]
        Sq, info = S.so3_exp(rng.uniform(-np.pi, np.pi, 3)), np.array([0.33, 0.27, 0.11])
        assert est.AddRelativeRotationAnalytic(ta, tb, Sq, info)
        d[:4], d[4:7], t_a, t_b = Sq, info, int(ta * 1e9), int(tb * 1e9)
    else:
        pi, pj = np.array([1, 2, 1.0]), np.array([4, 5, 1.0])
        Sq, pc = S.so3_exp(rng.uniform(-0.4, 0.4, 3)), rng.uniform(-0.5, 0.5, 3)
        est.AddImageDepthAnalytic(pi, pj, Sq, pc, 0)
        d[:3], d[3:6], d[6:10], d[10:13], d[13], loss_a = pi, pj, Sq, pc, 450.0 / 1.5, 1.0
    L = est.layout()
    r_o, J_o = oracle_lib.block_jacobian(oracle, est, 0, apply_loss=True)
    knot_col = np.array([6 * (k - L.first_free_knot) if k >= L.first_free_knot else -1 for k in range(w.n_knots)], np.int32)
    lm_col = np.array([L.col_inv_depth if L.n_free_landmarks else -1], np.int32)
    r = np.zeros(6)
    J = np.zeros((6, L.D))
    nres = C.c_int(0)
    ip = C.POINTER(C.c_int)
    kq, kp = np.ascontiguousarray(w.knot_q), np.ascontiguousarray(w.knot_p)
    rc = emul.emul_small(w.n_knots, dptr(kq), dptr(kp), w.t0_ns, w.dt_ns, kind, t_a, t_b, dptr(d), loss_a, lm,
                         float(w.t_offset_ns[0]), float(w.inv_depth[0]), knot_col.ctypes.data_as(ip), int(L.col_t_offset[0]),
                         lm_col.ctypes.data_as(ip), L.D, dptr(r), dptr(J), C.byref(nres))
    assert rc == 0 and nres.value == len(r_o)
    if loss_a > 0:  # emul_small returns the uncorrected rows; the kernel applies the Corrector's sqrt(rho') afterwards
        sq = float(r[:nres.value] @ r[:nres.value])
        sc = np.sqrt(1.0 / (1.0 + sq / loss_a ** 2))
        r, J = r * sc, J * sc
    scale = max(1.0, np.abs(J_o).max())
    assert np.abs(r[:nres.value] - r_o).max() < 1e-11 * max(1.0, np.abs(r_o).max())
    assert np.abs(J[:nres.value] - J_o).max() < 1e-10 * scale, np.abs(J[:nres.value] - J_o).max()


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_small_factor_rows(oracle, emul, kind, seed):
    _small_case(oracle, emul, kind, seed)


@pytest.mark.parametrize("kind", [0, 1])
def test_small_factor_rows_time_offset_free(oracle, emul, kind):
    _small_case(oracle, emul, kind, 2, unlock_toff=True, fixed=-1)


# ---- every small analytic factor: the device evaluators compiled for the host against the oracle ----
def _factor_names():
    import factor_cases as FC
    from test_oracle_autodiff import reference_test_trajectory
    return [c["name"] for c in FC.cases(reference_test_trajectory(1))]


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("name", _factor_names())
def test_single_factor_evaluators(oracle, emul, name, seed):
    import factor_cases as FC
    from clic_b200._abi import FactorDesc
    from test_oracle_autodiff import reference_test_trajectory
    w = reference_test_trajectory(seed)
    est = E.TrajectoryEstimator(w, E.default_options(oracle), oracle)
    c = next(c for c in FC.cases(w, seed) if c["name"] == name)
    r_o, J_o = FC.evaluate(est, c)
    f = FactorDesc()
    f.kind, f.geo_type, f.t_a_ns, f.t_b_ns = int(c["kind"]), int(c["geo"]), int(c["t_a"]), int(c["t_b"])
    for i, v in enumerate(np.asarray(c["d"], np.float64).ravel()):
        f.d[i] = v
    for i, v in enumerate(np.asarray(c["x"], np.float64).ravel()):
        f.x[i] = v
    pre = FC.preintegration_struct(c["pre"]) if c["pre"] is not None else None
    if pre is not None:
        f.pre = C.pointer(pre)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    emul.emul_factor.argtypes = [C.c_int, dp, dp, C.c_int64, C.c_int64, C.POINTER(FactorDesc), dp, dp, C.c_int, dp, dp, ip]
    kq, kp = np.ascontiguousarray(w.knot_q), np.ascontiguousarray(w.knot_p)
    r = np.zeros(15)
    J = np.zeros((15, J_o.shape[1]))
    nres = C.c_int(0)
    rc = emul.emul_factor(w.n_knots, dptr(kq), dptr(kp), w.t0_ns, w.dt_ns, C.byref(f), dptr(np.ascontiguousarray(w.S_CtoI)),
                          dptr(np.ascontiguousarray(w.p_CinI)), J_o.shape[1], dptr(r), dptr(J), C.byref(nres))
    assert rc == 0 and nres.value == len(r_o)
    m = nres.value
    assert np.abs(r[:m] - r_o).max() <= 1e-12 * max(1.0, np.abs(r_o).max())
    assert np.abs(J[:m] - J_o).max() <= 1e-11 * max(1.0, np.abs(J_o).max()), np.abs(J[:m] - J_o).max()
