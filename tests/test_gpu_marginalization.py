"""Marginalization parity (BASELINE config 4): a 40-knot window with a previous prior, Schur complement of the 20
oldest knots (+ last bias, gravity, time offset) as TrajectoryManager::UpdateLIOPrior does
(trajectory_manager.cpp:319-465).  Compared: m / n, the kept-block list and A = J0^T J0, g = J0^T r0 of the new prior
(J0 itself has sign / rotation freedom: SURVEY.md §7 hard part 5), then a solve that uses the new prior."""
import numpy as np
import pytest

from clic_b200 import estimator as E
from clic_b200 import synthetic as S
from clic_b200._abi import BLOCK_BIAS_ACCEL, BLOCK_BIAS_GYRO, BLOCK_KNOT_POS, BLOCK_KNOT_ROT
from test_gpu_parity import rel, rel_scaled
from test_gpu_solve import rot_err

pytestmark = pytest.mark.gpu


def previous_prior(lib, sw, knots, rng):
    kinds, ids, x0 = [], [], []
    for k in knots:
        kinds += [BLOCK_KNOT_ROT, BLOCK_KNOT_POS]
        ids += [k, k]
        x0 += [S.qmul(sw.window.knot_q[k], S.so3_exp(rng.normal(0, 2e-3, 3))),
               np.concatenate([sw.window.knot_p[k] + rng.normal(0, 2e-3, 3), [0]])]
    kinds += [BLOCK_BIAS_GYRO, BLOCK_BIAS_ACCEL]
    ids += [0, 0]
    x0 += [np.concatenate([sw.window.bias[0, :3] + 1e-4, [0]]), np.concatenate([sw.window.bias[0, 3:] - 1e-4, [0]])]
    n = 6 * len(knots) + 6
    J0 = 40.0 * np.eye(n) + rng.normal(size=(n, n)) * 2.0
    r0 = rng.normal(size=n) * 0.3
    return E.Prior.create(lib, J0, r0, kinds, ids, np.array(x0)), kinds, ids


def marg_estimator(lib, sw, prior, kinds, ids, later=20):
    opt = E.default_options(lib)
    opt.is_marg_state = 1
    opt.marg_bias_param = 0          # UpdateLIOPrior, trajectory_manager.cpp:326-333
    opt.marg_gravity_param = 1
    opt.marg_t_offset_param = 1
    opt.ctrl_to_be_opt_now = 0
    opt.ctrl_to_be_opt_later = later
    est = E.TrajectoryEstimator(sw.window, opt, lib)
    drop = [i for i, (kd, idv) in enumerate(zip(kinds, ids))
            if (kd in (BLOCK_KNOT_ROT, BLOCK_KNOT_POS) and idv < later) or kd in (BLOCK_BIAS_GYRO, BLOCK_BIAS_ACCEL)]
    est.AddMarginalizationFactor(prior, drop)
    L = sw.lidar
    est.AddLoamMeasurementAnalytic(L["t_point"], L["point"], L["geo_type"], L["geo_plane"], L["geo_normal"],
                                   L["geo_point"], L["S_GtoM"], L["p_GinM"], L["S_LtoI"], L["p_LinI"], L["weight"], True)
    M = sw.imu
    est.AddIMUMeasurementAnalytic(M["timestamp"], M["gyro"], M["accel"], 1, M["info_vec"], True)
    return est


def make(scale=0.25):
    return S.make_window(40, n_lidar=int(20000 * scale), n_imu=int(700 * max(scale, 0.5)), seed=20230219,
                         init_noise=(1e-3, 2e-3), line_fraction=0.2, name="C4")


def test_c4_schur_of_20_knots(cuda, oracle):
    sw = make()
    out = {}
    for name, lib in (("g", cuda), ("o", oracle)):
        rng = np.random.default_rng(7)
        p0, kinds, ids = previous_prior(lib, sw, range(0, 24), rng)
        est = marg_estimator(lib, sw, p0, kinds, ids)
        pr = est.SaveMarginalizationInfo()
        assert pr is not None
        out[name] = (pr.dims(), pr.read(), pr.normal())
    (dg, rg, (Ag, gg)), (do, ro, (Ao, go)) = out["g"], out["o"]
    assert dg == do, (dg, do)                      # n, n_blocks, m
    n, nb, m = do
    assert m >= 20 * 6 and n >= 6
    for k in ("kind", "id", "offset"):
        assert np.array_equal(rg[k], ro[k]), k       # same kept blocks in the same order
    assert np.allclose(rg["x0"], ro["x0"], rtol=0, atol=0)
    # Schur complement: conditioning-limited agreement (cancellation A_rr - A_rm A_mm^-1 A_mr)
    assert rel_scaled(Ag, Ao) < 1e-7, rel_scaled(Ag, Ao)
    assert rel(Ag, Ao) < 1e-8, rel(Ag, Ao)
    assert np.abs(gg - go).max() / np.abs(go).max() < 1e-7
    # prior energy at the linearisation point: 1/2 r0^T r0
    eg, eo = 0.5 * rg["r0"] @ rg["r0"], 0.5 * ro["r0"] @ ro["r0"]
    assert abs(eg - eo) <= 1e-7 * abs(eo)


def test_c4_fixed_lag_step_solve_with_new_prior(cuda, oracle):
    """One full fixed-lag step: marginalise, then solve the next window using the new prior (knots <= 19 fixed)."""
    sw = make(0.1)
    res = {}
    for name, lib in (("g", cuda), ("o", oracle)):
        rng = np.random.default_rng(9)
        p0, kinds, ids = previous_prior(lib, sw, range(0, 24), rng)
        pr = marg_estimator(lib, sw, p0, kinds, ids).SaveMarginalizationInfo()
        opt = E.default_options(lib)
        opt.lock_ab = opt.lock_wb = 0
        est = E.TrajectoryEstimator(sw.window, opt, lib)
        est.SetFixedIndex(19)                      # lidar_prior_ctrl_id.first - 1 (trajectory_manager.cpp:653)
        est.AddMarginalizationFactor(pr)
        S.add_all_factors(est, sw)
        res[name] = (est.Solve(6), est.read_state())
    (sg, xg), (so, xo) = res["g"], res["o"]
    for k in ("iterations", "num_successful_steps", "termination"):
        assert sg[k] == so[k], (sg, so)
    assert np.abs(xg["knot_p"] - xo["knot_p"]).max() < 1e-7
    assert rot_err(xg["knot_q"], xo["knot_q"]).max() < 1e-7


def test_nothing_to_keep(cuda, oracle):
    """Everything dropped -> SaveMarginalizationInfo yields nullptr (trajectory_estimator.cpp:241-248)."""
    sw = S.make_window(10, n_lidar=200, n_imu=0, seed=3, init_noise=(1e-4, 1e-4))
    for lib in (cuda, oracle):
        opt = E.default_options(lib)
        opt.is_marg_state = 1
        opt.ctrl_to_be_opt_later = 100
        est = E.TrajectoryEstimator(sw.window, opt, lib)
        L = sw.lidar
        est.AddLoamMeasurementAnalytic(L["t_point"], L["point"], L["geo_type"], L["geo_plane"], L["geo_normal"],
                                       L["geo_point"], L["S_GtoM"], L["p_GinM"], L["S_LtoI"], L["p_LinI"], L["weight"],
                                       True)
        assert est.SaveMarginalizationInfo() is None
