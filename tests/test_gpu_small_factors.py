"""The remaining adders of TrajectoryEstimator — IMUPoseFactor, LocalVelocityFactor,
RelativeOrientationFactor, ImageDepthFactor — through the C ABI on the GPU against the oracle: H, b, cost <= 1e-9, solved
poses <= 1e-7, with and without the streaming factors around them, time offset locked and free, both pass paths.
(The oracle's restatement of each is pinned to auto-diff on the reference test's fixtures, tests/test_oracle_autodiff.py;
the one-thread device evaluators to the oracle on the CPU, tests/test_device_math_host.py.)"""
import os

import numpy as np
import pytest

from clic_b200 import estimator as E
from clic_b200 import synthetic as S
from test_gpu_parity import REL, rel, rel_scaled
from test_gpu_solve import rot_err

pytestmark = pytest.mark.gpu


def add_small(est, sw, rng, depth=True):
    w = sw.window
    hi = (w.n_knots - 3) * w.dt_ns * 1e-9
    t = np.sort(rng.uniform(0.02, hi - 0.02, 5))
    sp = S.SplineNP(w.t0_ns, w.dt_ns, sw.gt_knot_q, sw.gt_knot_p)
    tn = (t * 1e9).astype(np.int64)
    q = np.array([S.qmul(qq, S.so3_exp(rng.normal(0, 0.01, 3))) for qq in sp.rotation(tn)])
    p = sp.position(tn) + rng.normal(0, 0.02, (len(t), 3))
    dropped = est.AddPoseMeasurementAnalytic(np.r_[t, hi + 1.0], np.r_[p, [[0, 0, 0]]], np.r_[q, [[0, 0, 0, 1.0]]],
                                             [30.0, 30.0, 30.0, 10.0, 10.0, 10.0])
    dropped += est.AddLocalVelocityMeasurementAnalytic(t[:3] + 0.013, rng.normal(0, 1.0, (3, 3)), 4.0)
    ok = est.AddRelativeRotationAnalytic(t[0], t[3], S.so3_exp(rng.normal(0, 0.3, 3)), [20.0, 25.0, 30.0])
    ok2 = est.AddRelativeRotationAnalytic(t[1], t[1] + 0.004, S.so3_exp(rng.normal(0, 0.01, 3)), [20.0, 25.0, 30.0])
    out = est.AddRelativeRotationAnalytic(-1.0, t[1], [0, 0, 0, 1.0], [1.0, 1.0, 1.0])
    if depth and len(w.inv_depth):
        for lm in (0, 1):
            est.AddImageDepthAnalytic([0.1 * lm, 0.05, 1.0], [0.12 * lm + 0.03, 0.02, 1.0], S.so3_exp(rng.normal(0, 0.05, 3)),
                                      rng.normal(0, 0.2, 3), lm)
    return dropped, ok, ok2, out


def build(lib, sw, streams=True, unlock_toff=False, depth=True, fixed=1):
    opt = E.default_options(lib)
    opt.lock_ab = opt.lock_wb = 0
    if unlock_toff:
        opt.lock_t_offset[0] = 0
        opt.t_offset_padding_ns = 10_000_000
    est = E.TrajectoryEstimator(sw.window, opt, lib)
    est.SetFixedIndex(fixed)
    if streams:
        S.add_all_factors(est, sw)
    flags = add_small(est, sw, np.random.default_rng(5), depth)
    return est, flags


def compare(cuda, oracle, sw, iters=6, **kw):
    eg, fg = build(cuda, sw, **kw)
    eo, fo = build(oracle, sw, **kw)
    assert fg == fo and fg[0] == 1 and fg[1] and fg[2] and not fg[3]
    assert eg.layout().D == eo.layout().D
    Hg, bg, cg, f_g = eg.Evaluate()
    Ho, bo, co, f_o = eo.Evaluate()
    assert rel(Hg, Ho) < REL and rel_scaled(Hg, Ho) < REL, (rel(Hg, Ho), rel_scaled(Hg, Ho))
    assert rel(bg, bo) < REL, rel(bg, bo)
    assert abs(cg - co) <= REL * abs(co) and abs(f_g - f_o) <= REL * max(abs(f_o), 1.0)
    sg, so = eg.Solve(iters), eo.Solve(iters)
    xg, xo = eg.read_state(), eo.read_state()
    for k in ("iterations", "num_successful_steps", "num_unsuccessful_steps", "termination"):
        assert sg[k] == so[k], (k, sg, so)
    assert np.abs(xg["knot_p"] - xo["knot_p"]).max() < 1e-7
    assert rot_err(xg["knot_q"], xo["knot_q"]).max() < 1e-7
    if xo["inv_depth"].size:
        assert np.abs(xg["inv_depth"] - xo["inv_depth"]).max() < 1e-7
    assert abs(sg["final_cost"] - so["final_cost"]) <= 1e-7 * abs(so["final_cost"])
    return Hg


def test_small_factors_alone(cuda, oracle):
    """No streaming factor at all: pose / velocity / relative-rotation / depth factors carry the whole solve."""
    sw = S.make_window(10, n_lidar=0, n_imu=0, n_landmarks=3, obs_per_landmark=2, seed=31)
    sw.lidar = dict(sw.lidar)
    compare(cuda, oracle, sw, streams=False, fixed=-1)


def test_small_factors_with_streams_per_stream_path(cuda, oracle):
    H = compare(cuda, oracle, S.config(3, 0.05), fixed=1)
    assert H.shape[0] >= 60


def test_small_factors_time_offset_free(cuda, oracle):
    compare(cuda, oracle, S.config(2, 0.05), unlock_toff=True, depth=False)


def test_small_factors_inside_the_fused_pass(cuda, oracle):
    """Forced fused pass: the small factors are evaluated into the addend system before the pass kernel."""
    prev = os.environ.get("CLIC_FUSED")
    os.environ["CLIC_FUSED"] = "1"
    try:
        sw = S.config(3, 0.1)
        eg, _ = build(cuda, sw)
        assert eg.pass_info()["fused"]
        eg.close()
        compare(cuda, oracle, sw)
    finally:
        if prev is None:
            os.environ.pop("CLIC_FUSED", None)
        else:
            os.environ["CLIC_FUSED"] = prev


def test_marginalization_estimator_refuses_them(cuda):
    sw = S.config(1, 0.1)
    opt = E.default_options(cuda)
    opt.is_marg_state = 1
    est = E.TrajectoryEstimator(sw.window, opt, cuda)
    with pytest.raises(Exception):
        est.AddPoseMeasurementAnalytic([0.1], [[0, 0, 0]], [[0, 0, 0, 1.0]], np.ones(6))


def test_residual_summary_by_type(cuda, oracle):
    """clic_residual_summary (ResidualSummary, trajectory_estimator.h:42-98): per-type costs GPU vs oracle, and their sum
    = cost + fixed cost of the pass."""
    sw = S.config(3, 0.1)
    out = {}
    for name, lib in (("g", cuda), ("o", oracle)):
        est, _ = build(lib, sw, fixed=2)
        est.AddGravityFactor(sw.window.gravity + np.array([0.01, 0.0, -0.02]), [40.0, 40.0, 40.0])
        H, b, c, f = est.Evaluate()
        cost, cnt = est.residual_summary()
        out[name] = (cost, cnt, c + f)
        est.close()
    (cg, ng, tg), (co, no, to) = out["g"], out["o"]
    assert abs(cg.sum() - tg) <= 1e-9 * abs(tg) and abs(co.sum() - to) <= 1e-9 * abs(to)
    for t in (1, 2, 3, 4, 11, 14):
        assert co[t] > 0 and abs(cg[t] - co[t]) <= 1e-9 * abs(co[t]), (t, cg[t], co[t])
    assert ng[1] > 0 and ng[4] > 0 and ng[11] > 0 and ng[2] == no[2] == 1 and ng[3] == no[3] == 1 and ng[14] == no[14]
