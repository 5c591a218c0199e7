"""LM solve parity: after the same number of iterations the knots solved on the GPU agree with the oracle's within
1e-7 m and 1e-7 rad (BASELINE.json north_star), and both follow the same accept/reject/termination path."""
import numpy as np
import pytest

from clic_b200 import estimator as E
from clic_b200 import synthetic as S
from test_gpu_parity import build

pytestmark = pytest.mark.gpu

POS_TOL, ROT_TOL = 1e-7, 1e-7


def rot_err(qa, qb):
    """Angle of qa^-1 qb per knot [rad]."""
    d = S.qmul(S.qconj(qa), qb)
    return 2.0 * np.arctan2(np.linalg.norm(d[:, :3], axis=1), np.abs(d[:, 3]))


def solve_both(cuda, oracle, sw, iters, **kw):
    out = {}
    for name, lib in (("g", cuda), ("o", oracle)):
        est, _ = build(lib, sw, **kw)
        out[name] = (est.Solve(iters), est.read_state())
    (sg, xg), (so, xo) = out["g"], out["o"]
    for k in ("iterations", "num_successful_steps", "num_unsuccessful_steps", "termination", "termination_reason"):
        assert sg[k] == so[k], (k, sg, so)
    assert abs(sg["initial_cost"] - so["initial_cost"]) <= 1e-9 * abs(so["initial_cost"])
    assert abs(sg["final_cost"] - so["final_cost"]) <= 1e-7 * abs(so["final_cost"])
    assert np.abs(xg["knot_p"] - xo["knot_p"]).max() < POS_TOL
    assert rot_err(xg["knot_q"], xo["knot_q"]).max() < ROT_TOL
    if xo["bias"].size:
        assert np.abs(xg["bias"] - xo["bias"]).max() < 1e-7
    assert np.abs(xg["t_offset_ns"] - xo["t_offset_ns"]).max() < 1e-1  # ns
    return sg, xg


def test_c1_solve(cuda, oracle):
    """BASELINE config 1: one analytic solve of the 10-knot LiDAR-inertial window."""
    sw = S.config(1)
    s, x = solve_both(cuda, oracle, sw, 8, bias_factor=False)
    assert s["final_cost"] < 0.05 * s["initial_cost"]
    assert np.abs(x["knot_p"][1:-1] - sw.gt_knot_p[1:-1]).max() < 0.02  # converged towards the ground truth


def test_c1_solve_fixed_and_bias(cuda, oracle):
    solve_both(cuda, oracle, S.config(1), 8, fixed=2, unlock_bias=True)


def test_c1_solve_50_iterations_converges(cuda, oracle):
    s, _ = solve_both(cuda, oracle, S.config(1), 50, bias_factor=False)
    assert s["termination"] == 0  # CONVERGENCE before the cap


def test_c2_solve_time_offset_bounded(cuda, oracle):
    """BASELINE config 2 at 1/10 size: Cauchy IMU + online LiDAR-IMU time offset (bounded parameter, projected line
    search path)."""
    solve_both(cuda, oracle, S.config(2, 0.1), 8, unlock_bias=True, unlock_toff=True)


def test_single_iteration_and_zero(cuda, oracle):
    sw = S.make_window(10, n_lidar=500, n_imu=60, line_fraction=0.3, seed=21)
    solve_both(cuda, oracle, sw, 1)
    solve_both(cuda, oracle, sw, 0)


def test_40_knot_solve(cuda, oracle):
    solve_both(cuda, oracle, S.config(4, 0.25), 6, unlock_bias=True)


def test_c3_solve(cuda, oracle):
    """BASELINE config 3 at 1/10 size: LiDAR-inertial-camera solve."""
    solve_both(cuda, oracle, S.config(3, 0.1), 8, unlock_bias=True)


def test_c3_solve_free_inverse_depths(cuda, oracle):
    """Config 3 with the landmarks' inverse depths optimised together with the trajectory (D = 72 + 20)."""
    sw = S.config(3, 0.1)
    rng = np.random.default_rng(12)
    sw.window.inv_depth = sw.window.inv_depth * (1.0 + rng.normal(0, 0.05, len(sw.window.inv_depth)))
    out = {}
    for name, lib in (("g", cuda), ("o", oracle)):
        est, _ = build(lib, sw, unlock_bias=True, fixed_depth=False)
        out[name] = (est.Solve(8), est.read_state())
    (sg, xg), (so, xo) = out["g"], out["o"]
    for k in ("iterations", "num_successful_steps", "num_unsuccessful_steps", "termination", "termination_reason"):
        assert sg[k] == so[k], (k, sg, so)
    assert abs(sg["final_cost"] - so["final_cost"]) <= 1e-7 * abs(so["final_cost"])
    assert np.abs(xg["knot_p"] - xo["knot_p"]).max() < POS_TOL
    assert rot_err(xg["knot_q"], xo["knot_q"]).max() < ROT_TOL
    assert np.abs(xg["inv_depth"] - xo["inv_depth"]).max() < 1e-7
    assert np.abs(xg["inv_depth"] - sw.window.inv_depth).max() > 1e-6       # they did move
