"""The fused pass (clic_b200/csrc/pass_fused.cuh: every factor stream, the record reduction, the assembly — and on
several GPUs the exchange — in one persistent launch) against the per-stream kernels of the same library
(CLIC_FUSED=0) and, through the other GPU test files, against the oracle (the fused pass is the default path).
The two paths partition the factors over warps differently, so they agree to rounding, not bitwise."""
import os

import numpy as np
import pytest

from clic_b200 import estimator as E
from clic_b200 import synthetic as S
from test_gpu_parity import build, rel, rel_scaled

pytestmark = pytest.mark.gpu


class pass_mode:
    """Estimators whose layout is built inside this block run the chosen pass: "0" one kernel per stream (round-1
    path), "1" the fused pass whenever applicable (by default the library picks by problem size)."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = os.environ.get("CLIC_FUSED")
        os.environ["CLIC_FUSED"] = self.mode

    def __exit__(self, *a):
        if self.prev is None:
            os.environ.pop("CLIC_FUSED", None)
        else:
            os.environ["CLIC_FUSED"] = self.prev


def per_stream_kernels():
    return pass_mode("0")


@pytest.fixture(autouse=True)
def force_fused():
    with pass_mode("1"):
        yield


def both_paths(cuda, sw, solve_iters=None, **kw):
    est, _ = build(cuda, sw, **kw)
    info = est.pass_info()
    assert info["fused"], info
    assert sum(info["part_ctas"]) <= info["grid"] <= 148 * 2
    out_f = est.Evaluate()
    sol_f = (est.Solve(solve_iters), est.read_state()) if solve_iters is not None else None
    est.close()
    with per_stream_kernels():
        est, _ = build(cuda, sw, **kw)
        assert not est.pass_info()["fused"]
        out_s = est.Evaluate()
        sol_s = (est.Solve(solve_iters), est.read_state()) if solve_iters is not None else None
        est.close()
    Hf, bf, cf, ff = out_f
    Hs, bs, cs, fs = out_s
    assert rel(Hf, Hs) < 1e-12 and rel_scaled(Hf, Hs) < 1e-11, (rel(Hf, Hs), rel_scaled(Hf, Hs))
    assert rel(bf, bs) < 1e-11, rel(bf, bs)
    assert abs(cf - cs) <= 1e-12 * abs(cs) and abs(ff - fs) <= 1e-12 * max(abs(fs), 1.0)
    assert np.array_equal(Hf, Hf.T)
    if solve_iters is not None:
        (sf, xf), (ss, xs) = sol_f, sol_s
        for k in ("iterations", "num_successful_steps", "num_unsuccessful_steps", "termination", "termination_reason"):
            assert sf[k] == ss[k], (k, sf, ss)
        assert np.abs(xf["knot_p"] - xs["knot_p"]).max() < 1e-9
        assert np.abs(xf["knot_q"] - xs["knot_q"]).max() < 1e-9
    return info


def test_c1_fused_equals_per_stream(cuda):
    both_paths(cuda, S.config(1), solve_iters=8, bias_factor=False)


def test_c3_all_streams_bias_and_fixed_knots(cuda):
    info = both_paths(cuda, S.config(3, 0.1), solve_iters=6, fixed=2, unlock_bias=True)
    p = info["part_ctas"]
    assert p[1] > 0 or p[0] > 0  # a LiDAR partition
    assert p[3] > 0 and p[4] > 0  # IMU and visual (pose packs) partitions


def test_c2_time_offset_bounded_line_search(cuda):
    """Residual-only passes (JAC = false instantiation) of the bounded problem's line search."""
    both_paths(cuda, S.config(2, 0.1), solve_iters=6, unlock_bias=True, unlock_toff=True)


def test_free_inverse_depths(cuda):
    sw = S.config(3, 0.05)
    est, _ = build(cuda, sw, unlock_bias=True, fixed_depth=False)
    assert est.pass_info()["fused"]
    Hf, bf, cf, _ = est.Evaluate()
    est.close()
    with per_stream_kernels():
        est, _ = build(cuda, sw, unlock_bias=True, fixed_depth=False)
        Hs, bs, cs, _ = est.Evaluate()
        est.close()
    assert rel_scaled(Hf, Hs) < 1e-10 and rel(bf, bs) < 1e-10 and abs(cf - cs) <= 1e-12 * abs(cs)


def test_tiny_and_ragged_windows(cuda):
    for n_l, n_i, seed in ((1, 0, 3), (33, 7, 4), (0, 40, 5), (700, 130, 6)):
        sw = S.make_window(10, n_lidar=n_l, n_imu=n_i, line_fraction=0.4, seed=seed)
        both_paths(cuda, sw, unlock_bias=n_i > 0)


def test_fused_pass_is_deterministic(cuda):
    sw = S.config(3, 0.1)
    runs = []
    for _ in range(3):
        est, _ = build(cuda, sw, unlock_bias=True)
        runs.append(est.Evaluate())
        est.close()
    for H, b, c, f in runs[1:]:
        assert np.array_equal(H, runs[0][0]) and np.array_equal(b, runs[0][1]) and c == runs[0][2] and f == runs[0][3]


def test_free_inverse_depth_system_is_deterministic(cuda):
    """The inverse-depth columns are gathered per landmark in a fixed order (rho_side_kernel), not scattered with
    atomics: rebuilt estimators and repeated passes give bit-identical systems on both layouts."""
    sw = S.config(3, 0.5)
    for ctx in (None, per_stream_kernels):
        runs = []
        for _ in range(3):
            if ctx is None:
                est, _ = build(cuda, sw, unlock_bias=True, fixed_depth=False, unlock_cam_toff=True)
                runs += [est.Evaluate(), est.Evaluate()]
                est.close()
            else:
                with ctx():
                    est, _ = build(cuda, sw, unlock_bias=True, fixed_depth=False, unlock_cam_toff=True)
                    runs += [est.Evaluate(), est.Evaluate()]
                    est.close()
        assert runs[0][0].shape[0] > 6 * sw.window.knot_q.shape[0] + 90  # the landmark columns are in the system
        for H, b, c, f in runs[1:]:
            assert np.array_equal(H, runs[0][0]) and np.array_equal(b, runs[0][1]) and c == runs[0][2]


def test_two_batches_of_one_kind_fall_back(cuda, oracle):
    """More than one batch per stream kind is not fused; both layouts give the same system."""
    sw = S.config(1)
    opt = E.default_options(cuda)
    est = E.TrajectoryEstimator(sw.window, opt, cuda)
    est.SetFixedIndex(-1)
    S.add_all_factors(est, sw, bias_factor=False)
    S.add_all_factors(est, sw, bias_factor=False)  # the same measurements twice: every sum doubles
    assert not est.pass_info()["fused"]
    H2, b2, c2, _ = est.Evaluate()
    est.close()
    est, _ = build(cuda, sw, bias_factor=False)
    H1, b1, c1, _ = est.Evaluate()
    est.close()
    assert rel(H2, 2 * H1) < 1e-12 and rel(b2, 2 * b1) < 1e-11 and abs(c2 - 2 * c1) < 1e-12 * abs(c1)


def test_stamps_cover_every_partition(cuda):
    sw = S.config(3, 0.1)
    est, _ = build(cuda, sw, unlock_bias=True)
    info = est.pass_info()
    est.pass_stamps(True)
    est.Evaluate()
    st = est.pass_stamps(False)
    est.close()
    assert st is not None and st.shape == (info["grid"], 8)
    assert (st[:, 1] >= st[:, 0]).all() and (st[:, 2] >= st[:, 1]).all() and (st[:, 5] >= st[:, 4]).all()
    assert (st[:, 4] >= st[:, 2]).all()


def test_automatic_choice_by_problem_size(cuda):
    """Unset CLIC_FUSED: small (production-size) windows keep the per-stream launches, large ones take the fused pass."""
    prev = os.environ.pop("CLIC_FUSED", None)
    try:
        est, _ = build(cuda, S.config(1), bias_factor=False)
        small = est.pass_info()["fused"]
        est.close()
        est, _ = build(cuda, S.config(3), unlock_bias=True)
        large = est.pass_info()["fused"]
        est.close()
    finally:
        if prev is not None:
            os.environ["CLIC_FUSED"] = prev
    assert not small and large


def test_fused_pass_against_the_oracle_on_small_windows(cuda, oracle):
    """The other GPU test files reach the fused pass only at BASELINE's full sizes (the automatic choice); here it is
    forced on the small and ragged windows too and held to the same bars against the oracle."""
    from test_gpu_parity import compare_eval
    from test_gpu_solve import solve_both
    compare_eval(cuda, oracle, S.config(1), bias_factor=False)
    compare_eval(cuda, oracle, S.config(1), fixed=3, unlock_bias=True)
    compare_eval(cuda, oracle, S.config(2, 0.1), unlock_bias=True, unlock_toff=True)
    compare_eval(cuda, oracle, S.config(3, 0.1), unlock_bias=True, unlock_cam_toff=True)
    compare_eval(cuda, oracle, S.make_window(10, n_lidar=3000, n_imu=0, line_fraction=1.0, seed=5))
    solve_both(cuda, oracle, S.config(1), 8, fixed=2, unlock_bias=True)
    solve_both(cuda, oracle, S.config(3, 0.1), 8, unlock_bias=True)
    solve_both(cuda, oracle, S.config(2, 0.1), 8, unlock_bias=True, unlock_toff=True)


def test_unsorted_large_streams_in_the_fused_pass(cuda, oracle):
    """More than 32 k time-UNSORTED LiDAR and IMU records per batch go to the kernels as they are: every warp meets many
    intervals per slab, runs out of record slots and takes the atomics fallback — for the IMU stream with the expansion of
    its compact accumulation basis on the way (flush_record<..., ImuCompactXform>)."""
    sw = S.make_window(10, n_lidar=40000, n_imu=40000, line_fraction=0.3, seed=18)
    rng = np.random.default_rng(1)
    perm = rng.permutation(40000)
    for k in ("t_point", "point", "geo_type", "geo_plane", "geo_normal", "geo_point"):
        sw.lidar[k] = np.ascontiguousarray(sw.lidar[k][perm])
    perm = rng.permutation(40000)
    for k in ("timestamp", "gyro", "accel"):
        sw.imu[k] = np.ascontiguousarray(sw.imu[k][perm])
    eg, _ = build(cuda, sw, unlock_bias=True)
    assert eg.pass_info()["fused"]
    eo, _ = build(oracle, sw, unlock_bias=True)
    Hg, bg, cg, _ = eg.Evaluate()
    Ho, bo, co, _ = eo.Evaluate()
    assert rel(Hg, Ho) < 1e-9 and rel_scaled(Hg, Ho) < 1e-9 and rel(bg, bo) < 1e-9 and abs(cg - co) <= 1e-9 * abs(co)
    Hg2, bg2, _, _ = eg.Evaluate()   # the fallback blocks are re-armed: a second pass gives the same system (to rounding:
    assert rel(Hg2, Hg) < 1e-12 and rel(bg2, bg) < 1e-12  # the atomics order is the one non-deterministic one)
    eg.close()
    eo.close()
