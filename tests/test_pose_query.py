"""batched Trajectory::GetSensorPose / UndistortScan(InG) (src/spline/trajectory.cpp:35-115), the step
before the hot path each iteration.  CPU: the oracle restatement vs an independent numpy forward spline (SplineNP).
GPU: the CUDA kernels through the C ABI vs the oracle."""
import numpy as np
import pytest

from clic_b200 import estimator as E
from clic_b200 import synthetic as S


def _window(seed=5, K=14):
    sw = S.make_window(K, n_lidar=50, n_imu=0, seed=seed)
    return sw, sw.lidar["S_LtoI"], sw.lidar["p_LinI"]


def _queries(sw, n, rng, toff=0.0):
    w = sw.window
    lo, hi = w.t0_ns * 1e-9, (w.t0_ns + (w.n_knots - 3) * w.dt_ns) * 1e-9
    t = rng.uniform(lo - toff * 1e-9 + 1e-6, hi - toff * 1e-9 - 1e-6, n)
    t[0] = lo - toff * 1e-9 + 1e-9            # first valid nanoseconds
    return np.sort(t)


def test_oracle_pose_query_matches_numpy_spline(oracle):
    sw, S_LtoI, p_LinI = _window()
    w = sw.window
    rng = np.random.default_rng(0)
    toff = 250_000.0
    t = _queries(sw, 400, rng, toff)
    est = E.TrajectoryEstimator(w, E.default_options(oracle), oracle)
    q, p, bad = est.GetSensorPoses(t, S_LtoI, p_LinI, toff)
    assert bad == 0
    sp = S.SplineNP(w.t0_ns, w.dt_ns, w.knot_q, w.knot_p)
    t_ns = (t * 1e9 + toff).astype(np.int64)
    R, pos = sp.rotation(t_ns), sp.position(t_ns)
    q_ref = S.qmul(R, np.broadcast_to(S_LtoI, R.shape))
    p_ref = S.qrot(R, np.broadcast_to(p_LinI, pos.shape)) + pos
    sign = np.sign(np.sum(q * q_ref, axis=1))[:, None]
    assert np.abs(q - sign * q_ref).max() < 1e-13
    assert np.abs(p - p_ref).max() < 1e-12
    # out-of-range times are counted (the reference asserts) and give the identity
    t_bad = np.array([t[0] - 1.0, t[-1] + 1.0, t[5]])
    q2, p2, bad2 = est.GetSensorPoses(t_bad, S_LtoI, p_LinI, toff)
    assert bad2 == 2 and np.array_equal(q2[0], [0, 0, 0, 1]) and not p2[1].any() and np.array_equal(q2[2], q[5])
    # undistortion = pose applied to the float point, stored as float
    xyz = rng.normal(0, 20, (len(t), 3)).astype(np.float32)
    out, bad3 = est.UndistortScan(t, xyz, S_LtoI, p_LinI, toff)
    assert bad3 == 0
    ref = (S.qrot(q, xyz.astype(np.float64)) + p).astype(np.float32)
    assert np.allclose(out, ref, rtol=3e-7, atol=1e-6)
    out_t, _ = est.UndistortScan(t, xyz, S_LtoI, p_LinI, toff, target_t_s=t[17])
    qt, pt = q[17], p[17]
    ref_t = S.qrot(S.qconj(qt), ref.astype(np.float64) - pt)
    assert np.abs(out_t - ref_t).max() < 2e-5      # ref went through float once more


@pytest.mark.gpu
@pytest.mark.parametrize("toff", [0.0, -130_000.4])
def test_gpu_pose_query_and_undistort(cuda, oracle, toff):
    sw, S_LtoI, p_LinI = _window(seed=9, K=20)
    w = sw.window
    rng = np.random.default_rng(1)
    t = _queries(sw, 20000, rng, toff)
    t[100] = t[-1] + 5.0                             # out of range
    xyz = rng.normal(0, 25, (len(t), 3)).astype(np.float32)
    xyz[7, 0] = np.nan                               # pcl_isnan(raw_p.x) -> skipped by the reference
    res = {}
    for name, lib in (("g", cuda), ("o", oracle)):
        est = E.TrajectoryEstimator(w, E.default_options(lib), lib)
        res[name] = (est.GetSensorPoses(t, S_LtoI, p_LinI, toff), est.UndistortScan(t, xyz, S_LtoI, p_LinI, toff),
                     est.UndistortScan(t, xyz, S_LtoI, p_LinI, toff, target_t_s=t[3000]))
        est.close()
    (qg, pg, bg), (ug, ubg), (vg, vbg) = res["g"]
    (qo, po, bo), (uo, ubo), (vo, vbo) = res["o"]
    assert bg == bo == 1 and ubg == ubo == 2 and vbg == vbo == 2
    assert np.abs(qg - qo).max() < 1e-13 and np.abs(pg - po).max() < 1e-12
    ok = ~np.isnan(uo[:, 0])
    assert np.array_equal(np.isnan(ug), np.isnan(uo)) and np.array_equal(np.isnan(vg), np.isnan(vo))
    # float outputs: identical up to the last float bit (the double results differ by ~1e-15 relative)
    assert np.abs(ug[ok] - uo[ok]).max() <= 4e-6 and (ug[ok] != uo[ok]).mean() < 1e-3
    assert np.abs(vg[ok] - vo[ok]).max() <= 4e-6 and (vg[ok] != vo[ok]).mean() < 1e-3
