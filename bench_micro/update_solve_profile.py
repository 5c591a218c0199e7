"""Where one production-size Solve(8) spends its time (the problem of bench_micro/odometry_update.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bench_micro"))
import numpy as np
import odometry_update as U
from clic_b200 import association as A, estimator as E, synthetic as S

lib = E.cuda_library()
sw = S.make_window(U.K, n_lidar=10, n_imu=125, seed=41, init_noise=(2e-3, 5e-3))
room = S.make_feature_scene(n_map_surf=50000, n_map_corner=8000, n_surf=1500, n_corner=1500, seed=8)
map_s = (room.map_surf.astype(np.float64) - U.SHIFT).astype(np.float32)
map_c = (room.map_corner.astype(np.float64) - U.SHIFT).astype(np.float32)
hi = (U.K - 3) * sw.window.dt_ns * 1e-9
gt = S.SplineNP(sw.window.t0_ns, sw.window.dt_ns, sw.gt_knot_q, sw.gt_knot_p)
rng = np.random.default_rng(11)
Lx, M = sw.lidar, sw.imu
scan = {}
for kind, pts_M in (("surf", room.surf_M), ("corner", room.corner_M)):
    tq = rng.uniform(1e-6, hi - 1e-6, len(pts_M))
    t_ns = (tq * 1e9).astype(np.int64)
    R, pos = gt.rotation(t_ns), gt.position(t_ns)
    q = S.qmul(R, np.broadcast_to(Lx["S_LtoI"], R.shape))
    p = S.qrot(R, np.broadcast_to(Lx["p_LinI"], pos.shape)) + pos
    scan[kind] = (S.qrot(S.qconj(q), pts_M.astype(np.float64) - U.SHIFT - p).astype(np.float32), tq)
fm = A.FeatureMap(map_s, map_c, lib)
spec = U.previous_prior(lib, sw, range(0, 8), np.random.default_rng(7))
names = ["lidar", "imu", "visual", "reduce_assemble", "prior_bias", "lm_step", "allreduce", "marginalize"]
for rep in range(4):
    opt = E.default_options(lib); opt.lock_ab = opt.lock_wb = 0
    est = E.TrajectoryEstimator(sw.window, opt, lib); est.SetFixedIndex(-1)
    est.AddMarginalizationFactor(spec[0], [])
    est.AddIMUMeasurementAnalytic(M["timestamp"], M["gyro"], M["accel"], 1, M["info_vec"], count_dropped=False)
    est.AddBiasFactor(0, 1, 1.0, S.BIAS_INFO_VEC / np.sqrt(0.1 / S.IMU_FREQ))
    fm.AddLoamFromScan(est, scan["corner"][0], scan["corner"][1], scan["surf"][0], scan["surf"][1], U.I4, U.Z3,
                       Lx["S_LtoI"], Lx["p_LinI"], 50.0, cor_downsample=2)
    est.synchronize()
    prof = rep == 3
    if prof: est.profile_enable(True)
    t0 = time.perf_counter(); summ = est.Solve(8); dt = time.perf_counter() - t0
    print("Solve(8): %.3f ms wall, %d iterations, D=%d" % (dt * 1e3, summ["iterations"], est.layout().D))
    if prof:
        ms, cnt = est.profile_read()
        print({n: (round(float(m), 4), int(c)) for n, m, c in zip(names, ms, cnt) if c})
    est.close()
