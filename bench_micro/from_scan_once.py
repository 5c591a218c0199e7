"""A few clic_add_loam_from_scan calls at the production shape (for ncu / quick timing)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from clic_b200 import association as A, estimator as E, synthetic as S
lib = E.cuda_library()
sw = S.make_window(12, n_lidar=10, n_imu=0, seed=32)
w = sw.window
S_LtoI, p_LinI = sw.lidar["S_LtoI"], sw.lidar["p_LinI"]
hi = (w.n_knots - 3) * w.dt_ns * 1e-9
shift = np.array([9.0, 7.5, 1.5])
room = S.make_feature_scene(n_map_surf=50000, n_map_corner=8000, n_surf=1500, n_corner=1500, seed=8)
rng = np.random.default_rng(11)
est0 = E.TrajectoryEstimator(w, E.default_options(lib), lib)
scan = {}
for kind, pts_M in (("surf", room.surf_M), ("corner", room.corner_M)):
    t = rng.uniform(1e-6, hi - 1e-6, len(pts_M))
    q, p, _ = est0.GetSensorPoses(t, S_LtoI, p_LinI, 0.0)
    scan[kind] = (S.qrot(S.qconj(q), pts_M.astype(np.float64) - shift - p).astype(np.float32), t)
est0.close()
fm = A.FeatureMap((room.map_surf.astype(np.float64) - shift).astype(np.float32),
                  (room.map_corner.astype(np.float64) - shift).astype(np.float32), lib)
I4, z3 = np.array([0, 0, 0, 1.0]), np.zeros(3)
ts = []
for r in range(8):
    est = E.TrajectoryEstimator(w, E.default_options(lib), lib); est.SetFixedIndex(2)
    t0 = time.perf_counter()
    c = fm.AddLoamFromScan(est, scan["corner"][0], scan["corner"][1], scan["surf"][0], scan["surf"][1], I4, z3,
                           S_LtoI, p_LinI, 50.0, cor_downsample=2)
    ts.append((time.perf_counter() - t0) * 1e3)
    est.close()
print(c, [round(x, 3) for x in ts])
fm.close()
