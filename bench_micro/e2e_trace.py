"""One e2e step (create, add from pinned host buffers, Evaluate, close) on C5 under CLIC_TRACE=1: host time of every
traced ABI call.  CLIC_TRACE=1 python bench_micro/e2e_trace.py"""
import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench as B
from clic_b200 import estimator as E, synthetic as S
lib = E.cuda_library()
sw = B.make_workload(1.0)
src = type(sw)(window=sw.window, gt_knot_q=sw.gt_knot_q, gt_knot_p=sw.gt_knot_p, name=sw.name)
keep = []
for name in ("lidar", "imu", "image"):
    d = dict(getattr(sw, name))
    if name == "lidar":
        for k in ("t_point", "point", "geo_type", "geo_plane", "geo_normal", "geo_point"):
            d.pop(k)
        d.update(S.split_lidar(sw.lidar))
    for k, v in d.items():
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] > 16:
            t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory(); keep.append(t); d[k] = t.numpy()
    setattr(src, name, d)
for rep in range(4):
    opt = E.default_options(lib); opt.lock_ab = opt.lock_wb = 0
    est = E.TrajectoryEstimator(src.window, opt, lib); est.SetFixedIndex(-1)
    if rep == 3: print("---- traced step", file=sys.stderr)
    S.add_all_factors(est, src, count_dropped=False, small_first=True)
    est.Evaluate(); est.close()
