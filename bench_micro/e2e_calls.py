"""Wall time of every call of one e2e step (create, the add calls, Evaluate, close) on the C5 window or a 1/N shard, from
pinned host buffers: where the fixed part of the step goes."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench as B
from clic_b200 import estimator as E, synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--shard", type=int, default=1)
args = ap.parse_args()
lib = E.cuda_library()
sw = B.shard(B.make_workload(1.0), 0, args.shard)
src = type(sw)(window=sw.window, gt_knot_q=sw.gt_knot_q, gt_knot_p=sw.gt_knot_p, name=sw.name)
keep, nbytes = [], 0
for name in ("lidar", "imu", "image"):
    d = dict(getattr(sw, name))
    if name == "lidar":
        for k in ("t_point", "point", "geo_type", "geo_plane", "geo_normal", "geo_point"):
            d.pop(k)
        d.update(S.split_lidar(sw.lidar))
    for k, v in d.items():
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] > 16:
            t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
            keep.append(t)
            d[k] = t.numpy()
            nbytes += d[k].nbytes
    setattr(src, name, d)
print("shard 1/%d: H2D %.1f MB per step" % (args.shard, nbytes / 1e6))
orig = {}
times = {}
def wrap(cls, name):
    f = getattr(cls, name)
    def g(self, *a, **k):
        t0 = time.perf_counter()
        r = f(self, *a, **k)
        times.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
        return r
    setattr(cls, name, g)
for n in ("__init__", "AddLoamPlanes", "AddLoamLines", "AddLoamMeasurementAnalytic", "AddIMUMeasurementAnalytic",
          "AddImageFeatureAnalytic", "AddBiasFactor", "Evaluate", "close", "SetFixedIndex"):
    if hasattr(E.TrajectoryEstimator, n):
        wrap(E.TrajectoryEstimator, n)
tot = []
for rep in range(12):
    t0 = time.perf_counter()
    opt = E.default_options(lib); opt.lock_ab = opt.lock_wb = 0
    est = E.TrajectoryEstimator(src.window, opt, lib); est.SetFixedIndex(-1)
    S.add_all_factors(est, src, count_dropped=False)
    est.Evaluate()
    est.close()
    tot.append((time.perf_counter() - t0) * 1e3)
print("step ms: median %.3f" % np.median(tot[2:]))
for k, v in times.items():
    n = len(v) // 12
    if n == 0:
        continue
    v = np.array(v[-12 * n:]).reshape(12, n)[2:]
    print("  %-28s %d call(s)/step  median %.3f ms" % (k, n, np.median(v.sum(axis=1))))
