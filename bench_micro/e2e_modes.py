"""e2e step (create, add every batch from pinned host memory, Evaluate, close) on C5 or a 1/N shard under the two pass
schedules and the two upload orders: the fused pass starts after the LAST byte arrives; the per-stream kernels run under
the uploads still in flight and leave only the last batch's kernel exposed."""
import argparse, gc, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench as B
from clic_b200 import estimator as E, synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--shard", type=int, default=1)
ap.add_argument("--steps", type=int, default=30)
args = ap.parse_args()
lib = E.cuda_library()
sw = B.shard(B.make_workload(1.0), 0, args.shard)
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
            t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
            keep.append(t)
            d[k] = t.numpy()
    setattr(src, name, d)


def step(small_first, solve=False):
    opt = E.default_options(lib)
    opt.lock_ab = opt.lock_wb = 0
    est = E.TrajectoryEstimator(src.window, opt, lib)
    est.SetFixedIndex(-1)
    S.add_all_factors(est, src, count_dropped=False, small_first=small_first)
    if solve:
        est.Solve(8)
        est.read_state()
        out = None
    else:
        out = est.Evaluate()
    est.close()
    return out


ref = None
gc.disable()
for mode in (os.environ.get("MODES", "default,0,1")).split(","):
    if mode == "default":
        os.environ.pop("CLIC_FUSED", None)
    else:
        os.environ["CLIC_FUSED"] = mode
    for small_first in (True, False):
        for solve in (False, True):
            for _ in range(4):
                out = step(small_first, solve)
            ts = []
            for _ in range(args.steps):
                t0 = time.perf_counter()
                out = step(small_first, solve)
                ts.append((time.perf_counter() - t0) * 1e3)
            if not solve:
                if ref is None:
                    ref = out
                dH = np.abs(out[0] - ref[0]).max() / np.abs(ref[0]).max()
            print("CLIC_FUSED=%-7s small_first=%-5s %-8s median %.3f ms  min %.3f%s" % (
                mode, small_first, "Solve(8)" if solve else "Evaluate", np.median(ts), np.min(ts),
                "" if solve else "  rel dH vs first %.1e" % dH), flush=True)
