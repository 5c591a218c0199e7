"""Factor-phase end time of every CTA of the fused pass on C5 (%globaltimer stamps, mean over 6 passes) and the time
warp 0 of each CTA is through its slabs: which CTAs a partition waits for.  python bench_micro/cta_ends.py"""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/bench_micro")
import numpy as np, torch, ctypes as C
import bench as B
from clic_b200 import estimator as E, synthetic as S
lib = E.cuda_library()
sw = B.make_workload(1.0)
split = type(sw)(window=sw.window, gt_knot_q=sw.gt_knot_q, gt_knot_p=sw.gt_knot_p, name=sw.name)
split.imu, split.image = sw.imu, sw.image
split.lidar = {k: v for k, v in sw.lidar.items() if k not in ("t_point", "point", "geo_type", "geo_plane", "geo_normal", "geo_point")}
split.lidar.update(S.split_lidar(sw.lidar))
opt = E.default_options(lib); opt.lock_ab = opt.lock_wb = 0
est = E.TrajectoryEstimator(sw.window, opt, lib); est.SetFixedIndex(-1)
S.add_all_factors(est, split, count_dropped=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5): est.linearize_async(1)
torch.cuda.synchronize()
rows = []
for rep in range(6):
    est.pass_stamps(True); flush.fill_(1); torch.cuda.synchronize()
    est.linearize_async(1); torch.cuda.synchronize()
    st = est.pass_stamps(False).astype(np.int64)
    t0 = st[:, 0].min()
    rows.append((st[:, 7] - t0) / 1e3)   # warp 0 slabs done
    rows.append((st[:, 1] - t0) / 1e3)
a = np.array(rows)
np.set_printoptions(linewidth=250, precision=0, suppress=True)
print("slabs-done (warp0) per CTA, 3 reps:"); print(a[0::2][:3])
print("F end per CTA, mean over reps:"); print(a[1::2].mean(0))
print("std over reps of F end per CTA (mean):", a[1::2].std(0).mean())
print("corr between reps:", np.corrcoef(a[1::2])[0, 1:])
