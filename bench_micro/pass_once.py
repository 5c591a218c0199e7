"""A few fused passes over the C5 window (or a 1/N shard), nothing else: the target of the ncu captures
(launch list, --set full, FP64 instruction counts).  python bench_micro/pass_once.py [--shard N] [--passes K]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from clic_b200 import estimator as E  # noqa: E402
from clic_b200 import synthetic as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shard", type=int, default=1)
ap.add_argument("--passes", type=int, default=3)
ap.add_argument("--scale", type=float, default=1.0)
args = ap.parse_args()
lib = E.cuda_library()
sw = B.shard(B.make_workload(args.scale), 0, args.shard)
split = type(sw)(window=sw.window, gt_knot_q=sw.gt_knot_q, gt_knot_p=sw.gt_knot_p, name=sw.name)
split.imu, split.image = sw.imu, sw.image
split.lidar = {k: v for k, v in sw.lidar.items()
               if k not in ("t_point", "point", "geo_type", "geo_plane", "geo_normal", "geo_point")}
split.lidar.update(S.split_lidar(sw.lidar))
opt = E.default_options(lib)
opt.lock_ab = opt.lock_wb = 0
est = E.TrajectoryEstimator(sw.window, opt, lib)
est.SetFixedIndex(-1)
S.add_all_factors(est, split, count_dropped=False)
print("pass:", est.pass_info(), "factors:", sw.n_factors)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(args.passes):
    flush.fill_(1)
    torch.cuda.synchronize()
    est.linearize_async(1)
    est.synchronize()
est.close()
