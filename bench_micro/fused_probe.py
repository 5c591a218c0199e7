"""Fused pass vs per-stream kernels on the C5 window (or a 1/N shard of it): step time with CUDA events (L2 flushed),
and the per-partition phase times of the fused kernel from its %globaltimer stamps (tuning aid for the cost model).
  python bench_micro/fused_probe.py [--shard N] [--scale S] [--steps K]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from clic_b200 import estimator as E  # noqa: E402
from clic_b200 import synthetic as S  # noqa: E402

KINDS = ["lidar_mixed", "lidar_planes", "lidar_lines", "imu", "visual_packed", "visual_chain"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shard", type=int, default=1, help="time rank 0's shard of an N-way split (no exchange)")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between steps (code and data)")
    ap.add_argument("--no-visual", action="store_true", help="LiDAR + IMU only")
    args = ap.parse_args()
    lib = E.cuda_library()
    stream = torch.cuda.Stream()
    sw = B.shard(B.make_workload(args.scale, with_visual=not args.no_visual), 0, args.shard)
    split = type(sw)(window=sw.window, gt_knot_q=sw.gt_knot_q, gt_knot_p=sw.gt_knot_p, name=sw.name)
    split.imu, split.image = sw.imu, sw.image
    split.lidar = {k: v for k, v in sw.lidar.items()
                   if k not in ("t_point", "point", "geo_type", "geo_plane", "geo_normal", "geo_point")}
    split.lidar.update(S.split_lidar(sw.lidar))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def run(tag):
        opt = E.default_options(lib)
        opt.lock_ab = opt.lock_wb = 0
        opt.stream = C.c_void_p(stream.cuda_stream)
        est = E.TrajectoryEstimator(sw.window, opt, lib)
        est.SetFixedIndex(-1)
        S.add_all_factors(est, split, count_dropped=False)
        info = est.pass_info()
        with torch.cuda.stream(stream):
            for _ in range(5):
                est.linearize_async(1)
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for a, b in ev:
                if not args.no_flush:
                    flush.fill_(1)
                a.record(stream)
                est.linearize_async(1)
                b.record(stream)
            torch.cuda.synchronize()
            ms = np.array([a.elapsed_time(b) for a, b in ev])
            print(f"{tag}: fused={info['fused']} grid={info['grid']} parts={info['part_ctas']}  step ms "
                  f"min {ms.min():.4f} median {np.median(ms):.4f} max {ms.max():.4f}  n_factors {sw.n_factors}")
            if info["fused"]:
                est.pass_stamps(True)
                if not args.no_flush:
                    flush.fill_(1)
                est.linearize_async(1)
                torch.cuda.synchronize()
                st = est.pass_stamps(False).astype(np.int64)
                t0 = st[:, 0].min()
                us = (st[:, :6] - t0) / 1e3
                names = ["start", "F end", "barrier", "staged", "A end", "end"]
                print("  whole kernel %.2f us | " % us[:, 5].max() +
                      " | ".join("%s max %.2f" % (n, us[:, i].max()) for i, n in enumerate(names)))
                # partitions in job order: LiDAR batches in add order (planes, lines), IMU, visual
                c0 = 0
                for kind in (0, 1, 2, 3, 4, 5):
                    n = info["part_ctas"][kind]
                    if n == 0:
                        continue
                    seg = us[c0:c0 + n, 1]
                    print("    %-14s CTAs %3d..%3d  factor phase ends  min %.2f  mean %.2f  max %.2f us" %
                          (KINDS[kind], c0, c0 + n - 1, seg.min(), seg.mean(), seg.max()))
                    body = (st[c0:c0 + n, 6:8] - t0) / 1e3  # warp 0 of each CTA: first slab begins, last slab done
                    print("    %-14s   warp 0: slabs begin mean %.2f  slabs done mean %.2f max %.2f  -> finish + write "
                          "mean %.2f us" % ("", body[:, 0].mean(), body[:, 1].mean(), body[:, 1].max(),
                                            (seg - body[:, 1]).mean()))
                    c0 += n
                print("    phase A per CTA (A end - staged): mean %.2f max %.2f us; staging (staged - barrier) mean %.2f; "
                      "barrier wait of the last CTA %.2f us" % (
                          (us[:-1, 4] - us[:-1, 3]).mean(), (us[:-1, 4] - us[:-1, 3]).max(),
                          (us[:-1, 3] - us[:-1, 2]).mean(), us[:, 2].max() - us[:, 1].max()))
            if not info["fused"]:
                est.profile_enable(True)
                for _ in range(args.steps):
                    if not args.no_flush:
                        flush.fill_(1)
                    est.linearize_async(1)
                torch.cuda.synchronize()
                pm, pn = est.profile_read()
                est.profile_enable(False)
                print("  per-class event spans (us/step): lidar %.1f imu %.1f visual %.1f reduce+assemble %.1f" % tuple(
                    1e3 * pm[i] / args.steps for i in range(4)))
        H, b, c, f = est.Evaluate()
        est.close()
        return H, b, c

    Hf, bf, cf = run("fused     ")
    os.environ["CLIC_FUSED"] = "0"
    Hs, bs, cs = run("per-stream")
    print("fused vs per-stream: rel H %.2e b %.2e cost %.2e" % (
        np.abs(Hf - Hs).max() / np.abs(Hs).max(), np.abs(bf - bs).max() / np.abs(bs).max(), abs(cf - cs) / abs(cs)))


if __name__ == "__main__":
    main()
