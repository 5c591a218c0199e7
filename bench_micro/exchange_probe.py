"""Per-rank phase stamps of the fused pass on the N-way sharded C5 window (torchrun, one rank per GPU): where the
in-kernel exchange waits.  Prints, per rank and relative to that rank's own kernel start: end of the factor phase,
barrier, end of the assembly + exchange, end of the kernel — and the step time by CUDA events.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
      bench_micro/exchange_probe.py [--no-flush]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from clic_b200 import estimator as E  # noqa: E402
from clic_b200 import synthetic as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = E.cuda_library()
    stream = torch.cuda.Stream()
    sw = B.shard(B.make_workload(1.0), rank, world)
    split = type(sw)(window=sw.window, gt_knot_q=sw.gt_knot_q, gt_knot_p=sw.gt_knot_p, name=sw.name)
    split.imu, split.image = sw.imu, sw.image
    split.lidar = {k: v for k, v in sw.lidar.items()
                   if k not in ("t_point", "point", "geo_type", "geo_plane", "geo_normal", "geo_point")}
    split.lidar.update(S.split_lidar(sw.lidar))
    opt = E.default_options(lib)
    opt.lock_ab = opt.lock_wb = 0
    opt.device = local
    opt.stream = C.c_void_p(stream.cuda_stream)
    est = E.TrajectoryEstimator(sw.window, opt, lib)
    est.SetFixedIndex(-1)
    S.add_all_factors(est, split, count_dropped=False, bias_factor=(rank == 0))
    est.comm_init_torch(dist, rank, world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows, ms = [], []
    with torch.cuda.stream(stream):
        for _ in range(5):
            est.linearize_async(1)
        torch.cuda.synchronize()
        dist.barrier()
        est.pass_stamps(True)
        for _ in range(args.steps):
            if not args.no_flush:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            est.linearize_async(1)
            b.record(stream)
            torch.cuda.synchronize()
            st = est.pass_stamps(True).astype(np.int64)
            t0 = st[:, 0].min()
            rows.append([st[:, 0].max() - t0, st[:, 1].max() - t0, st[:, 2].max() - t0, st[:, 4].max() - t0,
                         st[:, 5].max() - t0, t0])
            ms.append(a.elapsed_time(b))
    r = np.array(rows, dtype=np.float64)
    med = np.median(r[:, :5], axis=0) / 1e3
    mine = torch.tensor(list(med) + [float(np.median(ms)) * 1e3], device="cuda", dtype=torch.float64)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    # kernel start of every rank relative to rank 0's, per step (globaltimer is one clock per GPU: only indicative)
    starts = torch.tensor(r[:, 5], device="cuda", dtype=torch.float64)
    alls = [torch.zeros_like(starts) for _ in range(world)]
    dist.all_gather(alls, starts)
    if rank == 0:
        info = est.pass_info()
        print(f"{world} ranks, flush={not args.no_flush}, partition {info['part_ctas']}")
        print("rank  start_max  F_end  barrier  A+X_end  end   | step by events (us)   [medians over %d steps]" % args.steps)
        for k, t in enumerate(allr):
            v = t.cpu().numpy()
            print("%4d  %8.2f  %6.2f  %6.2f  %7.2f  %6.2f | %7.2f" % (k, v[0], v[1], v[2], v[3], v[4], v[5]))
        s0 = alls[0].cpu().numpy()
        skew = np.array([np.median(a.cpu().numpy() - s0) for a in alls]) / 1e3
        print("median kernel-start offset vs rank 0 by %globaltimer (us; clocks of different GPUs, indicative):",
              np.round(skew - skew.min(), 2))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
