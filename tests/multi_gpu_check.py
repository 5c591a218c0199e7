"""Run under torchrun (one rank per GPU): every rank holds a contiguous-in-time shard of the factors, H|b|cost are
all-reduced (NCCL) once per pass, every rank solves redundantly.  Checks against the full-window oracle on rank 0
and that all ranks end bit-identical.  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
from bench import shard  # noqa: E402
from clic_b200 import estimator as E  # noqa: E402
from clic_b200 import synthetic as S  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = E.cuda_library()
    sw = S.make_window(10, n_lidar=20000, n_imu=2000, n_landmarks=60, obs_per_landmark=8, line_fraction=0.3, seed=77)
    mine = shard(sw, rank, world)

    def build(lib_, data, bias_factor, device=0):
        opt = E.default_options(lib_)
        opt.lock_ab = opt.lock_wb = 0
        opt.device = device
        est = E.TrajectoryEstimator(data.window, opt, lib_)
        est.SetFixedIndex(1)
        S.add_all_factors(est, data, bias_factor=bias_factor)
        return est

    est = build(lib, mine, bias_factor=(rank == 0), device=local)   # the single bias factor lives on rank 0
    est.comm_init_torch(dist, rank, world)
    if rank == 0:
        print("pass: fused=%s exchange=%s" % (est.pass_info()["fused"],
                                               "kernel" if lib.clic_comm_p2p_ready() else "nccl"))
    H, b, cost, _ = est.Evaluate()
    summary = est.Solve(8)
    x = est.read_state()
    ok = True
    if rank == 0:
        olib = oracle_lib.load()
        eo = build(olib, sw, True)
        Ho, bo, co, _ = eo.Evaluate()
        so = eo.Solve(8)
        xo = eo.read_state()
        d = np.sqrt(np.maximum(np.diag(Ho), 1e-300))
        eh = (np.abs(H - Ho) / np.outer(d, d)).max()
        eb = np.abs(b - bo).max() / np.abs(bo).max()
        ec = abs(cost - co) / abs(co)
        ep = np.abs(x["knot_p"] - xo["knot_p"]).max()
        print(f"[{world} ranks] H {eh:.2e} b {eb:.2e} cost {ec:.2e} | solve iters {summary['iterations']}/{so['iterations']} "
              f"pos diff {ep:.2e}")
        ok = eh < 1e-9 and eb < 1e-9 and ec < 1e-9 and ep < 1e-7 and summary["iterations"] == so["iterations"]
    # every rank must hold the same state bit for bit (same all-reduced H on every rank, redundant solve)
    t = torch.from_numpy(np.concatenate([x["knot_q"].ravel(), x["knot_p"].ravel(), x["bias"].ravel()])).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(t, ref))
    flag = torch.tensor([1 if (ok and same) else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if flag.item() == 1 else "FAIL", "bit-identical ranks:", same)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
