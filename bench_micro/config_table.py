"""per-config pass time / LM rate on one B200 next to the CPU restatement on the box's host cores.
C1-C4 are the parity-test configurations (small: launch-latency-bound on a GPU); C5 is the bench workload."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import oracle_lib
from clic_b200 import estimator as E, synthetic as S

cuda, oracle = E.cuda_library(), oracle_lib.load()
cores = os.cpu_count() or 1


def build(lib, sw, c):
    opt = E.default_options(lib)
    if c in (3, 5):
        opt.lock_ab = opt.lock_wb = 0
    if c == 2:
        opt.lock_t_offset[0] = 0
        opt.t_offset_padding_ns = 10_000_000
    est = E.TrajectoryEstimator(sw.window, opt, lib)
    est.SetFixedIndex(-1)
    S.add_all_factors(est, sw, bias_factor=(c in (3, 5)))
    return est


def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)


print("| config | factors L/I/V | D | CPU 1-thr pass ms | CPU %d-thr pass ms | GPU pass us | GPU factor-evals/s | GPU LM it/s | CPU 1-thr LM it/s |" % cores)
for c in (1, 2, 3, 4, 5):
    sw = S.config(c, 1.0)
    nL, nI, nV = len(sw.lidar.get("t_point", ())), len(sw.imu.get("timestamp", ())), len(sw.image.get("t_i", ()))
    eg = build(cuda, sw, c)
    D = eg.layout().D
    def gpass():
        eg.linearize_async(20); eg.synchronize()
    t_g = timed(gpass, 5) / 20
    x0 = eg.read_state()
    def gsolve():
        eg.write_state(x0["knot_q"], x0["knot_p"], x0["bias"], x0["t_offset_ns"], None)
        return eg.Solve(8)
    summ = gsolve()
    t0 = time.perf_counter(); summ = gsolve(); t_gs = time.perf_counter() - t0
    eg.close()
    small = S.config(c, 0.05) if c == 5 else sw
    scale = sw.n_factors / small.n_factors
    res = {}
    for th in (1, cores):
        oracle.clic_oracle_set_threads(th)
        eo = build(oracle, small, c)
        res[th] = timed(lambda: eo.linearize_async(1), 2 if c == 5 else 3) * scale
        if th == 1 and c != 5:
            xo = eo.read_state()
            t0 = time.perf_counter(); so = eo.Solve(8); res["lm"] = so["iterations"] / (time.perf_counter() - t0)
        eo.close()
    oracle.clic_oracle_set_threads(1)
    print("| C%d | %d/%d/%d | %d | %.2f%s | %.2f | %.1f | %.3g | %.0f (%d it) | %s |" % (
        c, nL, nI, nV, D, res[1] * 1e3, " (5%% sample x20)" if c == 5 else "", res[cores] * 1e3, t_g * 1e6,
        sw.n_factors / t_g, summ["iterations"] / t_gs, summ["iterations"], ("%.1f" % res["lm"]) if "lm" in res else "-"))
