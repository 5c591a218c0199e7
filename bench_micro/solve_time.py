"""Solve(8) wall time on C5 (and C1, C3): median / min over repetitions, per LM iteration."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench as B
from clic_b200 import estimator as E, synthetic as S
lib = E.cuda_library()
for name, sw in (("C1", S.config(1)), ("C3", S.config(3)), ("C5", B.make_workload(1.0))):
    opt = E.default_options(lib)
    opt.lock_ab = opt.lock_wb = 0 if name != "C1" else 1
    est = E.TrajectoryEstimator(sw.window, opt, lib); est.SetFixedIndex(-1)
    S.add_all_factors(est, sw, bias_factor=(name != "C1"), count_dropped=False)
    x0 = est.read_state()
    ts = []
    for rep in range(25):
        est.write_state(x0["knot_q"], x0["knot_p"], x0["bias"], x0["t_offset_ns"], None)
        est.synchronize()
        t0 = time.perf_counter(); summ = est.Solve(8); ts.append(time.perf_counter() - t0)
    ts = np.array(ts[3:]) * 1e6
    it = summ["iterations"]
    print("%s: Solve(8) %d iterations: median %.1f us (min %.1f) -> %.1f us / iteration, %.0f LM it/s" % (
        name, it, np.median(ts), ts.min(), np.median(ts) / it, it / np.median(ts) * 1e6))
    est.close()
