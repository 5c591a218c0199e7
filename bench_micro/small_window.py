"""Production-size windows (BASELINE C1-C3): wall time of Solve(8) and per-launch picture."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from clic_b200 import estimator as E, synthetic as S
lib = E.cuda_library()
for c in (1, 2, 3):
    sw = S.config(c, 1.0)
    opt = E.default_options(lib)
    if c == 3: opt.lock_ab = opt.lock_wb = 0
    if c == 2:
        opt.lock_t_offset[0] = 0; opt.t_offset_padding_ns = 10_000_000
    est = E.TrajectoryEstimator(sw.window, opt, lib); est.SetFixedIndex(-1)
    S.add_all_factors(est, sw, bias_factor=(c == 3))
    x0 = est.read_state()
    ts = []
    for rep in range(6):
        est.write_state(x0["knot_q"], x0["knot_p"], x0["bias"], x0["t_offset_ns"], None)
        est.synchronize()
        t0 = time.perf_counter(); summ = est.Solve(8); ts.append(time.perf_counter() - t0)
    print("C%d: D=%d Solve(8) %.3f ms, %d iterations -> %.0f LM it/s, %.1f us / iteration" % (
        c, est.layout().D, min(ts) * 1e3, summ["iterations"], summ["iterations"] / min(ts), min(ts) / summ["iterations"] * 1e6))
    est.close()
