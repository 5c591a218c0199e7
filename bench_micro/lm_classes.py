"""CUDA-event spans per kernel class (pass / LM kernels) of Solve(8) on C5, with launch counts.
python bench_micro/lm_classes.py"""
import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import bench as B
from clic_b200 import estimator as E, synthetic as S
lib = E.cuda_library()
sw = B.make_workload(1.0)
opt = E.default_options(lib); opt.lock_ab = opt.lock_wb = 0
est = E.TrajectoryEstimator(sw.window, opt, lib); est.SetFixedIndex(-1)
S.add_all_factors(est, sw, count_dropped=False)
x0 = est.read_state()
for rep in range(3):
    est.write_state(x0["knot_q"], x0["knot_p"], x0["bias"], x0["t_offset_ns"], None); est.Solve(8)
est.profile_enable(True)
n = 10
for rep in range(n):
    est.write_state(x0["knot_q"], x0["knot_p"], x0["bias"], x0["t_offset_ns"], None); summ = est.Solve(8)
pm, pn = est.profile_read()
print("per Solve(8): class ms", np.round(np.array(pm) / n * 1e3, 1), "us; launches per class", np.array(pn) / n, "iterations", summ["iterations"])
