"""Entry / exit of every LM kernel of one Solve(8) on C5 on the device clock.  Needs the developer build
(-DCLIC_LM_TIMELINE) loaded through CLIC_CUDA_LIB; prints the timeline on stderr."""
import os, sys
sys.path.insert(0, "/root/repo")
import bench as B
from clic_b200 import estimator as E, synthetic as S
lib = E.cuda_library()
sw = B.make_workload(1.0)
opt = E.default_options(lib); opt.lock_ab = opt.lock_wb = 0
est = E.TrajectoryEstimator(sw.window, opt, lib); est.SetFixedIndex(-1)
S.add_all_factors(est, sw, count_dropped=False)
x0 = est.read_state()
for rep in range(4):
    est.write_state(x0["knot_q"], x0["knot_p"], x0["bias"], x0["t_offset_ns"], None)
    if rep == 3: os.environ["CLIC_LM_TIMELINE"] = "1"
    est.Solve(8)
