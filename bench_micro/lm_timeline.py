"""Phase timeline of the last lm_step_kernel launch of a production-size Solve(8).  Needs the developer build:
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --compiler-options "-fPIC -fvisibility=hidden" \
       -shared -DCLIC_LM_TIMING -o bench_micro/_libclic_lmtiming.so clic_b200/csrc/clic_cuda.cu
  CLIC_CUDA_LIB=$PWD/bench_micro/_libclic_lmtiming.so python bench_micro/lm_timeline.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bench_micro"))
import numpy as np
from clic_b200 import estimator as E, synthetic as S
lib = E.cuda_library()
for K, nb in ((14, 2), (10, 2)):
    sw = S.make_window(K, n_lidar=700, n_imu=125, seed=41, init_noise=(2e-3, 5e-3), line_fraction=0.5)
    opt = E.default_options(lib); opt.lock_ab = opt.lock_wb = 0
    est = E.TrajectoryEstimator(sw.window, opt, lib); est.SetFixedIndex(-1)
    S.add_all_factors(est, sw)
    summ = est.Solve(3)
    st = (C.c_longlong * 16)()
    lib.clic_debug_lm_stamps(st)
    t = np.array(st[:9], dtype=np.int64)
    names = ["gs/diagonal", "build A + rhs row", "cholesky", "backward solve", "step write+check", "model cost", "delta+plus", "tail", ""]
    d = np.diff(t)
    print("D=%d total %d cycles" % (est.layout().D, t[8] - t[0]))
    for n, c in zip(names[:-1], d):
        print("  %-16s %7d cycles  %5.1f%%" % (n, c, 100.0 * c / (t[8] - t[0])))
    print("  inside the look-ahead Cholesky (sums over panels): panel solve %d, next block column %d, diag block (warp 0) %d, diag || trailing incl. barrier %d" % tuple(st[9:13]))
    print("  inside warp_diag_block (sums): load %d, eliminate %d, store %d" % tuple(st[13:16]))
    est.close()
