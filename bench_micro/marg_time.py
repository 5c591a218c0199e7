"""BASELINE C4: wall time of SaveMarginalizationInfo (factor kernels in marg mode + assembly + Schur complement +
prior square-root) on the GPU next to the CPU restatement (eigen-decomposition route, as the reference)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib
from clic_b200 import estimator as E
import test_gpu_marginalization as T

for name, lib, reps in (("gpu", E.cuda_library(), 5), ("cpu oracle 1 thread", oracle_lib.load(), 1)):
    sw = T.make(1.0)
    ts = []
    for rep in range(reps + 1):
        rng = np.random.default_rng(7)
        p0, kinds, ids = T.previous_prior(lib, sw, range(0, 24), rng)
        est = T.marg_estimator(lib, sw, p0, kinds, ids)
        if hasattr(est, "synchronize"):
            try: est.synchronize()
            except Exception: pass
        t0 = time.perf_counter()
        pr = est.SaveMarginalizationInfo()
        ts.append(time.perf_counter() - t0)
        dims = pr.dims()
    print("%s: SaveMarginalizationInfo n=%d m=%d: %.2f ms (best of %d)" % (name, dims[0], dims[2], min(ts[1:] or ts) * 1e3, reps))
