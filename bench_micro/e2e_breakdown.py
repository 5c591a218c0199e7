"""Where does the e2e step (host buffers -> H, b) spend its time?  Phases separated by stream syncs (so their sum is
a little above the pipelined e2e number bench.py reports)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from clic_b200 import estimator as E, synthetic as S

lib = E.cuda_library()
sw = S.config(5, 1.0)
src = type(sw)(window=sw.window, gt_knot_q=sw.gt_knot_q, gt_knot_p=sw.gt_knot_p, name=sw.name)
keep, nbytes = [], 0
for name in ("lidar", "imu", "image"):
    d = dict(getattr(sw, name))
    if name == "lidar":
        for k in ("t_point", "point", "geo_type", "geo_plane", "geo_normal", "geo_point"):
            d.pop(k)
        d.update(S.pack_lidar(sw.lidar))
    for k, v in d.items():
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] > 16:
            t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
            keep.append(t)
            d[k] = t.numpy()
            nbytes += d[k].nbytes
    setattr(src, name, d)
print("H2D MB per step", nbytes / 1e6)
# raw pinned H2D bandwidth for the same bytes
dev = [torch.empty_like(t, device="cuda") for t in keep]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for a, b in zip(dev, keep):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("raw cudaMemcpyAsync of the same pinned arrays: %.3f ms = %.1f GB/s" % (dt * 1e3, nbytes / dt / 1e9))
rows = []
for rep in range(8):
    t = [time.perf_counter()]
    est = E.TrajectoryEstimator(src.window, E.default_options(lib), lib); est.SetFixedIndex(1)
    est.synchronize(); t.append(time.perf_counter())
    S.add_all_factors(est, src, count_dropped=False)
    t.append(time.perf_counter())          # host time to enqueue
    est.synchronize(); t.append(time.perf_counter())
    est.Evaluate(); t.append(time.perf_counter())
    est.close(); t.append(time.perf_counter())
    rows.append(np.diff(t) * 1e3)
print("ms: create | add (enqueue) | add (drain) | evaluate | close")
print(np.array(rows).round(3))

# ---- the bench's pipelined loop (no syncs between phases, shared torch stream, a resident estimator alive) ----
import ctypes as C
stream = torch.cuda.Stream()
def make(srcs, shared):
    opt = E.default_options(lib)
    opt.lock_ab = opt.lock_wb = 0
    if shared:
        opt.stream = C.c_void_p(stream.cuda_stream)
    est = E.TrajectoryEstimator(sw.window, opt, lib); est.SetFixedIndex(-1)
    return est
for shared in (True, False):
    resident = make(sw, shared); S.add_all_factors(resident, sw); resident.Evaluate()
    rows = []
    with torch.cuda.stream(stream):
        for rep in range(30):
            t = [time.perf_counter()]
            est = make(src, shared); t.append(time.perf_counter())
            S.add_all_factors(est, src, count_dropped=False); t.append(time.perf_counter())
            est.Evaluate(); t.append(time.perf_counter())
            est.close(); t.append(time.perf_counter())
            rows.append(np.diff(t) * 1e3)
    rows = np.array(rows)
    print("pipelined loop, shared torch stream =", shared, ": ms create | add | evaluate | close ; total median %.3f max %.3f" % (np.median(rows.sum(1)), rows.sum(1).max()))
    print(rows.round(2)[np.argsort(-rows.sum(1))[:6]])
    resident.close()
