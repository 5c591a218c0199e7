"""Why is loam_kernel 80 us under ncu (isolated launches) and 100 us in the back-to-back bench loop?
Runs the C5 fused pass (a) back to back, (b) with idle gaps between steps, sampling SM clock / power through NVML."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pynvml as N
from clic_b200 import estimator as E, synthetic as S

N.nvmlInit()
h = N.nvmlDeviceGetHandleByIndex(0)


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.halt = False
        self.rows = []

    def run(self):
        while not self.halt:
            self.rows.append((N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), N.nvmlDeviceGetPowerUsage(h) / 1000.0,
                              N.nvmlDeviceGetCurrentClocksEventReasons(h)))
            time.sleep(0.002)


lib = E.cuda_library()
sw = S.config(5, 1.0)
est = E.TrajectoryEstimator(sw.window, E.default_options(lib), lib)
est.SetFixedIndex(1)
S.add_all_factors(est, sw, count_dropped=False)
for _ in range(5):
    est.linearize_async(1)
torch.cuda.synchronize()
for name, gap, steps in (("back-to-back", 0.0, 2000), ("5 ms idle gaps", 0.005, 200), ("back-to-back again", 0.0, 2000)):
    s = Sampler(); s.start()
    est.profile_enable(True)
    for _ in range(steps):
        est.linearize_async(1)
        if gap:
            est.synchronize(); time.sleep(gap)
    est.synchronize()
    ms, n = est.profile_read()
    est.profile_enable(False)
    s.halt = True; s.join()
    r = np.array([(a, b) for a, b, _ in s.rows])
    reasons = set(x for _, _, x in s.rows)
    print(name, "per-class us/step:", [round(1e3 * v / steps, 1) for v in ms],
          "| SM MHz min/med/max", r[:, 0].min(), np.median(r[:, 0]), r[:, 0].max(), "| W med/max", np.median(r[:, 1]),
          r[:, 1].max(), "| reasons", [hex(x) for x in reasons], "| samples", len(r))
