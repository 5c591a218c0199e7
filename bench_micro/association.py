"""SURVEY.md §8f-3 measurement: LidarOdometry::FindCorrespondence on the device vs the CPU restatement.
Production shape: <= 1500 corner + <= 1500 surface features per call (the reference's own stride) against a
down-sampled feature map of 10^4-10^5 points."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from clic_b200 import association as A  # noqa: E402
from clic_b200 import estimator as E  # noqa: E402
from clic_b200 import synthetic as S  # noqa: E402


def main():
    import oracle_lib  # the checker / CPU baseline, never the product
    cuda, oracle = E.cuda_library(), oracle_lib.load()
    rows = []
    for ms, mc, ns, nc in ((20000, 3000, 1500, 1500), (50000, 8000, 1500, 1500), (100000, 15000, 1500, 1500),
                           (200000, 30000, 6000, 6000)):
        sc = S.make_feature_scene(n_map_surf=ms, n_map_corner=mc, n_surf=ns, n_corner=nc, seed=3)
        args = (sc.corner_L, sc.corner_t, sc.corner_M, sc.surf_L, sc.surf_t, sc.surf_M)
        fm = A.FeatureMap(sc.map_surf, sc.map_corner, cuda)
        for _ in range(3):
            c = fm.FindCorrespondence(*args)
        call, kern = [], []
        for _ in range(20):
            t0 = time.perf_counter()
            c = fm.FindCorrespondence(*args)
            call.append((time.perf_counter() - t0) * 1e3)
            kern.append(fm.launches()[1])
        fm.close()
        fo = A.FeatureMap(sc.map_surf, sc.map_corner, oracle)
        t0 = time.perf_counter()
        co = fo.FindCorrespondence(*args)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        fo.close()
        same = all(np.array_equal(getattr(c, k), getattr(co, k)) for k in ("line_geo", "plane_geo", "line_src", "plane_src"))
        nq = A.cuda_library().clic_correspondence_capacity(ns) + A.cuda_library().clic_correspondence_capacity(nc)
        rows.append({"map_surf": ms, "map_corner": mc, "features": ns + nc, "searched": int(nq),
                     "lines": len(c.line_t), "planes": len(c.plane_t), "bit_identical_to_oracle": bool(same),
                     "gpu_kernels_ms": round(float(np.median(kern)), 4), "gpu_call_ms": round(float(np.median(call)), 4),
                     "cpu_bruteforce_1thread_ms": round(cpu_ms, 2)})
        print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
