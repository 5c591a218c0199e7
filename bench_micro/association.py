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
        oracle.clic_oracle_set_association_search(1)
        fo = A.FeatureMap(sc.map_surf, sc.map_corner, oracle)
        t0 = time.perf_counter()
        fo.FindCorrespondence(*args)
        cpu_grid_ms = (time.perf_counter() - t0) * 1e3
        fo.close()
        oracle.clic_oracle_set_association_search(0)
        same = all(np.array_equal(getattr(c, k), getattr(co, k)) for k in ("line_geo", "plane_geo", "line_src", "plane_src"))
        nq = A.cuda_library().clic_correspondence_capacity(ns) + A.cuda_library().clic_correspondence_capacity(nc)
        rows.append({"map_surf": ms, "map_corner": mc, "features": ns + nc, "searched": int(nq),
                     "lines": len(c.line_t), "planes": len(c.plane_t), "bit_identical_to_oracle": bool(same),
                     "gpu_kernels_ms": round(float(np.median(kern)), 4), "gpu_call_ms": round(float(np.median(call)), 4),
                     "cpu_bruteforce_1thread_ms": round(cpu_ms, 2), "cpu_grid_1thread_ms": round(cpu_grid_ms, 2)})
        print(json.dumps(rows[-1]), flush=True)


def pipeline():
    """One inner iteration's producer (GetLoamFeatureAssociation + the LiDAR adders): fused device call vs the same
    steps through host arrays vs the CPU restatement.  Wall clock of the host calls, estimator state resident."""
    import oracle_lib
    cuda, oracle = E.cuda_library(), oracle_lib.load()
    sw = S.make_window(12, n_lidar=10, n_imu=0, seed=32)
    w = sw.window
    S_LtoI, p_LinI = sw.lidar["S_LtoI"], sw.lidar["p_LinI"]
    hi = (w.n_knots - 3) * w.dt_ns * 1e-9
    shift = np.array([9.0, 7.5, 1.5])
    I4, z3 = np.array([0, 0, 0, 1.0]), np.zeros(3)
    for ms, mc, ns, nc in ((50000, 8000, 1500, 1500), (100000, 15000, 3000, 3000)):
        room = S.make_feature_scene(n_map_surf=ms, n_map_corner=mc, n_surf=ns, n_corner=nc, seed=8)
        map_s = (room.map_surf.astype(np.float64) - shift).astype(np.float32)
        map_c = (room.map_corner.astype(np.float64) - shift).astype(np.float32)
        rng = np.random.default_rng(11)
        res = {}
        scan = None
        for name, lib in (("cuda", cuda), ("oracle", oracle)):
            est0 = E.TrajectoryEstimator(w, E.default_options(lib), lib)
            if scan is None:
                scan = {}
                for kind, pts_M in (("surf", room.surf_M), ("corner", room.corner_M)):
                    t = rng.uniform(1e-6, hi - 1e-6, len(pts_M))
                    q, p, _ = est0.GetSensorPoses(t, S_LtoI, p_LinI, 0.0)
                    scan[kind] = (S.qrot(S.qconj(q), pts_M.astype(np.float64) - shift - p).astype(np.float32), t)
            est0.close()
            if name == "oracle":
                oracle.clic_oracle_set_association_search(1)  # grid-accelerated exact search for the timed CPU arm
            fm = A.FeatureMap(map_s, map_c, lib)
            oracle.clic_oracle_set_association_search(0)
            reps = 12 if name == "cuda" else 3
            fused, split = [], []
            for r in range(reps):
                est = E.TrajectoryEstimator(w, E.default_options(lib), lib)
                est.SetFixedIndex(2)
                t0 = time.perf_counter()
                counts = fm.AddLoamFromScan(est, scan["corner"][0], scan["corner"][1], scan["surf"][0], scan["surf"][1],
                                            I4, z3, S_LtoI, p_LinI, 50.0, cor_downsample=2)
                fused.append((time.perf_counter() - t0) * 1e3)
                est.close()
                est = E.TrajectoryEstimator(w, E.default_options(lib), lib)
                est.SetFixedIndex(2)
                t0 = time.perf_counter()
                cM, _ = est.UndistortScan(scan["corner"][1], scan["corner"][0], S_LtoI, p_LinI)
                sM, _ = est.UndistortScan(scan["surf"][1], scan["surf"][0], S_LtoI, p_LinI)
                c = fm.FindCorrespondence(scan["corner"][0], scan["corner"][1], cM, scan["surf"][0], scan["surf"][1], sM)
                tt = np.concatenate([c.line_t, c.plane_t])
                order = np.argsort(tt, kind="stable")[::2]
                li, pi = order[order < len(c.line_t)], order[order >= len(c.line_t)] - len(c.line_t)
                est.AddLoamLines(c.line_t[li], c.line_point[li], c.line_geo[li], I4, z3, S_LtoI, p_LinI, 50.0)
                est.AddLoamPlanes(c.plane_t[pi], c.plane_point[pi], c.plane_geo[pi], I4, z3, S_LtoI, p_LinI, 50.0)
                split.append((time.perf_counter() - t0) * 1e3)
                est.close()
            fm.close()
            res[name] = (counts, float(np.median(fused[1:])), float(np.median(split[1:])))
        print(json.dumps({"map_surf": ms, "map_corner": mc, "scan_features": ns + nc, "kept_lines_planes": res["cuda"][0][:2],
                          "same_counts_as_oracle": res["cuda"][0] == res["oracle"][0],
                          "gpu_fused_call_ms": round(res["cuda"][1], 3), "gpu_via_host_arrays_ms": round(res["cuda"][2], 3),
                          "cpu_restatement_ms": round(res["oracle"][1], 2)}), flush=True)


if __name__ == "__main__":
    main()
    pipeline()
