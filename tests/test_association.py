"""LiDAR data association, LidarOdometry::SetTargetMap / FindCorrespondence
(src/lidar_odometry/lidar_odometry.cpp:472-657) — the producer of the hot path's input.
CPU: the oracle restatement vs independent numpy definitions (sorted float32 distances, float64 least squares / eigh).
GPU: the CUDA search + fits through the C ABI vs the oracle, bit for bit (indices, counts, every output double)."""
import numpy as np
import pytest

from clic_b200 import association as A
from clic_b200 import estimator as E
from clic_b200 import synthetic as S


NOT_FOUND = np.iinfo(np.int32).max  # the probe's mark for "no neighbour within the 1 m^2 gate"


def _run(lib, sc, corner_map=True):
    fm = A.FeatureMap(sc.map_surf, sc.map_corner if corner_map else None, lib)
    c = fm.FindCorrespondence(sc.corner_L, sc.corner_t, sc.corner_M, sc.surf_L, sc.surf_t, sc.surf_M)
    nn = (fm.neighbours(0), fm.neighbours(1))
    fm.close()
    return c, nn


def _np_knn5(map_xyz, q):
    """Exact 5-NN under flann::L2_Simple<float>: float32 ((dx*dx + dy*dy) + dz*dz), ties by index (stable sort)."""
    d = (q[None, :] - map_xyz).astype(np.float32)
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    order = np.argsort(d2, kind="stable")[:5]
    return order, d2[order]


def test_oracle_association_matches_numpy(oracle):
    sc = S.make_feature_scene(n_map_surf=6000, n_map_corner=1500, n_surf=300, n_corner=120, seed=1)
    c, (nn_c, nn_s) = _run(oracle, sc)
    assert len(c.plane_t) > 150 and len(c.line_t) > 40
    # neighbours
    for i in range(0, len(sc.surf_t), 7):
        if sc.surf_t[i] < 0:
            assert (nn_s[i] == -1).all()
            continue
        idx, d2 = _np_knn5(sc.map_surf, sc.surf_M[i])
        assert (nn_s[i] == np.where(d2 <= 1.0, idx, NOT_FOUND)).all()
    for i in range(0, len(sc.corner_t), 5):
        idx, d2 = _np_knn5(sc.map_corner, sc.corner_M[i])
        assert (nn_c[i] == np.where(d2 <= 1.0, idx, NOT_FOUND)).all()
    # planes: float64 least squares over the same 5 points, unit normal, all within 0.2 m; gates as the reference
    kept = set(c.plane_src.tolist())
    for i in range(len(sc.surf_t)):
        if sc.surf_t[i] < 0:
            assert i not in kept
            continue
        idx, d2 = _np_knn5(sc.map_surf, sc.surf_M[i])
        if d2[4] > 1.0:
            assert i not in kept
            continue
        P = sc.map_surf[idx].astype(np.float64)
        x = np.linalg.lstsq(P, -np.ones(5), rcond=None)[0]
        pl = np.append(x, 1.0) / np.linalg.norm(x)
        resid = np.abs(P @ pl[:3] + pl[3])
        if abs(resid.max() - 0.2) < 1e-3:
            continue  # too close to the gate to call in float32
        assert (i in kept) == (resid.max() <= 0.2)
        if i in kept:
            o = int(np.nonzero(c.plane_src == i)[0][0])
            cond = np.linalg.cond(P)
            assert np.abs(c.plane_geo[o] - pl).max() < 2e-6 * cond
            assert c.plane_t[o] == sc.surf_t[i] and (c.plane_point[o] == sc.surf_L[i].astype(np.float64)).all()
    # lines: float64 eigh of the covariance
    kept = set(c.line_src.tolist())
    for i in range(len(sc.corner_t)):
        idx, d2 = _np_knn5(sc.map_corner, sc.corner_M[i])
        if d2[4] > 1.0:
            assert i not in kept
            continue
        P = sc.map_corner[idx].astype(np.float64)
        cp = P.mean(axis=0)
        w, v = np.linalg.eigh((P - cp).T @ (P - cp) / 5)
        if abs(w[2] - 3 * w[1]) < 1e-4 * w[2]:
            continue
        assert (i in kept) == (w[2] > 3 * w[1])
        if i in kept:
            o = int(np.nonzero(c.line_src == i)[0][0])
            assert np.abs(c.line_geo[o, :3] - cp).max() < 1e-12 + 1e-15 * np.abs(cp).max()
            d = c.line_geo[o, 3:]
            assert abs(np.linalg.norm(d) - 1) < 1e-5
            assert 1 - abs(d @ v[:, 2]) < 1e-5 * w[2] / (w[2] - w[1])
    # reference order: sources ascending within each type
    assert (np.diff(c.line_src) > 0).all() and (np.diff(c.plane_src) > 0).all()


def test_oracle_association_stride_and_gates(oracle):
    sc = S.make_feature_scene(n_map_surf=4000, n_map_corner=600, n_surf=3400, n_corner=1700, seed=2)
    c, (nn_c, nn_s) = _run(oracle, sc)
    assert len(nn_s) == (3400 + 2) // 3 and (c.plane_src % 3 == 0).all()      # step = 3400 / 1500 + 1 = 3
    assert len(nn_c) == (1700 + 1) // 2 and (c.line_src % 2 == 0).all()       # step = 2
    # no corner map, or one of at most 10 points: no line records (lidar_odometry.cpp:502)
    c2, (nn_c2, _) = _run(oracle, sc, corner_map=False)
    assert len(c2.line_t) == 0 and len(nn_c2) == 0 and len(c2.plane_t) == len(c.plane_t)
    sc.map_corner = sc.map_corner[:10]
    c3, _ = _run(oracle, sc)
    assert len(c3.line_t) == 0


def _assert_identical(a, b):
    for name in ("line_t", "line_point", "line_geo", "line_src", "plane_t", "plane_point", "plane_geo", "plane_src"):
        x, y = getattr(a, name), getattr(b, name)
        assert x.shape == y.shape, name
        assert np.array_equal(x, y), (name, np.abs(x - y).max())


@pytest.mark.gpu
@pytest.mark.parametrize("sizes", [(20000, 3000, 1200, 400), (50000, 8000, 1500, 1500), (3000, 500, 4000, 3100)])
def test_gpu_association_bit_exact(cuda, oracle, sizes):
    ms, mc, ns, nc = sizes
    sc = S.make_feature_scene(n_map_surf=ms, n_map_corner=mc, n_surf=ns, n_corner=nc, seed=ms % 97)
    cg, nng = _run(cuda, sc)
    co, nno = _run(oracle, sc)
    assert len(cg.plane_t) > 0.5 * len(nng[1]) and len(cg.line_t) > 0.3 * len(nng[0])
    for g, o in zip(nng, nno):
        assert np.array_equal(g, o)
    _assert_identical(cg, co)


@pytest.mark.gpu
def test_gpu_association_edge_cases(cuda, oracle):
    sc = S.make_feature_scene(n_map_surf=2500, n_map_corner=400, n_surf=200, n_corner=90, seed=9)
    # duplicated map points (exact distance ties), a NaN feature, features exactly on map points
    sc.map_surf = np.concatenate([sc.map_surf, sc.map_surf[:800]])
    sc.surf_M[3] = np.nan
    sc.surf_M[10:20] = sc.map_surf[100:110]
    for corner_map in (True, False):
        cg, nng = _run(cuda, sc, corner_map)
        co, nno = _run(oracle, sc, corner_map)
        for g, o in zip(nng, nno):
            assert np.array_equal(g, o)
        _assert_identical(cg, co)
    # tiny maps: fewer than 5 surface points, at most 10 corner points
    sc.map_surf = sc.map_surf[:4]
    sc.map_corner = sc.map_corner[:10]
    cg, _ = _run(cuda, sc)
    assert len(cg.plane_t) == 0 and len(cg.line_t) == 0
    # empty scan
    fm = A.FeatureMap(sc.map_surf, None, cuda)
    e3, e0 = np.zeros((0, 3), np.float32), np.zeros(0)
    c = fm.FindCorrespondence(e3, e0, e3, e3, e0, e3)
    assert len(c.plane_t) == 0 and len(c.line_t) == 0
    fm.close()


@pytest.mark.gpu
def test_gpu_undistort_associate_linearize_chain(cuda, oracle):
    """The reference's inner iteration (odometry_manager.cpp:321-339): undistort the scan with the current
    trajectory, associate it with the map, add the correspondences as LiDAR factors, linearise."""
    sw = S.make_window(12, n_lidar=10, n_imu=0, seed=31)
    w = sw.window
    S_LtoI, p_LinI = sw.lidar["S_LtoI"], sw.lidar["p_LinI"]
    rng = np.random.default_rng(3)
    hi = (w.n_knots - 3) * w.dt_ns * 1e-9
    room = S.make_feature_scene(n_map_surf=30000, n_map_corner=4000, n_surf=1400, n_corner=500, seed=4)
    shift = np.array([9.0, 7.5, 1.5])  # map frame = G shifted so that the trajectory runs inside the room
    I4 = np.array([0, 0, 0, 1.0])
    ests = {name: E.TrajectoryEstimator(w, E.default_options(lib), lib) for name, lib in (("cuda", cuda), ("oracle", oracle))}
    scan = {}
    for kind, pts_M in (("surf", room.surf_M), ("corner", room.corner_M)):
        n = len(pts_M)
        t = np.sort(rng.uniform(1e-6, hi - 1e-6, n))
        q, p, bad = ests["cuda"].GetSensorPoses(t, S_LtoI, p_LinI, 0.0)
        assert bad == 0
        pts_L = S.qrot(S.qconj(q), pts_M.astype(np.float64) - shift - p).astype(np.float32)
        in_G = {}
        for name, est in ests.items():
            in_G[name], bad = est.UndistortScan(t, pts_L, S_LtoI, p_LinI)
            assert bad == 0
        assert np.abs(in_G["cuda"] - in_G["oracle"]).max() < 4e-6  # float32 outputs: at most an ulp apart at |x| < 16
        # both associations get the device's undistorted scan: bit-exact outputs need bit-equal inputs
        scan[kind] = (pts_L, t, (in_G["cuda"].astype(np.float64) + shift).astype(np.float32))
    out = {}
    for name, lib in (("cuda", cuda), ("oracle", oracle)):
        est = ests[name]
        est.SetFixedIndex(2)
        fm = A.FeatureMap(room.map_surf, room.map_corner, lib)
        c = fm.FindCorrespondence(scan["corner"][0], scan["corner"][1], scan["corner"][2],
                                  scan["surf"][0], scan["surf"][1], scan["surf"][2])
        fm.close()
        assert len(c.plane_t) > 800 and len(c.line_t) > 150
        # planes / lines live in the shifted map frame: p_M = p_G + shift  (S_GtoM = identity, p_GinM = shift)
        est.AddLoamPlanes(c.plane_t, c.plane_point, c.plane_geo, I4, shift, S_LtoI, p_LinI, 50.0)
        est.AddLoamLines(c.line_t, c.line_point, c.line_geo, I4, shift, S_LtoI, p_LinI, 50.0)
        out[name] = (c, est.Evaluate())
        est.close()
    _assert_identical(out["cuda"][0], out["oracle"][0])
    Hc, bc, cc, _ = out["cuda"][1]
    Ho, bo, co, _ = out["oracle"][1]
    d = np.sqrt(np.maximum(np.diag(Ho), 1e-300))
    assert (np.abs(Hc - Ho) / np.outer(d, d)).max() < 1e-9
    assert (np.abs(bc - bo) / d).max() <= 1e-9 * (np.abs(bo) / d).max()
    assert abs(cc - co) <= 1e-9 * abs(co)


def _scan_from_trajectory(est, room, S_LtoI, p_LinI, shift, hi, rng):
    """The room's scan features as raw LiDAR points taken along the estimator's current trajectory, in scan
    (not time) order, like the ring-ordered feature clouds the reference extracts."""
    scan = {}
    for kind, pts_M in (("surf", room.surf_M), ("corner", room.corner_M)):
        n = len(pts_M)
        t = rng.uniform(1e-6, hi - 1e-6, n)
        q, p, bad = est.GetSensorPoses(t, S_LtoI, p_LinI, 0.0)
        assert bad == 0
        scan[kind] = (S.qrot(S.qconj(q), pts_M.astype(np.float64) - shift - p).astype(np.float32), t)
    return scan


@pytest.mark.gpu
@pytest.mark.parametrize("ds", [1, 2, 3])
def test_gpu_add_loam_from_scan(cuda, oracle, ds):
    """clic_add_loam_from_scan: undistortion, association, time sort, down-sampling and the factor batches without
    leaving the device, against the same chain composed from the oracle's pieces."""
    sw = S.make_window(12, n_lidar=10, n_imu=0, seed=32)
    w = sw.window
    S_LtoI, p_LinI = sw.lidar["S_LtoI"], sw.lidar["p_LinI"]
    hi = (w.n_knots - 3) * w.dt_ns * 1e-9
    room = S.make_feature_scene(n_map_surf=30000, n_map_corner=4000, n_surf=1700, n_corner=600, seed=5 + ds)
    shift = np.array([9.0, 7.5, 1.5])
    I4 = np.array([0, 0, 0, 1.0])
    out = {}
    scan = None
    for name, lib in (("cuda", cuda), ("oracle", oracle)):
        est = E.TrajectoryEstimator(w, E.default_options(lib), lib)
        est.SetFixedIndex(2)
        if scan is None:
            scan = _scan_from_trajectory(est, room, S_LtoI, p_LinI, shift, hi, np.random.default_rng(11))
        # the map frame is G + shift: give the map in G instead, so that the undistorted scan (in G) meets it
        fm = A.FeatureMap((room.map_surf.astype(np.float64) - shift).astype(np.float32),
                          (room.map_corner.astype(np.float64) - shift).astype(np.float32), lib)
        counts = fm.AddLoamFromScan(est, scan["corner"][0], scan["corner"][1], scan["surf"][0], scan["surf"][1],
                                    I4, np.zeros(3), S_LtoI, p_LinI, 50.0, cor_downsample=ds)
        out[name] = (counts, est.Evaluate())
        fm.close()
        est.close()
    assert out["cuda"][0] == out["oracle"][0]
    nl, npl, nd = out["cuda"][0]
    assert nd == 0 and nl > 150 // ds and npl > 600 // ds
    Hc, bc, cc, _ = out["cuda"][1]
    Ho, bo, co, _ = out["oracle"][1]
    d = np.sqrt(np.maximum(np.diag(Ho), 1e-300))
    assert (np.abs(Hc - Ho) / np.outer(d, d)).max() < 1e-9
    assert (np.abs(bc - bo) / d).max() <= 1e-9 * (np.abs(bo) / d).max()
    assert abs(cc - co) <= 1e-9 * abs(co)


def test_oracle_grid_search_equals_exhaustive_search(oracle):
    """The grid-accelerated search that only the timed CPU baseline uses returns the same records and neighbours."""
    sc = S.make_feature_scene(n_map_surf=8000, n_map_corner=1200, n_surf=900, n_corner=400, seed=13)
    sc.map_surf = np.concatenate([sc.map_surf, sc.map_surf[:300]])     # exact ties
    sc.surf_M[5] = np.nan
    ref, nn_ref = _run(oracle, sc)
    oracle.clic_oracle_set_association_search(1)
    try:
        grid, nn_grid = _run(oracle, sc)
    finally:
        oracle.clic_oracle_set_association_search(0)
    _assert_identical(ref, grid)
    for a, b in zip(nn_ref, nn_grid):
        assert np.array_equal(a, b)
