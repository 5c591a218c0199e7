"""Parity of the CUDA path (through the C ABI) against the oracle on the same seeded synthetic windows.
Bar (BASELINE.json north_star): indices bit-exact; H and b within 1e-9 relative (FP64)."""
import numpy as np
import pytest

from clic_b200 import estimator as E
from clic_b200 import synthetic as S

pytestmark = pytest.mark.gpu

REL = 1e-9


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def rel_scaled(Hg, Ho):
    """Scale-invariant error of H: |dH_ij| / sqrt(H_ii H_jj) (max-norm alone is dominated by the largest block)."""
    d = np.sqrt(np.maximum(np.diag(Ho), 1e-300))
    return (np.abs(Hg - Ho) / np.outer(d, d)).max()


def build(lib, sw, fixed=-1, unlock_bias=False, unlock_toff=False, bias_factor=True, pad=10_000_000,
          unlock_cam_toff=False, fixed_depth=True, image_weight=None):
    opt = E.default_options(lib)
    if image_weight is not None:
        opt.image_sqrt_info = image_weight
    if unlock_bias:
        opt.lock_ab = opt.lock_wb = 0
    if unlock_toff:
        opt.lock_t_offset[0] = 0
        opt.t_offset_padding_ns = pad
    if unlock_cam_toff:
        opt.lock_t_offset[2] = 0
        opt.t_offset_padding_ns = pad
    est = E.TrajectoryEstimator(sw.window, opt, lib)
    est.SetFixedIndex(fixed)
    dropped = S.add_all_factors(est, sw, bias_factor=bias_factor, fixed_depth=fixed_depth)
    return est, dropped


def compare_eval(cuda, oracle, sw, **kw):
    eg, dg = build(cuda, sw, **kw)
    eo, do = build(oracle, sw, **kw)
    assert dg == do
    Lg, Lo = eg.layout(), eo.layout()
    assert Lg.D == Lo.D
    for stream in (0, 1, 2, 3):
        sg, ug = eg.indices(stream)
        so, uo = eo.indices(stream)
        assert np.array_equal(sg, so)          # feature-to-knot indexing: bit-exact
        assert np.array_equal(ug, uo)
    Hg, bg, cg, fg = eg.Evaluate()
    Ho, bo, co, fo = eo.Evaluate()
    assert rel(Hg, Ho) < REL, rel(Hg, Ho)
    assert rel_scaled(Hg, Ho) < REL, rel_scaled(Hg, Ho)
    assert rel(bg, bo) < REL, rel(bg, bo)
    assert abs(cg - co) <= REL * abs(co)
    assert abs(fg - fo) <= REL * max(abs(fo), 1.0)
    assert np.array_equal(Hg, Hg.T)
    # a large window takes the per-stream kernels on the first pass after the upload and the fused pass from then on
    # (run_pass: upload overlap): hold the second pass to the same bar
    H2, b2, c2, f2 = eg.Evaluate()
    assert rel(H2, Ho) < REL and rel_scaled(H2, Ho) < REL and rel(b2, bo) < REL and abs(c2 - co) <= REL * abs(co)
    assert rel(H2, Hg) < 1e-12 and rel(b2, bg) < 1e-12
    return Hg, bg


def test_c1_lidar_planes_imu(cuda, oracle):
    """BASELINE config 1: 10 knots, 2k point-to-plane + 200 IMU, everything but the knots locked."""
    compare_eval(cuda, oracle, S.config(1), bias_factor=False)


def test_c1_fixed_knots_and_bias(cuda, oracle):
    compare_eval(cuda, oracle, S.config(1), fixed=3, unlock_bias=True)


def test_lidar_only_lines(cuda, oracle):
    sw = S.make_window(10, n_lidar=3000, n_imu=0, line_fraction=1.0, seed=5)
    compare_eval(cuda, oracle, sw)


def test_c2_mixed_time_offset(cuda, oracle):
    """BASELINE config 2 at 1/10 size: planes+lines, Cauchy IMU, IMU time offset unlocked (arrow column)."""
    sw = S.config(2, 0.1)
    compare_eval(cuda, oracle, sw, unlock_bias=True, unlock_toff=True)


def test_large_rotations_random_knots(cuda, oracle):
    sw = S.make_window(12, n_lidar=1500, n_imu=150, line_fraction=0.4, seed=9, random_knots=True)
    compare_eval(cuda, oracle, sw, unlock_bias=True)


def test_ragged_and_out_of_window(cuda, oracle):
    """Counts that are not multiples of 32, measurements outside the window (silently dropped identically)."""
    sw = S.make_window(10, n_lidar=777, n_imu=45, line_fraction=0.5, seed=3)
    sw.lidar["t_point"] = sw.lidar["t_point"].copy()
    sw.lidar["t_point"][:5] = -0.01
    sw.lidar["t_point"][-7:] = 0.7 + np.arange(7) * 1e-3
    sw.imu["timestamp"] = sw.imu["timestamp"].copy()
    sw.imu["timestamp"][-3:] = 0.71
    compare_eval(cuda, oracle, sw)


def test_single_factor_and_tiny(cuda, oracle):
    sw = S.make_window(10, n_lidar=1, n_imu=1, seed=4)
    compare_eval(cuda, oracle, sw)


@pytest.mark.parametrize("n_lidar", [4000, 40000])
def test_unsorted_input(cuda, oracle, n_lidar):
    """Time-unsorted factors.  Up to 32 k records per batch the adder buckets them by knot interval (key-aligned
    storage); above that they go to the kernels as they are and exercise the multi-key warp loop and the atomics
    overflow path."""
    sw = S.make_window(10, n_lidar=n_lidar, n_imu=300, line_fraction=0.3, seed=8)
    rng = np.random.default_rng(0)
    perm = rng.permutation(n_lidar)
    for k in ("t_point", "point", "geo_type", "geo_plane", "geo_normal", "geo_point"):
        sw.lidar[k] = np.ascontiguousarray(sw.lidar[k][perm])
    perm = rng.permutation(300)
    for k in ("timestamp", "gyro", "accel"):
        sw.imu[k] = np.ascontiguousarray(sw.imu[k][perm])
    eg, _ = build(cuda, sw)
    eo, _ = build(oracle, sw)
    Hg, bg, cg, _ = eg.Evaluate()
    Ho, bo, co, _ = eo.Evaluate()
    assert rel(Hg, Ho) < REL and rel(bg, bo) < REL and abs(cg - co) <= REL * abs(co)


def test_40_knot_window(cuda, oracle):
    """BASELINE config 4 geometry (40 knots, D = 240) without the prior."""
    compare_eval(cuda, oracle, S.config(4, 0.25), fixed=-1)


def test_deterministic(cuda):
    sw = S.config(1)
    e1, _ = build(cuda, sw)
    H1, b1, c1, _ = e1.Evaluate()
    for _ in range(3):
        H2, b2, c2, _ = e1.Evaluate()
        assert np.array_equal(H1, H2) and np.array_equal(b1, b2) and c1 == c2


def test_prior_factor(cuda, oracle):
    """MarginalizationFactor contribution with a synthetic dense prior over knots + bias."""
    from clic_b200._abi import BLOCK_BIAS_ACCEL, BLOCK_BIAS_GYRO, BLOCK_KNOT_POS, BLOCK_KNOT_ROT
    sw = S.config(1)
    rng = np.random.default_rng(1)
    kinds, ids, x0 = [], [], []
    for k in range(2, 7):
        kinds += [BLOCK_KNOT_ROT, BLOCK_KNOT_POS]
        ids += [k, k]
        q = S.qmul(sw.window.knot_q[k], S.so3_exp(rng.normal(0, 0.01, 3)))
        x0 += [q, np.concatenate([sw.window.knot_p[k] + rng.normal(0, 0.01, 3), [0]])]
    kinds += [BLOCK_BIAS_GYRO, BLOCK_BIAS_ACCEL]
    ids += [1, 1]
    x0 += [np.concatenate([sw.window.bias[1, :3] + 1e-3, [0]]), np.concatenate([sw.window.bias[1, 3:] - 1e-3, [0]])]
    n = 5 * 6 + 6
    J0 = rng.normal(size=(n, n)) * 30
    r0 = rng.normal(size=n)
    res = {}
    for name, lib in (("g", cuda), ("o", oracle)):
        pr = E.Prior.create(lib, J0, r0, kinds, ids, np.array(x0))
        opt = E.default_options(lib)
        opt.lock_ab = opt.lock_wb = 0
        est = E.TrajectoryEstimator(sw.window, opt, lib)
        est.SetFixedIndex(3)
        est.AddMarginalizationFactor(pr)
        S.add_all_factors(est, sw)
        res[name] = est.Evaluate()
        A, g = pr.normal()
        res[name + "n"] = (A, g)
    assert rel(res["g"][0], res["o"][0]) < REL
    assert rel(res["g"][1], res["o"][1]) < REL
    assert abs(res["g"][2] - res["o"][2]) <= REL * abs(res["o"][2])
    assert rel(res["gn"][0], res["on"][0]) < 1e-12 and rel(res["gn"][1], res["on"][1]) < 1e-12


def test_c3_lidar_inertial_camera(cuda, oracle):
    """BASELINE config 3 at 1/10 size: LiDAR + 2k-scale visual reprojection factors (inverse-depth landmarks, fixed) +
    IMU with bias factor.  Visual factors create the only off-band H blocks."""
    sw = S.config(3, 0.1)
    assert len(sw.image["t_i"]) > 100
    compare_eval(cuda, oracle, sw, unlock_bias=True)


def test_visual_only_camera_time_offset(cuda, oracle):
    sw = S.make_window(10, n_lidar=0, n_imu=0, n_landmarks=60, obs_per_landmark=8, seed=41, time_margin_ns=3_000_000,
                       t_offset_ns=(0.0, 0.0, 150_000.3))
    compare_eval(cuda, oracle, sw, unlock_cam_toff=True, pad=1_000_000, bias_factor=False)


def test_visual_unsorted_and_fixed_knots(cuda, oracle):
    sw = S.make_window(10, n_lidar=300, n_imu=40, n_landmarks=80, obs_per_landmark=7, seed=43)
    perm = np.random.default_rng(2).permutation(len(sw.image["t_i"]))
    for k in ("t_i", "p_i", "t_j", "p_j", "landmark"):
        sw.image[k] = np.ascontiguousarray(sw.image[k][perm])
    compare_eval(cuda, oracle, sw, fixed=4)


@pytest.mark.gpu
@pytest.mark.parametrize("unlock_cam_toff", [False, True])
def test_visual_per_factor_chain_path(cuda, oracle, monkeypatch, unlock_cam_toff):
    """Camera frames share timestamps, so the visual kernel normally composes per-timestamp pose packs.  Per-feature
    times (rolling shutter: more distinct times than packs) take the per-factor chain kernel; so does the switch."""
    sw = S.make_window(10, n_lidar=0, n_imu=0, n_landmarks=90, obs_per_landmark=6, seed=47, time_margin_ns=3_000_000,
                       t_offset_ns=(0.0, 0.0, 90_000.6 if unlock_cam_toff else 0.0))
    kw = dict(unlock_cam_toff=unlock_cam_toff, pad=1_000_000, bias_factor=False)
    H_pack, b_pack = compare_eval(cuda, oracle, sw, **kw)
    monkeypatch.setenv("CLIC_NO_POSE_PACKS", "1")
    H_chain, b_chain = compare_eval(cuda, oracle, sw, **kw)
    monkeypatch.delenv("CLIC_NO_POSE_PACKS")
    assert rel(H_pack, H_chain) < 1e-11 and rel(b_pack, b_chain) < 1e-11
    rng = np.random.default_rng(5)
    n = len(sw.image["t_i"])
    sw.image["t_i"] = sw.image["t_i"] + rng.uniform(-4e-4, 4e-4, n)   # > 96 distinct times
    sw.image["t_j"] = sw.image["t_j"] + rng.uniform(-4e-4, 4e-4, n)
    compare_eval(cuda, oracle, sw, **kw)


@pytest.mark.gpu
def test_packed_lidar_input_is_identical(cuda):
    """clic_add_loam_packed (81 B/record) builds bit-identical normal equations to clic_add_loam."""
    sw = S.make_window(12, n_lidar=5000, n_imu=0, line_fraction=0.4, seed=77)
    L = sw.lidar
    out = []
    for packed in (False, True):
        est = E.TrajectoryEstimator(sw.window, E.default_options(cuda), cuda)
        est.SetFixedIndex(1)
        if packed:
            P = S.pack_lidar(L)
            nd = est.AddLoamMeasurementPacked(P["t_point"], P["point"], P["geo_type_u8"], P["geo6"], L["S_GtoM"],
                                              L["p_GinM"], L["S_LtoI"], L["p_LinI"], L["weight"])
        else:
            nd = est.AddLoamMeasurementAnalytic(L["t_point"], L["point"], L["geo_type"], L["geo_plane"],
                                                L["geo_normal"], L["geo_point"], L["S_GtoM"], L["p_GinM"],
                                                L["S_LtoI"], L["p_LinI"], L["weight"])
        out.append((nd,) + tuple(est.Evaluate()))
        est.close()
    assert out[0][0] == out[1][0]
    for a, b in zip(out[0][1:], out[1][1:]):
        assert np.array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.gpu
def test_type_uniform_lidar_batches(cuda, oracle):
    """clic_add_loam_planes + clic_add_loam_lines (what the C++ shim stages; type-uniform kernels) against the mixed
    stream and against the oracle fed the same two batches: same drops, same H, b, cost."""
    sw = S.make_window(12, n_lidar=6000, n_imu=0, line_fraction=0.35, seed=78)
    L = sw.lidar
    L["t_point"][5] = -1.0          # one dropped plane / line each
    L["t_point"][11] = 1e3
    sp = S.split_lidar(L)
    ext = (L["S_GtoM"], L["p_GinM"], L["S_LtoI"], L["p_LinI"], L["weight"])
    out = {}
    for name, lib, split in (("mixed", cuda, False), ("split", cuda, True), ("oracle", oracle, True)):
        est = E.TrajectoryEstimator(sw.window, E.default_options(lib), lib)
        est.SetFixedIndex(1)
        if split:
            nd = est.AddLoamPlanes(sp["pl_t"], sp["pl_point"], sp["pl_geo"], *ext)
            nd += est.AddLoamLines(sp["ln_t"], sp["ln_point"], sp["ln_geo"], *ext)
        else:
            nd = est.AddLoamMeasurementAnalytic(L["t_point"], L["point"], L["geo_type"], L["geo_plane"], L["geo_normal"],
                                                L["geo_point"], *ext)
        out[name] = (nd,) + tuple(est.Evaluate())
        est.close()
    assert out["mixed"][0] == out["split"][0] == out["oracle"][0] == 2
    for ref in ("mixed", "oracle"):
        assert rel(out["split"][1], out[ref][1]) < REL and rel(out["split"][2], out[ref][2]) < REL
        assert abs(out["split"][3] - out[ref][3]) <= REL * abs(out[ref][3])


@pytest.mark.gpu
@pytest.mark.parametrize("unlock_cam_toff", [False, True])
def test_c3_free_inverse_depths(cuda, oracle, unlock_cam_toff):
    """BASELINE config 3, fixed_depth = false variant: every landmark's inverse depth is a parameter (+200 columns;
    image_feature_factor.h:238-247).  The landmark columns are accumulated per factor (atomics into per-landmark side
    rows), everything else through the per-key tiles."""
    sw = S.config(3, 0.1)
    if unlock_cam_toff:
        sw.window.t_offset_ns = np.array([0.0, 0.0, 80_000.3])
    eg, _ = build(cuda, sw, unlock_bias=True, fixed_depth=False, unlock_cam_toff=unlock_cam_toff, pad=1_000_000)
    Lg = eg.layout()
    # a landmark whose observations were all dropped (time-offset padding at the window edge) adds no block
    assert 0 < Lg.n_free_landmarks <= len(sw.window.inv_depth) and Lg.D == Lg.col_inv_depth + Lg.n_free_landmarks
    compare_eval(cuda, oracle, sw, unlock_bias=True, fixed_depth=False, unlock_cam_toff=unlock_cam_toff, pad=1_000_000)
    # per-timestamp pose packs off: the per-factor chain kernel scatters the same landmark terms
    import os
    os.environ["CLIC_NO_POSE_PACKS"] = "1"
    try:
        compare_eval(cuda, oracle, sw, unlock_bias=True, fixed_depth=False, unlock_cam_toff=unlock_cam_toff, pad=1_000_000)
    finally:
        del os.environ["CLIC_NO_POSE_PACKS"]


@pytest.mark.gpu
def test_key_aligned_storage_is_a_layout_change_only(tmp_path):
    """Small time-sorted batches are stored with every knot interval's factors on a 32-factor slab boundary
    (build_key_gather).  It is a layout change only: the index probe must not move by a bit and H, b, cost only by the
    rounding of a different summation grouping, against the plain layout (CLIC_NO_KEY_ALIGN=1, read once per process,
    hence the two subprocesses)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from clic_b200 import estimator as E, synthetic as S
lib = E.cuda_library()
sw = S.make_window(14, n_lidar=700, n_imu=125, n_landmarks=12, obs_per_landmark=6, line_fraction=0.4, seed=77)
opt = E.default_options(lib); opt.lock_ab = opt.lock_wb = 0
est = E.TrajectoryEstimator(sw.window, opt, lib); est.SetFixedIndex(1)
S.add_all_factors(est, sw)
H, b, c, _ = est.Evaluate()
s0, u0 = est.indices(0); s1, u1 = est.indices(1)
np.savez(sys.argv[1], H=H, b=b, c=c, s0=s0, u0=u0, s1=s1, u1=u1)
''' % root
    out = {}
    for name, env in (("aligned", {}), ("plain", {"CLIC_NO_KEY_ALIGN": "1"})):
        path = str(tmp_path / (name + ".npz"))
        r = subprocess.run([sys.executable, "-c", code, path], env={**os.environ, **env}, capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[name] = np.load(path)
    for k in ("s0", "u0", "s1", "u1"):
        assert np.array_equal(out["aligned"][k], out["plain"][k]), k
    Ha, Hp = out["aligned"]["H"], out["plain"]["H"]
    d = np.sqrt(np.maximum(np.diag(Hp), 1e-300))
    assert (np.abs(Ha - Hp) / np.outer(d, d)).max() < 1e-13
    assert (np.abs(out["aligned"]["b"] - out["plain"]["b"]) / d).max() <= 1e-12 * (np.abs(out["plain"]["b"]) / d).max()
    assert abs(out["aligned"]["c"] - out["plain"]["c"]) <= 1e-13 * abs(out["plain"]["c"])


@pytest.mark.parametrize("weight", [200.0, 450.0])
def test_c3_image_weight_from_config(cuda, oracle, weight):
    """The visual factor's sqrt_info as every shipped config sets it (trajectory_manager.cpp:53-61: image_weight = 200 /
    450), not the class default; the option itself is pinned on the CPU against auto-diff
    (tests/test_oracle_autodiff.py::test_image_feature_weight_from_config)."""
    sw = S.config(3, 0.1)
    Hw, bw = compare_eval(cuda, oracle, sw, unlock_bias=True, image_weight=weight)
    Hd, bd = compare_eval(cuda, oracle, sw, unlock_bias=True)
    assert rel_scaled(Hw, Hd) > 1e-4  # the weight does change the system
