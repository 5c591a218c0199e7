"""Deterministic synthetic fixed-lag windows for parity tests and bench.py.

Host-side data generation only (numpy).  The vectorised forward spline below follows the forward evaluators
of the reference (src/spline/so3_spline.h:220-337, rd_spline.h:218-244) and is used solely to synthesise
measurements that are consistent with a ground-truth trajectory; it is not part of the timed path.
"""
from dataclasses import dataclass, field

import numpy as np

from ._abi import GEO_LINE, GEO_PLANE
from .estimator import Window

M_CUM = np.array([[6, 0, 0, 0], [5, 3, -3, 1], [1, 3, 3, -2], [0, 0, 0, 1]], dtype=np.float64) / 6.0
M_PLAIN = np.array([[1, -3, 3, -1], [4, 0, -6, 3], [1, 3, 3, -3], [0, 0, 0, 1]], dtype=np.float64) / 6.0


# ---------- quaternion helpers, (..., 4) in (x, y, z, w) order ----------
def qmul(a, b):
    ax, ay, az, aw = np.moveaxis(a, -1, 0)
    bx, by, bz, bw = np.moveaxis(b, -1, 0)
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def qconj(a):
    return a * np.array([-1.0, -1.0, -1.0, 1.0])


def qrot(q, v):
    qv = q[..., :3]
    uv = 2.0 * np.cross(qv, v)
    return v + q[..., 3:4] * uv + np.cross(qv, uv)


def so3_exp(w):
    th2 = np.sum(w * w, axis=-1, keepdims=True)
    th = np.sqrt(th2)
    small = th < 1e-10
    ths = np.where(small, 1.0, th)
    imag = np.where(small, 0.5 - th2 / 48.0, np.sin(0.5 * ths) / ths)
    real = np.where(small, 1.0 - th2 / 8.0, np.cos(0.5 * th))
    return np.concatenate([imag * w, real], axis=-1)


def so3_log(q):
    v, w = q[..., :3], q[..., 3:4]
    n2 = np.sum(v * v, axis=-1, keepdims=True)
    n = np.sqrt(n2)
    small = n < 1e-10
    ns = np.where(small, 1.0, n)
    f = np.where(small, 2.0 / w - 2.0 * n2 / (w * w * w), 2.0 * np.arctan(ns / w) / ns)
    return f * v


def quat_to_matrix(q):
    x, y, z, w = np.moveaxis(q, -1, 0)
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def matrix_to_quat(R):
    """Single 3x3 rotation matrix -> (x,y,z,w)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    return q / np.linalg.norm(q)


class SplineNP:
    """Vectorised forward evaluation of the uniform cubic SO(3) x R^3 spline."""

    def __init__(self, t0_ns, dt_ns, knot_q, knot_p):
        self.t0_ns, self.dt_ns = int(t0_ns), int(dt_ns)
        self.q, self.p = np.asarray(knot_q, np.float64), np.asarray(knot_p, np.float64)
        self.inv_dt = 1e9 / self.dt_ns

    def index(self, t_ns):
        st = np.asarray(t_ns, np.int64) - self.t0_ns
        s = st // self.dt_ns
        u = (st % self.dt_ns).astype(np.float64) / float(self.dt_ns)
        return s, u

    @staticmethod
    def _pows(u, deriv):
        one, zero = np.ones_like(u), np.zeros_like(u)
        if deriv == 0:
            return np.stack([one, u, u * u, u * u * u], -1)
        if deriv == 1:
            return np.stack([zero, one, 2 * u, 3 * u * u], -1)
        if deriv == 2:
            return np.stack([zero, zero, 2 * one, 6 * u], -1)
        return np.stack([zero, zero, zero, 6 * one], -1)

    def position(self, t_ns, deriv=0):
        s, u = self.index(t_ns)
        c = (self._pows(u, deriv) @ M_PLAIN.T) * self.inv_dt ** deriv
        out = np.zeros(u.shape + (3,))
        for i in range(4):
            out += c[..., i:i + 1] * self.p[s + i]
        return out

    def rotation(self, t_ns):
        s, u = self.index(t_ns)
        c = self._pows(u, 0) @ M_CUM.T
        r = self.q[s]
        for i in range(3):
            d = so3_log(qmul(qconj(self.q[s + i]), self.q[s + i + 1]))
            r = qmul(r, so3_exp(c[..., i + 1:i + 2] * d))
        return r

    def omega_body(self, t_ns):
        s, u = self.index(t_ns)
        c = self._pows(u, 0) @ M_CUM.T
        dc = (self._pows(u, 1) @ M_CUM.T) * self.inv_dt
        w = np.zeros(u.shape + (3,))
        for i in range(3):
            d = so3_log(qmul(qconj(self.q[s + i]), self.q[s + i + 1]))
            w = qrot(so3_exp(-c[..., i + 1:i + 2] * d), w) + dc[..., i + 1:i + 2] * d
        return w


# ---------- window generation ----------
def gt_motion(t):
    t = np.asarray(t, np.float64)
    p = np.stack([2 * np.sin(0.7 * t), 1.5 * np.cos(0.5 * t), 0.3 * np.sin(1.1 * t)], -1)
    w = np.stack([0.3 * np.sin(0.9 * t), 0.2 * np.cos(0.6 * t), 0.5 * t], -1)
    return so3_exp(w), p


def unit_sphere(rng, n):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


# config/ct_odometry_ntu.yaml:27-31 + OptWeight::LoadWeight (src/utils/opt_weight.h:114-128)
IMU_FREQ = 385.0
SIGMA_W, SIGMA_WB, SIGMA_A, SIGMA_AB = 6.1e-5, 3.4e-6, 1.4e-3, 3.9e-4
_SQRT_DT = np.sqrt(1.0 / IMU_FREQ)
IMU_INFO_VEC = np.array([1.0 / (SIGMA_W / _SQRT_DT)] * 3 + [1.0 / (SIGMA_A / _SQRT_DT)] * 3)
BIAS_INFO_VEC = np.array([1.0 / SIGMA_WB] * 3 + [1.0 / SIGMA_AB] * 3)
LIDAR_WEIGHT = 50.0  # config/ct_odometry_ntu.yaml:34
# config/ct_odometry_ntu.yaml:39-45
CAM_ROT = np.array([[0.02183084, -0.01312053, 0.99967558], [0.99975965, 0.00230088, -0.02180248],
                    [-0.00201407, 0.99991127, 0.01316761]])
CAM_TRANS = np.array([0.00552943, -0.12431302, 0.01614686])


@dataclass
class SyntheticWindow:
    window: Window            # initial estimate (what the estimator is created with)
    gt_knot_q: np.ndarray
    gt_knot_p: np.ndarray
    lidar: dict = field(default_factory=dict)
    imu: dict = field(default_factory=dict)
    image: dict = field(default_factory=dict)
    name: str = ""

    @property
    def n_factors(self):
        return (len(self.lidar.get("t_point", ())) + len(self.imu.get("timestamp", ())) +
                len(self.image.get("t_i", ())))


def make_window(n_knots=10, n_lidar=2000, n_imu=200, n_landmarks=0, obs_per_landmark=10, line_fraction=0.0,
                seed=20230216, dt_ns=100_000_000, random_knots=False, init_noise=(0.02, 0.05), n_bias=2,
                t_offset_ns=(0.0, 0.0, 0.0), bias_true=(0.01, -0.02, 0.015, 0.05, -0.03, 0.04), name="",
                time_margin_ns=0, meas_seed=None):
    """One synthetic LiDAR(-inertial(-camera)) window.  `time_margin_ns` keeps measurements that far inside the
    valid range (needed when a time offset is unlocked: the padding test of MeasuredTimeToNs)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    K = n_knots
    t0_ns = 0
    if random_knots:  # genRandomTrajectory statistics (so3_spline.h:149-159, rd_spline.h:156-165)
        gt_q = so3_exp(rng.uniform(-np.pi, np.pi, size=(K, 3)))
        gt_p = rng.uniform(-5, 5, size=(K, 3))
    else:
        tk = (t0_ns + (np.arange(K) - 1) * dt_ns) * 1e-9
        gt_q, gt_p = gt_motion(tk)
    gt = SplineNP(t0_ns, dt_ns, gt_q, gt_p)
    lo = t0_ns + time_margin_ns
    hi = t0_ns + (K - 3) * dt_ns - 2 - time_margin_ns  # inclusive; valid iff t < t0 + (K-3)*dt - 1 + 1

    S_LtoI = so3_exp(np.array([0.01, -0.02, 0.03]))
    p_LinI = np.array([-0.05, 0.0, 0.055])  # config/ntu/lidar_ntu.yaml:10
    S_CtoI = matrix_to_quat(CAM_ROT)
    p_CinI = CAM_TRANS.copy()
    gravity = np.array([0.0, 0.0, 9.8])  # test_analytic_factor.cpp:1020
    bias_true = np.asarray(bias_true, np.float64)
    toff = np.asarray(t_offset_ns, np.float64)

    # initial estimate
    init_q = qmul(gt_q, so3_exp(rng.normal(0, init_noise[0], size=(K, 3))))
    init_p = gt_p + rng.normal(0, init_noise[1], size=(K, 3))
    bias0 = np.tile(bias_true, (n_bias, 1)) + rng.normal(0, 1e-3, size=(n_bias, 6)) if n_bias else np.zeros((0, 6))

    sw = SyntheticWindow(window=None, gt_knot_q=gt_q, gt_knot_p=gt_p, name=name)
    if meas_seed is not None:  # same window state, different measurements (one shard per rank under weak scaling)
        rng = np.random.Generator(np.random.PCG64(meas_seed))

    if n_lidar:
        t_ns = np.sort(rng.integers(lo, hi + 1, size=n_lidar, dtype=np.int64))
        t_s = t_ns.astype(np.float64) * 1e-9
        t_eff = (t_s * 1e9).astype(np.int64)  # what the estimator will see after its own truncation
        p_L = unit_sphere(rng, n_lidar) * rng.uniform(2, 50, size=(n_lidar, 1))
        R, p = gt.rotation(t_eff), gt.position(t_eff)
        p_G = qrot(R, qrot(S_LtoI, p_L) + p_LinI) + p
        geo_type = np.where(rng.uniform(size=n_lidar) < line_fraction, GEO_LINE, GEO_PLANE).astype(np.int32)
        nrm = unit_sphere(rng, n_lidar)
        d = -np.sum(nrm * p_G, axis=1) + rng.normal(0, 0.02, size=n_lidar)
        geo_plane = np.concatenate([nrm, d[:, None]], axis=1)
        ldir = unit_sphere(rng, n_lidar)
        geo_point = p_G + rng.uniform(-2, 2, size=(n_lidar, 1)) * ldir + rng.normal(0, 0.02, size=(n_lidar, 3))
        sw.lidar = dict(t_point=t_s, point=p_L, geo_type=geo_type, geo_plane=geo_plane, geo_normal=ldir,
                        geo_point=geo_point, S_GtoM=np.array([0, 0, 0, 1.0]), p_GinM=np.zeros(3), S_LtoI=S_LtoI,
                        p_LinI=p_LinI, weight=LIDAR_WEIGHT)

    if n_imu:
        t_ns = np.linspace(lo + 1000, hi - 1000, n_imu).astype(np.int64)
        t_s = t_ns.astype(np.float64) * 1e-9
        t_eff = (t_s * 1e9).astype(np.int64) + np.int64(toff[0])
        w = gt.omega_body(t_eff)
        a = qrot(qconj(gt.rotation(t_eff)), gt.position(t_eff, 2) + gravity)
        gyro = w + bias_true[:3] + rng.normal(0, SIGMA_W / _SQRT_DT, size=(n_imu, 3))
        accel = a + bias_true[3:] + rng.normal(0, SIGMA_A / _SQRT_DT, size=(n_imu, 3))
        sw.imu = dict(timestamp=t_s, gyro=gyro, accel=accel, bias_slot=n_bias - 1, info_vec=IMU_INFO_VEC.copy())

    inv_depth = np.zeros(0)
    if n_landmarks:
        # frames every 0.05 s; a landmark is anchored in one frame (inverse depth) and observed in the others
        frame_ns = np.arange(lo + 2000, hi - 2000, 50_000_000, dtype=np.int64)
        nf = len(frame_ns)
        n_obs = min(obs_per_landmark, nf - 1)
        fi = rng.integers(0, nf, size=n_landmarks)
        depth = rng.uniform(3, 30, size=n_landmarks)
        xy = rng.uniform(-0.6, 0.6, size=(n_landmarks, 2))
        inv_depth = 1.0 / depth * (1.0 + rng.normal(0, 0.02, size=n_landmarks))
        eff = lambda t: (t.astype(np.float64) * 1e-9 * 1e9).astype(np.int64) + np.int64(toff[2])
        ti_ns = frame_ns[fi]
        x_ci = np.concatenate([xy, np.ones((n_landmarks, 1))], axis=1) * depth[:, None]
        P = qrot(gt.rotation(eff(ti_ns)), qrot(S_CtoI, x_ci) + p_CinI) + gt.position(eff(ti_ns))
        lm = np.repeat(np.arange(n_landmarks, dtype=np.int32), n_obs)
        fj = (np.repeat(fi, n_obs) + 1 + np.tile(np.arange(n_obs), n_landmarks)) % nf  # the n_obs frames after fi, cyclic
        tj_ns = frame_ns[fj]
        x_cj = qrot(qconj(S_CtoI), qrot(qconj(gt.rotation(eff(tj_ns))), P[lm] - gt.position(eff(tj_ns))) - p_CinI)
        ok = x_cj[:, 2] > 0.5
        obs = x_cj[:, :2] / x_cj[:, 2:3] + rng.normal(0, 1.0 / 450.0, size=(len(lm), 2))
        lm, tj_ns, obs = lm[ok], tj_ns[ok], obs[ok]
        ones = np.ones((len(lm), 1))
        sw.image = dict(t_i=ti_ns[lm].astype(np.float64) * 1e-9, p_i=np.concatenate([xy[lm], ones], axis=1),
                        t_j=tj_ns.astype(np.float64) * 1e-9, p_j=np.concatenate([obs, ones], axis=1), landmark=lm)

    sw.window = Window(t0_ns=t0_ns, dt_ns=dt_ns, knot_q=init_q, knot_p=init_p, S_LtoI=S_LtoI, p_LinI=p_LinI,
                       S_CtoI=S_CtoI, p_CinI=p_CinI, t_offset_ns=toff.copy(), gravity=gravity, bias=bias0,
                       inv_depth=inv_depth)
    return sw


def pack_lidar(L):
    """PointCorrespondence field arrays -> the packed input of clic_add_loam_packed."""
    plane = L["geo_type"] == GEO_PLANE
    geo6 = np.where(plane[:, None], np.concatenate([L["geo_plane"], np.zeros((len(plane), 2))], axis=1),
                    np.concatenate([L["geo_point"], L["geo_normal"]], axis=1))
    return dict(t_point=L["t_point"], point=L["point"], geo_type_u8=L["geo_type"].astype(np.uint8),
                geo6=np.ascontiguousarray(geo6))


def split_lidar(L):
    """PointCorrespondence field arrays -> a plane batch and a line batch (what the C++ shim stages), each time-sorted."""
    plane = L["geo_type"] == GEO_PLANE
    line = ~plane
    return dict(pl_t=np.ascontiguousarray(L["t_point"][plane]), pl_point=np.ascontiguousarray(L["point"][plane]),
                pl_geo=np.ascontiguousarray(L["geo_plane"][plane]),
                ln_t=np.ascontiguousarray(L["t_point"][line]), ln_point=np.ascontiguousarray(L["point"][line]),
                ln_geo=np.ascontiguousarray(np.concatenate([L["geo_point"][line], L["geo_normal"][line]], axis=1)))


def add_all_factors(est, sw, marg=False, fixed_depth=True, bias_factor=True, count_dropped=True, small_first=False):
    """Feed a synthetic window to an estimator in the reference's call order (trajectory_manager.cpp:682-750: LiDAR,
    image, IMU, bias) or, small_first, in the upload order of the C++ shim: plane batch (the big asynchronous copy
    goes first, the host-side staging of the visual batch runs under it), image, IMU, line batch (the kernel left
    exposed after the last byte arrives is the smallest one).  The normal equations do not depend on the order."""
    def lidar(part="all"):
        L = sw.lidar
        if "pl_t" in L:
            ext = (L["S_GtoM"], L["p_GinM"], L["S_LtoI"], L["p_LinI"], L["weight"], marg, count_dropped)
            n = 0
            if part in ("all", "planes") and len(L["pl_t"]):
                n += est.AddLoamPlanes(L["pl_t"], L["pl_point"], L["pl_geo"], *ext)
            if part in ("all", "lines") and len(L["ln_t"]):
                n += est.AddLoamLines(L["ln_t"], L["ln_point"], L["ln_geo"], *ext)
            return n
        if part == "lines":
            return 0
        if len(L.get("t_point", ())) and "geo6" in L:
            return est.AddLoamMeasurementPacked(L["t_point"], L["point"], L["geo_type_u8"], L["geo6"], L["S_GtoM"],
                                                L["p_GinM"], L["S_LtoI"], L["p_LinI"], L["weight"], marg, count_dropped)
        if len(L.get("t_point", ())):
            return est.AddLoamMeasurementAnalytic(L["t_point"], L["point"], L["geo_type"], L["geo_plane"],
                                                  L["geo_normal"], L["geo_point"], L["S_GtoM"], L["p_GinM"],
                                                  L["S_LtoI"], L["p_LinI"], L["weight"], marg, count_dropped)
        return 0

    def image():
        I = sw.image
        if not len(I.get("t_i", ())):
            return 0
        return est.AddImageFeatureAnalytic(I["t_i"], I["p_i"], I["t_j"], I["p_j"], I["landmark"], fixed_depth, marg,
                                           count_dropped)

    def imu():
        M = sw.imu
        if not len(M.get("timestamp", ())):
            return 0
        return est.AddIMUMeasurementAnalytic(M["timestamp"], M["gyro"], M["accel"], M["bias_slot"], M["info_vec"], marg,
                                             count_dropped)

    if small_first:
        dropped = lidar("planes") + image() + imu() + lidar("lines")
    else:
        dropped = lidar() + image() + imu()
    nb = sw.window.bias.shape[0]
    if bias_factor and nb >= 2:
        # trajectory_manager.cpp:738-750: cov = delta_time/dt * dt^2, sqrt_info = bias_info_vec / sqrt(cov), dt arg = 1
        dt = 1.0 / IMU_FREQ
        cov = 0.1 / dt * (dt * dt)
        est.AddBiasFactor(nb - 2, nb - 1, 1.0, BIAS_INFO_VEC / np.sqrt(cov), marg, False)
    return dropped


# BASELINE.json configs (SURVEY.md §8d "Sizes"); `scale` shrinks the factor counts for CPU-sized parity cases.
def config(c, scale=1.0, seed=None):
    seed = 20230215 + c if seed is None else seed
    n = lambda x: max(1, int(round(x * scale)))
    if c == 1:
        return make_window(10, n(2000), n(200), seed=seed, name="C1")
    if c == 2:
        return make_window(10, n(50000), n(200), line_fraction=0.3, seed=seed, name="C2",
                           t_offset_ns=(1.0e6, 0.0, 0.0), time_margin_ns=12_000_000)
    if c == 3:
        return make_window(10, n(50000), n(200), n_landmarks=n(200), obs_per_landmark=10, line_fraction=0.3,
                           seed=seed, name="C3")
    if c == 4:
        return make_window(40, n(20000), n(700), seed=seed, name="C4")
    if c == 5:
        return make_window(10, n(1_000_000), n(200_000), n_landmarks=n(2000), obs_per_landmark=10,
                           line_fraction=0.3, seed=seed, name="C5")
    raise ValueError(c)


# ---- LiDAR association scenes: a box room with poles, sampled as a feature map and as a scan ----
@dataclass
class FeatureScene:
    map_surf: np.ndarray      # (Ms, 3) float32, map frame
    map_corner: np.ndarray    # (Mc, 3) float32
    surf_L: np.ndarray        # (Ns, 3) float32, scan surface features in the LiDAR frame
    surf_M: np.ndarray        # the same features in the map frame (lf_cur_in_M)
    surf_t: np.ndarray        # (Ns,) double timestamps
    corner_L: np.ndarray
    corner_M: np.ndarray
    corner_t: np.ndarray


def _room_surface(rng, n, size, noise):
    """Points on the six faces of an axis-aligned box [0,size] with Gaussian noise along the face normal."""
    sx, sy, sz = size
    areas = np.array([sy * sz, sy * sz, sx * sz, sx * sz, sx * sy, sx * sy])
    face = rng.choice(6, n, p=areas / areas.sum())
    p = rng.uniform(0, 1, (n, 3)) * np.array(size)
    axis = face // 2
    side = face % 2
    p[np.arange(n), axis] = side * np.array(size)[axis] + rng.normal(0, noise, n)
    return p


def _room_edges(rng, n, size, noise, n_poles=6):
    """Points on the 12 box edges and on a few vertical poles inside the room."""
    sx, sy, sz = size
    segs = []
    for a in (0, sx):
        for b in (0, sy):
            segs.append(((a, b, 0), (a, b, sz)))
    for a in (0, sx):
        for c in (0, sz):
            segs.append(((a, 0, c), (a, sy, c)))
    for b in (0, sy):
        for c in (0, sz):
            segs.append(((0, b, c), (sx, b, c)))
    prng = np.random.default_rng(77)
    for _ in range(n_poles):
        x, y = prng.uniform(2, sx - 2), prng.uniform(2, sy - 2)
        segs.append(((x, y, 0), (x, y, sz)))
    segs = np.array(segs, dtype=np.float64)
    length = np.linalg.norm(segs[:, 1] - segs[:, 0], axis=1)
    which = rng.choice(len(segs), n, p=length / length.sum())
    u = rng.uniform(0, 1, (n, 1))
    return segs[which, 0] * (1 - u) + segs[which, 1] * u + rng.normal(0, noise, (n, 3))


def make_feature_scene(n_map_surf=20000, n_map_corner=3000, n_surf=1200, n_corner=400, seed=0, size=(20.0, 16.0, 6.0),
                       noise=0.01, outlier_fraction=0.05, negative_time_fraction=0.02):
    """A feature map and one scan of the same room.  The scan's map-frame copy carries a small pose error (what the
    association sees between the two inner iterations); a few features are far from the map, a few carry the
    negative timestamp the reference skips."""
    rng = np.random.default_rng(20230215 + seed)
    map_surf = _room_surface(rng, n_map_surf, size, noise).astype(np.float32)
    map_corner = _room_edges(rng, n_map_corner, size, noise).astype(np.float32)
    q_err = so3_exp(np.array([0.002, -0.001, 0.003]))
    p_err = np.array([0.02, -0.015, 0.01])
    q_LtoM = so3_exp(np.array([0.1, -0.05, 0.7]))
    p_LinM = np.array([9.0, 7.5, 1.5])

    def scan(pts):
        n = len(pts)
        out = rng.uniform(0, 1, n) < outlier_fraction
        pts = pts + out[:, None] * rng.normal(0, 3.0, (n, 3))
        in_M = qrot(np.broadcast_to(q_err, (n, 4)), pts) + p_err
        in_L = qrot(np.broadcast_to(qconj(q_LtoM), (n, 4)), pts - p_LinM)
        t = np.sort(rng.uniform(0.0, 0.1, n))
        t[rng.uniform(0, 1, n) < negative_time_fraction] = -1.0
        return in_L.astype(np.float32), in_M.astype(np.float32), t

    sL, sM, st = scan(_room_surface(rng, n_surf, size, noise))
    cL, cM, ct = scan(_room_edges(rng, n_corner, size, noise))
    return FeatureScene(map_surf, map_corner, sL, sM, st, cL, cM, ct)
