"""Inputs for the single-factor tests (clic_evaluate_factor): the values src/app/test_analytic_factor.cpp:903-1010
hard-codes, on the random 12-knot trajectory of that test.  Each case carries the C-ABI description of the factor AND
the torch-f64 residual of the independent auto-diff lineage (tests/autodiff_ref.py), so that
  oracle == autodiff      (CPU, tests/test_oracle_factors_extra.py) and
  CUDA   == oracle        (GPU, tests/test_gpu_factors_extra.py)
are checked on the same numbers."""
import numpy as np

import autodiff_ref as AD
from clic_b200 import _abi
from clic_b200 import synthetic as S

S_TO_NS = 1e9


def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def quat_to_matrix(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def midpoint_preintegration(acc0, gyr0, ba, bg, samples, noise=(0.08, 0.004, 4e-5, 2e-6), gravity_norm=-9.8):
    """IntegrationBase::push_back / midPointIntegration (src/visual_odometry/integration_base.h:74-190) in numpy:
    test-data generation only (the integration is the caller's job, not part of the estimator's hot path).
    noise = (ACC_N, GYR_N, ACC_W, GYR_W)."""
    acc_n, gyr_n, acc_w, gyr_w = noise
    Q = np.zeros((18, 18))
    for i, v in enumerate((acc_n, gyr_n, acc_n, gyr_n, acc_w, gyr_w)):
        Q[3 * i:3 * i + 3, 3 * i:3 * i + 3] = v * v * np.eye(3)
    dp, dv, dq = np.zeros(3), np.zeros(3), np.array([0, 0, 0, 1.0])
    J, P, sum_dt = np.eye(15), np.zeros((15, 15)), 0.0
    acc0, gyr0 = np.asarray(acc0, float), np.asarray(gyr0, float)
    for dt, acc1, gyr1 in samples:
        acc1, gyr1 = np.asarray(acc1, float), np.asarray(gyr1, float)
        w = 0.5 * (gyr0 + gyr1) - bg
        half = 0.5 * w * dt                     # Utility::deltaQ: (1, theta/2), not normalised by the reference either
        dq1 = S.qmul(dq, np.array([half[0], half[1], half[2], 1.0]))
        R0, R1 = quat_to_matrix(dq / np.linalg.norm(dq)), quat_to_matrix(dq1 / np.linalg.norm(dq1))
        a0, a1 = acc0 - ba, acc1 - ba
        un_acc = 0.5 * (R0 @ a0 + R1 @ a1)
        dp1 = dp + dv * dt + 0.5 * un_acc * dt * dt
        dv1 = dv + un_acc * dt
        Rw, Ra0, Ra1, I = hat(w), hat(a0), hat(a1), np.eye(3)
        F = np.zeros((15, 15))
        F[0:3, 0:3] = I
        F[0:3, 3:6] = -0.25 * R0 @ Ra0 * dt * dt - 0.25 * R1 @ Ra1 @ (I - Rw * dt) * dt * dt
        F[0:3, 6:9] = I * dt
        F[0:3, 9:12] = -0.25 * (R0 + R1) * dt * dt
        F[0:3, 12:15] = 0.25 * R1 @ Ra1 * dt ** 3
        F[3:6, 3:6] = I - Rw * dt
        F[3:6, 12:15] = -I * dt
        F[6:9, 3:6] = -0.5 * R0 @ Ra0 * dt - 0.5 * R1 @ Ra1 @ (I - Rw * dt) * dt
        F[6:9, 6:9] = I
        F[6:9, 9:12] = -0.5 * (R0 + R1) * dt
        F[6:9, 12:15] = 0.5 * R1 @ Ra1 * dt * dt
        F[9:12, 9:12] = I
        F[12:15, 12:15] = I
        G = np.zeros((15, 18))
        G[0:3, 0:3] = 0.25 * R0 * dt * dt
        G[0:3, 3:6] = -0.125 * R1 @ Ra1 * dt ** 3
        G[0:3, 6:9] = 0.25 * R1 * dt * dt
        G[0:3, 9:12] = G[0:3, 3:6]
        G[3:6, 3:6] = 0.5 * I * dt
        G[3:6, 9:12] = 0.5 * I * dt
        G[6:9, 0:3] = 0.5 * R0 * dt
        G[6:9, 3:6] = -0.25 * R1 @ Ra1 * dt * dt
        G[6:9, 6:9] = 0.5 * R1 * dt
        G[6:9, 9:12] = G[6:9, 3:6]
        G[9:12, 12:15] = I * dt
        G[12:15, 15:18] = I * dt
        J = F @ J
        P = F @ P @ F.T + G @ Q @ G.T
        dp, dv, dq = dp1, dv1, dq1 / np.linalg.norm(dq1)
        sum_dt += dt
        acc0, gyr0 = acc1, gyr1
    sqrt_info = np.linalg.cholesky(np.linalg.inv(P)).T  # LLT(covariance^-1).matrixL().transpose()
    return dict(sum_dt=sum_dt, delta_p=dp, delta_q=dq, delta_v=dv, linearized_ba=np.asarray(ba, float),
                linearized_bg=np.asarray(bg, float), jacobian=J, sqrt_info=sqrt_info, gravity_norm=gravity_norm)


def preintegration_struct(pre):
    p = _abi.PreIntegration()
    p.sum_dt, p.gravity_norm = float(pre["sum_dt"]), float(pre["gravity_norm"])
    for name in ("delta_p", "delta_q", "delta_v", "linearized_ba", "linearized_bg"):
        for i, v in enumerate(pre[name]):
            getattr(p, name)[i] = float(v)
    for i, v in enumerate(np.asarray(pre["jacobian"]).ravel()):
        p.jacobian[i] = float(v)
    for i, v in enumerate(np.asarray(pre["sqrt_info"]).ravel()):
        p.sqrt_info[i] = float(v)
    return p


def reference_preintegration():
    """test_analytic_factor.cpp:919-927."""
    return midpoint_preintegration([0.11, 0.12, -9.83], [0.1, 0.2, -0.1], np.array([0.01, 0.02, -0.01]),
                                   np.array([0.01, 0.20, -0.01]),
                                   [(0.025, [0.1, 0.2, -9.83], [0.1, 0.2, 0.3]), (0.46, [0.6, 0.4, -9.83], [0.3, 0.1, 0.17])])


def ns(t):
    return int(t * S_TO_NS)


def cases(w, seed=0):
    """All single-factor cases on window w (E.Window): list of dicts
    name, kind, t_a, t_b, geo, d, x, pre (dict or None), fn (torch residual), extras (autodiff variables, in x order),
    x_cols (tangent width of each extra block)."""
    rng = np.random.default_rng(100 + seed)
    wd = dict(t0_ns=w.t0_ns, dt_ns=w.dt_ns, knot_q=w.knot_q, knot_p=w.knot_p)
    rand_rot = lambda: S.so3_exp(rng.uniform(-np.pi, np.pi, size=3))
    out = []
    toff = 1.0e6
    info6 = np.array([0.33, 0.27, 0.11, 0.7, 0.9, 1.3])
    # IMUPoseFactor (:604-660)
    pose_p, pose_q = rng.uniform(-2, 2, 3), rand_rot()
    out.append(dict(name="pose", kind=_abi.FACTOR_POSE, t_a=ns(0.55), d=np.concatenate([pose_p, pose_q, info6]), x=[toff],
                    fn=AD.pose_residual(wd, ns(0.55), pose_p, pose_q, info6), extras=[[toff]], x_cols=[1],
                    # the analytic time-offset column of this factor is not the derivative (quirk, see
                    # tests/test_oracle_autodiff.py::test_pose_factor): compared on the knots only
                    skip_extra=[0]))
    # LocalVelocityFactor (:662-720)
    lv = np.array([0.2, 0.4, 0.5])
    out.append(dict(name="local_velocity", kind=_abi.FACTOR_LOCAL_VELOCITY, t_a=ns(0.23), d=np.concatenate([lv, [0.9]]),
                    x=[toff], fn=AD.local_velocity_residual(wd, ns(0.23), lv, 0.9), extras=[[toff]], x_cols=[1]))
    # RelativeOrientationFactor (:722-780), both time pairs of the test
    S_BtoA, info3 = rand_rot(), np.array([0.33, 0.27, 0.11])
    for i, (ta, tb) in enumerate([(0.11, 0.33), (0.21, 0.83)]):
        out.append(dict(name=f"relative_rotation{i}", kind=_abi.FACTOR_RELATIVE_ROTATION, t_a=ns(ta), t_b=ns(tb),
                        d=np.concatenate([S_BtoA, info3]), x=[],
                        fn=AD.relative_rotation_residual(wd, ns(ta), ns(tb), S_BtoA, info3), extras=[], x_cols=[]))
    # ImageDepthFactor
    p_i, p_j = np.array([1, 2, 1.0]), np.array([4, 5, 1.0])
    S_ij, p_ij = S.so3_exp(rng.uniform(-0.3, 0.3, 3)), rng.uniform(-0.5, 0.5, 3)
    out.append(dict(name="image_depth", kind=_abi.FACTOR_IMAGE_DEPTH, d=np.concatenate([p_i, p_j, S_ij, p_ij, [300.0]]),
                    x=[0.51], fn=AD.image_depth_residual(p_i, p_j, S_ij, p_ij, 300.0), extras=[[0.51]], x_cols=[1]))
    # PreIntegrationFactor (:1192-1210 of the test list; times :926-927)
    pre = reference_preintegration()
    bias_x = np.array([0.012, 0.19, -0.008, 0.02, 0.21, -0.02, 0.011, 0.018, -0.012, 0.02, 0.03, -0.02])
    out.append(dict(name="preintegration", kind=_abi.FACTOR_PREINTEGRATION, t_a=ns(0.21), t_b=ns(0.695), d=[], x=bias_x,
                    pre=pre, fn=AD.preintegration_residual(wd, ns(0.21), ns(0.695), pre),
                    extras=[bias_x[0:3], bias_x[3:6], bias_x[6:9], bias_x[9:12]], x_cols=[3, 3, 3, 3]))
    # Image3D2DFactor
    p_G = np.array([3.1, 2.8, 1.3])
    out.append(dict(name="image_3d2d", kind=_abi.FACTOR_IMAGE_3D2D, t_a=ns(0.34), d=np.concatenate([p_j, [300.0]]),
                    x=np.concatenate([p_G, [toff]]),
                    fn=AD.image_3d2d_residual(wd, ns(0.34), p_j, w.S_CtoI, w.p_CinI, 300.0), extras=[p_G, [toff]],
                    x_cols=[3, 1]))
    # ImageFeatureOnePoseFactor
    S_IitoG, p_IiinG = rand_rot(), rng.uniform(-3, 3, 3)
    out.append(dict(name="image_one_pose", kind=_abi.FACTOR_IMAGE_ONE_POSE, t_a=ns(0.34),
                    d=np.concatenate([p_i, S_IitoG, p_IiinG, p_j, [300.0]]), x=[0.51],
                    fn=AD.image_one_pose_residual(wd, p_i, S_IitoG, p_IiinG, ns(0.34), p_j, w.S_CtoI, w.p_CinI, 300.0),
                    extras=[[0.51]], x_cols=[1]))
    # EpipolarFactor (:964-973)
    x_i, x_k, S_Ck, p_Ck = np.array([0.1, 0.2, 1]), np.array([1.3, 1.2, 1]), rand_rot(), np.array([1.3, 1.2, 1.0])
    out.append(dict(name="epipolar", kind=_abi.FACTOR_EPIPOLAR, t_a=ns(0.53),
                    d=np.concatenate([x_i, x_k, S_Ck, p_Ck, [0.11]]), x=[],
                    fn=AD.epipolar_residual(wd, ns(0.53), x_i, x_k, S_Ck, p_Ck, 0.11, w.S_CtoI, w.p_CinI), extras=[],
                    x_cols=[]))
    # the two LiDAR factors (:977-1000), plane and line
    point = np.array([1, 2, 3.0])
    plane = np.array([1, 2, 3, 4.0]) / np.linalg.norm([1, 2, 3, 4.0])
    normal = np.array([3, 2, 1.0]) / np.linalg.norm([3, 2, 1.0])
    gpoint = np.array([1.5, 1.0, 2.0])
    S_Im, p_Im = rand_rot(), rng.uniform(-3, 3, 3)
    for geo in (1, 0):
        d = np.concatenate([point, plane, normal, gpoint, [1.3], w.S_LtoI, w.p_LinI])
        out.append(dict(name=f"loam_opt_map_pose_geo{geo}", kind=_abi.FACTOR_LOAM_OPT_MAP_POSE, t_a=ns(0.55), geo=geo, d=d,
                        x=np.concatenate([S_Im, p_Im]),
                        fn=AD.loam_opt_map_pose_residual(wd, ns(0.55), point, geo, plane, normal, gpoint, 1.3, w.S_LtoI,
                                                         w.p_LinI, S_Im),
                        extras=[np.zeros(3), p_Im], x_cols=[3, 3]))
        for i, (tp, tm) in enumerate([(0.55, 0.12), (0.31, 0.36)]):  # two segments / one segment
            out.append(dict(name=f"relative_loam_geo{geo}_{i}", kind=_abi.FACTOR_RELATIVE_LOAM, t_a=ns(tp), t_b=ns(tm),
                            geo=geo, d=d, x=[],
                            fn=AD.relative_loam_residual(wd, ns(tp), ns(tm), point, geo, plane, normal, gpoint, 1.3,
                                                         w.S_LtoI, w.p_LinI), extras=[], x_cols=[]))
    for c in out:
        c.setdefault("t_a", 0)
        c.setdefault("t_b", 0)
        c.setdefault("geo", 0)
        c.setdefault("pre", None)
        c.setdefault("skip_extra", [])
    return out


def evaluate(est, c, want_jacobian=True):
    pre = preintegration_struct(c["pre"]) if c["pre"] is not None else None
    return est.evaluate_factor(c["kind"], c["t_a"], c["t_b"], c["d"], c["x"], c["geo"], pre, want_jacobian)
