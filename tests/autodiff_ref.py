"""Independent auto-diff lineage for the oracle: residual functions written from the reference's AUTO-DIFF factors
(src/estimator/factor/auto_diff/*.h + src/spline/ceres_spline_helper.h:103-218), differentiated with
torch-float64 autograd at delta = 0 of the tangent R <- R exp(dtheta), p <- p + dp.  This is the same property the
reference's only test asserts (analytic == auto-diff, src/app/test_analytic_factor.cpp:846-900), with
torch.autograd standing in for ceres::Jet.
"""
import numpy as np
import torch

torch.set_default_dtype(torch.float64)

M_CUM = torch.tensor([[6, 0, 0, 0], [5, 3, -3, 1], [1, 3, 3, -2], [0, 0, 0, 1]], dtype=torch.float64) / 6.0
M_PLAIN = torch.tensor([[1, -3, 3, -1], [4, 0, -6, 3], [1, 3, 3, -3], [0, 0, 0, 1]], dtype=torch.float64) / 6.0


def qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                        aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def qconj(a):
    return a * torch.tensor([-1.0, -1.0, -1.0, 1.0])


def qrot(q, v):
    qv = q[:3]
    uv = 2.0 * torch.linalg.cross(qv, v)
    return v + q[3] * uv + torch.linalg.cross(qv, uv)


def qexp(w):
    th2 = (w * w).sum()
    small = th2 < 1e-16
    th2s = torch.where(small, torch.ones_like(th2), th2)
    th = torch.sqrt(th2s)
    imag = torch.where(small, 0.5 - th2 / 48.0, torch.sin(0.5 * th) / th)
    real = torch.where(small, 1.0 - th2 / 8.0, torch.cos(0.5 * th))
    return torch.cat([imag * w, real.reshape(1)])


def qlog(q):
    v, w = q[:3], q[3]
    n2 = (v * v).sum()
    small = n2 < 1e-24
    n2s = torch.where(small, torch.ones_like(n2), n2)
    n = torch.sqrt(n2s)
    f = torch.where(small, 2.0 / w - 2.0 * n2 / (w * w * w), 2.0 * torch.atan(n / w) / n)
    return f * v


def pows(u, deriv):
    one, zero = torch.ones_like(u), torch.zeros_like(u)
    if deriv == 0:
        return torch.stack([one, u, u * u, u * u * u])
    if deriv == 1:
        return torch.stack([zero, one, 2 * u, 3 * u * u])
    if deriv == 2:
        return torch.stack([zero, zero, 2 * one, 6 * u])
    return torch.stack([zero, zero, zero, 6 * one])


class Spline:
    """q: (K,4) perturbed quaternions (list of tensors), p: (K,3)."""

    def __init__(self, t0_ns, dt_ns, q, p):
        self.t0_ns, self.dt_ns, self.q, self.p = t0_ns, dt_ns, q, p
        self.inv_dt = 1e9 / dt_ns

    def su(self, t_ns_int, t_frac=None):
        """t = t_ns_int (python int) + t_frac (differentiable ns)."""
        st = t_ns_int - self.t0_ns
        s = st // self.dt_ns
        u = torch.tensor(float(st % self.dt_ns) / float(self.dt_ns))
        if t_frac is not None:
            u = u + t_frac / float(self.dt_ns)
        return s, u

    def pos(self, s, u, deriv=0):
        c = (M_PLAIN @ pows(u, deriv)) * self.inv_dt ** deriv
        return sum(c[i] * self.p[s + i] for i in range(4))

    def rot_vel(self, s, u):
        """R(t) quaternion and body angular velocity (ceres_spline_helper.h evaluate_lie)."""
        c = M_CUM @ pows(u, 0)
        dc = (M_CUM @ pows(u, 1)) * self.inv_dt
        r = self.q[s]
        w = torch.zeros(3)
        for i in range(3):
            d = qlog(qmul(qconj(self.q[s + i]), self.q[s + i + 1]))
            e = qexp(c[i + 1] * d)
            r = qmul(r, e)
            w = qrot(qconj(e), w) + dc[i + 1] * d
        return r, w


def perturbed_spline(t0_ns, dt_ns, knot_q, knot_p, dth, dp):
    K = knot_q.shape[0]
    q = [qmul(torch.tensor(knot_q[k]), qexp(dth[k])) for k in range(K)]
    p = [torch.tensor(knot_p[k]) + dp[k] for k in range(K)]
    return Spline(t0_ns, dt_ns, q, p)


def jac(fn, K, extras):
    """Jacobian of fn(dth (K,3), dp (K,3), *extras) at dth = dp = 0 and the given extra values.
    Returns residual, J_knots (m, K, 6) [dtheta, dp per knot], [J_extra...]."""
    dth0, dp0 = torch.zeros(K, 3), torch.zeros(K, 3)
    ex = [torch.tensor(np.atleast_1d(np.asarray(e, np.float64))) for e in extras]
    r = fn(dth0, dp0, *ex)
    J = torch.autograd.functional.jacobian(fn, (dth0, dp0, *ex))
    Jk = torch.cat([J[0], J[1]], dim=-1)  # (m, K, 6)
    return r.detach().numpy(), Jk.detach().numpy(), [j.detach().numpy() for j in J[2:]]


# ---- residual definitions ----
def lidar_residual(win, t_ns, point, geo_type, geo_plane, geo_normal, geo_point, S_GtoM, p_GinM, S_LtoI, p_LinI, weight):
    """Point-to-plane / point-to-line distance (same geometry as auto_diff/lidar_feature_factor.h:84-104 with the map
    frame given explicitly as in analytic LoamFeatureFactor)."""
    def fn(dth, dp):
        sp = perturbed_spline(win["t0_ns"], win["dt_ns"], win["knot_q"], win["knot_p"], dth, dp)
        s, u = sp.su(t_ns)
        R, _ = sp.rot_vel(s, u)
        p_I = qrot(torch.tensor(S_LtoI), torch.tensor(point)) + torch.tensor(p_LinI)
        p_M = qrot(torch.tensor(S_GtoM), qrot(R, p_I) + sp.pos(s, u)) + torch.tensor(p_GinM)
        if geo_type == 1:
            dist = (p_M * torch.tensor(geo_plane[:3])).sum() + geo_plane[3]
        else:
            dist = torch.linalg.norm(torch.linalg.cross(p_M - torch.tensor(geo_point), torch.tensor(geo_normal)))
        return (weight * dist).reshape(1)
    return fn


def imu_residual(win, t_ns, gyro, accel, info_vec):
    """auto_diff/trajectory_value_factor.h:237-272; extras = (gyro_bias, accel_bias, gravity, t_offset_ns).
    The offset is split like the analytic factor: integer part moves the index, the autograd variable is the
    perturbation around it."""
    def fn(dth, dp, bg, ba, g, toff):
        sp = perturbed_spline(win["t0_ns"], win["dt_ns"], win["knot_q"], win["knot_p"], dth, dp)
        toff_int = int(toff.detach()[0])  # (int64) truncation, trajectory_value_factor.h:175
        s, u = sp.su(t_ns + toff_int, toff[0] - toff.detach()[0])
        R, w = sp.rot_vel(s, u)
        a_w = sp.pos(s, u, 2)
        r_g = w - torch.tensor(gyro) + bg
        r_a = qrot(qconj(R), a_w + g) - torch.tensor(accel) + ba
        return torch.cat([r_g, r_a]) * torch.tensor(info_vec)
    return fn


def image_residual(win, ti_ns, p_i, tj_ns, p_j, S_CtoI, p_CinI, sqrt_info=450.0 / 1.5):
    """auto_diff/image_feature_factor.h:48-101; extras = (inv_depth, t_offset_ns)."""
    def fn(dth, dp, rho, toff):
        sp = perturbed_spline(win["t0_ns"], win["dt_ns"], win["knot_q"], win["knot_p"], dth, dp)
        toff_int = int(toff.detach()[0])
        frac = toff[0] - toff.detach()[0]
        si, ui = sp.su(ti_ns + toff_int, frac)
        sj, uj = sp.su(tj_ns + toff_int, frac)
        Ri, _ = sp.rot_vel(si, ui)
        Rj, _ = sp.rot_vel(sj, uj)
        qc = torch.tensor(S_CtoI)
        x_ci = torch.tensor(p_i) / rho[0]
        p_Ii = qrot(qc, x_ci) + torch.tensor(p_CinI)
        p_G = qrot(Ri, p_Ii) + sp.pos(si, ui)
        p_Ij = qrot(qconj(Rj), p_G - sp.pos(sj, uj))
        x_j = qrot(qconj(qc), p_Ij - torch.tensor(p_CinI))
        return sqrt_info * (x_j[:2] / x_j[2] - torch.tensor(p_j[:2]))
    return fn


def pose_residual(win, t_ns, position, orientation, info_vec):
    """auto_diff/trajectory_value_factor.h:377-434 (IMUPoseFactor); extras = (t_offset_ns,)."""
    def fn(dth, dp, toff):
        sp = perturbed_spline(win["t0_ns"], win["dt_ns"], win["knot_q"], win["knot_p"], dth, dp)
        toff_int = int(toff.detach()[0])
        s, u = sp.su(t_ns + toff_int, toff[0] - toff.detach()[0])
        R, _ = sp.rot_vel(s, u)
        r_R = qlog(qmul(R, qconj(torch.tensor(orientation))))
        r_p = sp.pos(s, u) - torch.tensor(position)
        return torch.cat([r_R, r_p]) * torch.tensor(info_vec)
    return fn


def local_velocity_residual(win, t_ns, local_v, weight):
    """auto_diff/trajectory_value_factor.h:490-547 (IMULocalVelocityFactor); extras = (t_offset_ns,)."""
    def fn(dth, dp, toff):
        sp = perturbed_spline(win["t0_ns"], win["dt_ns"], win["knot_q"], win["knot_p"], dth, dp)
        toff_int = int(toff.detach()[0])
        s, u = sp.su(t_ns + toff_int, toff[0] - toff.detach()[0])
        R, _ = sp.rot_vel(s, u)
        return weight * (qrot(R, torch.tensor(local_v)) - sp.pos(s, u, 1))
    return fn


def relative_rotation_residual(win, ta_ns, tb_ns, S_BtoA, info_vec):
    """auto_diff/trajectory_value_factor.h:770-840 (RelativeOrientationFactor)."""
    def fn(dth, dp):
        sp = perturbed_spline(win["t0_ns"], win["dt_ns"], win["knot_q"], win["knot_p"], dth, dp)
        sa, ua = sp.su(ta_ns)
        sb, ub = sp.su(tb_ns)
        Ra, _ = sp.rot_vel(sa, ua)
        Rb, _ = sp.rot_vel(sb, ub)
        return torch.tensor(info_vec) * qlog(qmul(qmul(qconj(Rb), Ra), torch.tensor(S_BtoA)))
    return fn


def image_depth_residual(p_i, p_j, S_CitoCj, p_CiinCj, sqrt_info):
    """auto_diff/image_feature_factor.h:310-357 (ImageDepthFactor); extras = (inv_depth,).  No spline involved: the
    knot arguments are ignored."""
    def fn(dth, dp, rho):
        x_ci = torch.tensor(p_i) / rho[0]
        x_j = qrot(torch.tensor(S_CitoCj), x_ci) + torch.tensor(p_CiinCj)
        return sqrt_info * (x_j[:2] / x_j[2] - torch.tensor(p_j[:2])) + 0.0 * dth.sum() + 0.0 * dp.sum()
    return fn


# ---- the analytic factors the reference only exercises from its unit test ----
def _loam_distance(p, geo_type, geo_plane, geo_normal, geo_point):
    if geo_type == 1:
        return (p * torch.tensor(geo_plane[:3])).sum() + geo_plane[3]
    return torch.linalg.norm(torch.linalg.cross(p - torch.tensor(geo_point), torch.tensor(geo_normal)))


def preintegration_residual(win, ta_ns, tb_ns, pre):
    """auto_diff/trajectory_value_factor.h:703-756 + IntegrationBase::evaluate333 (integration_base.h:207-252);
    pre: dict(sum_dt, delta_p, delta_q, delta_v, linearized_ba, linearized_bg, jacobian (15,15), sqrt_info (15,15),
    gravity_norm); extras = (bg_a, bg_b, ba_a, ba_b)."""
    Jm = torch.tensor(pre["jacobian"])
    O_P, O_R, O_V, O_BA, O_BG = 0, 3, 6, 9, 12
    dt = float(pre["sum_dt"])
    g = torch.tensor([0.0, 0.0, -float(pre["gravity_norm"])])

    def fn(dth, dp, bg_a, bg_b, ba_a, ba_b):
        sp = perturbed_spline(win["t0_ns"], win["dt_ns"], win["knot_q"], win["knot_p"], dth, dp)
        sa, ua = sp.su(ta_ns)
        sb, ub = sp.su(tb_ns)
        Ra, _ = sp.rot_vel(sa, ua)
        Rb, _ = sp.rot_vel(sb, ub)
        Pa, Pb = sp.pos(sa, ua), sp.pos(sb, ub)
        Va, Vb = sp.pos(sa, ua, 1), sp.pos(sb, ub, 1)
        dba = ba_a - torch.tensor(pre["linearized_ba"])
        dbg = bg_a - torch.tensor(pre["linearized_bg"])
        cdp = torch.tensor(pre["delta_p"]) + Jm[O_P:O_P + 3, O_BA:O_BA + 3] @ dba + Jm[O_P:O_P + 3, O_BG:O_BG + 3] @ dbg
        cdq = qmul(torch.tensor(pre["delta_q"]), qexp(Jm[O_R:O_R + 3, O_BG:O_BG + 3] @ dbg))
        cdq = cdq / torch.linalg.norm(cdq)
        cdv = torch.tensor(pre["delta_v"]) + Jm[O_V:O_V + 3, O_BA:O_BA + 3] @ dba + Jm[O_V:O_V + 3, O_BG:O_BG + 3] @ dbg
        r_p = qrot(qconj(Ra), 0.5 * dt * dt * g + Pb - Pa - Va * dt) - cdp
        r_R = qlog(qmul(qmul(qconj(cdq), qconj(Ra)), Rb))
        r_v = qrot(qconj(Ra), g * dt + Vb - Va) - cdv
        r = torch.cat([r_p, r_R, r_v, ba_b - ba_a, bg_b - bg_a])
        return torch.tensor(pre["sqrt_info"]) @ r
    return fn


def image_3d2d_residual(win, tj_ns, p_j, S_CtoI, p_CinI, sqrt_info):
    """auto_diff/image_feature_factor.h (Image3D2DFactor); extras = (p_G, t_offset_ns)."""
    def fn(dth, dp, p_G, toff):
        sp = perturbed_spline(win["t0_ns"], win["dt_ns"], win["knot_q"], win["knot_p"], dth, dp)
        toff_int = int(toff.detach()[0])
        sj, uj = sp.su(tj_ns + toff_int, toff[0] - toff.detach()[0])
        Rj, _ = sp.rot_vel(sj, uj)
        qc = torch.tensor(S_CtoI)
        p_Ij = qrot(qconj(Rj), p_G - sp.pos(sj, uj))
        x_j = qrot(qconj(qc), p_Ij - torch.tensor(p_CinI))
        return sqrt_info * (x_j[:2] / x_j[2] - torch.tensor(p_j[:2]))
    return fn


def image_one_pose_residual(win, p_i, S_IitoG, p_IiinG, tj_ns, p_j, S_CtoI, p_CinI, sqrt_info):
    """auto_diff/image_feature_factor.h (ImageFeatureOnePoseFactor); extras = (inv_depth,)."""
    def fn(dth, dp, rho):
        sp = perturbed_spline(win["t0_ns"], win["dt_ns"], win["knot_q"], win["knot_p"], dth, dp)
        sj, uj = sp.su(tj_ns)
        Rj, _ = sp.rot_vel(sj, uj)
        qc = torch.tensor(S_CtoI)
        p_Ii = qrot(qc, torch.tensor(p_i) / rho[0]) + torch.tensor(p_CinI)
        p_G = qrot(torch.tensor(S_IitoG), p_Ii) + torch.tensor(p_IiinG)
        p_Ij = qrot(qconj(Rj), p_G - sp.pos(sj, uj))
        x_j = qrot(qconj(qc), p_Ij - torch.tensor(p_CinI))
        return sqrt_info * (x_j[:2] / x_j[2] - torch.tensor(p_j[:2]))
    return fn


def epipolar_residual(win, ti_ns, x_i, x_k, S_GtoCk, p_CkinG, weight, S_CtoI, p_CinI):
    """auto_diff/image_feature_factor.h (EpipolarFactor): x_k . ([t]x R x_i)."""
    def fn(dth, dp):
        sp = perturbed_spline(win["t0_ns"], win["dt_ns"], win["knot_q"], win["knot_p"], dth, dp)
        si, ui = sp.su(ti_ns)
        Ri, _ = sp.rot_vel(si, ui)
        qk = torch.tensor(S_GtoCk)
        Rxi = qrot(qk, qrot(Ri, qrot(torch.tensor(S_CtoI), torch.tensor(x_i))))
        t = qrot(qk, qrot(Ri, torch.tensor(p_CinI)) + sp.pos(si, ui) - torch.tensor(p_CkinG))
        return (weight * (torch.tensor(x_k) * torch.linalg.cross(t, Rxi)).sum()).reshape(1)
    return fn


def loam_opt_map_pose_residual(win, t_ns, point, geo_type, geo_plane, geo_normal, geo_point, weight, S_LtoI, p_LinI,
                               S_ImtoG):
    """auto_diff/lidar_feature_factor.h (LoamFeatureOptMapPoseFactor); extras = (dtheta_map (right perturbation of
    S_ImtoG), p_IminG)."""
    def fn(dth, dp, dth_map, p_map):
        sp = perturbed_spline(win["t0_ns"], win["dt_ns"], win["knot_q"], win["knot_p"], dth, dp)
        s, u = sp.su(t_ns)
        R, _ = sp.rot_vel(s, u)
        ql = torch.tensor(S_LtoI)
        qm = qmul(torch.tensor(S_ImtoG), qexp(dth_map))
        p_IK = qrot(ql, torch.tensor(point)) + torch.tensor(p_LinI)
        p_G = qrot(R, p_IK) + sp.pos(s, u)
        p_Im = qrot(qconj(qm), p_G - p_map)          # the map IMU frame
        p_M = qrot(qconj(ql), p_Im - torch.tensor(p_LinI))  # ... and its LiDAR frame
        return (weight * _loam_distance(p_M, geo_type, geo_plane, geo_normal, geo_point)).reshape(1)
    return fn


def relative_loam_residual(win, t_point_ns, t_map_ns, point, geo_type, geo_plane, geo_normal, geo_point, weight, S_LtoI,
                           p_LinI):
    """RalativeLoamFeatureFactor (analytic only in the reference; residual from lidar_feature_factor.h:398-455): the
    point of scan time t_point expressed in the LiDAR frame of time t_map."""
    def fn(dth, dp):
        sp = perturbed_spline(win["t0_ns"], win["dt_ns"], win["knot_q"], win["knot_p"], dth, dp)
        si, ui = sp.su(t_point_ns)
        sj, uj = sp.su(t_map_ns)
        Ri, _ = sp.rot_vel(si, ui)
        Rj, _ = sp.rot_vel(sj, uj)
        ql = torch.tensor(S_LtoI)
        p_Ii = qrot(ql, torch.tensor(point)) + torch.tensor(p_LinI)
        p_G = qrot(Ri, p_Ii) + sp.pos(si, ui)
        p_Ij = qrot(qconj(Rj), p_G - sp.pos(sj, uj))
        x_j = qrot(qconj(ql), p_Ij - torch.tensor(p_LinI))
        return (weight * _loam_distance(x_j, geo_type, geo_plane, geo_normal, geo_point)).reshape(1)
    return fn
