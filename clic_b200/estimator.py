"""Python host-side mirror of clic::TrajectoryEstimator (src/estimator/trajectory_estimator.h:100-278) on top of
the C ABI (include/clic_cuda.h).  Method names and argument meaning follow the reference; every Add* call is
batched over numpy arrays instead of one call per measurement.  There is no compute in this file: everything
runs inside libclic_cuda.so on the GPU, and construction fails loudly if that library or a GPU is missing.
"""
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _abi
from ._abi import ClicError, FactorDesc, Layout, Options, PreIntegration, Summary, WindowDesc, dptr, f64, i32, iptr

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libclic_cuda.so")
_lib = None


def cuda_library():
    """The product library.  No fallback: a missing .so is an error."""
    global _lib
    if _lib is None:
        path = os.environ.get("CLIC_CUDA_LIB", _LIB_PATH)  # developer override: kernel-tuning variants of the same library
        if path != _LIB_PATH:
            _lib = _abi.bind(path)
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise ImportError(
                f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        _lib = _abi.bind(_LIB_PATH)
    return _lib


@dataclass
class Window:
    """Trajectory window + non-spline states (clic_window_desc)."""
    t0_ns: int
    dt_ns: int
    knot_q: np.ndarray  # (K,4) xyzw
    knot_p: np.ndarray  # (K,3)
    knot_id0: int = 0
    S_LtoI: np.ndarray = field(default_factory=lambda: np.array([0, 0, 0, 1.0]))
    p_LinI: np.ndarray = field(default_factory=lambda: np.zeros(3))
    S_CtoI: np.ndarray = field(default_factory=lambda: np.array([0, 0, 0, 1.0]))
    p_CinI: np.ndarray = field(default_factory=lambda: np.zeros(3))
    t_offset_ns: np.ndarray = field(default_factory=lambda: np.zeros(3))
    gravity: np.ndarray = field(default_factory=lambda: np.array([0, 0, 9.8]))
    bias: np.ndarray = field(default_factory=lambda: np.zeros((0, 6)))
    bias_id0: int = 0
    inv_depth: np.ndarray = field(default_factory=lambda: np.zeros(0))

    @property
    def n_knots(self):
        return int(self.knot_q.shape[0])


def default_options(lib=None):
    lib = lib or cuda_library()
    o = Options()
    lib.clic_default_options(C.byref(o))
    return o


class Prior:
    """MarginalizationInfo after marginalize() (marginalization_factor.h:100-142)."""

    def __init__(self, lib, handle):
        self.lib, self.h = lib, handle

    @classmethod
    def create(cls, lib, J0, r0, kinds, ids, x0):
        J0, r0, x0 = f64(J0), f64(r0), f64(x0)
        kinds, ids = i32(kinds), i32(ids)
        h = C.c_void_p()
        _check(lib, lib.clic_prior_create(len(r0), dptr(J0), dptr(r0), len(kinds), iptr(kinds), iptr(ids), dptr(x0),
                                          C.byref(h)))
        return cls(lib, h)

    def dims(self):
        n, nb, m = C.c_int32(), C.c_int32(), C.c_int32()
        _check(self.lib, self.lib.clic_prior_dims(self.h, C.byref(n), C.byref(nb), C.byref(m)))
        return n.value, nb.value, m.value

    def read(self):
        n, nb, _ = self.dims()
        J0, r0 = np.zeros((n, n)), np.zeros(n)
        kind, idv, off = np.zeros(nb, np.int32), np.zeros(nb, np.int32), np.zeros(nb, np.int32)
        x0 = np.zeros((nb, 4))
        _check(self.lib, self.lib.clic_prior_read(self.h, dptr(J0), dptr(r0), iptr(kind), iptr(idv), iptr(off), dptr(x0)))
        return dict(J0=J0, r0=r0, kind=kind, id=idv, offset=off, x0=x0)

    def normal(self):
        n, _, _ = self.dims()
        A, g = np.zeros((n, n)), np.zeros(n)
        _check(self.lib, self.lib.clic_prior_normal(self.h, dptr(A), dptr(g)))
        return A, g

    def release(self):
        if self.h:
            self.lib.clic_prior_release(self.h)
            self.h = None


def _check(lib, status):
    if status != 0:
        raise ClicError(status, (lib.clic_last_error() or b"").decode())


class TrajectoryEstimator:
    """One estimator per solve, like the reference (trajectory_manager.cpp:637)."""

    def __init__(self, window: Window, options: Options = None, lib=None):
        self.lib = lib or cuda_library()
        self.options = options or default_options(self.lib)
        self.window = window
        self._keep = []  # arrays borrowed by clic_create
        d = WindowDesc()
        d.t0_ns, d.dt_ns = int(window.t0_ns), int(window.dt_ns)
        d.n_knots, d.knot_id0 = window.n_knots, int(window.knot_id0)
        kq, kp = f64(window.knot_q), f64(window.knot_p)
        bias, invd = f64(window.bias).reshape(-1, 6), f64(window.inv_depth).reshape(-1)
        self._keep += [kq, kp, bias, invd]
        d.knot_q, d.knot_p = dptr(kq), dptr(kp)
        for name in ("S_LtoI", "p_LinI", "S_CtoI", "p_CinI", "t_offset_ns", "gravity"):
            v = f64(getattr(window, name)).reshape(-1)
            getattr(d, name)[:] = list(v)
        d.n_bias, d.bias_id0 = bias.shape[0], int(window.bias_id0)
        d.bias = dptr(bias) if bias.size else None
        d.n_landmarks = invd.shape[0]
        d.inv_depth = dptr(invd) if invd.size else None
        self.h = C.c_void_p()
        _check(self.lib, self.lib.clic_create(C.byref(d), C.byref(self.options), C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.clic_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- trajectory_estimator.h:140 ----
    def SetFixedIndex(self, idx):
        _check(self.lib, self.lib.clic_set_fixed_index(self.h, int(idx)))

    # ---- trajectory_estimator.h:188-193 (batched) ----
    def AddLoamMeasurementAnalytic(self, t_point, point, geo_type, geo_plane, geo_normal, geo_point, S_GtoM, p_GinM,
                                   S_LtoI, p_LinI, weight, marg_this_factor=False, count_dropped=True):
        """count_dropped=False keeps the call asynchronous (no read-back of the number of dropped measurements)."""
        t_point, point = f64(t_point), f64(point)
        geo_type = i32(geo_type)
        geo_plane, geo_normal, geo_point = f64(geo_plane), f64(geo_normal), f64(geo_point)
        S_GtoM, p_GinM, S_LtoI, p_LinI = f64(S_GtoM), f64(p_GinM), f64(S_LtoI), f64(p_LinI)
        dropped = C.c_int64(0)
        _check(self.lib, self.lib.clic_add_loam(self.h, len(t_point), dptr(t_point), dptr(point), iptr(geo_type),
                                                dptr(geo_plane), dptr(geo_normal), dptr(geo_point), dptr(S_GtoM),
                                                dptr(p_GinM), dptr(S_LtoI), dptr(p_LinI), float(weight),
                                                int(bool(marg_this_factor)),
                                                C.byref(dropped) if count_dropped else None))
        return dropped.value

    def AddLoamMeasurementPacked(self, t_point, point, geo_type_u8, geo6, S_GtoM, p_GinM, S_LtoI, p_LinI, weight,
                                 marg_this_factor=False, count_dropped=True):
        """Packed input of the same factors: geo_type as uint8, geo6 = plane (n, d, 0, 0) | line (point, direction)."""
        t_point, point, geo6 = f64(t_point), f64(point), f64(geo6)
        geo_type_u8 = np.ascontiguousarray(geo_type_u8, dtype=np.uint8)
        S_GtoM, p_GinM, S_LtoI, p_LinI = f64(S_GtoM), f64(p_GinM), f64(S_LtoI), f64(p_LinI)
        dropped = C.c_int64(0)
        _check(self.lib, self.lib.clic_add_loam_packed(
            self.h, len(t_point), dptr(t_point), dptr(point), geo_type_u8.ctypes.data_as(C.POINTER(C.c_uint8)),
            dptr(geo6), dptr(S_GtoM), dptr(p_GinM), dptr(S_LtoI), dptr(p_LinI), float(weight),
            int(bool(marg_this_factor)), C.byref(dropped) if count_dropped else None))
        return dropped.value

    def _add_loam_typed(self, fn, t_point, point, geo, S_GtoM, p_GinM, S_LtoI, p_LinI, weight, marg, count_dropped):
        t_point, point, geo = f64(t_point), f64(point), f64(geo)
        S_GtoM, p_GinM, S_LtoI, p_LinI = f64(S_GtoM), f64(p_GinM), f64(S_LtoI), f64(p_LinI)
        dropped = C.c_int64(0)
        _check(self.lib, fn(self.h, len(t_point), dptr(t_point), dptr(point), dptr(geo), dptr(S_GtoM), dptr(p_GinM),
                            dptr(S_LtoI), dptr(p_LinI), float(weight), int(bool(marg)),
                            C.byref(dropped) if count_dropped else None))
        return dropped.value

    def AddLoamPlanes(self, t_point, point, plane4, S_GtoM, p_GinM, S_LtoI, p_LinI, weight, marg_this_factor=False,
                      count_dropped=True):
        """A time-sorted batch of plane features: plane4 = (n.x, n.y, n.z, d) per feature."""
        return self._add_loam_typed(self.lib.clic_add_loam_planes, t_point, point, plane4, S_GtoM, p_GinM, S_LtoI, p_LinI,
                                    weight, marg_this_factor, count_dropped)

    def AddLoamLines(self, t_point, point, line6, S_GtoM, p_GinM, S_LtoI, p_LinI, weight, marg_this_factor=False,
                     count_dropped=True):
        """A time-sorted batch of line features: line6 = (geo_point xyz, geo_normal xyz) per feature."""
        return self._add_loam_typed(self.lib.clic_add_loam_lines, t_point, point, line6, S_GtoM, p_GinM, S_LtoI, p_LinI,
                                    weight, marg_this_factor, count_dropped)

    # ---- trajectory_estimator.h:153-156 (batched; gravity is the window's) ----
    def AddIMUMeasurementAnalytic(self, timestamp, gyro, accel, bias_slot, info_vec, marg_this_factor=False,
                                  count_dropped=True):
        timestamp, gyro, accel, info_vec = f64(timestamp), f64(gyro), f64(accel), f64(info_vec)
        dropped = C.c_int64(0)
        _check(self.lib, self.lib.clic_add_imu(self.h, len(timestamp), dptr(timestamp), dptr(gyro), dptr(accel),
                                               int(bias_slot), dptr(info_vec), int(bool(marg_this_factor)),
                                               C.byref(dropped) if count_dropped else None))
        return dropped.value

    # ---- trajectory_estimator.h:159-162 ----
    def AddBiasFactor(self, slot_i, slot_j, dt, info_vec, marg_this_factor=False, marg_all_bias=False):
        info_vec = f64(info_vec)
        _check(self.lib, self.lib.clic_add_bias(self.h, int(slot_i), int(slot_j), float(dt), dptr(info_vec),
                                                int(bool(marg_this_factor)), int(bool(marg_all_bias))))

    # ---- trajectory_estimator.h:164-165 ----
    def AddGravityFactor(self, gravity, info_vec, marg_this_factor=False):
        g, info_vec = f64(gravity), f64(info_vec)
        _check(self.lib, self.lib.clic_add_gravity(self.h, dptr(g), dptr(info_vec), int(bool(marg_this_factor))))

    # ---- trajectory_estimator.h:150-151 (IMUPoseFactor) ----
    def AddPoseMeasurementAnalytic(self, timestamp, position, orientation, info_vec):
        """Batched over pose measurements (orientation (n,4) xyzw); returns the number dropped (outside the window)."""
        t, p, q, info_vec = f64(timestamp), f64(position), f64(orientation), f64(info_vec)
        nd = C.c_int64(0)
        _check(self.lib, self.lib.clic_add_pose(self.h, len(t), dptr(t), dptr(p), dptr(q), dptr(info_vec), C.byref(nd)))
        return nd.value

    # ---- trajectory_estimator.h:183-185 (LocalVelocityFactor) ----
    def AddLocalVelocityMeasurementAnalytic(self, timestamp, local_v, weight):
        t, v = f64(timestamp), f64(local_v)
        nd = C.c_int64(0)
        _check(self.lib, self.lib.clic_add_local_velocity(self.h, len(t), dptr(t), dptr(v), float(weight), C.byref(nd)))
        return nd.value

    # ---- trajectory_estimator.h:174-175 (RelativeOrientationFactor) ----
    def AddRelativeRotationAnalytic(self, ta, tb, S_BtoA, info_vec):
        S_BtoA, info_vec = f64(S_BtoA), f64(info_vec)
        added = C.c_int32(0)
        _check(self.lib, self.lib.clic_add_relative_rotation(self.h, float(ta), float(tb), dptr(S_BtoA), dptr(info_vec),
                                                             C.byref(added)))
        return bool(added.value)

    # ---- trajectory_estimator.h:201-204 (ImageDepthFactor) ----
    def AddImageDepthAnalytic(self, p_i, p_j, S_CitoCj, p_CiinCj, landmark):
        a = [f64(p_i), f64(p_j), f64(S_CitoCj), f64(p_CiinCj)]
        _check(self.lib, self.lib.clic_add_image_depth(self.h, *[dptr(x) for x in a], int(landmark)))

    # ---- trajectory_estimator.h:168-172 (PreIntegrationFactor) ----
    def AddPreIntegrationAnalytic(self, ta_ns, tb_ns, pre, slot_i, slot_j):
        """pre: a PreIntegration struct (make_preintegration).  Times are int64 ns on the trajectory clock."""
        added = C.c_int32(0)
        _check(self.lib, self.lib.clic_add_preintegration(self.h, int(ta_ns), int(tb_ns), C.byref(pre), int(slot_i),
                                                          int(slot_j), C.byref(added)))
        return bool(added.value)

    # ---- src/app/test_analytic_factor.cpp:846-900: one cost function, residual and Jacobian ----
    def evaluate_factor(self, kind, t_a_ns=0, t_b_ns=0, d=(), x=(), geo_type=0, pre=None, want_jacobian=True):
        """Returns (residual (m,), J (m, 6 K + extras)) of one analytic factor at the current window; the layout of d, x
        and of the Jacobian columns is documented at clic_evaluate_factor in include/clic_cuda.h."""
        f = FactorDesc()
        f.kind, f.geo_type, f.t_a_ns, f.t_b_ns = int(kind), int(geo_type), int(t_a_ns), int(t_b_ns)
        for i, v in enumerate(np.asarray(d, np.float64).ravel()):
            f.d[i] = v
        for i, v in enumerate(np.asarray(x, np.float64).ravel()):
            f.x[i] = v
        if pre is not None:
            f.pre = C.pointer(pre)
        n_res, n_cols = C.c_int32(0), C.c_int32(0)
        res = np.zeros(15)
        cap = 15 * (6 * self.window.n_knots + 16)
        jac = np.zeros(cap)
        _check(self.lib, self.lib.clic_evaluate_factor(self.h, C.byref(f), C.byref(n_res), dptr(res), C.byref(n_cols),
                                                       dptr(jac) if want_jacobian else None, cap))
        m, n = n_res.value, n_cols.value
        return res[:m].copy(), (jac[:m * n].reshape(m, n).copy() if want_jacobian else None)

    # ---- trajectory_estimator.cpp:609-664 ----
    def AddGlobalVelocityMeasurement(self, timestamp, global_v, vel_weight):
        """IMUGlobalVelocityFactor at one timestamp; returns False when it falls outside the window."""
        v = f64(global_v)
        added = C.c_int32(0)
        _check(self.lib, self.lib.clic_add_global_velocity(self.h, float(timestamp), dptr(v), float(vel_weight),
                                                           C.byref(added)))
        return bool(added.value)

    # ---- trajectory_estimator.h:196-199 (batched) ----
    def AddImageFeatureAnalytic(self, ti, pi, tj, pj, landmark, fixed_depth=True, marg_this_feature=False,
                                count_dropped=True):
        ti, pi, tj, pj = f64(ti), f64(pi), f64(tj), f64(pj)
        landmark = i32(landmark)
        dropped = C.c_int64(0)
        _check(self.lib, self.lib.clic_add_image(self.h, len(ti), dptr(ti), dptr(pi), dptr(tj), dptr(pj),
                                                 iptr(landmark), int(bool(fixed_depth)), int(bool(marg_this_feature)),
                                                 C.byref(dropped) if count_dropped else None))
        return dropped.value

    # ---- trajectory_estimator.h:207-209 ----
    def AddMarginalizationFactor(self, prior: Prior, drop_blocks=None):
        drop = i32(drop_blocks if drop_blocks is not None else [])
        _check(self.lib, self.lib.clic_add_prior(self.h, prior.h, len(drop), iptr(drop) if len(drop) else None))

    def layout(self):
        L = Layout()
        _check(self.lib, self.lib.clic_get_layout(self.h, C.byref(L)))
        return L

    def Evaluate(self, want_H=True):
        """One fused residual+Jacobian+normal-equation pass -> (H, b, cost, fixed_cost)."""
        D = self.layout().D
        H = np.zeros((D, D)) if want_H else None
        b = np.zeros(D) if want_H else None
        cost, fcost = C.c_double(0), C.c_double(0)
        _check(self.lib, self.lib.clic_evaluate(self.h, dptr(H), dptr(b), C.byref(cost), C.byref(fcost)))
        return H, b, cost.value, fcost.value

    def indices(self, stream):
        n = C.c_int64(0)
        s0, u0 = np.zeros(1, np.int32), np.zeros(1)
        _check(self.lib, self.lib.clic_get_indices(self.h, stream, 0, iptr(s0), dptr(u0), C.byref(n)))
        s, u = np.zeros(max(n.value, 1), np.int32), np.zeros(max(n.value, 1))
        _check(self.lib, self.lib.clic_get_indices(self.h, stream, n.value, iptr(s), dptr(u), C.byref(n)))
        return s[:n.value], u[:n.value]

    # ---- trajectory_estimator.h:225-226 ----
    def Solve(self, max_iterations=50):
        s = Summary()
        _check(self.lib, self.lib.clic_solve(self.h, int(max_iterations), C.byref(s)))
        return s.as_dict()

    def read_state(self):
        K = self.window.n_knots
        nb, nl = self._keep[2].shape[0], self._keep[3].shape[0]
        q, p = np.zeros((K, 4)), np.zeros((K, 3))
        bias, toff, invd = np.zeros((max(nb, 1), 6)), np.zeros(3), np.zeros(max(nl, 1))
        _check(self.lib, self.lib.clic_read_state(self.h, dptr(q), dptr(p), dptr(bias), dptr(toff), dptr(invd)))
        return dict(knot_q=q, knot_p=p, bias=bias[:nb], t_offset_ns=toff, inv_depth=invd[:nl])

    def write_state(self, knot_q=None, knot_p=None, bias=None, t_offset_ns=None, inv_depth=None):
        a = [f64(knot_q), f64(knot_p), f64(bias), f64(t_offset_ns), f64(inv_depth)]
        _check(self.lib, self.lib.clic_write_state(self.h, *[dptr(x) for x in a]))

    # ---- trajectory_estimator.h:234-235 ----
    def SaveMarginalizationInfo(self):
        h = C.c_void_p()
        st = self.lib.clic_marginalize(self.h, C.byref(h))
        if st == 8:  # CLIC_ERR_NOTHING_TO_KEEP: the reference yields nullptr (trajectory_estimator.cpp:241-248)
            return None
        _check(self.lib, st)
        return Prior(self.lib, h)

    # ---- multi-GPU: this rank's estimator holds a shard of the factors ----
    def comm_init_torch(self, dist, rank, world):
        """Create / attach the process NCCL communicator; the unique id travels over torch.distributed."""
        import torch
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            _check(self.lib, self.lib.clic_comm_unique_id(ident))
        t = torch.tensor(list(ident), dtype=torch.uint8, device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.broadcast(t, src=0)
        ident = (C.c_uint8 * 128)(*t.cpu().tolist())
        _check(self.lib, self.lib.clic_comm_init(self.h, int(world), int(rank), ident))
        # one-shot all-reduce over NVLink peer memory (CUDA IPC between the ranks of one node); NCCL stays the fallback
        import os
        if (world > 1 and dist.get_backend() == "nccl" and not os.environ.get("CLIC_NO_P2P")
                and not self.lib.clic_comm_p2p_ready()):
            mine = (C.c_uint8 * 64)()
            _check(self.lib, self.lib.clic_comm_p2p_export(self.h, mine))
            local = torch.tensor(list(mine), dtype=torch.uint8, device="cuda")
            gathered = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(gathered, local)
            flat = (C.c_uint8 * (64 * world))(*torch.cat(gathered).cpu().tolist())
            rc = self.lib.clic_comm_p2p_import(self.h, int(world), int(rank), flat)
            ok = torch.tensor([1 if rc == 0 else 0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)        # every rank must take the same path
            if int(ok.item()) == 0:  # e.g. no CUDA IPC between the processes: keep ncclAllReduce on every rank
                self.lib.clic_comm_p2p_disable()

    # ---- measurement hooks ----
    def linearize_async(self, reps=1):
        _check(self.lib, self.lib.clic_linearize_async(self.h, int(reps)))

    def synchronize(self):
        _check(self.lib, self.lib.clic_synchronize(self.h))

    def profile_enable(self, on=True):
        _check(self.lib, self.lib.clic_profile_enable(self.h, int(on)))

    # ---- Trajectory::GetSensorPose / UndistortScan(InG) batched (src/spline/trajectory.cpp:35-115) ----
    def GetSensorPoses(self, t_s, S_StoI, p_SinI, t_offset_ns=0.0):
        """pose_S_to_G at every timestamp (seconds): returns q (n,4 xyzw), p (n,3), number of out-of-range times."""
        t_s, S_StoI, p_SinI = f64(t_s), f64(S_StoI), f64(p_SinI)
        n = len(t_s)
        q, p, bad = np.zeros((n, 4)), np.zeros((n, 3)), C.c_int64(0)
        _check(self.lib, self.lib.clic_query_poses(self.h, n, dptr(t_s), dptr(S_StoI), dptr(p_SinI), float(t_offset_ns),
                                                   dptr(q), dptr(p), C.byref(bad)))
        return q, p, bad.value

    def UndistortScan(self, t_s, xyz, S_LtoI, p_LinI, t_offset_ns=0.0, target_t_s=None):
        """Raw LiDAR points (float32 xyz, double timestamps) into G (target_t_s None: UndistortScanInG) or into the
        LiDAR frame at target_t_s (UndistortScan)."""
        t_s, S_LtoI, p_LinI = f64(t_s), f64(S_LtoI), f64(p_LinI)
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        n = len(t_s)
        out, bad = np.zeros((n, 3), np.float32), C.c_int64(0)
        fp = C.POINTER(C.c_float)
        _check(self.lib, self.lib.clic_undistort_scan(
            self.h, n, dptr(t_s), xyz.ctypes.data_as(fp), dptr(S_LtoI), dptr(p_LinI), float(t_offset_ns),
            1 if target_t_s is None else 0, 0.0 if target_t_s is None else float(target_t_s), out.ctypes.data_as(fp),
            C.byref(bad)))
        return out, bad.value

    def profile_read(self):
        ms, cnt = np.zeros(8), np.zeros(8, np.int64)
        _check(self.lib, self.lib.clic_profile_read(self.h, dptr(ms), cnt.ctypes.data_as(_abi.c_int64_p)))
        return ms, cnt

    def residual_summary(self):
        """Per-ResidualType cost and factor count of the solve problem at the current state (clic_residual_summary)."""
        cost, cnt = np.zeros(16), np.zeros(16, np.int64)
        _check(self.lib, self.lib.clic_residual_summary(self.h, dptr(cost), cnt.ctypes.data_as(_abi.c_int64_p)))
        return cost, cnt

    def pass_info(self):
        """How the pass executes: dict(fused, grid, part_ctas[6]) — see clic_pass_info."""
        fused, grid = C.c_int32(0), C.c_int32(0)
        part = (C.c_int32 * 6)()
        _check(self.lib, self.lib.clic_pass_info(self.h, C.byref(fused), C.byref(grid), part))
        return dict(fused=bool(fused.value), grid=grid.value, part_ctas=list(part))

    def pass_stamps(self, enable=True):
        """Per-CTA %globaltimer stamps (ns) of the last fused pass recorded while enabled: (grid, 8) uint64 array
        [start, end of factor phase, after the grid barrier, after key staging, end of assembly, end, warp 0 begins its slabs, warp 0 is through its slabs], or None;
        then sets the switch."""
        n = C.c_int32(0)
        buf = np.zeros((1024, 8), np.uint64)
        _check(self.lib, self.lib.clic_pass_stamps(self.h, int(bool(enable)), buf.size,
                                                   buf.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(n)))
        return buf[:n.value].copy() if n.value > 0 and buf[:n.value].any() else None

    def num_kernel_launches(self):
        n = C.c_int64(0)
        _check(self.lib, self.lib.clic_num_kernel_launches(self.h, C.byref(n)))
        return n.value
