"""ctypes view of include/clic_cuda.h.

The same prototypes bind the product library (clic_b200/lib/libclic_cuda.so) and — from tests/ and
bench.py's CPU-baseline leg only — the oracle library, which exports identical symbols.
"""
import ctypes as C
import os

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_float_p = C.POINTER(C.c_float)


class WindowDesc(C.Structure):
    _fields_ = [
        ("t0_ns", C.c_int64),
        ("dt_ns", C.c_int64),
        ("n_knots", C.c_int32),
        ("knot_id0", C.c_int32),
        ("knot_q", c_double_p),
        ("knot_p", c_double_p),
        ("S_LtoI", C.c_double * 4),
        ("p_LinI", C.c_double * 3),
        ("S_CtoI", C.c_double * 4),
        ("p_CinI", C.c_double * 3),
        ("t_offset_ns", C.c_double * 3),
        ("gravity", C.c_double * 3),
        ("n_bias", C.c_int32),
        ("bias_id0", C.c_int32),
        ("bias", c_double_p),
        ("n_landmarks", C.c_int32),
        ("reserved0", C.c_int32),
        ("inv_depth", c_double_p),
    ]


class Options(C.Structure):
    _fields_ = [
        ("lock_traj", C.c_int32),
        ("lock_ab", C.c_int32),
        ("lock_wb", C.c_int32),
        ("lock_g", C.c_int32),
        ("lock_t_offset", C.c_int32 * 3),
        ("is_marg_state", C.c_int32),
        ("t_offset_padding_ns", C.c_int64),
        ("ctrl_to_be_opt_now", C.c_int32),
        ("ctrl_to_be_opt_later", C.c_int32),
        ("marg_bias_param", C.c_int32),
        ("marg_gravity_param", C.c_int32),
        ("marg_t_offset_param", C.c_int32),
        ("device", C.c_int32),
        ("stream", C.c_void_p),
        ("deterministic", C.c_int32),
        ("reserved0", C.c_int32),
        ("image_sqrt_info", C.c_double),
        ("reserved", C.c_int32 * 4),
    ]


class Summary(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32),
        ("num_successful_steps", C.c_int32),
        ("num_unsuccessful_steps", C.c_int32),
        ("termination", C.c_int32),
        ("termination_reason", C.c_int32),
        ("num_jacobian_evals", C.c_int32),
        ("num_residual_evals", C.c_int32),
        ("reserved", C.c_int32),
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("fixed_cost", C.c_double),
        ("final_radius", C.c_double),
        ("final_gradient_max_norm", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class Layout(C.Structure):
    _fields_ = [
        ("D", C.c_int32),
        ("first_free_knot", C.c_int32),
        ("n_free_knots", C.c_int32),
        ("col_bias_gyro", C.c_int32),
        ("col_bias_accel", C.c_int32),
        ("col_t_offset", C.c_int32 * 3),
        ("col_inv_depth", C.c_int32),
        ("n_free_landmarks", C.c_int32),
    ]


class PreIntegration(C.Structure):
    """clic_preintegration (include/clic_cuda.h): IntegrationBase reduced to what PreIntegrationFactor reads."""
    _fields_ = [
        ("sum_dt", C.c_double),
        ("delta_p", C.c_double * 3),
        ("delta_q", C.c_double * 4),
        ("delta_v", C.c_double * 3),
        ("linearized_ba", C.c_double * 3),
        ("linearized_bg", C.c_double * 3),
        ("jacobian", C.c_double * 225),
        ("sqrt_info", C.c_double * 225),
        ("gravity_norm", C.c_double),
    ]


class FactorDesc(C.Structure):
    """clic_factor_desc (include/clic_cuda.h)."""
    _fields_ = [
        ("kind", C.c_int32),
        ("geo_type", C.c_int32),
        ("t_a_ns", C.c_int64),
        ("t_b_ns", C.c_int64),
        ("d", C.c_double * 32),
        ("x", C.c_double * 16),
        ("pre", C.POINTER(PreIntegration)),
    ]


(FACTOR_POSE, FACTOR_LOCAL_VELOCITY, FACTOR_RELATIVE_ROTATION, FACTOR_IMAGE_DEPTH, FACTOR_PREINTEGRATION,
 FACTOR_IMAGE_3D2D, FACTOR_IMAGE_ONE_POSE, FACTOR_EPIPOLAR, FACTOR_LOAM_OPT_MAP_POSE, FACTOR_RELATIVE_LOAM) = range(10)

SENSOR_IMU, SENSOR_LIDAR, SENSOR_CAMERA = 0, 1, 2
GEO_LINE, GEO_PLANE = 0, 1
(BLOCK_KNOT_ROT, BLOCK_KNOT_POS, BLOCK_BIAS_GYRO, BLOCK_BIAS_ACCEL, BLOCK_GRAVITY, BLOCK_T_OFFSET,
 BLOCK_INV_DEPTH) = range(7)

STATUS_NAMES = {
    0: "CLIC_OK", 1: "CLIC_ERR_BAD_HANDLE", 2: "CLIC_ERR_BAD_ARGUMENT", 3: "CLIC_ERR_CUDA",
    4: "CLIC_ERR_NO_DEVICE", 5: "CLIC_ERR_NCCL", 6: "CLIC_ERR_NOT_POSITIVE_DEFINITE",
    7: "CLIC_ERR_UNSUPPORTED", 8: "CLIC_ERR_NOTHING_TO_KEEP",
}

# every symbol include/clic_cuda.h declares: name -> (restype, argtypes)
H = C.c_void_p  # opaque handles
PROTOTYPES = {
    "clic_last_error": (C.c_char_p, []),
    "clic_backend": (C.c_char_p, []),
    "clic_default_options": (None, [C.POINTER(Options)]),
    "clic_create": (C.c_int, [C.POINTER(WindowDesc), C.POINTER(Options), C.POINTER(H)]),
    "clic_destroy": (C.c_int, [H]),
    "clic_set_fixed_index": (C.c_int, [H, C.c_int32]),
    "clic_add_loam": (C.c_int, [H, C.c_int64, c_double_p, c_double_p, c_int32_p, c_double_p, c_double_p, c_double_p,
                                c_double_p, c_double_p, c_double_p, c_double_p, C.c_double, C.c_int32, c_int64_p]),
    "clic_add_loam_packed": (C.c_int, [H, C.c_int64, c_double_p, c_double_p, C.POINTER(C.c_uint8), c_double_p, c_double_p,
                                       c_double_p, c_double_p, c_double_p, C.c_double, C.c_int32, c_int64_p]),
    "clic_add_loam_planes": (C.c_int, [H, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                       c_double_p, C.c_double, C.c_int32, c_int64_p]),
    "clic_add_loam_lines": (C.c_int, [H, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                      c_double_p, C.c_double, C.c_int32, c_int64_p]),
    "clic_add_imu": (C.c_int, [H, C.c_int64, c_double_p, c_double_p, c_double_p, C.c_int32, c_double_p, C.c_int32,
                               c_int64_p]),
    "clic_add_bias": (C.c_int, [H, C.c_int32, C.c_int32, C.c_double, c_double_p, C.c_int32, C.c_int32]),
    "clic_add_gravity": (C.c_int, [H, c_double_p, c_double_p, C.c_int32]),
    "clic_add_pose": (C.c_int, [H, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_int64_p]),
    "clic_add_local_velocity": (C.c_int, [H, C.c_int64, c_double_p, c_double_p, C.c_double, c_int64_p]),
    "clic_add_relative_rotation": (C.c_int, [H, C.c_double, C.c_double, c_double_p, c_double_p, c_int32_p]),
    "clic_add_image_depth": (C.c_int, [H, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int32]),
    "clic_add_preintegration": (C.c_int, [H, C.c_int64, C.c_int64, C.POINTER(PreIntegration), C.c_int32, C.c_int32,
                                          c_int32_p]),
    "clic_evaluate_factor": (C.c_int, [H, C.POINTER(FactorDesc), c_int32_p, c_double_p, c_int32_p, c_double_p, C.c_int64]),
    "clic_add_global_velocity": (C.c_int, [H, C.c_double, c_double_p, C.c_double, c_int32_p]),
    "clic_add_image": (C.c_int, [H, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p, C.c_int32,
                                 C.c_int32, c_int64_p]),
    "clic_add_prior": (C.c_int, [H, H, C.c_int32, c_int32_p]),
    "clic_get_layout": (C.c_int, [H, C.POINTER(Layout)]),
    "clic_evaluate": (C.c_int, [H, c_double_p, c_double_p, c_double_p, c_double_p]),
    "clic_get_indices": (C.c_int, [H, C.c_int32, C.c_int64, c_int32_p, c_double_p, c_int64_p]),
    "clic_solve": (C.c_int, [H, C.c_int32, C.POINTER(Summary)]),
    "clic_read_state": (C.c_int, [H, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "clic_write_state": (C.c_int, [H, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "clic_query_poses": (C.c_int, [H, C.c_int64, c_double_p, c_double_p, c_double_p, C.c_double, c_double_p, c_double_p,
                                   c_int64_p]),
    "clic_undistort_scan": (C.c_int, [H, C.c_int64, c_double_p, C.POINTER(C.c_float), c_double_p, c_double_p, C.c_double,
                                      C.c_int32, C.c_double, C.POINTER(C.c_float), c_int64_p]),
    "clic_feature_map_create": (C.c_int, [c_float_p, C.c_int64, c_float_p, C.c_int64, C.POINTER(H)]),
    "clic_feature_map_destroy": (None, [H]),
    "clic_correspondence_capacity": (C.c_int64, [C.c_int64]),
    "clic_find_correspondence": (C.c_int, [H, C.c_int64, c_float_p, c_double_p, c_float_p,
                                           C.c_int64, c_float_p, c_double_p, c_float_p,
                                           c_int64_p, c_double_p, c_double_p, c_double_p, c_int64_p,
                                           c_int64_p, c_double_p, c_double_p, c_double_p, c_int64_p]),
    "clic_add_loam_from_scan": (C.c_int, [H, H, C.c_int64, c_float_p, c_double_p, C.c_int64, c_float_p, c_double_p,
                                          c_double_p, c_double_p, c_double_p, c_double_p, C.c_double, C.c_int32,
                                          C.c_double, C.c_int32, c_int64_p, c_int64_p, c_int64_p]),
    "clic_correspondence_neighbours": (C.c_int, [H, C.c_int32, C.c_int64, c_int32_p, c_int64_p]),
    "clic_feature_map_launches": (C.c_int, [H, c_int64_p, c_double_p]),
    "clic_marginalize": (C.c_int, [H, C.POINTER(H)]),
    "clic_prior_create": (C.c_int, [C.c_int32, c_double_p, c_double_p, C.c_int32, c_int32_p, c_int32_p, c_double_p,
                                    C.POINTER(H)]),
    "clic_prior_dims": (C.c_int, [H, c_int32_p, c_int32_p, c_int32_p]),
    "clic_prior_read": (C.c_int, [H, c_double_p, c_double_p, c_int32_p, c_int32_p, c_int32_p, c_double_p]),
    "clic_prior_normal": (C.c_int, [H, c_double_p, c_double_p]),
    "clic_prior_release": (C.c_int, [H]),
    "clic_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "clic_comm_init": (C.c_int, [H, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]),
    "clic_comm_p2p_export": (C.c_int, [H, C.POINTER(C.c_uint8)]),
    "clic_comm_p2p_import": (C.c_int, [H, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]),
    "clic_comm_p2p_ready": (C.c_int, []),
    "clic_comm_p2p_disable": (C.c_int, []),
    "clic_linearize_async": (C.c_int, [H, C.c_int32]),
    "clic_synchronize": (C.c_int, [H]),
    "clic_residual_summary": (C.c_int, [H, c_double_p, c_int64_p]),
    "clic_pass_info": (C.c_int, [H, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "clic_pass_stamps": (C.c_int, [H, C.c_int32, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]),
    "clic_num_kernel_launches": (C.c_int, [H, c_int64_p]),
    "clic_profile_enable": (C.c_int, [H, C.c_int32]),
    "clic_profile_read": (C.c_int, [H, c_double_p, c_int64_p]),
}


class ClicError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {msg}")
        self.status = status


def bind(path):
    """dlopen `path` and attach the prototypes of every symbol the header declares (raises if one is missing)."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    lib._clic_path = path
    return lib


def dptr(a):
    """double* of a C-contiguous float64 array (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int32_p)


def f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)
