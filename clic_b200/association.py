"""Python host-side mirror of the association half of clic::LidarOdometry
(src/lidar_odometry/lidar_odometry.cpp:472-657: SetTargetMap + FindCorrespondence) on top of the C ABI.
No compute here: the search and the fits run inside libclic_cuda.so on the GPU."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi
from ._abi import c_float_p, c_int64_p, dptr
from .estimator import _check, cuda_library


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return None if a is None else a.ctypes.data_as(c_float_p)


@dataclass
class Correspondences:
    """Type-uniform batches in the reference's order (corner features first), ready for AddLoamLines / AddLoamPlanes."""
    line_t: np.ndarray
    line_point: np.ndarray   # (n, 3) feature in the LiDAR frame
    line_geo: np.ndarray     # (n, 6) geo_point, geo_normal
    line_src: np.ndarray     # index of the source corner feature
    plane_t: np.ndarray
    plane_point: np.ndarray
    plane_geo: np.ndarray    # (n, 4) geo_plane
    plane_src: np.ndarray


class FeatureMap:
    """SetTargetMap: the down-sampled surface / corner feature map (float32 xyz) kept resident on the device."""

    def __init__(self, surface_xyz, corner_xyz=None, lib=None):
        self.lib = lib or cuda_library()
        surf = _f32(surface_xyz).reshape(-1, 3)
        corner = None if corner_xyz is None else _f32(corner_xyz).reshape(-1, 3)
        self.h = _abi.H()
        _check(self.lib, self.lib.clic_feature_map_create(_fp(surf), len(surf), _fp(corner),
                                                          0 if corner is None else len(corner), C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.clic_feature_map_destroy(self.h)
            self.h = _abi.H()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def FindCorrespondence(self, corner_xyz_L, corner_t, corner_xyz_M, surf_xyz_L, surf_t, surf_xyz_M):
        """lf_cur (features in the LiDAR frame + timestamps) and lf_cur_in_M (the same features in the map frame)."""
        cL, cM, sL, sM = (_f32(a).reshape(-1, 3) for a in (corner_xyz_L, corner_xyz_M, surf_xyz_L, surf_xyz_M))
        ct = np.ascontiguousarray(corner_t, dtype=np.float64)
        st = np.ascontiguousarray(surf_t, dtype=np.float64)
        nc, ns = len(ct), len(st)
        cap_c = self.lib.clic_correspondence_capacity(nc)
        cap_s = self.lib.clic_correspondence_capacity(ns)
        lt, lp, lg, ls = np.zeros(cap_c), np.zeros((cap_c, 3)), np.zeros((cap_c, 6)), np.zeros(cap_c, np.int64)
        pt, pp, pg, ps = np.zeros(cap_s), np.zeros((cap_s, 3)), np.zeros((cap_s, 4)), np.zeros(cap_s, np.int64)
        nl, npl = C.c_int64(0), C.c_int64(0)
        _check(self.lib, self.lib.clic_find_correspondence(
            self.h, nc, _fp(cL), dptr(ct), _fp(cM), ns, _fp(sL), dptr(st), _fp(sM),
            C.byref(nl), dptr(lt), dptr(lp), dptr(lg), ls.ctypes.data_as(c_int64_p),
            C.byref(npl), dptr(pt), dptr(pp), dptr(pg), ps.ctypes.data_as(c_int64_p)))
        a, b = nl.value, npl.value
        return Correspondences(lt[:a], lp[:a], lg[:a], ls[:a], pt[:b], pp[:b], pg[:b], ps[:b])

    def AddLoamFromScan(self, est, corner_xyz, corner_t, surf_xyz, surf_t, S_GtoM, p_GinM, S_LtoI, p_LinI, weight,
                        cor_downsample=1, t_offset_ns=0.0, marg_this_factor=False):
        """GetLoamFeatureAssociation + one AddLoamMeasurementAnalytic per kept correspondence, on the device:
        returns (n_lines, n_planes, n_dropped)."""
        cL, sL = _f32(corner_xyz).reshape(-1, 3), _f32(surf_xyz).reshape(-1, 3)
        ct = np.ascontiguousarray(corner_t, dtype=np.float64)
        st = np.ascontiguousarray(surf_t, dtype=np.float64)
        f = [np.ascontiguousarray(a, dtype=np.float64) for a in (S_GtoM, p_GinM, S_LtoI, p_LinI)]
        nl, npl, nd = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        _check(self.lib, self.lib.clic_add_loam_from_scan(
            est.h, self.h, len(ct), _fp(cL), dptr(ct), len(st), _fp(sL), dptr(st), dptr(f[0]), dptr(f[1]), dptr(f[2]),
            dptr(f[3]), float(t_offset_ns), int(cor_downsample), float(weight), int(bool(marg_this_factor)),
            C.byref(nl), C.byref(npl), C.byref(nd)))
        return nl.value, npl.value, nd.value

    def neighbours(self, kind):
        """(n_selected, 5) neighbour indices of the last call; kind 0 corner, 1 surface."""
        n = C.c_int64(0)
        _check(self.lib, self.lib.clic_correspondence_neighbours(self.h, kind, 0, None, C.byref(n)))
        idx = np.zeros((n.value, 5), np.int32)
        if n.value:
            _check(self.lib, self.lib.clic_correspondence_neighbours(self.h, kind, n.value, _abi.iptr(idx), C.byref(n)))
        return idx

    def launches(self):
        """(kernels launched so far, device ms of the last call's kernels)."""
        n, ms = C.c_int64(0), C.c_double(0)
        _check(self.lib, self.lib.clic_feature_map_launches(self.h, C.byref(n), C.byref(ms)))
        return n.value, ms.value
