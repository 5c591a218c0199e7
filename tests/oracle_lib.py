"""Loads oracle/_build/libclic_oracle.so (building it if g++ is available and it is stale/missing).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use this."""
import ctypes as C
import os
import subprocess

import numpy as np

from clic_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "_build", "libclic_oracle.so")
_lib = None


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp"))]
    srcs.append(os.path.join(ROOT, "include", "clic_cuda.h"))
    return any(os.path.getmtime(s) > t for s in srcs)


def load():
    global _lib
    if _lib is None:
        if _stale():
            subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
        _lib = _abi.bind(SO)
        _lib.clic_oracle_set_threads.restype = C.c_int
        _lib.clic_oracle_set_threads.argtypes = [C.c_int]
        _lib.clic_oracle_num_blocks.restype = C.c_int64
        _lib.clic_oracle_num_blocks.argtypes = [C.c_void_p, C.c_int32]
        _lib.clic_oracle_block_jacobian.restype = C.c_int
        _lib.clic_oracle_block_jacobian.argtypes = [C.c_void_p, C.c_int64, C.c_int32, _abi.c_int32_p, _abi.c_double_p,
                                                    _abi.c_double_p]
        _lib.clic_oracle_last_marg_system.restype = C.c_int
        _lib.clic_oracle_last_marg_system.argtypes = [_abi.c_double_p, _abi.c_double_p, _abi.c_int32_p, _abi.c_int32_p]
    return _lib


def last_marg_system(lib):
    """The assembled [dropped | kept] system of the oracle's last marginalization, before the Schur complement:
    (A (pos, pos), b (pos,), m)."""
    pos, m = C.c_int32(0), C.c_int32(0)
    assert lib.clic_oracle_last_marg_system(None, None, C.byref(pos), C.byref(m)) == 0
    A, b = np.zeros((pos.value, pos.value)), np.zeros(pos.value)
    assert lib.clic_oracle_last_marg_system(_abi.dptr(A), _abi.dptr(b), C.byref(pos), C.byref(m)) == 0
    return A, b, m.value


def schur_extended_precision(A, b, m):
    """A' = A_rr - A_rm A_mm^-1 A_mr, b' = b_r - A_rm A_mm^-1 b_m by symmetric elimination of the first m unknowns in
    80-bit long double with diagonal (Jacobi) scaling: the yardstick both Schur routes are measured against."""
    ld = np.longdouble
    M = np.zeros((A.shape[0], A.shape[0] + 1), ld)
    M[:, :-1] = 0.5 * (A.astype(ld) + A.T.astype(ld))
    M[:, -1] = b.astype(ld)
    d = np.sqrt(np.maximum(np.diag(M[:, :-1]), ld(1e-300)))
    M[:, :-1] /= np.outer(d, d)
    M[:, -1] /= d
    for k in range(m):
        piv = M[k, k]
        f = M[k + 1:, k] / piv
        M[k + 1:, k:] -= np.outer(f, M[k, k:])
    Ap = M[m:, m:-1] * np.outer(d[m:], d[m:])
    bp = M[m:, -1] * d[m:]
    return np.asarray(Ap, np.float64), np.asarray(bp, np.float64)


def block_jacobian(lib, est, block, apply_loss=False):
    D = est.layout().D
    nres = C.c_int32(0)
    res = np.zeros(1024)
    J = np.zeros((1024, max(D, 1)))
    st = lib.clic_oracle_block_jacobian(est.h, block, int(apply_loss), C.byref(nres), _abi.dptr(res), None)
    assert st == 0
    J = np.zeros((nres.value, D))
    st = lib.clic_oracle_block_jacobian(est.h, block, int(apply_loss), C.byref(nres), _abi.dptr(res), _abi.dptr(J))
    assert st == 0
    return res[:nres.value].copy(), J
