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
    return _lib


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
