"""ctypes loader of oracle/libgf_oracle.so (plain-C restatement of BatchLSIGF).
TEST INFRASTRUCTURE ONLY -- see oracle/gf_oracle.c."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libgf_oracle.so")


def build():
    src = os.path.join(HERE, "gf_oracle.c")
    if not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "libgf_oracle.so"], check=True, capture_output=True)
    return LIB


def _call(fn, x, S, h, b, out):
    B, G, N = x.shape
    F, E, K, G2 = h.shape
    assert E == 1 and G2 == G
    S = np.ascontiguousarray(S.reshape(B, N, N))
    assert S.dtype in (np.float32, np.float64)
    x = np.ascontiguousarray(x, np.float32)
    h = np.ascontiguousarray(h, np.float32)
    bp = None if b is None else np.ascontiguousarray(b, np.float32).ravel()
    vp = C.c_void_p
    fn.argtypes = [vp, vp, C.c_int, vp, vp, vp] + [C.c_int] * 5
    fn.restype = C.c_int
    rc = fn(x.ctypes.data, S.ctypes.data, int(S.dtype == np.float64), h.ctypes.data,
            None if bp is None else bp.ctypes.data, out.ctypes.data, B, N, G, F, K)
    assert rc == 0
    return out


def graph_filter_f32(x, S, h, b=None):
    lib = C.CDLL(build())
    return _call(lib.gf_oracle_f32, x, S, h, b, np.empty((x.shape[0], h.shape[0], x.shape[2]), np.float32))


def graph_filter_f64(x, S, h, b=None):
    lib = C.CDLL(build())
    return _call(lib.gf_oracle_f64, x, S, h, b, np.empty((x.shape[0], h.shape[0], x.shape[2]), np.float64))
