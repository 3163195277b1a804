"""ctypes access to oracle/_ref/libps_ref_core.so: the reference's OWN libMultiArray / libBoostMath code, compiled here
from /root/reference (oracle/ref_core.cpp, `make -C oracle ref`).  TEST INFRASTRUCTURE: used by
tests/test_oracle_vs_ref.py and tests/golden/make_ref_golden.py to pin the oracle's restatement; never by the product."""
import ctypes as C
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libps_ref_core.so")
_fp, _dp, _ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(PATH)
        L.ref_gaussian_filter.argtypes = [C.c_double, _dp, C.c_int]
        L.ref_eig2d.argtypes = [_dp, _dp, _dp]
        L.ref_hc_inverse.argtypes = [_dp, _dp]
        L.ref_hc_compose.argtypes = [C.c_int, C.c_double, C.c_double, _dp]
        L.ref_prod3.argtypes = [_dp, _dp, _dp]
        L.ref_map_point.argtypes = [_dp, C.c_double, C.c_double, _dp, _dp]
        L.ref_transformed_bbox.argtypes = [_dp, C.c_int, C.c_int, _dp]
        L.ref_transform_fixed.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, _dp, C.c_float, C.c_int]
        L.ref_transform_resize.argtypes = [_fp, C.c_int, C.c_int, _dp, C.c_float, C.c_int, _fp, C.c_long, _ip, _ip, _dp]
        L.ref_transform_resize.restype = C.c_long
        L.ref_gauss_filter_diag2d.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_double, C.c_double]
        L.ref_gauss_filter_2d.argtypes = [_fp, _fp, C.c_int, C.c_int, _dp, C.c_int]
        L.ref_gauss_filter_2d_offset.argtypes = [_fp, _fp, C.c_int, C.c_int, _dp, _dp, C.c_int]
        L.ref_filter_1d_wraparound.argtypes = [_fp, _fp, C.c_int, _fp, C.c_int]
        L.ref_pointwise.argtypes = [C.c_int, _fp, _fp, C.c_float, C.c_long]
        L.ref_min_max.argtypes = [_fp, C.c_long, _fp, _fp]
        L.ref_bins.argtypes = [C.c_int, C.c_float, C.c_float, C.c_uint, C.c_int, C.c_double]
        L.ref_bins.restype = C.c_double
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(_fp)


def _d(a):
    a = np.ascontiguousarray(a, np.float64)
    return a, a.ctypes.data_as(_dp)


def gaussian_filter(sigma):
    out = np.empty(4096, np.float64)
    n = lib().ref_gaussian_filter(float(sigma), out.ctypes.data_as(_dp), out.size)
    assert n > 0
    return out[:n].copy()


def eig2d(Cm):
    _c, pc = _d(Cm)
    V, E = np.empty(4), np.empty(4)
    lib().ref_eig2d(pc, V.ctypes.data_as(_dp), E.ctypes.data_as(_dp))
    return V.reshape(2, 2), E.reshape(2, 2)


def hc_inverse(T):
    _t, pt = _d(T)
    out = np.empty(9)
    lib().ref_hc_inverse(pt, out.ctypes.data_as(_dp))
    return out.reshape(3, 3)


def transformed_bbox(T, w, h):
    _t, pt = _d(T)
    out = np.empty(4)
    lib().ref_transformed_bbox(pt, int(w), int(h), out.ctypes.data_as(_dp))
    return out


def transform_fixed(grid, out_shape, T, default_value, method):
    g = np.ascontiguousarray(grid, np.float32)
    out = np.empty(out_shape, np.float32)
    _t, pt = _d(T)
    lib().ref_transform_fixed(_f(g), g.shape[0], g.shape[1], _f(out), out.shape[0], out.shape[1], pt, float(default_value),
                              int(method))
    return out


def gauss_filter_2d(grid, Cm, sparse):
    g = np.ascontiguousarray(grid, np.float32)
    out = np.empty_like(g)
    _c, pc = _d(Cm)
    lib().ref_gauss_filter_2d(_f(g), _f(out), g.shape[0], g.shape[1], pc, int(bool(sparse)))
    return out


def gauss_filter_2d_offset(grid, Cm, offset, sparse):
    g = np.ascontiguousarray(grid, np.float32)
    out = np.empty_like(g)
    _c, pc = _d(Cm)
    _o, po = _d(offset)
    lib().ref_gauss_filter_2d_offset(_f(g), _f(out), g.shape[0], g.shape[1], pc, po, int(bool(sparse)))
    return out


def filter_1d_wraparound(col, taps):
    c = np.ascontiguousarray(col, np.float32)
    f = np.ascontiguousarray(taps, np.float32)
    out = np.empty_like(c)
    lib().ref_filter_1d_wraparound(_f(c), _f(out), c.size, _f(f), f.size)
    return out


def pointwise(op, a):
    a = np.ascontiguousarray(a, np.float32).copy()
    lib().ref_pointwise(int(op), _f(a), None, 0.0, a.size)
    return a


def bins(mn, mx, n):
    """(rot_from_index(i) for all i, index_from_rot of each of those centres)."""
    L = lib()
    centres = np.array([L.ref_bins(0, mn, mx, n, i, 0.0) for i in range(n)])
    back = np.array([L.ref_bins(2, mn, mx, n, 0, float(c)) for c in centres])
    return np.concatenate([centres, back])
