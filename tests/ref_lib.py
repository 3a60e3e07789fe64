"""ctypes binding of oracle/_ref/libvtmref.so: the reference's own InterpolationFilter / RdCost object code
(scalar, SSE4.1 and AVX2 paths) behind oracle/ref_wrap.cpp.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_ref", "libvtmref.so")
_lib = None


def available():
    if not os.path.exists(_SO) and os.path.isdir("/root/reference/source/Lib/CommonLib"):
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO)
        L.ref_host_level.restype = C.c_int
        L.ref_init.argtypes = [C.c_int]
        L.ref_filter_hor.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 7
        L.ref_filter_ver.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 8
        L.ref_dist.restype = C.c_uint64
        L.ref_dist.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 5
        L.ref_dist_subshift.restype = C.c_int
        L.ref_dist_subshift.argtypes = [C.c_int] * 3
        L.ref_dmvr_cost.restype = C.c_uint64
        L.ref_dmvr_cost.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib = L
    return _lib


def init(level=-1):
    lib().ref_init(level)


def filter_hor(comp, src, x0, y0, w, h, frac, is_last, bd=10, clp=(0, 1023)):
    dst = np.zeros((h, w), dtype=np.int16)
    lib().ref_filter_hor(comp, src.ctypes.data + 2 * (y0 * src.shape[1] + x0), src.shape[1], dst.ctypes.data, w, w, h, frac,
                         int(is_last), bd, clp[0], clp[1])
    return dst


def filter_ver(comp, src, x0, y0, w, h, frac, is_first, is_last, bd=10, clp=(0, 1023)):
    dst = np.zeros((h, w), dtype=np.int16)
    lib().ref_filter_ver(comp, src.ctypes.data + 2 * (y0 * src.shape[1] + x0), src.shape[1], dst.ctypes.data, w, w, h, frac,
                         int(is_first), int(is_last), bd, clp[0], clp[1])
    return dst


def dist(org, cur, sub_shift_mode=0, hadamard=False, bd=10):
    return int(lib().ref_dist(org.ctypes.data, org.strides[0] // 2, cur.ctypes.data, cur.strides[0] // 2, org.shape[1],
                              org.shape[0], bd, sub_shift_mode, int(hadamard)))


def sub_shift(w, h, mode):
    return int(lib().ref_dist_subshift(w, h, mode))


def dmvr_cost(a, b, bd=10):
    return int(lib().ref_dmvr_cost(a.ctypes.data, a.strides[0] // 2, b.ctypes.data, b.strides[0] // 2, a.shape[1], a.shape[0], bd))
