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
        L.ref_filter_hor_x.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 9
        L.ref_filter_ver_x.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 10
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


def filter_hor_x(src, x0, y0, w, h, frac, is_last, bd=10, clp=(0, 1023), dmvr=False, alt_hpel=False):
    """luma filterHor with FILTER_DMVR (bilinear) or useAltHpelIf; src: 2-D int16 array, (x0, y0) the first output sample"""
    dst = np.zeros((h, w), dtype=np.int16)
    lib().ref_filter_hor_x(src.ctypes.data + 2 * (y0 * src.shape[1] + x0), src.shape[1], dst.ctypes.data, w, w, h, frac, int(is_last),
                           bd, clp[0], clp[1], int(dmvr), int(alt_hpel))
    return dst


def filter_ver_x(src, x0, y0, w, h, frac, is_first, is_last, bd=10, clp=(0, 1023), dmvr=False, alt_hpel=False):
    dst = np.zeros((h, w), dtype=np.int16)
    lib().ref_filter_ver_x(src.ctypes.data + 2 * (y0 * src.shape[1] + x0), src.shape[1], dst.ctypes.data, w, w, h, frac, int(is_first),
                           int(is_last), bd, clp[0], clp[1], int(dmvr), int(alt_hpel))
    return dst


def dist(org, cur, sub_shift_mode=0, hadamard=False, bd=10):
    return int(lib().ref_dist(org.ctypes.data, org.strides[0] // 2, cur.ctypes.data, cur.strides[0] // 2, org.shape[1],
                              org.shape[0], bd, sub_shift_mode, int(hadamard)))


def sub_shift(w, h, mode):
    return int(lib().ref_dist_subshift(w, h, mode))


def dmvr_cost(a, b, bd=10):
    return int(lib().ref_dmvr_cost(a.ctypes.data, a.strides[0] // 2, b.ctypes.data, b.strides[0] // 2, a.shape[1], a.shape[0], bd))


# ---- oracle/_ref/libvtmmap.so: the reference's own float-mapping source (compiled against oracle/eigen_func_shim) ----
_MAP_SO = os.path.join(ROOT, "oracle", "_ref", "libvtmmap.so")
_map = None


def map_available():
    if not os.path.exists(_MAP_SO) and os.path.isdir("/root/reference/source/Lib/CommonLib"):
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(_MAP_SO)


def maplib():
    global _map
    if _map is None:
        L = C.CDLL(_MAP_SO)
        L.refmap_create.restype = C.c_void_p
        L.refmap_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int]
        L.refmap_set_epipole.argtypes = [C.c_float] * 3
        L.refmap_rodrigues.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.refmap_reproject.argtypes = [C.c_void_p] + [C.c_int] * 11 + [C.c_void_p, C.c_void_p]
        L.refmap_mpa_cache.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refmap_set_epipole2.argtypes = [C.c_float] * 3
        L.refmap_mvp.argtypes = [C.c_void_p] + [C.c_int] * 18 + [C.c_void_p]
        _map = L
    return _map


class RefMap:
    """MVReprojection of the reference itself (MVReprojection.cpp + MotionModels/*.cpp + Projection.cpp + Coordinate.cpp)."""

    def __init__(self, w, h, offset4x4=1.0, ged_flavor=1):
        self.L, self.w, self.h = maplib(), w, h
        self.h_ = self.L.refmap_create(w, h, offset4x4, ged_flavor)

    def set_epipole(self, e):
        self.L.refmap_set_epipole(float(e[0]), float(e[1]), float(e[2]))

    def rodrigues(self, e):
        R = np.zeros(9, dtype=np.float32)
        self.L.refmap_rodrigues(self.h_, float(e[0]), float(e[1]), float(e[2]), R.ctypes.data)
        return R

    def reproject(self, pos, size, mv, model, comp=0, chroma_fmt=0, cur_poc=0, ref_poc=1):
        sx = 1 if (comp and chroma_fmt in (1, 2)) else 0
        sy = 1 if (comp and chroma_fmt == 1) else 0
        n = (size[1] // (4 >> sy)) * (size[0] // (4 >> sx))
        ox = np.zeros(n, dtype=np.int32)
        oy = np.zeros(n, dtype=np.int32)
        self.L.refmap_reproject(self.h_, pos[0], pos[1], size[0], size[1], mv[0], mv[1], model, comp, chroma_fmt, cur_poc, ref_poc,
                                ox.ctypes.data, oy.ctypes.data)
        return ox, oy

    def set_epipole2(self, e):
        """epipole of the (curPOC 7, *) pairs"""
        self.L.refmap_set_epipole2(float(e[0]), float(e[1]), float(e[2]))

    def mvp(self, pos, mv, model_orig, model_des, cand, cur, shift=(4, 4), poc_orig=0, poc_des=0):
        out = np.zeros(2, dtype=np.int32)
        self.L.refmap_mvp(self.h_, pos[0], pos[1], mv[0], mv[1], model_orig, model_des, shift[0], shift[1], poc_orig, poc_des,
                          cand[0], cand[1], cand[2], cand[3], cur[0], cur[1], cur[2], cur[3], out.ctypes.data)
        return int(out[0]), int(out[1])

    def mpa_cache(self, model):
        n = (self.w // 4) * (self.h // 4)
        px = np.zeros(n, dtype=np.float32)
        py = np.zeros(n, dtype=np.float32)
        v = np.zeros(n, dtype=np.uint8)
        self.L.refmap_mpa_cache(self.h_, model, px.ctypes.data, py.ctypes.data, v.ctypes.data)
        return px, py, v
