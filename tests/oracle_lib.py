"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")

# MotionModelID (reference TypeDef.h:865-880)
CLASSIC, MPA_FB, MPA_LR, MPA_TB, TAN, T3D, ROT, GED_X, GED_Y, GED_Z, GED_CAMPOSE = range(11)
MODEL_NAMES = {1: "MPA_FB", 2: "MPA_LR", 3: "MPA_TB", 4: "TAN", 5: "3DT", 6: "ROT", 7: "GED_X", 8: "GED_Y",
               9: "GED_Z", 10: "GED_CAMPOSE"}
CHROMA_400, CHROMA_420 = 0, 1
OOB_PRED, OOB_ME = 0, 1
DIST_SAD, DIST_HAD = 0, 1
HAD_SCALAR, HAD_SIMD = 0, 1


class OrcCfg(C.Structure):
    _fields_ = [("pic_w", C.c_int), ("pic_h", C.c_int), ("chroma_fmt", C.c_int), ("max_cu", C.c_int),
                ("offset4x4", C.c_float), ("ged_flavor", C.c_int), ("bit_depth", C.c_int)]


def build():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(ROOT, "oracle", "mm_oracle.c")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcCfg)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_rodrigues.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_fixed_to_float.restype = C.c_float
        L.orc_fixed_to_float.argtypes = [C.c_int, C.c_int]
        L.orc_set_rotation.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.orc_reproject.argtypes = [C.c_void_p] + [C.c_int] * 9 + [C.c_void_p, C.c_void_p]
        for f in (L.orc_mpa_cache_x, L.orc_mpa_cache_y, L.orc_mpa_cache_vip):
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p, C.c_int]
        L.orc_filter_hor.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 7
        L.orc_filter_ver.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 8
        L.orc_mc_block.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 14 + [C.c_void_p, C.c_int]
        L.orc_mc_block_classic.argtypes = [C.c_void_p] + [C.c_int] * 13 + [C.c_void_p, C.c_int]
        L.orc_mc_bdof.argtypes = [C.c_void_p] + [C.c_int] * 10 + [C.c_void_p]
        L.orc_sad.restype = C.c_uint64
        L.orc_sad.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_had.restype = C.c_uint64
        L.orc_had.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_me_block.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 6 + \
            [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]
        L.orc_add_avg.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 5
        L.orc_extend_border.argtypes = [C.c_void_p] + [C.c_int] * 5
        L.orc_extend_wrap_border.argtypes = [C.c_void_p] + [C.c_int] * 6
        L.orc_dmvr_costs.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 9 + \
            [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_dmvr_cost.restype = C.c_uint64
        L.orc_dmvr_cost.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_dmvr_decide.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_dmvr_classic_costs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 11 + [C.c_void_p]
        L.orc_dmvr_final_mc.argtypes = [C.c_void_p, C.c_int] + [C.c_int] * 15 + [C.c_void_p, C.c_int]
        L.orc_mc_block_classic_alt.argtypes = [C.c_void_p, C.c_int] + [C.c_int] * 10 + [C.c_void_p, C.c_int]
        _lib = L
    return _lib


class PaddedPlane:
    """int16 plane with margins; .origin is the address of sample (0,0)."""

    def __init__(self, w, h, margin_x, margin_y, align=8):
        self.w, self.h, self.mx, self.my = w, h, margin_x, margin_y
        self.stride = (w + 2 * margin_x + align - 1) // align * align
        self.buf = np.zeros((h + 2 * margin_y, self.stride), dtype=np.int16)

    @property
    def inner(self):
        return self.buf[self.my:self.my + self.h, self.mx:self.mx + self.w]

    @property
    def origin(self):
        return self.buf.ctypes.data + 2 * (self.my * self.stride + self.mx)

    def extend(self):
        lib().orc_extend_border(self.origin, self.stride, self.w, self.h, self.mx, self.my)

    def extend_wrap(self, xoffset):
        lib().orc_extend_wrap_border(self.origin, self.stride, self.w, self.h, self.mx, self.my, int(xoffset))


class Oracle:
    def __init__(self, pic_w, pic_h, chroma_fmt=CHROMA_420, max_cu=128, offset4x4=1.0, ged_flavor=1, bit_depth=10):
        self.L = lib()
        self.cfg = OrcCfg(pic_w, pic_h, chroma_fmt, max_cu, offset4x4, ged_flavor, bit_depth)
        self.h = self.L.orc_create(C.byref(self.cfg))
        self.pic_w, self.pic_h, self.bit_depth = pic_w, pic_h, bit_depth

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @staticmethod
    def rodrigues(epipole):
        e = (C.c_float * 3)(*epipole)
        r = (C.c_float * 9)()
        lib().orc_rodrigues(e, r)
        return np.array(list(r), dtype=np.float32)

    def set_rotation(self, model, R):
        r = (C.c_float * 9)(*[float(x) for x in R])
        self.L.orc_set_rotation(self.h, model, r)

    def reproject(self, pos, size, mv, model, comp=0, chroma_fmt=CHROMA_400):
        sx = 1 if (comp and chroma_fmt in (1, 2)) else 0
        sy = 1 if (comp and chroma_fmt == 1) else 0
        rows, cols = size[1] // (4 >> sy), size[0] // (4 >> sx)
        ox = np.zeros(rows * cols, dtype=np.int32)
        oy = np.zeros(rows * cols, dtype=np.int32)
        self.L.orc_reproject(self.h, pos[0], pos[1], size[0], size[1], mv[0], mv[1], model, comp, chroma_fmt,
                             ox.ctypes.data, oy.ctypes.data)
        return ox, oy  # column-major (rows x cols)

    def mpa_cache(self, model):
        n = (self.pic_w // 4) * (self.pic_h // 4)
        x = np.ctypeslib.as_array(C.cast(self.L.orc_mpa_cache_x(self.h, model), C.POINTER(C.c_float)), (n,)).copy()
        y = np.ctypeslib.as_array(C.cast(self.L.orc_mpa_cache_y(self.h, model), C.POINTER(C.c_float)), (n,)).copy()
        v = np.ctypeslib.as_array(C.cast(self.L.orc_mpa_cache_vip(self.h, model), C.POINTER(C.c_uint8)), (n,)).copy()
        return x, y, v

    def mc_block(self, plane, comp, chroma_fmt, pos, size, mv, model, bi=False, clp=(0, 1023), oob=OOB_PRED):
        dst = np.zeros((size[1], size[0]), dtype=np.int16)
        self.L.orc_mc_block(self.h, plane.origin, plane.stride, comp, chroma_fmt, pos[0], pos[1], size[0], size[1],
                            mv[0], mv[1], model, int(bi), clp[0], clp[1], oob, dst.ctypes.data, size[0])
        return dst

    def mc_block_classic(self, plane, comp, chroma_fmt, pos, size, mv, bi=False, clp=(0, 1023)):
        dst = np.zeros((size[1], size[0]), dtype=np.int16)
        self.L.orc_mc_block_classic(plane.origin, plane.stride, comp, chroma_fmt, pos[0], pos[1], size[0], size[1],
                                    mv[0], mv[1], int(bi), self.bit_depth, clp[0], clp[1], dst.ctypes.data, size[0])
        return dst

    def mc_bdof(self, plane, pos, size, mv, clp=(0, 1023)):
        tmp = np.zeros((size[1] + 4, size[0] + 4), dtype=np.int16)
        self.L.orc_mc_bdof(plane.origin, plane.stride, pos[0], pos[1], size[0], size[1], mv[0], mv[1], self.bit_depth, clp[0], clp[1],
                           tmp.ctypes.data)
        return tmp

    def me_block(self, plane, org, org_stride_or_none, pos, size, model, mvs, mv_prec_shift=4, dist=DIST_SAD,
                 sub_shift=0, had_mode=HAD_SIMD, clp=(0, 1023)):
        """org: 2-D int16 array holding the original block (h x w view, any stride)."""
        mvs = np.asarray(mvs, dtype=np.int32).reshape(-1, 2)
        mh = np.ascontiguousarray(mvs[:, 0])
        mv = np.ascontiguousarray(mvs[:, 1])
        out = np.zeros(len(mvs), dtype=np.uint64)
        org_stride = org.strides[0] // 2
        self.L.orc_me_block(self.h, plane.origin, plane.stride, org.ctypes.data, org_stride, pos[0], pos[1], size[0],
                            size[1], model, len(mvs), mh.ctypes.data, mv.ctypes.data, mv_prec_shift, dist, sub_shift,
                            had_mode, clp[0], clp[1], out.ctypes.data)
        return out


def dmvr_costs(o, plane0, plane1, pos, size, mv0, mv1, model, R0=None, R1=None, clp=(0, 1023), simd=1):
    out = np.zeros(25, dtype=np.uint64)
    r0 = None if R0 is None else np.ascontiguousarray(R0, dtype=np.float32)
    r1 = None if R1 is None else np.ascontiguousarray(R1, dtype=np.float32)
    lib().orc_dmvr_costs(o.h, plane0.origin, plane0.stride, plane1.origin, plane1.stride, pos[0], pos[1], size[0], size[1],
                         mv0[0], mv0[1], mv1[0], mv1[1], model, None if r0 is None else r0.ctypes.data,
                         None if r1 is None else r1.ctypes.data, clp[0], clp[1], simd, out.ctypes.data)
    return out


def dmvr_classic_costs(plane0, plane1, org0, frac0, org1, frac1, size, bd=10):
    """xProcessDMVR's bilinear search of one sub-PU: org = xPrefetch's Rec_offset (luma), frac = clipped merge MV & 15."""
    out = np.zeros(25, dtype=np.uint64)
    lib().orc_dmvr_classic_costs(plane0.origin, plane0.stride, plane1.origin, plane1.stride, org0[0], org0[1], frac0[0], frac0[1],
                                 org1[0], org1[1], frac1[0], frac1[1], size[0], size[1], bd, out.ctypes.data)
    return out


def dmvr_final_mc(plane, is_luma, org, size, padsize, dint, frac, alt_hpel=False, bio=False, bd=10, clp=(0, 1023)):
    """xFinalPaddedMCForDMVR of one component of one list; bio: returns the (h + 4) x (w + 4) BDOF buffer."""
    w, h = size
    dst = np.zeros((h + 4, w + 4) if bio else (h, w), dtype=np.int16)
    lib().orc_dmvr_final_mc(plane.origin, plane.stride, int(is_luma), org[0], org[1], w, h, padsize, dint[0], dint[1], frac[0], frac[1],
                            int(alt_hpel), int(bio), bd, clp[0], clp[1], dst.ctypes.data, dst.shape[1])
    return dst


def mc_block_classic_alt(plane, pos, size, mv, bi=False, bd=10, clp=(0, 1023)):
    dst = np.zeros((size[1], size[0]), dtype=np.int16)
    lib().orc_mc_block_classic_alt(plane.origin, plane.stride, pos[0], pos[1], size[0], size[1], mv[0], mv[1], int(bi), bd, clp[0], clp[1],
                                   dst.ctypes.data, size[0])
    return dst


def dmvr_decide(costs25, w, h):
    c = np.ascontiguousarray(costs25, dtype=np.uint64)
    d = np.zeros(2, dtype=np.int32)
    m = np.zeros(1, dtype=np.uint64)
    nz = lib().orc_dmvr_decide(c.ctypes.data, w, h, d.ctypes.data, m.ctypes.data)
    return int(nz), (int(d[0]), int(d[1])), int(m[0])


def sad(org, cur, sub_shift=0):
    return lib().orc_sad(org.ctypes.data, org.strides[0] // 2, cur.ctypes.data, cur.strides[0] // 2, org.shape[1],
                         org.shape[0], sub_shift)


def had(org, cur, had_mode=HAD_SIMD):
    return lib().orc_had(org.ctypes.data, org.strides[0] // 2, cur.ctypes.data, cur.strides[0] // 2, org.shape[1],
                         org.shape[0], had_mode)
