"""ctypes mirror of include/mmgpu.h.

Names follow the reference: blocks (pu.blocks[compID]), Mv in 1/16 pel, MotionModelID values,
reference slots holding Picture::getRecoBuf planes with margins.
"""
import ctypes as C
import os
import numpy as np
from .build import library_path

CLASSIC, MPA_FB, MPA_LR, MPA_TB, TAN, T3D, ROT, GED_X, GED_Y, GED_Z, GED_CAMPOSE = range(11)
MODEL_NAMES = {0: "CLASSIC", 1: "MPA_FB", 2: "MPA_LR", 3: "MPA_TB", 4: "TAN", 5: "3DT", 6: "ROT", 7: "GED_X",
               8: "GED_Y", 9: "GED_Z", 10: "GED_CAMPOSE"}
OOB_PRED, OOB_ME = 0, 1
DIST_SAD, DIST_HAD = 0, 1
HAD_SCALAR, HAD_SIMD = 0, 1

EXPORTS = [
    "mmgpu_create", "mmgpu_destroy", "mmgpu_last_error", "mmgpu_sync", "mmgpu_stream", "mmgpu_launch_count",
    "mmgpu_kernel_timing_enable", "mmgpu_kernel_timing_read",
    "mmgpu_set_ref", "mmgpu_set_ref_device", "mmgpu_set_ref_unpadded", "mmgpu_rodrigues", "mmgpu_fixed_to_float",
    "mmgpu_set_rotation", "mmgpu_reproject", "mmgpu_mc", "mmgpu_add_avg", "mmgpu_set_org", "mmgpu_set_org_device",
    "mmgpu_set_ref_unpadded_wrap", "mmgpu_me_batch", "mmgpu_me_batch_submit", "mmgpu_me_batch_wait", "mmgpu_plan_create", "mmgpu_plan_destroy", "mmgpu_me_batch_device", "mmgpu_plan_samples",
    "mmgpu_mc_batch_device", "mmgpu_dmvr_costs", "mmgpu_dmvr_decide", "mmgpu_mc_jobs", "mmgpu_selftest_tan_sincos", "mmgpu_me_batch_device_multiref", "mmgpu_set_ref_unpadded_device", "mmgpu_set_dist_oob_mode", "mmgpu_mc_jobs_device", "mmgpu_set_refs_unpadded_device", "mmgpu_mvp_convert", "mmgpu_mc_bdof",
    "mmgpu_dmvr_classic_costs", "mmgpu_dmvr_final_mc", "mmgpu_mc_alt_hpel", "mmgpu_pull_refs_unpadded", "mmgpu_multicast_copy",
]


class MmGpuError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("pic_w", C.c_int32), ("pic_h", C.c_int32), ("chroma_fmt", C.c_int32), ("bit_depth", C.c_int32),
                ("max_cu", C.c_int32), ("margin", C.c_int32), ("offset4x4", C.c_float), ("ged_flavor", C.c_int32),
                ("max_refs", C.c_int32), ("device", C.c_int32), ("libm_fma", C.c_int32), ("had_mode", C.c_int32)]


class Mv(C.Structure):
    _fields_ = [("hor", C.c_int32), ("ver", C.c_int32)]


class Block(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("model", C.c_int32),
                ("ref_slot", C.c_int32), ("rot_slot", C.c_int32), ("cand_first", C.c_int32), ("cand_count", C.c_int32),
                ("org_x", C.c_int32), ("org_y", C.c_int32), ("reserved", C.c_int32)]


BLOCK_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("model", "<i4"), ("ref_slot", "<i4"),
                        ("rot_slot", "<i4"), ("cand_first", "<i4"), ("cand_count", "<i4"), ("org_x", "<i4"),
                        ("org_y", "<i4"), ("reserved", "<i4")])
MV_DTYPE = np.dtype([("hor", "<i4"), ("ver", "<i4")])
JOB_DTYPE = np.dtype([("comp", "<i4"), ("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("model", "<i4"), ("bi", "<i4"),
                      ("ref_slot0", "<i4"), ("rot_slot0", "<i4"), ("mv0_hor", "<i4"), ("mv0_ver", "<i4"),
                      ("ref_slot1", "<i4"), ("rot_slot1", "<i4"), ("mv1_hor", "<i4"), ("mv1_ver", "<i4")])
DMVR_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("model", "<i4"), ("ref_slot0", "<i4"),
                       ("ref_slot1", "<i4"), ("rot_slot0", "<i4"), ("rot_slot1", "<i4"), ("mv0_hor", "<i4"),
                       ("mv0_ver", "<i4"), ("mv1_hor", "<i4"), ("mv1_ver", "<i4")])
DMVR_CLASSIC_DTYPE = np.dtype([("w", "<i4"), ("h", "<i4"), ("ref_slot0", "<i4"), ("ref_slot1", "<i4"), ("org_x0", "<i4"), ("org_y0", "<i4"),
                               ("org_x1", "<i4"), ("org_y1", "<i4"), ("frac_x0", "<i4"), ("frac_y0", "<i4"), ("frac_x1", "<i4"), ("frac_y1", "<i4")])
DMVR_MC_JOB_DTYPE = np.dtype([("comp", "<i4"), ("ref_slot", "<i4"), ("org_x", "<i4"), ("org_y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("pad", "<i4"),
                              ("dint_x", "<i4"), ("dint_y", "<i4"), ("frac_x", "<i4"), ("frac_y", "<i4"), ("alt_hpel", "<i4"), ("bio", "<i4"),
                              ("dst_offset", "<i4"), ("dst_stride", "<i4")])

_lib = None


def load_library():
    """Load libmmgpu.so; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise MmGpuError("libmmgpu.so is not built: run __graft_entry__.build() (nvcc, sm_100a)")
    L = C.CDLL(path)
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    L.mmgpu_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.mmgpu_destroy.argtypes = [vp]
    L.mmgpu_last_error.restype = C.c_char_p
    L.mmgpu_last_error.argtypes = [vp]
    L.mmgpu_sync.argtypes = [vp]
    L.mmgpu_stream.restype = vp
    L.mmgpu_stream.argtypes = [vp]
    L.mmgpu_launch_count.restype = u64
    L.mmgpu_launch_count.argtypes = [vp]
    L.mmgpu_kernel_timing_enable.argtypes = [vp, i32]
    L.mmgpu_kernel_timing_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(u64)]
    L.mmgpu_selftest_tan_sincos.argtypes = [vp, C.c_uint32, C.POINTER(u64)]
    for f in (L.mmgpu_set_ref, L.mmgpu_set_ref_device, L.mmgpu_set_ref_unpadded):
        f.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(C.c_int32)]
    L.mmgpu_set_ref_unpadded_wrap.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(C.c_int32), i32]
    L.mmgpu_set_ref_unpadded_device.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(C.c_int32), i32, vp]
    L.mmgpu_set_refs_unpadded_device.argtypes = [vp, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(vp), C.POINTER(C.c_int32), i32, vp]
    L.mmgpu_rodrigues.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.mmgpu_fixed_to_float.restype = C.c_float
    L.mmgpu_fixed_to_float.argtypes = [C.c_int32, C.c_int32]
    L.mmgpu_set_rotation.argtypes = [vp, i32, C.POINTER(C.c_float)]
    L.mmgpu_reproject.argtypes = [vp, i32, i32, i32, i32, i32, i32, Mv, i32, i32, vp, vp]
    L.mmgpu_pull_refs_unpadded.argtypes = [vp, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(vp), C.POINTER(C.c_int32), i32, vp]
    L.mmgpu_multicast_copy.argtypes = [vp, vp, vp, C.c_size_t, i32, vp]
    L.mmgpu_mvp_convert.argtypes = [vp, i32, i32, Mv] + [i32] * 14 + [C.POINTER(Mv)]
    L.mmgpu_mc.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, Mv, i32, i32, i32, i32, i32, i32, vp, i32]
    L.mmgpu_add_avg.argtypes = [vp, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32]
    L.mmgpu_set_org.argtypes = [vp, vp, i32, i32, i32]
    L.mmgpu_set_org_device.argtypes = [vp, vp, i32, i32, i32]
    L.mmgpu_me_batch.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp]
    L.mmgpu_me_batch_submit.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp]
    L.mmgpu_me_batch_wait.argtypes = [vp, i32]
    L.mmgpu_plan_create.argtypes = [vp, i32, vp, i32, C.POINTER(vp)]
    L.mmgpu_plan_destroy.argtypes = [vp]
    L.mmgpu_me_batch_device.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp]
    L.mmgpu_set_dist_oob_mode.argtypes = [vp, i32]
    L.mmgpu_me_batch_device_multiref.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), vp]
    L.mmgpu_plan_samples.restype = u64
    L.mmgpu_plan_samples.argtypes = [vp]
    L.mmgpu_mc_batch_device.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32]
    L.mmgpu_dmvr_costs.argtypes = [vp, i32, vp, i32, i32, vp]
    L.mmgpu_dmvr_classic_costs.argtypes = [vp, i32, vp, vp]
    L.mmgpu_dmvr_final_mc.argtypes = [vp, i32, vp, vp, i32]
    L.mmgpu_mc_alt_hpel.argtypes = [vp, i32, i32, i32, i32, i32, Mv, i32, i32, i32, vp, i32]
    L.mmgpu_mc_bdof.argtypes = [vp, i32, i32, i32, i32, i32, Mv, i32, i32, vp]
    L.mmgpu_mc_jobs.argtypes = [vp, i32, vp, i32, i32, C.POINTER(vp), C.POINTER(C.c_int32)]
    L.mmgpu_mc_jobs_device.argtypes = [vp, i32, vp, i32, i32, C.POINTER(vp), C.POINTER(C.c_int32)]
    L.mmgpu_dmvr_decide.argtypes = [vp, i32, i32, vp, vp]
    _lib = L
    return L


def exported_symbols():
    """Names include/mmgpu.h declares that the loaded library really exports."""
    L = load_library()
    return [n for n in EXPORTS if hasattr(L, n)]


def rodrigues(epipole):
    """GeodesicMotionModel::setEpipole (GeodesicMotionModel.cpp:14-46): epipole -> row-major 3x3."""
    L = load_library()
    e = (C.c_float * 3)(*[float(v) for v in epipole])
    r = (C.c_float * 9)()
    L.mmgpu_rodrigues(e, r)
    return np.array(list(r), dtype=np.float32)


def fixed_to_float(fixed, precision=24):
    return float(load_library().mmgpu_fixed_to_float(int(fixed), int(precision)))


def dmvr_decide(costs25, w, h):
    """Host decision logic of xProcessDMVRProjected: -> (notZeroCost, (dmv_hor, dmv_ver) in 1/16 pel, minCost)."""
    L = load_library()
    c = np.ascontiguousarray(costs25, dtype=np.uint64)
    d = np.zeros(2, dtype=np.int32)
    m = np.zeros(1, dtype=np.uint64)
    nz = L.mmgpu_dmvr_decide(c.ctypes.data, int(w), int(h), d.ctypes.data, m.ctypes.data)
    return int(nz), (int(d[0]), int(d[1])), int(m[0])


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)


class MmPlan:
    def __init__(self, gpu, handle, n_blocks, n_cands):
        self.gpu, self.h, self.n_blocks, self.n_cands = gpu, handle, n_blocks, n_cands

    @property
    def samples(self):
        return int(self.gpu.L.mmgpu_plan_samples(self.h))

    def close(self):
        if self.h:
            self.gpu.L.mmgpu_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MmGpu:
    """One context per InterPrediction / InterSearch instance (MVReprojection::init on the device)."""

    def __init__(self, pic_w, pic_h, chroma_fmt=1, bit_depth=10, max_cu=128, margin=288, offset4x4=1.0, ged_flavor=1,
                 max_refs=4, device=0, libm_fma=-1, had_mode=-1):
        self.L = load_library()
        self.cfg = Config(pic_w, pic_h, chroma_fmt, bit_depth, max_cu, margin, offset4x4, ged_flavor, max_refs, device,
                          libm_fma, had_mode)
        h = C.c_void_p()
        rc = self.L.mmgpu_create(C.byref(self.cfg), C.byref(h))
        if rc != 0:
            raise MmGpuError("mmgpu_create failed (%d): %s" % (rc, self.L.mmgpu_last_error(None).decode()))
        self.h = h
        self._keep = {}

    def close(self):
        if getattr(self, "h", None):
            self.L.mmgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise MmGpuError("mmgpu error %d: %s" % (rc, self.L.mmgpu_last_error(self.h).decode()))

    # -- plumbing ------------------------------------------------------------------------------
    def sync(self):
        self._chk(self.L.mmgpu_sync(self.h))

    @property
    def stream(self):
        return int(self.L.mmgpu_stream(self.h))

    @property
    def launch_count(self):
        return int(self.L.mmgpu_launch_count(self.h))

    def selftest_tan_sincos(self, stride=1):
        """Checksum of the device's float(sin/cos((double)x)) over every stride-th float of |x| <= pi/2 (mmgpu.h)."""
        v = C.c_uint64()
        self._chk(self.L.mmgpu_selftest_tan_sincos(self.h, int(stride), C.byref(v)))
        return int(v.value)

    def kernel_timing(self, on=True):
        self._chk(self.L.mmgpu_kernel_timing_enable(self.h, int(on)))

    def kernel_timing_read(self):
        """(device ms of k_map launches, device ms of k_fir launches, number of batches) since the last read"""
        m, f, n = C.c_double(), C.c_double(), C.c_uint64()
        self._chk(self.L.mmgpu_kernel_timing_read(self.h, C.byref(m), C.byref(f), C.byref(n)))
        return m.value, f.value, int(n.value)

    # -- reference pictures --------------------------------------------------------------------
    def _planes(self, origins, strides):
        n = len(origins)
        p = (C.c_void_p * 3)(*([_ptr(o) for o in origins] + [None] * (3 - n)))
        s = (C.c_int32 * 3)(*(list(strides) + [0] * (3 - n)))
        return p, s

    def set_ref(self, slot, poc, origins, strides):
        """origins[c]: HOST address of sample (0,0) of component c (margins around it are uploaded)."""
        p, s = self._planes(origins, strides)
        self._chk(self.L.mmgpu_set_ref(self.h, slot, poc, p, s))

    def set_ref_device(self, slot, poc, origins, strides, keep=None):
        p, s = self._planes(origins, strides)
        self._keep[("ref", slot)] = keep
        self._chk(self.L.mmgpu_set_ref_device(self.h, slot, poc, p, s))

    def set_ref_unpadded(self, slot, poc, planes):
        """planes: list of 2-D int16 numpy arrays without margins; padding happens on the device."""
        p, s = self._planes([a.ctypes.data for a in planes], [a.strides[0] // 2 for a in planes])
        self._chk(self.L.mmgpu_set_ref_unpadded(self.h, slot, poc, p, s))

    def set_ref_unpadded_wrap(self, slot, poc, planes, wrap_offset):
        """like set_ref_unpadded with Picture::extendWrapBorder's horizontally wrapped margins (wrap_offset in luma samples)."""
        p, s = self._planes([a.ctypes.data for a in planes], [a.strides[0] // 2 for a in planes])
        self._chk(self.L.mmgpu_set_ref_unpadded_wrap(self.h, slot, poc, p, s, int(wrap_offset)))

    def set_ref_unpadded_device(self, slot, poc, ptrs, strides, wrap_offset=0, stream=None, keep=None):
        """Un-padded planes already in device memory (ptrs: device addresses): D2D copy + border extension on `stream`
        (a raw cudaStream_t value; None = the context's stream)."""
        p, s = self._planes(list(ptrs), list(strides))
        self._keep[("refu", slot)] = keep
        self._chk(self.L.mmgpu_set_ref_unpadded_device(self.h, slot, poc, p, s, int(wrap_offset), C.c_void_p(stream) if stream else None))

    def set_refs_unpadded_device(self, slots, pocs, ptrs, strides, wrap_offset=0, stream=None, keep=None):
        """n pictures at once (ptrs / strides: n x 3 device addresses / sample strides): one launch copies and pads them all."""
        n = len(slots)
        sl = (C.c_int32 * n)(*[int(v) for v in slots])
        pc = (C.c_int32 * n)(*[int(v) for v in pocs])
        flat_p = [int(v) for row in ptrs for v in (list(row) + [0] * (3 - len(row)))]
        flat_s = [int(v) for row in strides for v in (list(row) + [0] * (3 - len(row)))]
        pp = (C.c_void_p * (3 * n))(*flat_p)
        ss = (C.c_int32 * (3 * n))(*flat_s)
        self._keep[("refus", tuple(slots))] = keep
        self._chk(self.L.mmgpu_set_refs_unpadded_device(self.h, n, sl, pc, pp, ss, int(wrap_offset), C.c_void_p(stream) if stream else None))

    def set_rotation(self, rot_slot, R):
        r = (C.c_float * 9)(*[float(v) for v in R])
        self._chk(self.L.mmgpu_set_rotation(self.h, rot_slot, r))

    def pull_refs_unpadded(self, slots, pocs, ptrs, strides, max_ctas=24, stream=None, keep=None):
        """set_refs_unpadded_device with the sources in PEER memory: persistent kernels of max_ctas CTAs pull the rows over NVLink and
        write the padded planes (exchange + border extension in one pass)."""
        n = len(slots)
        sl = (C.c_int32 * n)(*[int(v) for v in slots])
        pc = (C.c_int32 * n)(*[int(v) for v in pocs])
        flat_p = [int(v) for row in ptrs for v in (list(row) + [0] * (3 - len(row)))]
        flat_s = [int(v) for row in strides for v in (list(row) + [0] * (3 - len(row)))]
        pp = (C.c_void_p * (3 * n))(*flat_p)
        ss = (C.c_int32 * (3 * n))(*flat_s)
        self._keep[("refus", tuple(slots))] = keep
        self._chk(self.L.mmgpu_pull_refs_unpadded(self.h, n, sl, pc, pp, ss, int(max_ctas), C.c_void_p(stream) if stream else None))

    def multicast_copy(self, src_ptr, dst_multicast_ptr, nbytes, max_ctas=16, stream=None):
        """local device memory -> multicast address of symmetric memory (one multimem.st per 16 bytes; the switch fans out)"""
        self._chk(self.L.mmgpu_multicast_copy(self.h, C.c_void_p(int(src_ptr)), C.c_void_p(int(dst_multicast_ptr)), int(nbytes), int(max_ctas),
                                              C.c_void_p(stream) if stream else None))

    def mvp_convert(self, pos, mv, model_orig, model_desired, cand, cur, shift=(4, 4), rot_slots=(0, 0)):
        """MVReprojection::motionVectorInDesiredMotionModel; cand / cur = (x, y, w, h) of the candidate's / the current block."""
        out = Mv()
        self._chk(self.L.mmgpu_mvp_convert(self.h, pos[0], pos[1], Mv(mv[0], mv[1]), model_orig, model_desired, shift[0], shift[1],
                                           rot_slots[0], rot_slots[1], cand[0], cand[1], cand[2], cand[3], cur[0], cur[1], cur[2], cur[3],
                                           C.byref(out)))
        return out.hor, out.ver

    # -- a1 ------------------------------------------------------------------------------------
    def reproject(self, pos, size, mv, model, comp=0, chroma_fmt=0, rot_slot=0):
        sx = 1 if (comp and chroma_fmt in (1, 2)) else 0
        sy = 1 if (comp and chroma_fmt == 1) else 0
        n = (size[1] // (4 >> sy)) * (size[0] // (4 >> sx))
        ox = np.zeros(n, dtype=np.int32)
        oy = np.zeros(n, dtype=np.int32)
        self._chk(self.L.mmgpu_reproject(self.h, comp, chroma_fmt, pos[0], pos[1], size[0], size[1], Mv(mv[0], mv[1]),
                                         model, rot_slot, ox.ctypes.data, oy.ctypes.data))
        return ox, oy

    # -- a5/a6/a8 ------------------------------------------------------------------------------
    def mc(self, ref_slot, comp, chroma_fmt, pos, size, mv, model, rot_slot=0, bi=False, clp=(0, 1023), oob=OOB_PRED):
        dst = np.zeros((size[1], size[0]), dtype=np.int16)
        self._chk(self.L.mmgpu_mc(self.h, ref_slot, comp, chroma_fmt, pos[0], pos[1], size[0], size[1], Mv(mv[0], mv[1]),
                                  model, rot_slot, int(bi), clp[0], clp[1], oob, dst.ctypes.data, size[0]))
        return dst

    def mc_bdof(self, ref_slot, pos, size, mv, clp=(0, 1023)):
        """Classic bi MC of a luma block into the BDOF work buffer: returns the (h + 4) x (w + 4) buffer (outer frame untouched = 0)."""
        w, h = size
        dst = np.zeros((h + 4, w + 4), dtype=np.int16)
        self._chk(self.L.mmgpu_mc_bdof(self.h, ref_slot, pos[0], pos[1], w, h, Mv(mv[0], mv[1]), clp[0], clp[1], dst.ctypes.data))
        return dst

    def add_avg(self, src0, src1, clp=(0, 1023)):
        h, w = src0.shape
        dst = np.zeros((h, w), dtype=np.int16)
        self._chk(self.L.mmgpu_add_avg(self.h, src0.ctypes.data, src0.strides[0] // 2, src1.ctypes.data,
                                       src1.strides[0] // 2, dst.ctypes.data, w, w, h, clp[0], clp[1]))
        return dst

    # -- distortion ----------------------------------------------------------------------------
    def set_org(self, org):
        """org: 2-D int16 numpy array (pcPatternKey block or the whole current picture)."""
        self._keep["org"] = org
        self._chk(self.L.mmgpu_set_org(self.h, org.ctypes.data, org.strides[0] // 2, org.shape[1], org.shape[0]))

    def set_org_device(self, ptr, stride, w, h, keep=None):
        self._keep["org"] = keep
        self._chk(self.L.mmgpu_set_org_device(self.h, ptr, stride, w, h))

    def set_dist_oob_mode(self, oob):
        """OOB_ME (motion search, default) or OOB_PRED (distortions of decided MVs: merge SATD, xGetInterPredictionError)."""
        self._chk(self.L.mmgpu_set_dist_oob_mode(self.h, int(oob)))

    def me_batch(self, blocks, mvs, mv_prec_shift=4, dist=DIST_SAD, sub_shift=0, clp=(0, 1023), out=None):
        """blocks: numpy array of BLOCK_DTYPE; mvs: numpy array of MV_DTYPE -> uint64 distortions."""
        blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        mvs = np.ascontiguousarray(mvs, dtype=MV_DTYPE)
        if out is None:
            out = np.zeros(len(mvs), dtype=np.uint64)
        self._chk(self.L.mmgpu_me_batch(self.h, len(blocks), blocks.ctypes.data, len(mvs), mvs.ctypes.data, mv_prec_shift,
                                        dist, sub_shift, clp[0], clp[1], out.ctypes.data))
        return out

    def me_batch_submit(self, slot, org, blocks, mvs, out, mv_prec_shift=4, dist=DIST_SAD, sub_shift=0, clp=(0, 1023)):
        """Pipelined me_batch: org (2-D int16), blocks (BLOCK_DTYPE), mvs (MV_DTYPE) and out (uint64) are numpy arrays
        (pinned for real overlap) that must stay alive and untouched until me_batch_wait(slot)."""
        assert org.dtype == np.int16 and blocks.dtype == BLOCK_DTYPE and mvs.dtype == MV_DTYPE and out.dtype == np.uint64
        assert blocks.flags.c_contiguous and mvs.flags.c_contiguous and out.flags.c_contiguous and len(out) == len(mvs)
        h, w = org.shape
        self._chk(self.L.mmgpu_me_batch_submit(self.h, slot, org.ctypes.data, org.strides[0] // 2, w, h, len(blocks),
                                               blocks.ctypes.data, len(mvs), mvs.ctypes.data, mv_prec_shift, dist, sub_shift,
                                               clp[0], clp[1], out.ctypes.data))
        self._keep[("submit", slot)] = (org, blocks, mvs, out)

    def me_batch_wait(self, slot):
        self._chk(self.L.mmgpu_me_batch_wait(self.h, slot))
        return self._keep.pop(("submit", slot))[3]

    def mc_jobs(self, jobs, planes, clp=(0, 1023)):
        """Whole-picture reconstruction shape: jobs (JOB_DTYPE) write their predictions into `planes`
        (list of 2-D int16 numpy arrays Y[,Cb,Cr], modified in place)."""
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        p, s = self._planes([a.ctypes.data for a in planes], [a.strides[0] // 2 for a in planes])
        self._chk(self.L.mmgpu_mc_jobs(self.h, len(jobs), jobs.ctypes.data, clp[0], clp[1], p, s))
        return planes

    def mc_jobs_device(self, jobs, ptrs, strides, clp=(0, 1023)):
        """mc_jobs with the destination planes in device memory (ptrs: device addresses); asynchronous on the context's stream."""
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        self._keep["jobs"] = jobs
        p, s = self._planes(list(ptrs), list(strides))
        self._chk(self.L.mmgpu_mc_jobs_device(self.h, len(jobs), jobs.ctypes.data, clp[0], clp[1], p, s))

    def dmvr_costs(self, subpus, clp=(0, 1023)):
        """subpus: numpy array of DMVR_DTYPE -> (n, 25) uint64 xDMVRCost values at m_pSearchOffset."""
        subpus = np.ascontiguousarray(subpus, dtype=DMVR_DTYPE)
        out = np.zeros((len(subpus), 25), dtype=np.uint64)
        self._chk(self.L.mmgpu_dmvr_costs(self.h, len(subpus), subpus.ctypes.data, clp[0], clp[1], out.ctypes.data))
        return out

    def dmvr_classic_costs(self, subpus):
        """subpus: DMVR_CLASSIC_DTYPE -> (n, 25) uint64 xDMVRCost of xProcessDMVR's bilinear search (index 12 = centre, raw)."""
        subpus = np.ascontiguousarray(subpus, dtype=DMVR_CLASSIC_DTYPE)
        out = np.zeros((len(subpus), 25), dtype=np.uint64)
        self._chk(self.L.mmgpu_dmvr_classic_costs(self.h, len(subpus), subpus.ctypes.data, out.ctypes.data))
        return out

    def dmvr_final_mc(self, jobs, dst_samples):
        """jobs: DMVR_MC_JOB_DTYPE -> flat int16 buffer holding every job's output at its dst_offset (xFinalPaddedMCForDMVR)."""
        jobs = np.ascontiguousarray(jobs, dtype=DMVR_MC_JOB_DTYPE)
        out = np.zeros(int(dst_samples), dtype=np.int16)
        self._chk(self.L.mmgpu_dmvr_final_mc(self.h, len(jobs), jobs.ctypes.data, out.ctypes.data, int(dst_samples)))
        return out

    def mc_alt_hpel(self, ref_slot, pos, size, mv, bi=False, clp=(0, 1023)):
        """Classic luma xPredInterBlk with the alternative half-pel filter (pu.cu->imv == IMV_HPEL)."""
        w, h = size
        dst = np.zeros((h, w), dtype=np.int16)
        self._chk(self.L.mmgpu_mc_alt_hpel(self.h, ref_slot, pos[0], pos[1], w, h, Mv(mv[0], mv[1]), int(bi), clp[0], clp[1], dst.ctypes.data, w))
        return dst

    def plan(self, blocks, n_cands):
        blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        h = C.c_void_p()
        self._chk(self.L.mmgpu_plan_create(self.h, len(blocks), blocks.ctypes.data, n_cands, C.byref(h)))
        return MmPlan(self, h, len(blocks), n_cands)

    def me_batch_device(self, plan, d_mvs, d_dist, mv_prec_shift=4, dist=DIST_SAD, sub_shift=0, clp=(0, 1023)):
        self._chk(self.L.mmgpu_me_batch_device(self.h, plan.h, _ptr(d_mvs), mv_prec_shift, dist, sub_shift, clp[0],
                                               clp[1], _ptr(d_dist)))

    def me_batch_device_multiref(self, plan, d_mvs, d_dist, ref_slots, rot_slots=None, mv_prec_shift=4, dist=DIST_SAD, sub_shift=0,
                                 clp=(0, 1023)):
        """One mapping, len(ref_slots) interpolation + distortion passes; d_dist: len(ref_slots) * n_cands uint64, reference-major."""
        rs = (C.c_int32 * len(ref_slots))(*[int(v) for v in ref_slots])
        ro = (C.c_int32 * len(ref_slots))(*[int(v) for v in rot_slots]) if rot_slots is not None else None
        self._chk(self.L.mmgpu_me_batch_device_multiref(self.h, plan.h, _ptr(d_mvs), mv_prec_shift, dist, sub_shift, clp[0], clp[1],
                                                        len(ref_slots), rs, ro, _ptr(d_dist)))

    def mc_batch_device(self, plan, d_mvs, d_dst, dst_stride, mv_prec_shift=4, clp=(0, 1023)):
        self._chk(self.L.mmgpu_mc_batch_device(self.h, plan.h, _ptr(d_mvs), mv_prec_shift, clp[0], clp[1], _ptr(d_dst),
                                               dst_stride))
