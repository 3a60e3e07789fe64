"""CPU tests of the PRODUCT's host-compilable code and of the C-ABI library (no compute without a GPU)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "tests", "native")
BUILD = os.path.join(NATIVE, "_build")
CXXFLAGS = ["-O2", "-std=c++17", "-msse4.1", "-ffp-contract=off", "-fno-builtin"]


def _build(src, out, extra):
    os.makedirs(BUILD, exist_ok=True)
    deps = [os.path.join(NATIVE, src)] + [os.path.join(ROOT, "vvc-extension-mm_b200", "csrc", f)
                                          for f in ("mm_libm.cuh", "mm_map.cuh", "mm_htaps.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++"] + CXXFLAGS + extra + ["-x", "c++", os.path.join(NATIVE, src), "-o", out, "-lm"])
    return out


def test_libm_port_pinned_against_box_libm():
    """mm_libm.cuh (glibc 2.39 sinf/cosf/tanf/atanf/atan2f/acosf/asinf/roundf restated) == the real libm.
    Strided here (~8M inputs per function); the exhaustive 2^32 run is tests/golden/libm_pin_exhaustive_r01.log."""
    exe = _build("libm_pin.cpp", os.path.join(BUILD, "libm_pin"), ["-pthread"])
    out = subprocess.run([exe, "509", "8"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "total mismatches 0" in out.stdout


def test_fir_coefficient_table_equals_eight_tap_filter():
    """csrc/mm_htaps.h (k_fir's alignment-folded coefficient words, built on the host) reproduces the plain 8-tap sum for
    all 4 alignments x 16 phases x 4 outputs, and carries the vertical pass' words."""
    exe = _build("htaps_check.cpp", os.path.join(BUILD, "htaps_check"), [])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 mismatches" in out.stdout


def test_product_mapping_host_build_equals_oracle():
    """mm_map.cuh compiled for the host == oracle, all models, luma + chroma, poles/seam/zero-MV cases."""
    so = _build("map_host.cpp", os.path.join(BUILD, "libmaphost.so"), ["-fPIC", "-shared"])
    Hn = C.CDLL(so)
    Hn.mmhost_create.restype = C.c_void_p
    Hn.mmhost_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
    Hn.mmhost_destroy.argtypes = [C.c_void_p]
    Hn.mmhost_reproject.argtypes = [C.c_void_p] + [C.c_int] * 10 + [C.c_void_p] * 3
    import vvc_extension_mm_b200 as mm
    for (W, Hh) in ((2048, 1024), (4096, 2048), (1280, 640)):
        o = ol.Oracle(W, Hh)
        fma = 1  # this box's glibc picks __sinf_fma (checked by test_libm_port_pinned_against_box_libm)
        h = Hn.mmhost_create(W, Hh, 1.0, 1, fma)
        rng = np.random.default_rng(W)
        Rs = {7: mm.rodrigues([1, 0, 0]), 8: mm.rodrigues([0, 1, 0]), 9: mm.rodrigues([0, 0, 1]),
              10: mm.rodrigues([0.3, -0.5, 0.81])}
        assert np.array_equal(Rs[10].view(np.uint32), ol.Oracle.rodrigues([0.3, -0.5, 0.81]).view(np.uint32))
        o.set_rotation(10, Rs[10])
        sizes = [(8, 4), (4, 8), (8, 8), (16, 8), (8, 16), (16, 16), (32, 8), (64, 64), (128, 128), (4, 16), (64, 4)]
        bad = tot = 0
        for it in range(700):
            w, hh = sizes[rng.integers(len(sizes))]
            x = int(rng.integers(0, (W - w) // 4 + 1)) * 4
            y = int(rng.integers(0, (Hh - hh) // 4 + 1)) * 4
            if it % 7 == 0:
                y = 0 if it % 2 else Hh - hh
            if it % 9 == 0:
                x = 0 if it % 2 else W - w
            mv = (int(rng.integers(-900, 900)), int(rng.integers(-500, 500)))
            if it % 11 == 0: mv = (0, 0)
            if it % 13 == 0: mv = (mv[0], 0)
            if it % 17 == 0: mv = (0, mv[1])
            m = int(rng.integers(1, 11))
            for comp, fmt in ((0, 0), (1, 1)):
                px, py, pw, ph, cs = (x, y, w, hh, 0) if comp == 0 else (x // 2, y // 2, w // 2, hh // 2, 1)
                ox, oy = o.reproject((px, py), (pw, ph), mv, m, comp, fmt)
                gx = np.zeros_like(ox); gy = np.zeros_like(oy)
                R = Rs.get(m, np.zeros(9, np.float32)).astype(np.float32)
                Hn.mmhost_reproject(h, comp, cs, cs, px, py, pw, ph, mv[0], mv[1], m, R.ctypes.data, gx.ctypes.data, gy.ctypes.data)
                bad += int((ox != gx).sum() + (oy != gy).sum())
                tot += 2 * len(ox)
        Hn.mmhost_destroy(h)
        o.close()
        assert bad == 0 and tot > 100000


def test_abi_library_exports_every_declared_symbol():
    import vvc_extension_mm_b200 as mm
    mm.build_library()
    hdr = open(os.path.join(ROOT, "include", "mmgpu.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|void|const char\*|void\*|uint64_t|float)\s+(mmgpu_[a-z_0-9]+)\(", hdr, re.M)))
    assert len(declared) >= 20
    lib = C.CDLL(mm.library_path())
    for name in declared:
        assert hasattr(lib, name), "library does not export " + name
    assert sorted(mm.exported_symbols()) == declared


def test_no_cpu_fallback_without_gpu():
    """The product must fail loudly when there is no usable sm_100 device."""
    import torch
    import vvc_extension_mm_b200 as mm
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mm.MmGpuError):
        mm.MmGpu(2048, 1024)


def test_host_helpers_match_oracle():
    import vvc_extension_mm_b200 as mm
    for e in ([1, 0, 0], [0, 0, 1], [0, 0, -1], [0.2, 0.4, -0.7], [0, 0, 0], [-3.5, 2.0, 1.0]):
        a, b = mm.rodrigues(e), ol.Oracle.rodrigues(e)
        nan = np.isnan(a)
        assert np.array_equal(nan, np.isnan(b))          # NaN sign/payload is irrelevant (isNaN().select)
        assert np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32))
    for v in (0, 1, -1, 1 << 24, -(1 << 24) + 5, 12345678):
        assert mm.fixed_to_float(v, 24) == float(ol.lib().orc_fixed_to_float(v, 24))


def test_workload_partition_covers_frame():
    from vvc_extension_mm_b200 import workload as wl
    b, mv, prec = wl.frame_batch(2048, 1024, seed=3, cands_per_block=2)
    cover = np.zeros((1024, 2048), dtype=np.int32)
    for x, y, w, h in zip(b["x"], b["y"], b["w"], b["h"]):
        cover[y:y + h, x:x + w] += 1
    assert (cover == 1).all()
    rows = [wl.shard_ctu_rows(8, r, 4) for r in range(4)]
    assert sorted(sum(rows, [])) == list(range(8))


def _build_adapter():
    import vvc_extension_mm_b200 as mm
    mm.build_library()
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "adapter_check")
    src = os.path.join(NATIVE, "adapter_check.cpp")
    hdr = os.path.join(ROOT, "vvc-extension-mm_b200", "host", "mm_gpu_adapter.h")
    if not os.path.exists(exe) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(exe):
        subprocess.check_call(["g++", "-O2", "-std=c++11", src, "-o", exe, "-L", os.path.dirname(mm.library_path()), "-lmmgpu",
                               "-Wl,-rpath," + os.path.dirname(mm.library_path())])
    return exe


def test_cpp_host_adapter_builds_and_fails_loudly_without_gpu():
    """host/mm_gpu_adapter.h (the reference-shaped C++ host side) links against the C ABI; init() throws like a
    reference CHECK when no sm_100 device exists."""
    import torch
    exe = _build_adapter()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("nogpu-throws")


def test_dmvr_decide_host_equals_oracle():
    """mmgpu_dmvr_decide (host side of a12: integer-search bookkeeping + error surface) vs the oracle's restatement
    of InterPrediction.cpp:1958-2048, 2157-2175, 2502-2583, on random and crafted cost surfaces."""
    import vvc_extension_mm_b200 as mm
    rng = np.random.default_rng(7)
    n = 0
    for it in range(4000):
        kind = it % 5
        if kind == 0:
            c = rng.integers(0, 3000, size=25)
        elif kind == 1:                                   # bowl around a random minimum
            cx, cy = rng.integers(-2, 3, size=2)
            xs, ys = np.meshgrid(np.arange(-2, 3), np.arange(-2, 3))
            c = 500 + 90 * (xs - cx) ** 2 + 70 * (ys - cy) ** 2 + rng.integers(0, 40, size=(5, 5))
            c = c.reshape(-1)
        elif kind == 2:                                   # early termination: tiny centre cost
            c = rng.integers(0, 4000, size=25); c[12] = rng.integers(0, 300)
        elif kind == 3:                                   # ties / flat neighbours -> half-pel branches
            c = np.full(25, 1000); c[rng.integers(0, 25)] = 900; c[12] = 1300
        else:
            c = rng.integers(200, 100000, size=25)
        for (w, h) in ((16, 16), (8, 16), (16, 8)):
            a = mm.dmvr_decide(c, w, h)
            b = ol.dmvr_decide(c, w, h)
            assert a == b, (c, w, h, a, b)
            n += 1
    assert n == 12000


def test_search_enumerators_of_the_adapter_equal_oracle_restatement():
    """host/mm_gpu_adapter.h's tzDiamondRound / tz2Point / tzRaster list the points of one call of
    xTZ8PointDiamondSearch / xTZ2PointSearch / the raster scan in the reference's visiting order (tests/oracle_tz.py), for
    every distance 1..64, with and without corner points, start points inside / at / beyond the search-range border."""
    import oracle_tz as tz
    exe = os.path.join(BUILD, "tz_check")
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(NATIVE, "tz_check.cpp")
    hdr = os.path.join(ROOT, "vvc-extension-mm_b200", "host", "mm_gpu_adapter.h")
    if not os.path.exists(exe) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(exe):
        subprocess.check_call(["g++", "-O1", "-std=c++11", src, "-o", exe])
    rng = np.random.default_rng(14)
    cases, exp = [], []
    for _ in range(3000):
        sr = (-int(rng.integers(0, 70)), int(rng.integers(0, 70)), -int(rng.integers(0, 70)), int(rng.integers(0, 70)))
        pts = []
        visit = lambda x, y, nr, d: pts.extend((x, y, nr, d))  # noqa: E731
        kind = "DDDPR"[int(rng.integers(0, 5))]
        sx, sy = int(rng.integers(sr[0] - 3, sr[1] + 4)), int(rng.integers(sr[2] - 3, sr[3] + 4))
        if kind == "D":
            dist, corners = int(2 ** rng.integers(0, 7)), int(rng.integers(0, 2))
            tz.diamond(sr, sx, sy, dist, bool(corners), visit)
            cases.append("D %d %d %d %d %d %d %d %d" % (sr + (sx, sy, dist, corners)))
        elif kind == "P":
            nr = int(rng.integers(0, 9))
            tz.two_point(sr, sx, sy, nr, visit)
            cases.append("P %d %d %d %d %d %d %d" % (sr + (sx, sy, nr)))
        else:
            step = int(rng.integers(1, 9))
            tz.raster(sr, step, visit)
            cases.append("R %d %d %d %d %d" % (sr + (step,)))
        exp.append(pts)
    out = subprocess.run([exe], input="\n".join(cases) + "\n", capture_output=True, text=True)
    assert out.returncode == 0
    got = [[int(v) for v in line.split()] for line in out.stdout.split("\n")[:len(cases)]]
    assert got == exp


def test_mvp_conversion_host_build_equals_reference_source():
    """mm_mvp.cuh (MM-MVP conversion, the code behind mmgpu_mvp_convert) compiled for the host == the REFERENCE's
    own MVReprojection::motionVectorInDesiredMotionModel (oracle/_ref/libvtmmap.so: MVReprojection.cpp + MotionModels/*.cpp
    compiled unchanged), for every ordered pair of the 11 motion models, random positions / blocks / MVs, luma and chroma MV
    precisions, zero MVs, poles, the ERP seam, and two different epipoles for GEODESIC_CAMPOSE."""
    import ref_lib as rl
    if not rl.map_available():
        pytest.skip("oracle/_ref/libvtmmap.so not built (reference sources absent)")
    import vvc_extension_mm_b200 as mm
    so = _build("map_host.cpp", os.path.join(BUILD, "libmaphost.so"), ["-fPIC", "-shared"])
    Hn = C.CDLL(so)
    Hn.mmhost_create.restype = C.c_void_p
    Hn.mmhost_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
    Hn.mmhost_mvp.argtypes = [C.c_void_p] + [C.c_int] * 8 + [C.c_void_p, C.c_void_p] + [C.c_int] * 9 + [C.c_void_p]
    for (W, Hh, flavor) in ((2048, 1024, 1), (1280, 640, 0)):
        r = rl.RefMap(W, Hh, 1.0, flavor)
        e1, e2 = [0.3, -0.5, 0.81], [0.1, 0.9, -0.2]
        r.set_epipole(e1)
        r.set_epipole2(e2)
        R = {7: mm.rodrigues([1, 0, 0]), 8: mm.rodrigues([0, 1, 0]), 9: mm.rodrigues([0, 0, 1])}
        R1, R2 = mm.rodrigues(e1), mm.rodrigues(e2)
        h = Hn.mmhost_create(W, Hh, 1.0, flavor, 1)
        rng = np.random.default_rng(W)
        n = bad = nz = 0
        for mo in range(11):
            for md in range(11):
                for k in range(40):
                    cw, chh = [int(2 ** rng.integers(2, 8)) for _ in range(2)]
                    uw, uh = [int(2 ** rng.integers(2, 8)) for _ in range(2)]
                    cx, cy = int(rng.integers(0, (W - cw) // 4 + 1)) * 4, int(rng.integers(0, (Hh - chh) // 4 + 1)) * 4
                    ux, uy = int(rng.integers(0, (W - uw) // 4 + 1)) * 4, int(rng.integers(0, (Hh - uh) // 4 + 1)) * 4
                    if k % 7 == 0: uy = 0                      # pole
                    if k % 11 == 0: ux = W - uw                # ERP seam
                    pos = (ux + int(rng.integers(0, uw)), uy + int(rng.integers(0, uh)))
                    mv = (int(rng.integers(-600, 600)), int(rng.integers(-300, 300)))
                    if k % 13 == 0: mv = (0, 0)
                    if k % 17 == 0: mv = (mv[0], 0)
                    shift = (4, 4) if k % 3 else (5, 5)
                    po, pd = (7 if k % 2 else 0), (7 if k % 4 < 2 else 0)        # which epipole each side uses
                    exp = r.mvp(pos, mv, mo, md, (cx, cy, cw, chh), (ux, uy, uw, uh), shift, po, pd)
                    Ro = R.get(mo, R2 if po == 7 else R1)
                    Rd = R.get(md, R2 if pd == 7 else R1)
                    out = np.zeros(2, dtype=np.int32)
                    Hn.mmhost_mvp(h, pos[0], pos[1], mv[0], mv[1], mo, md, shift[0], shift[1], Ro.ctypes.data, Rd.ctypes.data,
                                  int(po == pd), cx, cy, cw, chh, ux, uy, uw, uh, out.ctypes.data)
                    n += 1
                    nz += int(exp != (0, 0))
                    if (int(out[0]), int(out[1])) != exp:
                        bad += 1
                        if bad < 5:
                            print("MVP mismatch", W, mo, md, pos, mv, shift, (cx, cy, cw, chh), (ux, uy, uw, uh), exp, out)
        assert bad == 0, (bad, n)
        assert nz > n // 2


def test_committed_integration_patch_is_what_the_documented_edits_produce():
    """oracle/patches/mmgpu.patch (what oracle/vtm.mk applies to the build-time copy of four reference units) must be exactly the
    diff that make_patch.py's asserted, documented edits produce from the reference sources -- no hand-edited patch, no stale one."""
    if not os.path.isdir("/root/reference/source/Lib/CommonLib"):
        pytest.skip("reference sources absent")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_patch", os.path.join(ROOT, "oracle", "patches", "make_patch.py"))
    mp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mp)
    want = "".join(mp.build_patch())
    have = open(os.path.join(ROOT, "oracle", "patches", "mmgpu.patch"), encoding="latin-1").read()
    assert want == have
    # every hook family is in it
    for hook in ("mmglue::predInterBlkMM", "mmglue::predInterBlkClassic", "mmglue::predInterBlkClassicBdof", "mmglue::predInterBlkClassicAltHpel",
                 "mmglue::mvpConvert", "mmglue::meEval", "mmglue::endCollect", "mmglue::dmvrRun", "mmglue::dmvrFinalRun", "mmglue::verifyCompare",
                 "mmglue::meVerify", "mmglue::invalidate", "mmglue::init"):
        assert hook in have, hook
