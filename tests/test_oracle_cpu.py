"""CPU tests (no GPU): the oracle against hand-derived known answers for the integer steps
(filter taps, rounding, clipping, SAD, Hadamard) and against the committed golden fixtures."""
import os

import numpy as np
import pytest

import oracle_lib as ol

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

LUMA = np.array([[0, 0, 0, 64, 0, 0, 0, 0], [0, 1, -3, 63, 4, -2, 1, 0], [-1, 2, -5, 62, 8, -3, 1, 0],
                 [-1, 3, -8, 60, 13, -4, 1, 0], [-1, 4, -10, 58, 17, -5, 1, 0], [-1, 4, -11, 52, 26, -8, 3, -1],
                 [-1, 3, -9, 47, 31, -10, 4, -1], [-1, 4, -11, 45, 34, -10, 4, -1], [-1, 4, -11, 40, 40, -11, 4, -1],
                 [-1, 4, -10, 34, 45, -11, 4, -1], [-1, 4, -10, 31, 47, -9, 3, -1], [-1, 3, -8, 26, 52, -11, 4, -1],
                 [0, 1, -5, 17, 58, -10, 4, -1], [0, 1, -4, 13, 60, -8, 3, -1], [0, 1, -3, 8, 62, -5, 2, -1],
                 [0, 1, -2, 4, 63, -3, 1, 0]])   # InterpolationFilter.cpp:82-100


def _fh(src, frac, is_last, w, h, x0, y0, luma=1):
    dst = np.zeros((h, w), dtype=np.int16)
    ol.lib().orc_filter_hor(luma, src.ctypes.data + 2 * (y0 * src.shape[1] + x0), src.shape[1], dst.ctypes.data, w, w, h,
                            frac, is_last, 10, 0, 1023)
    return dst


def _fv(src, frac, is_first, is_last, w, h, x0, y0, luma=1):
    dst = np.zeros((h, w), dtype=np.int16)
    ol.lib().orc_filter_ver(luma, src.ctypes.data + 2 * (y0 * src.shape[1] + x0), src.shape[1], dst.ctypes.data, w, w, h,
                            frac, is_first, is_last, 10, 0, 1023)
    return dst


def test_filter_taps_sum_to_64_and_impulse_response():
    assert (LUMA.sum(axis=1) == 64).all()
    src = np.zeros((16, 32), dtype=np.int16)
    src[8, 16] = 16                       # impulse: (16*c + 32) >> 6 == (c+2)>>2 ... use bi path instead: exact taps
    for frac in range(1, 16):
        # first, not last (10 bit): (sum - 32768) >> 2  ->  with impulse 64: 16*c - 8192
        s2 = np.zeros((16, 32), dtype=np.int16); s2[8, 16] = 64
        out = _fh(s2, frac, 0, 12, 1, 10, 8)
        # out[x] filters taps src[x-3..x+4]; impulse at column 16 -> tap index k = 16 - (10 + x) + 3
        for x in range(12):
            k = 16 - (10 + x) + 3
            c = LUMA[frac][k] if 0 <= k < 8 else 0
            assert out[0, x] == 16 * c - 8192
        outv = _fv(np.ascontiguousarray(s2.T), frac, 1, 0, 1, 12, 8, 10)
        assert np.array_equal(outv[:, 0], out[0])


def test_filter_constant_field_roundtrip_and_clip():
    src = np.full((32, 32), 700, dtype=np.int16)
    for frac in range(16):
        assert (_fh(src, frac, 1, 4, 4, 8, 8) == 700).all()          # (64*700+32)>>6
        assert (_fv(src, frac, 1, 1, 4, 4, 8, 8) == 700).all()
        assert (_fh(src, frac, 0, 4, 4, 8, 8) == 700 * 16 - 8192).all()
    # two-pass last: ((64 * t) + 512 + 8192*64) >> 10 with t = 16p - 8192 gives p back
    tmp = np.full((32, 32), 700 * 16 - 8192, dtype=np.int16)
    assert (_fv(tmp, 5, 0, 1, 4, 4, 8, 8) == 700).all()
    # clipping: overshoot of the sharpest phase on a 0/1023 edge is clipped to [0,1023]
    edge = np.zeros((32, 32), dtype=np.int16); edge[:, 16:] = 1023
    out = _fh(edge, 8, 1, 16, 1, 8, 8)
    assert out.min() == 0 and out.max() == 1023
    raw = _fh(edge, 8, 0, 16, 1, 8, 8).astype(np.int32)               # unclipped intermediate shows the overshoot
    assert raw.max() > 1023 * 16 - 8192


def test_filter_copy_cases():
    rng = np.random.default_rng(0)
    src = rng.integers(0, 1024, size=(16, 16)).astype(np.int16)
    assert np.array_equal(_fh(src, 0, 1, 4, 4, 4, 4), src[4:8, 4:8])                       # first & last: copy
    assert np.array_equal(_fh(src, 0, 0, 4, 4, 4, 4), (src[4:8, 4:8] << 4) - 8192)         # (p<<4) - 8192
    t = ((src.astype(np.int32) << 4) - 8192).astype(np.int16)
    assert np.array_equal(_fv(t, 0, 0, 1, 4, 4, 4, 4), src[4:8, 4:8])                      # (v+8192+8)>>4 clipped


def test_chroma_filter():
    src = np.full((16, 16), 300, dtype=np.int16)
    for frac in range(32):
        assert (_fh(src, frac, 1, 2, 2, 4, 4, luma=0) == 300).all()
        assert (_fv(src, frac, 1, 1, 2, 2, 4, 4, luma=0) == 300).all()


def _had4_literal(org, cur):
    """Literal transcription of RdCost::xCalcHADs4x4 (RdCost.cpp:2290-2389)."""
    diff = (org.astype(np.int64) - cur.astype(np.int64)).reshape(16)
    m = np.zeros(16, dtype=np.int64); d = np.zeros(16, dtype=np.int64)
    m[0] = diff[0] + diff[12]; m[1] = diff[1] + diff[13]; m[2] = diff[2] + diff[14]; m[3] = diff[3] + diff[15]
    m[4] = diff[4] + diff[8]; m[5] = diff[5] + diff[9]; m[6] = diff[6] + diff[10]; m[7] = diff[7] + diff[11]
    m[8] = diff[4] - diff[8]; m[9] = diff[5] - diff[9]; m[10] = diff[6] - diff[10]; m[11] = diff[7] - diff[11]
    m[12] = diff[0] - diff[12]; m[13] = diff[1] - diff[13]; m[14] = diff[2] - diff[14]; m[15] = diff[3] - diff[15]
    d[0] = m[0] + m[4]; d[1] = m[1] + m[5]; d[2] = m[2] + m[6]; d[3] = m[3] + m[7]
    d[4] = m[8] + m[12]; d[5] = m[9] + m[13]; d[6] = m[10] + m[14]; d[7] = m[11] + m[15]
    d[8] = m[0] - m[4]; d[9] = m[1] - m[5]; d[10] = m[2] - m[6]; d[11] = m[3] - m[7]
    d[12] = m[12] - m[8]; d[13] = m[13] - m[9]; d[14] = m[14] - m[10]; d[15] = m[15] - m[11]
    m[0] = d[0] + d[3]; m[1] = d[1] + d[2]; m[2] = d[1] - d[2]; m[3] = d[0] - d[3]
    m[4] = d[4] + d[7]; m[5] = d[5] + d[6]; m[6] = d[5] - d[6]; m[7] = d[4] - d[7]
    m[8] = d[8] + d[11]; m[9] = d[9] + d[10]; m[10] = d[9] - d[10]; m[11] = d[8] - d[11]
    m[12] = d[12] + d[15]; m[13] = d[13] + d[14]; m[14] = d[13] - d[14]; m[15] = d[12] - d[15]
    d[0] = m[0] + m[1]; d[1] = m[0] - m[1]; d[2] = m[2] + m[3]; d[3] = m[3] - m[2]
    d[4] = m[4] + m[5]; d[5] = m[4] - m[5]; d[6] = m[6] + m[7]; d[7] = m[7] - m[6]
    d[8] = m[8] + m[9]; d[9] = m[8] - m[9]; d[10] = m[10] + m[11]; d[11] = m[11] - m[10]
    d[12] = m[12] + m[13]; d[13] = m[12] - m[13]; d[14] = m[14] + m[15]; d[15] = m[15] - m[14]
    satd = int(np.abs(d).sum())
    satd -= abs(int(d[0])); satd += abs(int(d[0])) >> 2
    return (satd + 1) >> 1


def _hadamard(n):
    Hm = np.array([[1]], dtype=np.int64)
    while Hm.shape[0] < n:
        Hm = np.block([[Hm, Hm], [Hm, -Hm]])
    return Hm


def test_had_against_literal_4x4_and_matrix_definition():
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = rng.integers(0, 1024, size=(4, 4)).astype(np.int16)
        b = rng.integers(0, 1024, size=(4, 4)).astype(np.int16)
        assert ol.had(a, b) == _had4_literal(a, b)
    for (w, h) in [(8, 8), (16, 8), (8, 16), (8, 4), (4, 8)]:
        a = rng.integers(0, 1024, size=(h, w)).astype(np.int16)
        b = rng.integers(0, 1024, size=(h, w)).astype(np.int16)
        D = a.astype(np.int64) - b.astype(np.int64)
        T = _hadamard(h) @ D @ _hadamard(w)
        s = int(np.abs(T).sum()) - abs(int(T[0, 0])) + (abs(int(T[0, 0])) >> 2)
        if (w, h) == (8, 8):
            exp_simd = exp_scalar = (s + 2) >> 2
        elif w * h == 128:
            exp_simd = ((s * 0xb504f334) >> 32) >> 2
            exp_scalar = int(s / np.sqrt(16.0 * 8) * 2)
        else:
            exp_simd = ((s * 0xb504f334) >> 32) >> 1
            exp_scalar = int(s / np.sqrt(4.0 * 8) * 2)
        assert ol.had(a, b, ol.HAD_SIMD) == exp_simd
        assert ol.had(a, b, ol.HAD_SCALAR) == exp_scalar


def test_had_tiling_rule():
    rng = np.random.default_rng(2)
    a = rng.integers(0, 1024, size=(64, 64)).astype(np.int16)
    b = rng.integers(0, 1024, size=(64, 64)).astype(np.int16)
    for (w, h, tw, th) in [(32, 8, 16, 8), (8, 32, 8, 16), (32, 4, 8, 4), (4, 32, 4, 8), (32, 32, 8, 8), (64, 16, 16, 8),
                           (16, 64, 8, 16), (16, 4, 8, 4)]:
        tot = sum(ol.had(a[y:y + th, x:x + tw], b[y:y + th, x:x + tw]) for y in range(0, h, th) for x in range(0, w, tw))
        assert ol.had(a[:h, :w], b[:h, :w]) == tot


def test_sad_subshift():
    rng = np.random.default_rng(3)
    a = rng.integers(0, 1024, size=(32, 16)).astype(np.int16)
    b = rng.integers(0, 1024, size=(32, 16)).astype(np.int16)
    d = np.abs(a.astype(np.int64) - b.astype(np.int64))
    assert ol.sad(a, b, 0) == int(d.sum())
    assert ol.sad(a, b, 1) == int(d[::2].sum()) << 1
    assert ol.sad(a, b, 2) == int(d[::4].sum()) << 2


def test_identity_at_zero_mv_models(oracle_2k):
    """Quirk Q3: ROT/TAN/3DT (and chroma GED) return the input untouched for mv == 0."""
    o = oracle_2k
    for model in (ol.ROT, ol.TAN, ol.T3D):
        for pos, size in [((0, 0), (16, 16)), ((2044 - 12, 1020 - 4), (16, 8)), ((512, 640), (64, 32))]:
            x, y = o.reproject(pos, size, (0, 0), model)
            rows, cols = size[1] // 4, size[0] // 4
            ex = np.repeat((pos[0] + 4 * np.arange(cols)) * 16, rows)
            ey = np.tile((pos[1] + 4 * np.arange(rows)) * 16, cols)
            assert np.array_equal(x, ex) and np.array_equal(y, ey)


def test_translation_near_equator_is_close_to_classic(oracle_2k):
    """Sanity of the geometry: near the equator, a small MV moves every model by about that MV."""
    o = oracle_2k
    for model in (ol.TAN, ol.T3D, ol.ROT, ol.MPA_LR, ol.GED_X):
        x, y = o.reproject((512, 508), (8, 8), (32, -16), model)
        assert np.abs(x / 16.0 - (512 + 4 * np.array([0, 0, 1, 1])) - 2.0).max() < 0.2
        assert np.abs(y / 16.0 - (508 + 4 * np.array([0, 1, 0, 1])) + 1.0).max() < 0.2


def test_golden_positions(oracle_2k):
    o = oracle_2k
    o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues([1, 0, 0]))
    g = np.load(os.path.join(GOLD, "positions_2048x1024.npz"))
    off = 0
    for (model, comp, fmt, px, py, pw, ph, mh, mv, n) in g["cases"]:
        ox, oy = o.reproject((int(px), int(py)), (int(pw), int(ph)), (int(mh), int(mv)), int(model), int(comp), int(fmt))
        assert np.array_equal(ox, g["out"][0, off:off + n]) and np.array_equal(oy, g["out"][1, off:off + n])
        off += n


def test_golden_distortion(oracle_2k):
    from vvc_extension_mm_b200 import workload as wl
    o = oracle_2k
    o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues([1, 0, 0]))
    g = np.load(os.path.join(GOLD, "distortion_2048x1024.npz"))
    ref = wl.pad_frame(wl.synth_frame(2048, 1024, seed=1), 288, 1)
    cur = wl.synth_frame(2048, 1024, seed=2, shift=(2.0, 1.0))
    class P: pass
    pl = P(); pl.origin = ref[0].origin; pl.stride = ref[0].stride
    for (model, x, y, w, h, mh, mv), (sad, had) in zip(g["cases"], g["dist"]):
        org = cur[0][y:y + h, x:x + w]
        m = np.array([[mh, mv]])
        assert o.me_block(pl, org, None, (int(x), int(y)), (int(w), int(h)), int(model), m, dist=ol.DIST_SAD)[0] == sad
        assert o.me_block(pl, org, None, (int(x), int(y)), (int(w), int(h)), int(model), m, dist=ol.DIST_HAD)[0] == had


def test_extend_wrap_border_closed_form():
    """Picture::extendWrapBorder restated with the reference's loop order == its closed form: the first xoffset margin
    columns continue the row from the opposite side, the others replicate; top/bottom repeat the extended rows."""
    rng = np.random.default_rng(5)
    w, h, mx, my = 40, 24, 16, 8
    p = rng.integers(0, 1024, size=(h, w)).astype(np.int16)
    for xoff in (0, 5, 16, 40):
        pl = ol.PaddedPlane(w, h, mx, my)
        pl.inner[:] = p
        pl.extend_wrap(xoff)
        cols = []
        for X in range(-mx, w + mx):
            if X < 0:
                cols.append(X + xoff if (-1 - X) < xoff else 0)
            elif X >= w:
                cols.append(X - xoff if (X - w) < xoff else w - 1)
            else:
                cols.append(X)
        rows = np.clip(np.arange(-my, h + my), 0, h - 1)
        assert np.array_equal(pl.buf[:, :w + 2 * mx], p[rows][:, cols]), xoff
    pl = ol.PaddedPlane(w, h, mx, my); pl.inner[:] = p; pl.extend_wrap(0)
    q = ol.PaddedPlane(w, h, mx, my); q.inner[:] = p; q.extend()
    assert np.array_equal(pl.buf, q.buf)


def test_extend_border_equals_edge_padding():
    from vvc_extension_mm_b200 import workload as wl
    rng = np.random.default_rng(4)
    p = rng.integers(0, 1024, size=(24, 40)).astype(np.int16)
    pl = ol.PaddedPlane(40, 24, 16, 8)
    pl.inner[:] = p
    pl.extend()
    assert np.array_equal(pl.buf[:, :72], np.pad(p, ((8, 8), (16, 16)), mode="edge"))
    assert np.array_equal(wl.Padded(p, 16, 8).buf[:, :72], pl.buf[:, :72])
