"""End to end against the reference's own applications (north_star bar: bit-exact bitstream, decoded MD5).

Fixtures under tests/golden/e2e/ were produced in the build container by the STOCK reference EncoderApp (whole reference
compiled unchanged against the Eigen stand-in, oracle/vtm.mk) on a synthetic ERP sequence with all five motion-model
families active, and by the MMGPU_DUMP mode of the patched app running the reference's CPU path (tests/golden/make_e2e_golden.py):

  CPU tests   the oracle reproduces what the reference computed IN SITU: 12 485 xPredInterBlkMM calls (all models, Y/Cb/Cr)
              and 20 546 motion-search candidates (SAD + Hadamard through the reference's SIMD distFunc; MM searches at integer, half and
              quarter pel and the classic fractional refinement) of a real encode.
  GPU tests   libmmgpu.so reproduces the same through the C ABI; the patched DecoderApp (GPU MC) decodes the stock bitstream
              with 0 MD5 errors and byte-identical YUV; the patched EncoderApp (GPU MC + GPU candidate evaluation, the
              reference's own search control flow) emits a byte-identical bitstream.
"""
import hashlib
import lzma
import os
import subprocess
import zlib

import numpy as np
import pytest

import mc_trace as mt
import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "e2e")
VTM = os.path.join(ROOT, "oracle", "_ref", "vtm")
E2E = os.path.join(ROOT, "oracle", "_ref", "e2e")
W, H, MARGIN = 256, 128, 288
ENC_ARGS = ["-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", "3", "-q", "37", "--InputBitDepth=8", "--InputChromaFormat=420",
            "--MPA=1", "--GED=1", "--ROT=1", "--TAN=1", "--3DT=1", "--MMMVP=1", "--Projection=2", "--Epipole=-1,-1,1.0,0.0,0.0",
            "--SEIDecodedPictureHash=1"]


@pytest.fixture(scope="module")
def trace():
    z = np.load(os.path.join(GOLD, "trace_ldp_q37.npz"))
    refs = {}
    for k in z.files:
        if k.startswith("ref_"):
            _, poc, wrap, c = k.split("_")
            refs.setdefault((int(poc), int(wrap)), {})[int(c)] = z[k]
    refs = {k: [v[c] for c in sorted(v)] for k, v in refs.items()}
    return z, refs


MC_KEYS = ("comp", "fmt", "x", "y", "w", "h", "ref_poc", "wrap", "mv_hor", "mv_ver", "model", "bi", "clp_min", "clp_max", "cur_poc")


def test_oracle_equals_reference_encoder_mc_in_situ(trace):
    z, refs = trace
    o = ol.Oracle(W, H)
    o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues([float(v) for v in z["epipole"]]))
    padded = {k: mt.pad_planes(v, MARGIN, 1) for k, v in refs.items()}
    models = set()
    for row in z["mc"]:
        m = dict(zip(MC_KEYS, [int(v) for v in row[:15]]))
        got = o.mc_block(padded[(m["ref_poc"], m["wrap"])][m["comp"]], m["comp"], m["fmt"], (m["x"], m["y"]), (m["w"], m["h"]),
                         (m["mv_hor"], m["mv_ver"]), m["model"], bi=bool(m["bi"]), clp=(m["clp_min"], m["clp_max"]))
        assert zlib.crc32(got.tobytes()) == int(row[15]), m
        models.add(m["model"])
    assert models == {1, 2, 3, 4, 5, 6, 10} and len(z["mc"]) > 12000


def _searches(z):
    hdr, off, cand, dist = z["me_hdr"], z["me_org_off"], z["me_cand"], z["me_dist"]
    order = np.argsort(cand[:, 0], kind="stable")
    bounds = np.searchsorted(cand[order, 0], np.arange(len(hdr) + 1))
    for s in range(len(hdr)):
        h = dict(zip(("x", "y", "w", "h", "ref_poc", "wrap", "model", "cur_poc", "had", "sub_shift", "clp_min", "clp_max"),
                     [int(v) for v in hdr[s]]))
        org = z["me_org"][off[s]:off[s + 1]].reshape(h["h"], h["w"])
        idx = order[bounds[s]:bounds[s + 1]]
        yield h, org, cand[idx, 1:4], dist[idx]


def test_oracle_equals_reference_encoder_me_in_situ(trace):
    z, refs = trace
    o = ol.Oracle(W, H)
    o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues([float(v) for v in z["epipole"]]))
    padded = {k: mt.pad_planes(v, MARGIN, 1) for k, v in refs.items()}
    n = had = classic = 0
    for h, org, cand, dist in _searches(z):
        for shift in np.unique(cand[:, 2]):
            sel = cand[:, 2] == shift
            got = o.me_block(padded[(h["ref_poc"], h["wrap"])][0], org, None, (h["x"], h["y"]), (h["w"], h["h"]), h["model"],
                             cand[sel, :2], mv_prec_shift=int(shift), dist=ol.DIST_HAD if h["had"] else ol.DIST_SAD,
                             sub_shift=h["sub_shift"], had_mode=ol.HAD_SIMD, clp=(h["clp_min"], h["clp_max"]))
            assert np.array_equal(got, dist[sel]), h
            n += int(sel.sum())
            had += int(sel.sum()) if h["had"] else 0
            classic += int(sel.sum()) if h["model"] == 0 else 0
    assert n > 15000 and had > 1000 and classic > 100


# ---- GPU ----------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu_small(trace):
    import vvc_extension_mm_b200 as mm
    z, refs = trace
    g = mm.MmGpu(W, H, max_refs=4)
    g.set_rotation(0, mm.rodrigues([float(v) for v in z["epipole"]]))
    slots, keep = {}, []
    for i, (k, planes) in enumerate(sorted(refs.items())):
        pp = mt.pad_planes(planes, MARGIN, 1)
        g.set_ref(i, k[0], [p.origin for p in pp], [p.stride for p in pp])
        slots[k] = i
        keep.append(pp)
    yield g, slots
    g.close()


@pytest.mark.gpu
def test_gpu_equals_reference_encoder_mc_in_situ(trace, gpu_small):
    z, _ = trace
    g, slots = gpu_small
    for row in z["mc"]:
        m = dict(zip(MC_KEYS, [int(v) for v in row[:15]]))
        got = g.mc(slots[(m["ref_poc"], m["wrap"])], m["comp"], m["fmt"], (m["x"], m["y"]), (m["w"], m["h"]),
                   (m["mv_hor"], m["mv_ver"]), m["model"], bi=bool(m["bi"]), clp=(m["clp_min"], m["clp_max"]))
        assert zlib.crc32(np.ascontiguousarray(got).tobytes()) == int(row[15]), m


@pytest.mark.gpu
def test_gpu_equals_reference_encoder_me_in_situ(trace, gpu_small):
    import vvc_extension_mm_b200 as mm
    z, _ = trace
    g, slots = gpu_small
    n = 0
    for h, org, cand, dist in _searches(z):
        g.set_org(np.ascontiguousarray(org))
        for shift in np.unique(cand[:, 2]):
            sel = cand[:, 2] == shift
            blocks = np.zeros(1, dtype=mm.api.BLOCK_DTYPE)
            blocks[0] = (h["x"], h["y"], h["w"], h["h"], h["model"], slots[(h["ref_poc"], h["wrap"])], 0, 0, int(sel.sum()), 0, 0, 0)
            mvs = np.zeros(int(sel.sum()), dtype=mm.api.MV_DTYPE)
            mvs["hor"], mvs["ver"] = cand[sel, 0], cand[sel, 1]
            got = g.me_batch(blocks, mvs, mv_prec_shift=int(shift), dist=mm.DIST_HAD if h["had"] else mm.DIST_SAD,
                             sub_shift=h["sub_shift"], clp=(h["clp_min"], h["clp_max"]))
            assert np.array_equal(got, dist[sel]), h
            n += int(sel.sum())
    assert n > 15000


def _apps():
    need = [os.path.join(VTM, "DecoderApp_gpu"), os.path.join(VTM, "EncoderApp_gpu"), os.path.join(E2E, "encoder_lowdelay_P_vtm.cfg")]
    missing = [p for p in need if not os.path.exists(p)]
    if missing:
        pytest.fail("reference apps not built (python __graft_entry__.py builds them where /root/reference exists): %s" % missing)
    return need


def _yuv(tmp):
    p = os.path.join(tmp, "erp.yuv")
    with lzma.open(os.path.join(GOLD, "erp_256x128_3f.yuv.xz")) as f, open(p, "wb") as g:
        g.write(f.read())
    return p


@pytest.mark.gpu
def test_decoder_app_with_gpu_mc_md5_and_yuv(tmp_path):
    dec, _, _ = _apps()
    out = str(tmp_path / "dec.yuv")
    r = subprocess.run([dec, "-b", os.path.join(GOLD, "mm_ldp_q37.bin"), "-o", out], capture_output=True, text=True, timeout=900)
    log = r.stdout + r.stderr
    assert r.returncode == 0, log[-2000:]
    assert "[mmgpu] context" in log and "MC calls" in log, "GPU path not active"
    assert log.count("(OK)") == 3 and "***ERROR***" not in log, log[-2000:]      # DecLib.cpp:859-883
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == open(os.path.join(GOLD, "mm_ldp_q37.md5")).read().strip()


@pytest.mark.gpu
def test_encoder_app_with_gpu_me_bitstream_identical(tmp_path):
    _, enc, cfg = _apps()
    yuv = _yuv(str(tmp_path))
    bit = str(tmp_path / "enc.bin")
    r = subprocess.run([enc, "-c", cfg, "-i", yuv, "-b", bit] + ENC_ARGS, capture_output=True, text=True, timeout=1500)
    log = r.stdout + r.stderr
    assert r.returncode == 0, log[-2000:]
    assert "ME batches" in log, "GPU path not active"
    assert open(bit, "rb").read() == open(os.path.join(GOLD, "mm_ldp_q37.bin"), "rb").read(), log[-3000:]


# ---- random access (BASELINE config 1 shape: encoder_randomaccess_vtm.cfg -- B slices, bi-prediction with 14-bit
#      intermediates + addAvg, DMVR / BDOF / affine / AMVR on the classic side, MM-MVP) ---------------------------------
RA_ARGS = [a if a != "3" else "5" for a in ENC_ARGS]     # -f 5


@pytest.mark.gpu
def test_decoder_app_random_access_stream(tmp_path):
    dec, _, _ = _apps()
    out = str(tmp_path / "dec.yuv")
    r = subprocess.run([dec, "-b", os.path.join(GOLD, "mm_ra_q37.bin"), "-o", out], capture_output=True, text=True, timeout=900)
    log = r.stdout + r.stderr
    assert r.returncode == 0, log[-2000:]
    assert "[mmgpu] context" in log, "GPU path not active"
    assert log.count("(OK)") == 5 and "***ERROR***" not in log, log[-2000:]
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == open(os.path.join(GOLD, "mm_ra_q37.md5")).read().strip()


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("MMGPU_E2E_RA_ENCODE", "0") != "1",
                    reason="several minutes of encoder time: MMGPU_E2E_RA_ENCODE=1 (run once per round, log under profiles/)")
def test_encoder_app_random_access_bitstream_identical(tmp_path):
    _, enc, _ = _apps()
    cfg = os.path.join(E2E, "encoder_randomaccess_vtm.cfg")
    yuv = os.path.join(str(tmp_path), "erp.yuv")
    with lzma.open(os.path.join(GOLD, "erp_256x128_5f_seed2.yuv.xz")) as f, open(yuv, "wb") as g:
        g.write(f.read())
    bit = str(tmp_path / "enc.bin")
    r = subprocess.run([enc, "-c", cfg, "-i", yuv, "-b", bit] + RA_ARGS, capture_output=True, text=True, timeout=3000)
    log = r.stdout + r.stderr
    assert r.returncode == 0, log[-2000:]
    print(log[-1500:])
    assert open(bit, "rb").read() == open(os.path.join(GOLD, "mm_ra_q37.bin"), "rb").read(), log[-3000:]


@pytest.mark.gpu
def test_decoder_app_random_access_verify_mode(tmp_path):
    """MMGPU_VERIFY=1: classic DMVR (integer search costs and final padded MC) runs on the device AND through the reference's own
    code in the same process; the app stops at the first differing cost or sample.  The stats line says how many were compared."""
    dec, _, _ = _apps()
    out = str(tmp_path / "dec.yuv")
    env = dict(os.environ, MMGPU_VERIFY="1")
    r = subprocess.run([dec, "-b", os.path.join(GOLD, "mm_ra_q37.bin"), "-o", out], capture_output=True, text=True, timeout=900, env=env)
    log = r.stdout + r.stderr
    assert r.returncode == 0, log[-2000:]
    line = [l for l in log.splitlines() if "classic DMVR" in l]
    assert line, log[-2000:]
    print(line[0])
    assert log.count("(OK)") == 5 and "***ERROR***" not in log, log[-2000:]
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == open(os.path.join(GOLD, "mm_ra_q37.md5")).read().strip()
    n_sub = int(line[0].split("classic DMVR:")[1].split()[0])
    if n_sub:
        assert "VERIFY:" in line[0] and int(line[0].split("VERIFY:")[1].split()[0]) > 0


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("MMGPU_E2E_VERIFY_ENCODE", "0") != "1",
                    reason="a CPU-speed encode: MMGPU_E2E_VERIFY_ENCODE=1 (run once per round, log under profiles/)")
def test_encoder_app_verify_mode_every_hooked_call_against_the_reference(tmp_path):
    """MMGPU_VERIFY=1 on a whole low-delay encode: every hooked call (xPredInterBlkMM, classic xPredInterBlk incl. BDOF ring and
    alternative half-pel filter, every motion-search candidate's distortion, every MM-MVP conversion, classic DMVR) is computed on
    the device AND by the reference's own code in the same process and compared value by value; the first difference aborts."""
    _, enc, cfg = _apps()
    yuv = _yuv(str(tmp_path))
    bit = str(tmp_path / "enc.bin")
    env = dict(os.environ, MMGPU_VERIFY="1")
    r = subprocess.run([enc, "-c", cfg, "-i", yuv, "-b", bit] + ENC_ARGS, capture_output=True, text=True, timeout=3000, env=env)
    log = r.stdout + r.stderr
    assert r.returncode == 0, log[-3000:]
    lines = [l for l in log.splitlines() if l.startswith("[mmgpu]")]
    print("\n".join(lines))
    ver = [l for l in lines if "VERIFY (device vs" in l]
    assert ver, log[-2000:]
    n_mc = int(ver[0].split("all equal):")[1].split()[0])
    n_dist = int(ver[0].split("samples,")[1].split()[0])
    assert n_mc > 10000 and n_dist > 100000, ver[0]
    assert open(bit, "rb").read() == open(os.path.join(GOLD, "mm_ldp_q37.bin"), "rb").read(), log[-3000:]
