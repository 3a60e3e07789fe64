"""Generates tests/golden/e2e/ -- the fixtures of the end-to-end tests (run in THIS container, where /root/reference and
the apps built by `make -C oracle -f vtm.mk stock gpu` exist; the GPU box only reads the committed fixtures):

  erp_256x128_3f.yuv.xz     synthetic 8-bit 4:2:0 ERP sequence (tests/golden/make_sequence.py, seed 1)
  mm_ldp_q37.bin            bitstream of the STOCK reference EncoderApp (CPU), low-delay P, all five motion-model families
  mm_ldp_q37.md5            MD5 of the stock encoder's reconstruction (= what DecoderApp must output)
  trace_ldp_q37.npz         what the reference's CPU path computed during that encode: every 8th xPredInterBlkMM call
                            (parameters + CRC32 of the predicted block) and every 65th motion search (MM integer / half / quarter-pel searches and the classic fractional refinement) (original block,
                            every candidate MV with the distortion m_cDistParam.distFunc returned), plus the two
                            reference pictures -- written by the MMGPU_DUMP mode of the patched app running its CPU path.

usage: python tests/golden/make_e2e_golden.py        (about 8 minutes)"""
import hashlib
import lzma
import os
import subprocess
import sys
import zlib
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mc_trace as mt  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "e2e")
VTM = os.path.join(ROOT, "oracle", "_ref", "vtm")
CFG = "/root/reference/cfg/encoder_lowdelay_P_vtm.cfg"
W, H, FRAMES, QP = 256, 128, 3, 37
ENC_ARGS = ["-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", str(FRAMES), "-q", str(QP), "--InputBitDepth=8",
            "--InputChromaFormat=420", "--MPA=1", "--GED=1", "--ROT=1", "--TAN=1", "--3DT=1", "--MMMVP=1", "--Projection=2",
            "--Epipole=-1,-1,1.0,0.0,0.0", "--SEIDecodedPictureHash=1"]


def main():
    os.makedirs(OUT, exist_ok=True)
    tmp = "/tmp/e2e_golden"
    os.makedirs(tmp, exist_ok=True)
    yuv = os.path.join(tmp, "erp.yuv")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "golden", "make_sequence.py"), str(W), str(H), str(FRAMES), yuv, "1"])
    with open(yuv, "rb") as f, lzma.open(os.path.join(OUT, "erp_256x128_3f.yuv.xz"), "wb", preset=9) as g:
        g.write(f.read())
    bit = os.path.join(OUT, "mm_ldp_q37.bin")
    rec = os.path.join(tmp, "rec.yuv")
    subprocess.check_call([os.path.join(VTM, "EncoderApp"), "-c", CFG, "-i", yuv, "-b", bit, "-o", rec] + ENC_ARGS,
                          stdout=open(os.path.join(tmp, "enc.log"), "w"))
    with open(os.path.join(OUT, "mm_ldp_q37.md5"), "w") as f:
        f.write(hashlib.md5(open(rec, "rb").read()).hexdigest() + "\n")
    trace = os.path.join(tmp, "trace.bin")
    env = dict(os.environ, MMGPU="0", MMGPU_DUMP=trace, MMGPU_DUMP_ME="5")
    bit2 = os.path.join(tmp, "dump.bin")
    subprocess.check_call([os.path.join(VTM, "EncoderApp_gpu"), "-c", CFG, "-i", yuv, "-b", bit2] + ENC_ARGS, env=env,
                          stdout=open(os.path.join(tmp, "dump.log"), "w"))
    assert open(bit, "rb").read() == open(bit2, "rb").read(), "patched app (CPU path) and stock app disagree"
    compact(trace, os.path.join(OUT, "trace_ldp_q37.npz"))


def compact(trace, out, mc_every=8, me_every=13):
    refs = {}
    mc = []
    i = 0
    for rec in mt.read_trace(trace):
        if rec[0] == "R":
            refs[(rec[1], rec[2])] = rec[3]
        elif rec[0] == "M":
            if i % mc_every == 0:
                m = rec[1]
                mc.append([m[k] for k in ("comp", "fmt", "x", "y", "w", "h", "ref_poc", "wrap", "mv_hor", "mv_ver", "model", "bi",
                                          "clp_min", "clp_max", "cur_poc")] + [zlib.crc32(np.ascontiguousarray(rec[2]).tobytes())])
            i += 1
    hdrs, orgs, org_off, cands = [], [], [0], []
    for j, (hdr, org, cl, _) in enumerate(mt.searches(trace)):
        if j % me_every:
            continue
        s = len(hdrs)
        hdrs.append([hdr[k] for k in ("x", "y", "w", "h", "ref_poc", "wrap", "model", "cur_poc", "had", "sub_shift", "clp_min", "clp_max")])
        orgs.append(org.ravel())
        org_off.append(org_off[-1] + org.size)
        cands += [(s, a, b, c, d) for a, b, c, d in cl]
    arrays = {"mc": np.array(mc, dtype=np.int64), "me_hdr": np.array(hdrs, dtype=np.int32), "me_org": np.concatenate(orgs),
              "me_org_off": np.array(org_off, dtype=np.int64), "me_cand": np.array([c[:4] for c in cands], dtype=np.int32),
              "me_dist": np.array([c[4] for c in cands], dtype=np.uint64), "epipole": np.array([1.0, 0.0, 0.0], dtype=np.float32)}
    for (poc, wrap), planes in refs.items():
        for c, p in enumerate(planes):
            arrays["ref_%d_%d_%d" % (poc, wrap, c)] = p
    np.savez_compressed(out, **arrays)
    print("golden: %d MC calls, %d searches, %d candidates, %d reference pictures -> %s (%d bytes)"
          % (len(mc), len(hdrs), len(cands), len(refs), out, os.path.getsize(out)))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "compact":
        compact(sys.argv[2], os.path.join(OUT, "trace_ldp_q37.npz"))
    else:
        main()
