"""BASELINE config 5 as a sweep: device-resident MC + distortion step of a whole 2048x1024 frame tiled with ONE block size,
for block sizes 8x4 .. 128x128 x {1,2,4,9,16,25,64,81} candidates/block x {SAD, HAD}, motion models cycled over
MPA_FB/LR/TB, GED_CAMPOSE, ROT; plus the 4096x2048 five-model batch (config 3).  Writes a CSV (default
profiles/sweep_r02.csv): step ms, G sample-candidates/s, algorithmic GB/s and fraction of the measured HBM peak."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vvc_extension_mm_b200 as mm
from vvc_extension_mm_b200 import workload as wl

OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sweep_r02.csv")
PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
SIZES = [(8, 4), (4, 8), (8, 8), (16, 8), (8, 16), (16, 16), (32, 16), (32, 32), (64, 32), (64, 64), (128, 64), (128, 128)]
CANDS = [1, 2, 4, 9, 16, 25, 64, 81]


def run(g, W, H, blocks, mvs, prec, dist, iters):
    plan = g.plan(blocks, len(mvs))
    dm = torch.from_numpy(mvs.view(np.int32).reshape(-1, 2).copy()).cuda()
    dd = torch.zeros(len(mvs), dtype=torch.int64, device="cuda")
    stream = torch.cuda.ExternalStream(g.stream)
    for _ in range(3):
        g.me_batch_device(plan, dm.data_ptr(), dd.data_ptr(), mv_prec_shift=prec, dist=dist)
    g.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        g.me_batch_device(plan, dm.data_ptr(), dd.data_ptr(), mv_prec_shift=prec, dist=dist)
    e1.record(stream)
    g.sync()
    ms = e0.elapsed_time(e1) / iters
    wh = blocks["w"].astype(np.int64) * blocks["h"]
    samples = float((wh * blocks["cand_count"]).sum())
    alg = float((2 * wh + blocks["cand_count"] * (2 * wh + 8)).sum())
    del plan, dm, dd
    return ms, samples, alg


def main():
    rows = ["picture,block,candidates_per_block,dist,blocks,step_ms,G_sample_candidates_per_s,algorithmic_GBps,frac_of_measured_hbm_peak"]
    W, H = 2048, 1024
    g = mm.MmGpu(W, H, max_refs=2)
    g.set_rotation(0, mm.rodrigues([1, 0, 0]))
    ref = wl.pad_frame(wl.synth_frame(W, H, 100), 288, 1)
    g.set_ref(0, 0, [p.origin for p in ref], [p.stride for p in ref])
    g.set_org(wl.synth_plane(W, H, 7))
    for (bw, bh) in SIZES:
        for nc in CANDS:
            blocks, mvs, prec = wl.frame_batch(W, H, seed=1, cands_per_block=nc, mv_range=64, models=(1, 2, 3, 10, 6), fixed=(bw, bh))
            for name, dist in (("SAD", mm.DIST_SAD), ("HAD", mm.DIST_HAD)):
                ms, samples, alg = run(g, W, H, blocks, mvs, prec, dist, 10 if nc >= 25 else 30)
                rows.append("%dx%d,%dx%d,%d,%s,%d,%.4f,%.1f,%.1f,%.4f" % (W, H, bw, bh, nc, name, len(blocks), ms, samples / ms / 1e6,
                                                                          alg / ms / 1e6, alg / ms / 1e6 / PEAK))
                print(rows[-1], flush=True)
    g.close()
    W, H = 4096, 2048
    g = mm.MmGpu(W, H, max_refs=2)
    g.set_rotation(0, mm.rodrigues([1, 0, 0]))
    ref = wl.pad_frame(wl.synth_frame(W, H, 100), 288, 1)
    g.set_ref(0, 0, [p.origin for p in ref], [p.stride for p in ref])
    g.set_org(wl.synth_plane(W, H, 7))
    blocks, mvs, prec = wl.frame_batch(W, H, seed=1, cands_per_block=16, mv_range=64, models=(1, 4, 5, 6, 10))
    for name, dist in (("SAD", mm.DIST_SAD), ("HAD", mm.DIST_HAD)):
        ms, samples, alg = run(g, W, H, blocks, mvs, prec, dist, 30)
        rows.append("%dx%d,mixed 8x4..128x128 (MPA/TAN/3DT/ROT/GED),%d,%s,%d,%.4f,%.1f,%.1f,%.4f" % (W, H, 16, name, len(blocks), ms,
                    samples / ms / 1e6, alg / ms / 1e6, alg / ms / 1e6 / PEAK))
        print(rows[-1], flush=True)
    open(OUT, "w").write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
