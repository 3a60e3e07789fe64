"""Device-resident step time of a frame batch: python tools/step_time.py [W H [model,model,...]]
(default: the bench batch, 2048x1024, MPA_FB/LR/TB + GED_CAMPOSE + ROT; config 3 = 4096 2048 1,4,5,6,10)."""
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vvc_extension_mm_b200 as mm
from vvc_extension_mm_b200 import workload as wl
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 1024)
MODELS = tuple(int(m) for m in sys.argv[3].split(',')) if len(sys.argv) > 3 else (1, 2, 3, 10, 6)
g = mm.MmGpu(W, H, max_refs=8)
g.set_rotation(0, mm.rodrigues([1, 0, 0]))
for f in range(8):
    ref = wl.pad_frame(wl.synth_frame(W, H, 100 + f, shift=(0.37 * f, 0.11 * f)), 288, 1)
    g.set_ref(f, f, [p.origin for p in ref], [p.stride for p in ref])
blocks, mvs, prec = wl.frame_batch(W, H, seed=1, cands_per_block=16, mv_range=64, models=MODELS)
g.set_org(wl.synth_plane(W, H, 7))
plans = []
for f in range(8):
    b = blocks.copy(); b["ref_slot"] = f
    plans.append(g.plan(b, len(mvs)))
dm = torch.from_numpy(mvs.view(np.int32).reshape(-1, 2).copy()).cuda()
dd = torch.zeros(len(mvs), dtype=torch.int64, device="cuda")
stream = torch.cuda.ExternalStream(g.stream)
def run(n):
    for i in range(n):
        g.me_batch_device(plans[i % 8], dm.data_ptr(), dd.data_ptr(), mv_prec_shift=prec, dist=mm.DIST_HAD)
run(10); g.sync()
ref = dd.cpu().numpy().copy()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); run(200); e1.record(stream); g.sync()
ms = e0.elapsed_time(e1) / 200
samples = float((blocks["w"].astype(np.int64) * blocks["h"] * blocks["cand_count"]).sum())
print("%dx%d models %s: %d blocks, %d candidates, step %.4f ms = %.1f G sample-candidates/s  checksum %d"
      % (W, H, MODELS, len(blocks), len(mvs), ms, samples / ms / 1e6, int(ref.sum())))
