"""Final-MC shape of the batch path (mmgpu_mc_batch_device: one MV per block, predictions written to a device picture)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vvc_extension_mm_b200 as mm
from vvc_extension_mm_b200 import workload as wl
W, H = 2048, 1024
g = mm.MmGpu(W, H, max_refs=8)
g.set_rotation(0, mm.rodrigues([1, 0, 0]))
for f in range(8):
    ref = wl.pad_frame(wl.synth_frame(W, H, 100 + f, shift=(0.37 * f, 0.11 * f)), 288, 1)
    g.set_ref(f, f, [p.origin for p in ref], [p.stride for p in ref])
blocks, mvs, prec = wl.frame_batch(W, H, seed=1, cands_per_block=1, mv_range=64, models=(1, 2, 3, 10, 6))
plans = []
for f in range(8):
    b = blocks.copy(); b["ref_slot"] = f
    plans.append(g.plan(b, len(mvs)))
dm = torch.from_numpy(mvs.view(np.int32).reshape(-1, 2).copy()).cuda()
dst = torch.zeros((H, W), dtype=torch.int16, device="cuda")
stream = torch.cuda.ExternalStream(g.stream)
def run(n):
    for i in range(n):
        g.mc_batch_device(plans[i % 8], dm.data_ptr(), dst.data_ptr(), W, mv_prec_shift=prec)
run(10); g.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); run(200); e1.record(stream); g.sync()
ms = e0.elapsed_time(e1) / 200
print("mc_batch_device: %d blocks, whole %dx%d luma picture: %.4f ms = %.1f G samples/s, %.0f GB/s algorithmic (4 B/sample)"
      % (len(blocks), W, H, ms, W * H / ms / 1e6, 4.0 * W * H / ms / 1e6))
