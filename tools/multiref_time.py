"""Multi-reference batch: one k_map + R x k_fir (mmgpu_me_batch_device_multiref) vs R plain calls.  python tools/multiref_time.py [R]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vvc_extension_mm_b200 as mm
from vvc_extension_mm_b200 import workload as wl
W, H = 2048, 1024
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = mm.MmGpu(W, H, max_refs=R)
g.set_rotation(0, mm.rodrigues([1, 0, 0]))
for f in range(R):
    ref = wl.pad_frame(wl.synth_frame(W, H, 100 + f, shift=(0.37 * f, 0.11 * f)), 288, 1)
    g.set_ref(f, f, [p.origin for p in ref], [p.stride for p in ref])
blocks, mvs, prec = wl.frame_batch(W, H, seed=1, cands_per_block=16, mv_range=64, models=(1, 2, 3, 10, 6))
g.set_org(wl.synth_plane(W, H, 7))
plans = []
for f in range(R):
    b = blocks.copy(); b["ref_slot"] = f
    plans.append(g.plan(b, len(mvs)))
dm = torch.from_numpy(mvs.view(np.int32).reshape(-1, 2).copy()).cuda()
dd = torch.zeros(R * len(mvs), dtype=torch.int64, device="cuda")
stream = torch.cuda.ExternalStream(g.stream)
def plain():
    for f in range(R):
        g.me_batch_device(plans[f], dm.data_ptr(), dd.data_ptr() + 8 * f * len(mvs), mv_prec_shift=prec, dist=mm.DIST_HAD)
def multi():
    g.me_batch_device_multiref(plans[0], dm.data_ptr(), dd.data_ptr(), list(range(R)), [0] * R, mv_prec_shift=prec, dist=mm.DIST_HAD)
out = {}
for name, fn in (("plain", plain), ("multiref", multi)):
    for _ in range(5): fn()
    g.sync()
    chk = dd.cpu().numpy().copy()
    g.kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(50): fn()
    e1.record(stream); g.sync()
    km, kf, n = g.kernel_timing_read(); g.kernel_timing(False)
    out[name] = (e0.elapsed_time(e1) / 50, km / 50, kf / 50, chk)
    print("%-9s %d references: %.4f ms per %d-reference search step (k_map %.4f ms, k_fir %.4f ms)" % (name, R, out[name][0], R, km / 50, kf / 50))
print("identical results:", np.array_equal(out["plain"][3], out["multiref"][3]))
