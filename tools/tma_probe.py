"""Debug probe: one block, one candidate through k_fir; argv: model w h x y mvx mvy dist"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vvc_extension_mm_b200 as mm
from vvc_extension_mm_b200 import workload as wl
model, w, h, x, y, mvx, mvy, dist = [int(a) for a in sys.argv[1:9]]
W, H = 2048, 1024
g = mm.MmGpu(W, H, max_refs=2)
ref = wl.pad_frame(wl.synth_frame(W, H, seed=1), 288, 1)
g.set_ref(0, 0, [p.origin for p in ref], [p.stride for p in ref])
g.set_org(wl.synth_plane(W, H, 7))
blocks = np.zeros(1, dtype=mm.api.BLOCK_DTYPE)
blocks["x"], blocks["y"], blocks["w"], blocks["h"], blocks["model"] = x, y, w, h, model
blocks["org_x"], blocks["org_y"], blocks["cand_count"] = x, y, 1
mvs = np.zeros(1, dtype=mm.api.MV_DTYPE); mvs["hor"], mvs["ver"] = mvx, mvy
try:
    print(sys.argv[1:], "->", g.me_batch(blocks, mvs, mv_prec_shift=4, dist=dist))
except Exception as e:
    print(sys.argv[1:], "FAILED", str(e)[-60:])
