import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vvc_extension_mm_b200 as mm
from vvc_extension_mm_b200 import workload as wl
W,H=2048,1024
g=mm.MmGpu(W,H,max_refs=2)
g.set_rotation(0, mm.rodrigues([1,0,0]))
ref=wl.pad_frame(wl.synth_frame(W,H,1),288,1)
g.set_ref(0,0,[p.origin for p in ref],[p.stride for p in ref])
blocks,mvs,prec=wl.frame_batch(W,H,seed=1,cands_per_block=16,models=(1,2,3,10,6))
org=torch.from_numpy(wl.synth_plane(W,H,7)).pin_memory().numpy()
hb=torch.from_numpy(blocks.view(np.int32).reshape(-1,12).copy()).pin_memory().numpy().view(mm.api.BLOCK_DTYPE).reshape(-1)
hm=torch.from_numpy(mvs.view(np.int32).reshape(-1,2).copy()).pin_memory().numpy().view(mm.api.MV_DTYPE).reshape(-1)
out=torch.zeros(len(mvs),dtype=torch.int64).pin_memory().numpy().view(np.uint64)
for i in range(5): g.set_org(org); g.me_batch(hb,hm,mv_prec_shift=prec,dist=mm.DIST_HAD,out=out)
def t(f,n=50):
    g.sync(); t0=time.perf_counter()
    for i in range(n): f()
    g.sync(); return (time.perf_counter()-t0)/n*1e3
print('set_org+sync ms', t(lambda:(g.set_org(org), g.sync())))
print('me_batch ms', t(lambda: g.me_batch(hb,hm,mv_prec_shift=prec,dist=mm.DIST_HAD,out=out)))
print('both ms', t(lambda:(g.set_org(org), g.me_batch(hb,hm,mv_prec_shift=prec,dist=mm.DIST_HAD,out=out))))
plan=g.plan(blocks,len(mvs))
dm=torch.from_numpy(mvs.view(np.int32).reshape(-1,2).copy()).cuda(); dd=torch.zeros(len(mvs),dtype=torch.int64,device='cuda')
print('device only ms', t(lambda: g.me_batch_device(plan,dm.data_ptr(),dd.data_ptr(),mv_prec_shift=prec,dist=mm.DIST_HAD)))
print('plan create ms', t(lambda: g.plan(blocks,len(mvs)).close(), 20))
