"""Isolated timing of the fused exchange (mmgpu_pull_refs_unpadded) between two ranks: every rank pulls R un-padded 2048x1024 4:2:0
pictures from its peer's symmetric memory (nothing else running), per SM budget.
torchrun --nproc-per-node 2 tools/pull_time.py [R]"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vvc_extension_mm_b200 as mm
from vvc_extension_mm_b200 import workload as wl

W, H = 2048, 1024
R = int(sys.argv[1]) if len(sys.argv) > 1 else 7
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm
dev = torch.device("cuda", local)
geo = wl.pad_frame(wl.synth_frame(W, H, seed=100), 288, 1)
uw = [(p.w, p.h) for p in geo]
uoffs = np.concatenate([[0], np.cumsum([(2 * w_ * h_ + 255) // 256 * 256 for w_, h_ in uw])]).astype(np.int64)
n = int(uoffs[-1])
stage = symm.empty(n * R, dtype=torch.uint8, device=dev)
hdl = symm.rendezvous(stage, dist.group.WORLD)
peer = hdl.get_buffer(1 - rank, (n * R,), torch.uint8)
stage.random_(0, 4)
gpu = mm.MmGpu(W, H, max_refs=R, device=local)
st = torch.cuda.Stream(device=dev)
for src, name in ((peer, "peer (NVLink)"), (stage, "local")):
    for ctas in (4, 8, 16, 32, 64):
        args = (list(range(R)), list(range(R)), [[src.data_ptr() + r * n + int(uoffs[c]) for c in range(3)] for r in range(R)], [[w_ for w_, _ in uw]] * R)
        with torch.cuda.stream(st):
            gpu.pull_refs_unpadded(*args, max_ctas=ctas, stream=st.cuda_stream, keep=stage)
            hdl.barrier(channel=0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(10):
                gpu.pull_refs_unpadded(*args, max_ctas=ctas, stream=st.cuda_stream, keep=stage)
            e1.record(st)
            hdl.barrier(channel=0)
        st.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if rank == 0:
            print("%-14s %3d CTAs: %.4f ms per %d pictures = %.0f GB/s in, %.0f GB/s out" % (name, ctas, ms, R, R * n / ms / 1e6, R * 12.6e6 / ms / 1e6), flush=True)
dist.destroy_process_group()
