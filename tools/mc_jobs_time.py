"""Decoder-shape timing (BASELINE config 4): whole-picture reconstruction through mmgpu_mc_jobs (host planes in and out)."""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vvc_extension_mm_b200 as mm
from vvc_extension_mm_b200 import workload as wl
W, H = 2048, 1024
g = mm.MmGpu(W, H, max_refs=4)
g.set_rotation(0, mm.rodrigues([0.3, -0.5, 0.81])); g.set_rotation(2, mm.rodrigues([0.1, 0.2, -0.97]))
for slot, seed in ((0, 11), (3, 41)):
    ref = wl.pad_frame(wl.synth_frame(W, H, seed=seed), 288, 1)
    g.set_ref(slot, slot, [p.origin for p in ref], [p.stride for p in ref])
rng = np.random.default_rng(2026)
jobs = []
for (x, y, w, h) in wl.frame_partition(W, H, seed=77):
    model = int(rng.integers(0, 11)); bi = bool(rng.random() < 0.3) and (w * h > 32)
    mv0 = (int(rng.integers(-700, 700)), int(rng.integers(-400, 400))); mv1 = (int(rng.integers(-700, 700)), int(rng.integers(-400, 400)))
    for comp in range(3):
        cx, cy, cw, ch = (x, y, w, h) if comp == 0 else (x // 2, y // 2, w // 2, h // 2)
        jobs.append((comp, cx, cy, cw, ch, model, int(bi), 0, 0, mv0[0], mv0[1], 3, 2, mv1[0], mv1[1]))
jobs = np.array(jobs, dtype=mm.JOB_DTYPE)
got = [np.zeros((H, W), np.int16), np.zeros((H // 2, W // 2), np.int16), np.zeros((H // 2, W // 2), np.int16)]
for i in range(3): g.mc_jobs(jobs, got)
t0 = time.perf_counter()
for i in range(20): g.mc_jobs(jobs, got)
dt = (time.perf_counter() - t0) / 20
print("mc_jobs: %d jobs, %.3f ms per 2048x1024 4:2:0 picture (host planes in+out) = %.1f pictures/s" % (len(jobs), dt * 1e3, 1 / dt))
import torch
# the same call with the caller's planes in PINNED host memory (cudaHostRegister / pinned allocation): the read-back then is one
# direct copy-engine transfer per plane instead of the driver's staged pageable copy
pinned = [torch.zeros((H, W), dtype=torch.int16).pin_memory(), torch.zeros((H // 2, W // 2), dtype=torch.int16).pin_memory(),
          torch.zeros((H // 2, W // 2), dtype=torch.int16).pin_memory()]
gotp = [t.numpy() for t in pinned]
for i in range(3): g.mc_jobs(jobs, gotp)
t0 = time.perf_counter()
for i in range(20): g.mc_jobs(jobs, gotp)
dtp = (time.perf_counter() - t0) / 20
assert all(np.array_equal(a, b) for a, b in zip(got, gotp))
print("mc_jobs (pinned host planes): %.3f ms per picture = %.1f pictures/s" % (dtp * 1e3, 1 / dtp))
dev = [torch.zeros((H, W), dtype=torch.int16, device="cuda"), torch.zeros((H // 2, W // 2), dtype=torch.int16, device="cuda"),
       torch.zeros((H // 2, W // 2), dtype=torch.int16, device="cuda")]
ptrs, strides = [t.data_ptr() for t in dev], [W, W // 2, W // 2]
stream = torch.cuda.ExternalStream(g.stream)
for i in range(3): g.mc_jobs_device(jobs, ptrs, strides)
g.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for i in range(50): g.mc_jobs_device(jobs, ptrs, strides)
e1.record(stream); g.sync()
ms = e0.elapsed_time(e1) / 50
samples = W * H * 1.5
print("mc_jobs_device: %.4f ms per picture (device planes, job list from the host) = %.1f G samples/s, %.0f GB/s algorithmic (4 B/sample)"
      % (ms, samples / ms / 1e6, 4 * samples / ms / 1e6))
