#!/usr/bin/env python3
"""bench.py -- multi-model MC + SATD throughput of the B200 path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU arm (the reference's compiled AVX2 FIR / Hadamard
                                                              # objects driven like its per-candidate loop, all host threads)

A "step" is one pass of the hot path over one frame-wide batch: every block of a 2048x1024 ERP frame
(mixed VVC partition) x 16 candidate MVs, per block: map (MPA_FB/LR/TB, GED_CAMPOSE epipole (1,0,0), ROT)
-> 8-tap separable interpolation -> Hadamard SATD against the original.  Synthetic frames, 10 bit, 4:2:0.
`value`  : luma sample-candidates/s with blocks, MVs, references and originals resident in HBM.
`e2e`    : the same through the host C ABI (mmgpu_set_org + mmgpu_me_batch) from pinned host buffers,
           H2D of the original picture / blocks / MVs and D2H of the distortions inside the timed region.
N > 1    : weak scaling -- each step covers N frames; rank r evaluates CTU rows r, r+N, ... of all N frames;
           the reference picture of frame f reaches every GPU from rank f inside the timed region (NVLS multicast push up to
           4 GPUs, one ncclAllGather beyond; other forms by MMGPU_EXCHANGE).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, CTU = 2048, 1024, 128
CANDS = 16
# DRAM bytes and executed warp-instructions of one k_map / k_fir<DIST> launch of this workload come from the committed ncu
# capture (profiles/ncu_counts.json, written by tools/ncu_counts.py) and are used only while the kernel sources are the ones
# that capture was taken with (SHA-256 of csrc/ in the file); otherwise the line carries null for them.
POOL = 32            # synthetic frames resident per GPU (references cycle through them: 270 MB > L2)
MV_RANGE = 64
EPIPOLE = (1.0, 0.0, 0.0)
WORKLOAD = "2048x1024 ERP frame batch (Broadway shape, configs[1]): MPA_FB/LR/TB+GED_CAMPOSE+ROT, mixed 8x4..128x128 " \
           "partition, 16 integer-pel candidates/block, MC + Hadamard SATD (HadamardME)"


def models():
    import vvc_extension_mm_b200 as mm
    return (mm.MPA_FB, mm.MPA_LR, mm.MPA_TB, mm.GED_CAMPOSE, mm.ROT)


def algorithmic_bytes(blocks):
    """per block 2WH (original) + per candidate 2WH (reference samples) + 8 (distortion)."""
    wh = blocks["w"].astype(np.int64) * blocks["h"].astype(np.int64)
    return int((2 * wh + blocks["cand_count"].astype(np.int64) * (2 * wh + 8)).sum())


class ClockSampler:
    """nvidia-smi in loop mode (100 ms) while the GPU is under the bench's load (warm-up, timed steps, e2e steps)."""

    def __init__(self, index):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu,pcie.link.gen.current," \
            "pcie.link.width.current"
        self.cmd = ["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"]
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(self.cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            rows = [[c.strip() for c in l.split(",")] for l in out.splitlines() if l.strip()]
        rows = [r for r in rows if len(r) >= 7 and r[0].isdigit()]
        busy = [r for r in rows if r[6].isdigit() and int(r[6]) > 0] or rows
        sm = [int(r[0]) for r in busy]
        mx = [int(r[1]) for r in rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[2 + i].lower().startswith("active")})
        pcie = sorted({"gen%s x%s" % (r[7], r[8]) for r in busy if len(r) >= 9})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(rows), "samples_under_load": len(busy), "pcie_link_under_load": pcie}


def cpu_workload(seed=1):
    """The frame batch of one step on the host: oracle context, padded reference, original, blocks, MVs."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from vvc_extension_mm_b200 import workload as wl
    ref = wl.pad_frame(wl.synth_frame(W, H, seed), 288, 1)
    cur = wl.synth_frame(W, H, seed + 1, shift=(3.0, 1.0))
    blocks, mvs, prec = wl.frame_batch(W, H, seed=seed, cands_per_block=CANDS, mv_range=MV_RANGE, models=models())
    return ol, ref, cur, blocks, mvs, prec


def oracle_distortions(ref, org, blocks, mvs, prec, n_threads, epipole=EPIPOLE):
    """Distortions of a whole batch through the CPU oracle (the checker), multi-threaded over blocks."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from concurrent.futures import ThreadPoolExecutor
    R = ol.Oracle.rodrigues(list(epipole))
    out = np.zeros(len(mvs), dtype=np.uint64)
    tls = threading.local()

    class P:
        pass
    pl = P()
    pl.origin, pl.stride = ref.origin, ref.stride

    def run_block(b):
        if not hasattr(tls, "o"):
            tls.o = ol.Oracle(W, H)
            tls.o.set_rotation(ol.GED_CAMPOSE, R)
        sl = slice(int(b["cand_first"]), int(b["cand_first"]) + int(b["cand_count"]))
        o_y, o_x = int(b["org_y"]), int(b["org_x"])
        blk = org[o_y:o_y + int(b["h"]), o_x:o_x + int(b["w"])]
        m = np.stack([mvs["hor"][sl], mvs["ver"][sl]], axis=1)
        out[sl] = tls.o.me_block(pl, blk, None, (int(b["x"]), int(b["y"])), (int(b["w"]), int(b["h"])), int(b["model"]), m,
                                 mv_prec_shift=prec, dist=ol.DIST_HAD, had_mode=ol.HAD_SIMD)

    with ThreadPoolExecutor(max_workers=n_threads) as ex:
        list(ex.map(run_block, blocks))
    return out


_REFBENCH = os.path.join(ROOT, "oracle", "_ref", "librefbench.so")


def cpu_reference_throughput(n_threads, budget_s, seed=1):
    """CPU arm: the reference's per-candidate loop xMVReprojectionInterpolation + distFunc on the same frame batch.
    oracle/_ref/librefbench.so (kind "reference": the reference's own compiled InterpolationFilter / RdCost AVX2 code for
    the interpolation and the Hadamard SATD, driven per 4x4 like InterSearch.cpp:6330-6380; the float mapping is the C
    restatement because Eigen is absent) when it was built, else the pure oracle port (kind "port").
    Returns (sample-candidates/s, kind, description of the sample)."""
    import ctypes as C
    ol, ref, cur, blocks, mvs, prec = cpu_workload(seed)
    org = np.ascontiguousarray(cur[0])
    if os.path.exists(_REFBENCH):
        L = C.CDLL(_REFBENCH)
        L.refbench_run.restype = C.c_double
        L.refbench_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        cfg = ol.OrcCfg(W, H, 1, 128, 1.0, 1, 10)
        R = ol.Oracle.rodrigues(list(EPIPOLE))
        b = np.ascontiguousarray(blocks)
        m = np.ascontiguousarray(mvs)
        stats = np.zeros(5)
        v = L.refbench_run(C.byref(cfg), R.ctypes.data, ref[0].origin, ref[0].stride, org.ctypes.data, org.strides[0] // 2,
                           b.ctypes.data, len(b), m.ctypes.data, prec, 1, n_threads, float(budget_s), None, stats.ctypes.data)
        tot = max(1e-9, stats[2] + stats[3] + stats[4])
        desc = ("%d whole blocks of the same frame batch in %.1f s on %d thread(s); reference's compiled AVX2 filterHor/filterVer + "
                "setDistParam/distFunc per candidate, C restatement of the Eigen mapping (time split: mapping %.0f%%, interpolation "
                "%.0f%%, distortion %.0f%%)" % (int(stats[1]), stats[0], n_threads, 100 * stats[2] / tot, 100 * stats[3] / tot,
                                                  100 * stats[4] / tot))
        return float(v), "reference", desc, float(stats[0])
    t0 = time.perf_counter()
    done = nblk = 0
    step = 64 * n_threads
    i = 0
    while time.perf_counter() - t0 < budget_s:
        part = blocks[[(i + k) % len(blocks) for k in range(step)]].copy()
        first = 0
        for q in part:
            first += int(q["cand_count"])
        oracle_distortions(ref[0], org, part, mvs, prec, n_threads)
        done += int((part["w"].astype(np.int64) * part["h"] * part["cand_count"]).sum())
        nblk += step
        i += step
    dt = time.perf_counter() - t0
    return done / dt, "port", "%d whole blocks of the same frame batch in %.1f s, oracle port, %d thread(s)" % (nblk, dt, n_threads), dt


def bench_config(N, blocks_per_rank=None, cands_per_rank=None, samples=None, parallelism="1 gpu"):
    """config keys shared by both arms (the driver compares them)"""
    return {"workload": WORKLOAD, "frames_per_step": N, "blocks_per_rank": blocks_per_rank, "candidates_per_rank": cands_per_rank,
            "sample_candidates_per_step": samples,
            "l2": "references cycle through %d resident frames (%.0f MB of luma planes > 126 MB L2)" % (POOL, POOL * 8.4),
            "parallelism": parallelism}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on all host threads; rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    kind, desc = "port", ""
    for s in range(args.warmup + args.steps):
        v, kind, desc, dt = cpu_reference_throughput(cores, per_step, seed=1 + (s % 4))
        if s >= args.warmup:
            vals.append((v, dt))
    tot_s = sum(d for _, d in vals)
    value = sum(v * d for v, d in vals) / max(1e-9, tot_s)
    from vvc_extension_mm_b200 import workload as wl
    blocks, mvs, _ = wl.frame_batch(W, H, seed=1, cands_per_block=CANDS, mv_range=MV_RANGE, models=models())
    samples = float((blocks["w"].astype(np.int64) * blocks["h"] * blocks["cand_count"]).sum())
    line = {
        "impl": "reference", "metric": "multi-model MC+SATD luma samples/sec at 2048x1024 ERP", "value": value,
        "unit": "luma sample-candidates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * tot_s / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 mapping + s16/s32 FIR/SATD", "data": "synthetic",
        "config": bench_config(1, int(len(blocks)), int(len(mvs)), samples),
        "cpu_baseline": {"value": value, "unit": "luma sample-candidates/s", "cores": cores, "kind": kind,
                         "sample": "per step: " + desc},
        "e2e": {"value": value, "unit": "luma sample-candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def kernel_source_hash():
    import hashlib
    d = os.path.join(ROOT, "vvc-extension-mm_b200", "csrc")
    h = hashlib.sha256()
    # the translation units and headers k_map / k_fir are compiled from (the host ABI, MC / DMVR kernels do not change their SASS)
    files = ("mm_kernels_map.cu", "mm_kernels_fir.cu", "mm_fir.cuh", "mm_map.cuh", "mm_libm.cuh", "mm_sbpos.cuh", "mm_common.cuh", "mm_htaps.h",
             "mm_tan_exceptions.h")
    for f in files:
        if os.path.exists(os.path.join(d, f)):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


def ncu_counts():
    """Instruction / DRAM-byte counts of one k_map and one k_fir launch of this workload from the committed ncu capture,
    valid only for the kernel sources they were captured with."""
    try:
        c = json.load(open(os.path.join(ROOT, "profiles", "ncu_counts.json")))
        return c if c.get("csrc_sha256") == kernel_source_hash() and "MMGPU_LIB" not in os.environ else None
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of back-to-back steps for the sustained figure")
    ap.add_argument("--no-parity-check", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import vvc_extension_mm_b200 as mm
    from vvc_extension_mm_b200 import workload as wl
    from vvc_extension_mm_b200 import shard
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = world
    dev = torch.device("cuda", local)

    # ---- resident data -------------------------------------------------------------------------------------
    # A picture = one contiguous byte buffer holding the padded Y, Cb, Cr planes (one NCCL broadcast per picture).
    # Ring of landing slots: the references of the next RING - 1 steps arrive while the current step computes.  With 2 the exchange
    # of step s + 1 can only start when step s - 1 has finished on EVERY rank and must end before step s ends: its whole chain
    # (barrier, transfer, barrier) has to fit one step; with 3 it has two steps (MMGPU_RING).
    # Default: 3 for the multicast push (measured at N = 2 and 4: same device figure, e2e 374 instead of 312 G/s at N = 2), 2 for
    # the NCCL forms (the configuration their N = 8 numbers were measured with).
    RING_MAX = max(2, int(os.environ.get("MMGPU_RING", "3")))
    n_comm = RING_MAX * N if N > 1 else 0
    gpu = mm.MmGpu(W, H, max_refs=POOL + n_comm, device=local)
    gpu.set_rotation(0, mm.rodrigues(EPIPOLE))
    geo = wl.pad_frame(wl.synth_frame(W, H, seed=100), 288, 1)
    plane_bytes = [p.buf.nbytes for p in geo]
    offs = np.concatenate([[0], np.cumsum([(n + 255) // 256 * 256 for n in plane_bytes])]).astype(np.int64)
    pic_bytes = int(offs[-1])

    def new_picture():
        return torch.empty(pic_bytes, dtype=torch.uint8, device=dev)

    def register(slot, pic, poc):
        origins = [pic.data_ptr() + int(offs[c]) + 2 * (p.my * p.stride + p.mx) for c, p in enumerate(geo)]
        gpu.set_ref_device(slot, poc, origins, [p.stride for p in geo], keep=pic)

    # un-padded form of a picture (what crosses NVLink at N > 1: the margins are rebuilt on arrival by k_extend_border)
    uw = [(p.w, p.h) for p in geo]
    uoffs = np.concatenate([[0], np.cumsum([(2 * w_ * h_ + 255) // 256 * 256 for w_, h_ in uw])]).astype(np.int64)
    upic_bytes = int(uoffs[-1])
    pool_ref, pool_org_h, pool_unp = [], [], []
    for f in range(POOL):
        planes = geo if f == 0 else wl.pad_frame(wl.synth_frame(W, H, seed=100 + f, shift=(0.37 * f, 0.11 * f)), 288, 1)
        pic = new_picture()
        for c, p in enumerate(planes):
            pic[int(offs[c]):int(offs[c]) + plane_bytes[c]].copy_(torch.from_numpy(p.buf.view(np.uint8).reshape(-1)))
        pool_ref.append(pic)
        register(f, pic, f)
        if N > 1:
            up = torch.empty(upic_bytes, dtype=torch.uint8, device=dev)
            for c, p in enumerate(planes):
                inner = np.ascontiguousarray(p.inner)
                up[int(uoffs[c]):int(uoffs[c]) + inner.nbytes].copy_(torch.from_numpy(inner.view(np.uint8).reshape(-1)))
            pool_unp.append(up)
        pool_org_h.append(torch.from_numpy(wl.synth_plane(W, H, 5000 + f, shift=(0.37 * f + 2.5, 0.11 * f + 1.0))).pin_memory())
    pool_org_d = [t.to(dev) for t in pool_org_h]
    # landing buffers: parity p holds the N UN-PADDED pictures of a step back to back (filled by ONE all-gather: rank f
    # contributes the picture it owns, i.e. N simultaneous broadcasts over NVLink/NVSwitch); each is then copied into the
    # padded planes of its reference slot and its borders are extended on the device (mmgpu_set_ref_unpadded_device)
    # MMGPU_EXCHANGE_PADDED=1: round-1 form for comparison -- the padded pictures themselves are all-gathered straight into
    # the registered reference slots (twice the NVLink bytes, no padding kernel)
    # Exchange of the step's reference pictures.  Default: up to 4 GPUs the NVLS multicast push of padded pictures
    # (MMGPU_EXCHANGE=push) where symmetric memory has a multicast mapping, else ONE ncclAllGather of the padded pictures.  Other forms, for comparison:
    # MMGPU_EXCHANGE=nccl with MMGPU_EXCHANGE_PADDED=0 (un-padded all-gather + k_pad_planes), =pull (un-padded, SM / TMA pulls
    # fused with the border extension), =p2p (copy-engine pulls).
    explicit_x = "MMGPU_EXCHANGE" in os.environ
    # measured on B200 boxes: push wins at N = 2 (408 vs 396 G/s), ties at N = 4, loses at N = 8 where every form is bound
    # by the ~400 GB/s each GPU can take in over NVLink (1050 vs 1285 G/s for the all-gather, which spreads the arrivals better)
    want_x = os.environ.get("MMGPU_EXCHANGE", "push" if N <= 4 else "nccl").lower()
    padded_x = os.environ.get("MMGPU_EXCHANGE_PADDED", "1" if want_x in ("push", "nccl") and not explicit_x else "0") == "1" or want_x == "push"
    xbytes = pic_bytes if padded_x else upic_bytes
    exchange = shard.PictureExchange(xbytes, dev, mode=want_x, ring=RING_MAX) if N > 1 else None
    push = exchange is not None and exchange.mode == "push"
    RING = RING_MAX if ("MMGPU_RING" in os.environ or push) else 2
    LOOK = RING - 1
    if N > 1 and want_x == "push" and not push and explicit_x:
        raise SystemExit("MMGPU_EXCHANGE=push needs symmetric memory with a multicast mapping (%s)" % getattr(exchange, "err", "?"))
    if push:
        comm = [exchange.landing(p) for p in range(RING)]      # the landing slots live in the multicast-mapped symmetric allocation
    else:
        comm = [torch.zeros(N * xbytes, dtype=torch.uint8, device=dev) for _ in range(RING if N > 1 else 0)]
    for p, big in enumerate(comm):     # allocate + publish the planes of every landing slot once (addresses stay fixed)
        for f in range(N):
            if padded_x:
                register(POOL + p * N + f, big[f * pic_bytes:(f + 1) * pic_bytes], f)
            else:
                base = big.data_ptr() + f * upic_bytes
                gpu.set_ref_unpadded_device(POOL + p * N + f, f, [base + int(uoffs[c]) for c in range(3)], [w_ for w_, _ in uw], keep=big)
    gpu.sync()

    # ---- batch: this rank's CTU rows of the N frames of a step ------------------------------------------------
    rows = wl.shard_ctu_rows(H // CTU, rank, N)
    blocks_l, mvs_l = [], []
    for f in range(N):
        b, m, prec = wl.frame_batch(W, H, seed=1 + f, cands_per_block=CANDS, mv_range=MV_RANGE, models=models(),
                                    ref_slot=f, ctu_rows=rows)
        b["cand_first"] += sum(len(x) for x in mvs_l)
        b["org_y"] += f * H
        blocks_l.append(b)
        mvs_l.append(m)
    blocks = np.concatenate(blocks_l)     # ref_slot holds the frame index f here; step_plan() maps it to a slot
    mvs = np.concatenate(mvs_l)
    n_cands = len(mvs)
    samples_rank = int((blocks["w"].astype(np.int64) * blocks["h"] * blocks["cand_count"]).sum())
    alg_bytes_rank = algorithmic_bytes(blocks)
    d_mvs = torch.from_numpy(mvs.view(np.int32).reshape(-1, 2).copy()).to(dev)
    d_dist = torch.zeros(n_cands, dtype=torch.int64, device=dev)
    d_org = torch.empty((N * H, W), dtype=torch.int16, device=dev)
    # e2e uploads only what this rank evaluates: its CTU rows of the N original pictures, packed into one plane
    # (frame f, kept row k -> lines (f * len(rows) + k) * CTU ..); blocks of the e2e call point into that plane
    n_rows = len(rows)
    h_org = torch.empty((N * n_rows * CTU, W), dtype=torch.int16).pin_memory()
    row_pos = {r: k for k, r in enumerate(rows)}
    e2e_org_y = np.array([((int(b["org_y"]) // H) * n_rows + row_pos[(int(b["org_y"]) % H) // CTU]) * CTU + int(b["org_y"]) % CTU
                          for b in blocks], dtype=np.int32)
    h_blocks = torch.from_numpy(blocks.view(np.int32).reshape(-1, 12).copy()).pin_memory()
    h_mvs = torch.from_numpy(mvs.view(np.int32).reshape(-1, 2).copy()).pin_memory()
    h_dist = torch.zeros(n_cands, dtype=torch.int64).pin_memory()
    gpu.set_org_device(d_org.data_ptr(), W, W, N * H, keep=d_org)
    plans = {}
    stream = torch.cuda.ExternalStream(gpu.stream, device=dev)
    # the exchange stream of the hand-written forms has the higher priority: their few CTAs become resident as soon as CTAs of the
    # running step retire instead of queueing behind the step's 65 k pending CTAs; NCCL measured slower with it (MMGPU_COMM_PRIORITY)
    hand_written = exchange is not None and exchange.mode in ("push", "pull")
    comm_prio = int(os.environ.get("MMGPU_COMM_PRIORITY", "-1" if hand_written else "0"))
    comm_stream = torch.cuda.Stream(device=dev, priority=comm_prio) if N > 1 else None
    ready = [torch.cuda.Event() for _ in range(RING)]   # landing buffers of ring position p hold the references of a step
    freed = [torch.cuda.Event() for _ in range(RING)]   # the step that read ring position p has finished

    def slots_of(s):
        """reference slot of frame f of step s"""
        if N == 1:
            return np.array([s % POOL])
        return POOL + (s % RING) * N + np.arange(N)

    def step_plan(s):
        key = (s % POOL) if N == 1 else (s % RING)
        if key not in plans:
            b = blocks.copy()
            b["ref_slot"] = slots_of(s)[blocks["ref_slot"]]
            plans[key] = (gpu.plan(b, n_cands), b)
        return plans[key]

    pull = exchange is not None and exchange.mode == "pull"
    push_ctas = int(os.environ.get("MMGPU_PUSH_CTAS", "16"))
    pull_ctas = int(os.environ.get("MMGPU_PULL_CTAS", "24"))
    if exchange is not None and os.environ.get("MMGPU_EXCHANGE", "nccl").lower() == "pull" and (not pull or padded_x):
        raise SystemExit("MMGPU_EXCHANGE=pull needs symmetric memory and the un-padded exchange (%s)" % getattr(exchange, "err", "padded form selected"))

    def prefetch(s):
        """Rank f owns frame f of step s: it copies its picture into the landing buffer and NCCL-broadcasts it.
        Runs on the communication stream, one step ahead of the compute stream."""
        if N == 1:
            return
        p = s % RING
        if push:
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(freed[p])
                exchange.push(gpu, p, pool_ref[(s * N + rank) % POOL], max_ctas=push_ctas)
                ready[p].record(comm_stream)
            return
        if pull:
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(freed[p])
                exchange.publish(p, pool_unp[(s * N + rank) % POOL])
                gpu.pull_refs_unpadded([POOL + p * N + f for f in range(N)], list(range(N)),
                                       [[exchange.peer_ptr(f, p) + int(uoffs[c]) for c in range(3)] for f in range(N)],
                                       [[w_ for w_, _ in uw]] * N, max_ctas=pull_ctas, stream=comm_stream.cuda_stream, keep=exchange)
                exchange.done()
                ready[p].record(comm_stream)
            return
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(freed[p])
            exchange.allgather(comm[p], (pool_ref if padded_x else pool_unp)[(s * N + rank) % POOL])
            if not padded_x:
              gpu.set_refs_unpadded_device([POOL + p * N + f for f in range(N)], list(range(N)),
                                           [[comm[p].data_ptr() + f * upic_bytes + int(uoffs[c]) for c in range(3)] for f in range(N)],
                                           [[w_ for w_, _ in uw]] * N, stream=comm_stream.cuda_stream, keep=comm[p])
            ready[p].record(comm_stream)

    def prime():
        """every ring position free, then the exchanges of the first LOOK steps"""
        for p in range(RING):
            freed[p].record(stream)
        for k in range(LOOK):
            prefetch(k)

    def load_orgs(s, tgt, src_pool):
        for f in range(N):
            tgt[f * H:(f + 1) * H].copy_(src_pool[(s * N + f + 1) % POOL], non_blocking=True)

    def device_step(s):
        prefetch(s + LOOK)
        with torch.cuda.stream(stream):
            if N > 1:
                stream.wait_event(ready[s % RING])
            plan, _ = step_plan(s)
            gpu.me_batch_device(plan, d_mvs.data_ptr(), d_dist.data_ptr(), mv_prec_shift=prec, dist=mm.DIST_HAD)
            if N > 1:
                freed[s % RING].record(stream)

    # e2e: the host-facing call with every input in pinned HOST memory -- original picture, blocks and MVs go up, the
    # distortions come back, every step.  NSLOT batches are kept in flight (mmgpu_me_batch_submit / _wait) so that the
    # copies and the preparation of step s+1 overlap the kernels of step s and a short stall of the host thread does not
    # idle the GPU; step s's result is read before step s+NSLOT is submitted.
    NSLOT = 3
    h_dist2 = [h_dist] + [torch.zeros(n_cands, dtype=torch.int64).pin_memory() for _ in range(NSLOT - 1)]
    h_blocks2 = [h_blocks] + [h_blocks.clone().pin_memory() for _ in range(NSLOT - 1)]

    def e2e_submit(s):
        prefetch(s + LOOK)
        if N > 1:
            ready[s % RING].synchronize()
        _, b = step_plan(s)
        hb = h_blocks2[s % NSLOT].numpy().view(mm.api.BLOCK_DTYPE).reshape(-1)
        hb["ref_slot"] = b["ref_slot"]
        hb["org_y"] = e2e_org_y
        gpu.me_batch_submit(s % NSLOT, h_org.numpy(), hb, h_mvs.numpy().view(mm.api.MV_DTYPE).reshape(-1),
                            h_dist2[s % NSLOT].numpy().view(np.uint64), mv_prec_shift=prec, dist=mm.DIST_HAD)
        if N > 1:
            freed[s % RING].record(stream)

    def e2e_step(s):
        if s >= NSLOT:
            gpu.me_batch_wait(s % NSLOT)
        e2e_submit(s)

    def e2e_drain(k):
        for s in range(max(0, k - NSLOT), k):
            gpu.me_batch_wait(s % NSLOT)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, prep, drain=None):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        prime()                          # prologue: references of the first LOOK steps (steady state prefetches LOOK steps ahead)
        barrier()
        t0 = time.perf_counter()
        ev0.record(stream)
        for s in range(k):
            prep(s)
            fn(s)
        if drain is not None:
            drain(k)
        ev1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1000.0
        return ev0.elapsed_time(ev1), wall

    # org pictures of a step are copied into the contiguous N*H org plane outside the device timing for `value`
    # (resident in HBM when the step starts) and live in pinned host memory for e2e.
    with torch.cuda.stream(stream):
        load_orgs(0, d_org, pool_org_d)
    for f in range(N):
        for k, r in enumerate(rows):
            h_org[(f * n_rows + k) * CTU:(f * n_rows + k + 1) * CTU].copy_(pool_org_h[(f + 1) % POOL][r * CTU:(r + 1) * CTU])

    for s in range(POOL if N == 1 else RING):
        step_plan(s)          # plans (device-resident block / work lists) are built once, outside any timed region
    sampler = ClockSampler(local)
    sampler.start()
    prime()
    for s in range(args.warmup):
        device_step(s)
    barrier()
    gpu.kernel_timing(True)
    launches0 = gpu.launch_count
    dev_ms, _ = timed(device_step, args.steps, lambda s: None)
    launches = gpu.launch_count - launches0
    kmap_ms, k_ms, k_n = gpu.kernel_timing_read()
    gpu.kernel_timing(False)
    for s in range(2 * NSLOT):
        e2e_step(s)
    e2e_drain(2 * NSLOT)
    # self-check: the host-facing call and the device-resident call evaluate the same batch -> identical distortions;
    # with N > 1 the exchanged pictures are what their owners hold (every rank can regenerate every synthetic picture)
    prime()
    if N > 1:
        barrier()
        for f in range(N if not pull else 0):    # pull mode has no landing buffer: the oracle check of step 0 below covers every exchanged picture
            if not torch.equal(comm[0][f * xbytes:(f + 1) * xbytes], (pool_ref if padded_x else pool_unp)[f % POOL]):
                raise SystemExit("bench self-check failed: picture %d arrived corrupted on rank %d (%s)" % (f, rank, exchange.mode))
    device_step(0)
    barrier()
    chk_dev = d_dist.cpu().numpy().view(np.uint64).copy()
    prime()
    e2e_submit(0)
    gpu.me_batch_wait(0)
    if not np.array_equal(chk_dev, h_dist2[0].numpy().view(np.uint64)):
        raise SystemExit("bench self-check failed: e2e distortions differ from the device-resident path")
    # parity of the TIMED batch: every distortion of step 0 (this rank's blocks of all N frames) against the CPU oracle
    parity = None
    if not args.no_parity_check:
        exp = np.zeros(n_cands, dtype=np.uint64)
        for f in range(N):
            sel = (blocks["org_y"] // H) == f
            if not sel.any():
                continue
            bf = blocks[sel].copy()
            bf["org_y"] -= f * H
            fr = (0 * N + f) % POOL
            planes = geo if fr == 0 else wl.pad_frame(wl.synth_frame(W, H, seed=100 + fr, shift=(0.37 * fr, 0.11 * fr)), 288, 1)
            exp += oracle_distortions(planes[0], pool_org_h[(f + 1) % POOL].numpy(), bf, mvs, prec, os.cpu_count() or 1)
        bad = int((exp != chk_dev).sum())
        parity = {"distortions_checked": int(n_cands), "mismatches": bad, "checker": "CPU oracle (oracle/mm_oracle.c), step 0 of the timed batch"}
        if bad:
            raise SystemExit("bench parity check failed on rank %d: %d of %d distortions differ from the oracle" % (rank, bad, n_cands))
    barrier()
    e2e_dev_ms, e2e_wall_ms = timed(e2e_step, args.steps, lambda s: None, e2e_drain)
    # sustained leg: back-to-back device-resident steps for about args.sustain seconds (clock samples fall inside it)
    sus_steps = max(args.steps, int(args.sustain * 1000.0 / max(1e-3, dev_ms / args.steps)))
    if dist is not None:      # every rank must run the same number of steps (each step holds a collective)
        n_sus = torch.tensor([sus_steps], dtype=torch.int64, device=dev)
        dist.all_reduce(n_sus, op=dist.ReduceOp.MIN)
        sus_steps = int(n_sus.item())
    sus_ms, _ = timed(device_step, sus_steps, lambda s: None)
    clocks = sampler.stop()
    # what the box's PCIe path gives a plain pinned host->device copy (e2e moves h2d bytes per step over it)
    probe_h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    probe_d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    h2d_probe = 0.0
    for _ in range(4):
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        probe_d.copy_(probe_h, non_blocking=True)
        pe1.record()
        torch.cuda.synchronize()
        h2d_probe = max(h2d_probe, (64 << 20) / 1e9 / (pe0.elapsed_time(pe1) / 1000.0))
    del probe_h, probe_d
    e2e_ms = max(e2e_dev_ms, e2e_wall_ms)   # host work (plan building, copies) is part of the call: use the larger clock

    t = torch.tensor([dev_ms, e2e_ms, float(samples_rank), float(alg_bytes_rank), k_ms, kmap_ms, sus_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms, k_ms, kmap_ms, sus_ms = float(tmax[0]), float(tmax[1]), float(tmax[4]), float(tmax[5]), float(tmax[6])
        samples_all, bytes_all = float(tsum[2]), float(tsum[3])
    else:
        samples_all, bytes_all = float(samples_rank), float(alg_bytes_rank)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        value = samples_all * args.steps / (dev_ms / 1000.0)
        e2e_value = samples_all * args.steps / (e2e_ms / 1000.0)
        k_avg_ms = k_ms / max(1, k_n)
        achieved = (alg_bytes_rank / 1e9) / (k_avg_ms / 1000.0) if k_avg_ms > 0 else 0.0
        kmap_avg_ms = kmap_ms / max(1, k_n)
        path_ms = k_avg_ms + kmap_avg_ms
        path_achieved = (alg_bytes_rank / 1e9) / (path_ms / 1000.0) if path_ms > 0 else 0.0
        sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965)
        n_smsp = 4 * torch.cuda.get_device_properties(local).multi_processor_count
        counts = ncu_counts() if N == 1 else None
        issue = None
        if counts and k_avg_ms > 0 and kmap_avg_ms > 0:
            issue = {"unit": "warp-instructions/clk/SMSP", "peak": 1.0,
                     "k_fir": counts["k_fir"]["inst_executed"] / (k_avg_ms * 1e-3 * sm_hz * n_smsp),
                     "k_map": counts["k_map"]["inst_executed"] / (kmap_avg_ms * 1e-3 * sm_hz * n_smsp),
                     "note": "instruction counts from the committed ncu capture of one launch (%s), times measured live" % counts.get("source")}
        cpu = None
        if not args.no_cpu_baseline and N == 1:
            v, kind, desc, dt = cpu_reference_throughput(1, args.cpu_budget)
            cpu = {"value": v, "unit": "luma sample-candidates/s", "cores": 1, "kind": kind, "sample": desc}
        h2d = int(h_org.numel() * 2 + h_blocks.numel() * 4 + h_mvs.numel() * 4)
        line = {
            "metric": "multi-model MC+SATD luma samples/sec at 2048x1024 ERP", "value": value,
            "unit": "luma sample-candidates/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 mapping + s16/s32 FIR/SATD", "data": "synthetic",
            "config": bench_config(N, int(len(blocks)), int(n_cands), samples_all,
                                   ("ctu-rows x%d; the %d reference pictures of a step (%.1f MB each%s, one owner rank each) reach every GPU %s, "
                                    "prefetched %d step(s) ahead on a second stream"
                                    % (N, N, xbytes / 1e6, " with margins" if padded_x else
                                       (" un-padded; pulled row by row from the owners' symmetric memory by %d persistent CTAs that write the padded planes directly (k_pull_rows + k_pad_tb)" % pull_ctas
                                        if pull else " un-padded; margins rebuilt on arrival by one k_pad_planes launch"),
                                       "by SM loads over NVLink (mmgpu_pull_refs_unpadded)" if pull else
                                       ("by ONE multicast copy per owner (mmgpu_multicast_copy: multimem.st to the NVLS mapping of the landing slots, %d CTAs; the NVSwitch fans out)" % push_ctas) if push else
                                       ("by copy-engine pulls from symmetric (peer) memory over NVLink" if exchange.mode == "p2p" else "by one ncclAllGather"), LOOK))
                                   if N > 1 else "1 gpu"),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": counts["k_fir"]["dram_bytes"] if counts else None, "kernel": "k_fir<DIST>", "kernel_ms": k_avg_ms, "kernel_launches": k_n,
                         "path": {"kernels": "k_map + k_fir<DIST>", "k_map_ms": kmap_avg_ms, "ms": path_ms,
                                  "achieved": path_achieved, "frac": path_achieved / peak},
                         "algorithmic_bytes_per_launch": alg_bytes_rank, "issue": issue,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "luma sample-candidates/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(h_dist.numel() * 8), "ms_per_step": e2e_ms / args.steps,
                    "rank0_device_ms_per_step": e2e_dev_ms / args.steps, "rank0_host_ms_per_step": e2e_wall_ms / args.steps,
                    "h2d_probe_gbs": h2d_probe},
            "gpu_launches": int(launches),
            "parity": parity,
            "sustained": {"seconds": sus_ms / 1000.0, "steps": sus_steps, "ms_per_step": sus_ms / sus_steps,
                          "value": samples_all * sus_steps / (sus_ms / 1000.0)},
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
