"""Multi-GPU plumbing: the path shards by CTU rows of a frame against shared reference pictures.

One process per GPU; torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) moves the only
data that has to cross GPUs: a newly reconstructed reference picture (once per frame from the rank that holds it:
PictureExchange -- NCCL all-gather, NVLS multicast push, fused pull or copy-engine pulls over symmetric memory) and,
optionally, the per-candidate distortions (all-gather).  No collective sits inside the
MC/distortion kernels: blocks of different CTU rows are independent.
"""
import numpy as np
import torch
import torch.distributed as dist

from .workload import shard_ctu_rows  # noqa: F401  (re-exported)


def broadcast_planes(planes, src, group=None):
    """ncclBroadcast of the int16 planes of one reference picture.  NCCL/gloo have no int16 type: the bytes are
    sent as uint8 views of the same storage (no copy)."""
    for t in planes:
        dist.broadcast(t.view(torch.uint8), src=src, group=group)


def allgather_pictures(landing, mine, group=None):
    """Every rank owns one newly reconstructed picture (`mine`, a contiguous byte buffer with the padded planes);
    after the call `landing` (world * len(mine) bytes) holds all of them in rank order on every GPU: the N
    per-picture broadcasts of a step fused into one ncclAllGather."""
    dist.all_gather_into_tensor(landing, mine, group=group)


class PictureExchange:
    """The per-step exchange of newly reconstructed reference pictures: every rank contributes one picture, every rank
    ends up with all of them in `landing` (rank order).

    mode "nccl" (default): one ncclAllGather; also what the gloo CPU tests use.
    mode "p2p" (MMGPU_EXCHANGE=p2p): each rank publishes its picture in a symmetric-memory staging buffer and PULLS its
    peers' pictures with plain device-to-device copies, i.e. on the copy engines over NVLink / NVSwitch peer memory, so
    that no collective kernel occupies SMs while k_map / k_fir of the current step run; two device-side barriers (signal
    pads in peer memory) order publish -> pull -> reuse.  Measured on 8 B200 (bench.py, 12.6 MB pictures, 88 MB per
    rank and step): the overlapped ncclAllGather slows the kernels (k_fir 0.096 -> 0.125 ms, k_map 0.064 -> 0.095 ms) but
    finishes inside the step (0.232 ms per step, 1158 G sample-candidates/s); the copy-engine pulls leave the kernels
    almost alone (0.108 / 0.077 ms) but move only ~280 GB/s per GPU whether issued on 1, 4 or 7 streams, so the step
    becomes exchange-bound (0.31 ms, 847 G/s).  At 2 GPUs both are equal (360 vs 357 G/s).  Hence the default."""

    def __init__(self, pic_bytes, device, group=None, mode=None, ring=2):
        import os
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n = int(pic_bytes)
        self.mode = "nccl"
        want = (mode or os.environ.get("MMGPU_EXCHANGE", "nccl")).lower()
        if want in ("p2p", "pull", "push") and self.world > 1 and torch.device(device).type == "cuda":
            ok = 1
            try:
                import torch.distributed._symmetric_memory as symm
                # pull: `ring` staging pictures (ring position of the step), read in place by the peers' pull kernels
                # push: the landing slots themselves (ring x world pictures) are symmetric + multicast-mapped
                self.stage = symm.empty(self.n * (ring if want == "pull" else ring * self.world if want == "push" else 1), dtype=torch.uint8, device=device)
                self.hdl = symm.rendezvous(self.stage, group if group is not None else dist.group.WORLD)
                self.peers = [self.hdl.get_buffer(r, (self.stage.numel(),), torch.uint8) for r in range(self.world)]
                if want == "push" and not self.hdl.multicast_ptr:
                    raise RuntimeError("no multicast (NVLS) mapping for symmetric memory on this box")
            except Exception as e:  # noqa: BLE001 -- any failure means: use the NCCL path on every rank
                self.err = repr(e)
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 1 and want in ("pull", "push"):
                self.mode = want
            elif int(flag.item()) == 1:
                self.mode = "p2p"
                # one copy engine moves ~280 GB/s over NVLink (measured: 7 pulls of 12.6 MB in 0.32 ms); several pull
                # streams keep several engines busy
                k = max(1, min(self.world - 1, int(os.environ.get("MMGPU_EXCHANGE_STREAMS", "4"))))
                self.side = [torch.cuda.Stream(device=device) for _ in range(k)]
                self.fork = torch.cuda.Event()
                self.join = [torch.cuda.Event() for _ in range(k)]

    def allgather(self, landing, mine):
        """Enqueue the exchange on the current stream."""
        if self.mode == "nccl":
            dist.all_gather_into_tensor(landing, mine, group=self.group)
            return
        cur = torch.cuda.current_stream()
        self.stage.copy_(mine, non_blocking=True)
        landing[self.rank * self.n:(self.rank + 1) * self.n].copy_(mine, non_blocking=True)
        self.hdl.barrier(channel=0)                     # every rank's picture is published
        self.fork.record(cur)
        for i, st in enumerate(self.side):              # peers spread over the pull streams, ring order from the neighbour
            st.wait_event(self.fork)
            with torch.cuda.stream(st):
                for step in range(1 + i, self.world, len(self.side)):
                    r = (self.rank - step) % self.world
                    landing[r * self.n:(r + 1) * self.n].copy_(self.peers[r][:self.n], non_blocking=True)
                self.join[i].record(st)
        for ev in self.join:
            cur.wait_event(ev)
        self.hdl.barrier(channel=0)                     # every rank has pulled: staging buffers may be rewritten


    # ---- mode "pull" (MMGPU_EXCHANGE=pull): exchange fused with the border extension --------------------------------------
    # The owner publishes its UN-PADDED picture in symmetric memory; every rank's mmgpu_pull_refs_unpadded kernels read the
    # peers' pictures in place over NVLink (SM loads of a few persistent CTAs) and write the padded reference planes directly:
    # half the NVLink bytes of the padded all-gather, no landing buffer, no separate padding pass, and an SM budget the caller
    # chooses.  publish() -> [pull kernels, enqueued by the caller on the same stream] -> done().
    def publish(self, parity, mine):
        self.stage[parity * self.n:(parity + 1) * self.n].copy_(mine, non_blocking=True)
        self.hdl.barrier(channel=0)                     # every rank's picture of this step is in place

    def peer_ptr(self, r, parity):
        """device address (peer-mapped for r != rank) of rank r's published picture of that parity"""
        return self.peers[r].data_ptr() + parity * self.n

    def done(self):
        self.hdl.barrier(channel=0)                     # every rank has pulled: the staging picture may be rewritten


    # ---- mode "push" (MMGPU_EXCHANGE=push): NVLS multicast -------------------------------------------------------------------
    # The landing reference slots of every rank (2 parities x world padded pictures) are one symmetric allocation with a multicast
    # mapping.  The owner of a picture copies it ONCE to multicast_ptr + slot offset (mmgpu_multicast_copy, multimem.st): the
    # NVSwitch writes it into the slot of all ranks, the owner's included.  Per rank and step: one picture read, one picture of
    # NVLink egress, a few CTAs for a few microseconds -- instead of world - 1 pictures gathered by collective kernels.
    def landing(self, parity):
        """local view of the world landing pictures of that parity (register them as reference slots once)"""
        return self.stage[parity * self.world * self.n:(parity + 1) * self.world * self.n]

    def push(self, gpu, parity, mine, max_ctas=16):
        """Enqueue on the current stream: [all ranks done with the parity's slots] -> multicast copy -> [all pushes landed]."""
        self.hdl.barrier(channel=0)
        gpu.multicast_copy(mine.data_ptr(), self.hdl.multicast_ptr + (parity * self.world + self.rank) * self.n, self.n, max_ctas=max_ctas,
                           stream=torch.cuda.current_stream().cuda_stream)
        self.hdl.barrier(channel=0)


def shard_blocks(blocks, rank, world, ctu=128, pic_h=None):
    """Blocks (numpy BLOCK_DTYPE) whose CTU row belongs to `rank` (round-robin over CTU rows)."""
    rows = (blocks["y"] // ctu) % world
    return blocks[rows == rank]


def gather_distortions(local, group=None):
    """all_gather of per-rank distortion vectors of different lengths -> list of numpy uint64 arrays."""
    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    m = int(max(int(s) for s in sizes))
    pad = torch.zeros(m, dtype=torch.int64, device=local.device)
    pad[:local.numel()] = local.view(torch.int64)
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return [o[:int(s)].cpu().numpy().view(np.uint64) for o, s in zip(outs, sizes)]
