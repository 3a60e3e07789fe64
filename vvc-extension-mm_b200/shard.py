"""Multi-GPU plumbing: the path shards by CTU rows of a frame against shared reference pictures.

One process per GPU; torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) moves the only
data that has to cross GPUs: a newly reconstructed, padded reference picture (broadcast once per frame from the
rank that holds it) and, optionally, the per-candidate distortions (all-gather).  No collective sits inside the
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

    mode "p2p" (default on NVLink boxes): each rank publishes its picture in a symmetric-memory staging buffer and
    PULLS its peers' pictures with plain device-to-device copies, i.e. on the copy engines over NVLink / NVSwitch peer
    memory -- no collective kernel occupies SMs while k_map / k_fir of the current step run (an ncclAllGather
    overlapped with the kernels cost them 30-50 % at 8 GPUs).  Two device-side barriers (signal pads in peer memory)
    order publish -> pull -> reuse.  mode "nccl": one ncclAllGather (also the fallback when symmetric memory cannot be
    set up, and what the gloo CPU tests use)."""

    def __init__(self, pic_bytes, device, group=None, mode=None):
        import os
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n = int(pic_bytes)
        self.mode = "nccl"
        want = (mode or os.environ.get("MMGPU_EXCHANGE", "p2p")).lower()
        if want == "p2p" and self.world > 1 and torch.device(device).type == "cuda":
            ok = 1
            try:
                import torch.distributed._symmetric_memory as symm
                self.stage = symm.empty(self.n, dtype=torch.uint8, device=device)
                self.hdl = symm.rendezvous(self.stage, group if group is not None else dist.group.WORLD)
                self.peers = [self.hdl.get_buffer(r, (self.n,), torch.uint8) for r in range(self.world)]
            except Exception as e:  # noqa: BLE001 -- any failure means: use the NCCL path on every rank
                self.err = repr(e)
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 1:
                self.mode = "p2p"

    def allgather(self, landing, mine):
        """Enqueue the exchange on the current stream."""
        if self.mode == "nccl":
            dist.all_gather_into_tensor(landing, mine, group=self.group)
            return
        self.stage.copy_(mine, non_blocking=True)
        self.hdl.barrier(channel=0)                     # every rank's picture is published
        for step in range(self.world):
            r = (self.rank - step) % self.world         # start with the local copy, then walk the ring
            landing[r * self.n:(r + 1) * self.n].copy_(self.peers[r], non_blocking=True)
        self.hdl.barrier(channel=0)                     # every rank has pulled: staging buffers may be rewritten


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
