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
