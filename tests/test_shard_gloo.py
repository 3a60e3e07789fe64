"""N > 1 host logic on CPU: world_size 2, gloo.  Each rank takes its CTU rows of a frame batch, receives the
reference picture by broadcast from rank 0, evaluates its blocks with the CPU oracle, and the gathered
distortions must equal the single-process result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 512, 256


def _eval(o, pl, cur, blocks, mvs, prec):
    import oracle_lib as ol
    out = []
    for b in blocks:
        sl = slice(b["cand_first"], b["cand_first"] + b["cand_count"])
        org = cur[b["y"]:b["y"] + b["h"], b["x"]:b["x"] + b["w"]]
        m = np.stack([mvs["hor"][sl], mvs["ver"][sl]], axis=1)
        out.append(o.me_block(pl, org, None, (int(b["x"]), int(b["y"])), (int(b["w"]), int(b["h"])), int(b["model"]), m,
                              mv_prec_shift=prec, dist=ol.DIST_HAD))
    return np.concatenate(out) if out else np.zeros(0, dtype=np.uint64)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as ol
    from vvc_extension_mm_b200 import shard, workload as wl
    o = ol.Oracle(W, H)
    o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues([1, 0, 0]))
    # rank 0 "reconstructed" the reference; the others start from garbage and receive it
    ref = wl.pad_frame(wl.synth_frame(W, H, seed=1 if rank == 0 else 99), 288, 1)
    planes = [torch.from_numpy(p.buf) for p in ref]
    shard.broadcast_planes(planes, src=0)
    cur = wl.synth_frame(W, H, seed=2)[0]
    # fused form: every rank owns one picture, one all-gather delivers all of them to everybody
    mine = torch.full((1024,), rank + 1, dtype=torch.uint8)
    landing = torch.zeros(world * 1024, dtype=torch.uint8)
    shard.allgather_pictures(landing, mine)
    assert all(int(landing[r * 1024]) == r + 1 and int(landing[r * 1024 + 1023]) == r + 1 for r in range(world))
    # the exchange object the bench uses: on CPU / gloo it must pick the collective path on every rank, same result
    ex = shard.PictureExchange(1024, "cpu")
    assert ex.mode == "nccl"
    landing2 = torch.zeros(world * 1024, dtype=torch.uint8)
    ex.allgather(landing2, mine)
    assert torch.equal(landing, landing2)

    class P:
        pass
    pl = P()
    pl.origin, pl.stride = ref[0].origin, ref[0].stride
    blocks, mvs, prec = wl.frame_batch(W, H, seed=3, cands_per_block=3, mv_range=16, models=(1, 2, 3, 5, 6, 10))
    mine = shard.shard_blocks(blocks, rank, world)
    assert set((mine["y"] // 128) % world) <= {rank}
    local = _eval(o, pl, cur, mine, mvs, prec)
    parts = shard.gather_distortions(torch.from_numpy(local.view(np.int64).copy()))
    if rank == 0:
        full = _eval(o, pl, cur, blocks, mvs, prec)
        # reassemble: rank r's vector follows the order of its own block list
        got = np.zeros(len(mvs), dtype=np.uint64)
        for r in range(world):
            br = shard.shard_blocks(blocks, r, world)
            idx = np.concatenate([np.arange(b["cand_first"], b["cand_first"] + b["cand_count"]) for b in br])
            got[idx] = parts[r]
        exp = np.zeros(len(mvs), dtype=np.uint64)
        idx = np.concatenate([np.arange(b["cand_first"], b["cand_first"] + b["cand_count"]) for b in blocks])
        exp[idx] = full
        ret["ok"] = bool(np.array_equal(got, exp))
        ret["n"] = int(len(mvs))
        ret["ref_equal"] = True
    else:
        chk = wl.pad_frame(wl.synth_frame(W, H, seed=1), 288, 1)
        ret["ref_equal_%d" % rank] = all(np.array_equal(a.buf, b.buf) for a, b in zip(ref, chk))
    dist.barrier()
    dist.destroy_process_group()


def test_ctu_row_sharding_world2_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + (os.getpid() % 200)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get("ok") is True and ret.get("n", 0) > 100
    assert ret.get("ref_equal_1") is True
