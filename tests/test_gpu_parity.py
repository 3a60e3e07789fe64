"""Parity tests proper: the sm_100a path, called through the C ABI (ctypes), against the CPU oracle on
the same seeded inputs.  Bit-exact everywhere (integer positions, samples, distortions).
Reference lines each case follows are quoted in oracle/mm_oracle.c next to the restated function."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import vvc_extension_mm_b200 as mm
from vvc_extension_mm_b200 import workload as wl

pytestmark = pytest.mark.gpu

W, H = 2048, 1024
SIZES = [(8, 4), (4, 8), (8, 8), (16, 8), (8, 16), (16, 16), (32, 8), (4, 32), (32, 32), (64, 16), (16, 64), (64, 64),
         (128, 64), (64, 128), (128, 128), (16, 4), (128, 4), (4, 128)]
EPIPOLE = [0.3, -0.5, 0.81]


@pytest.fixture(scope="module")
def env(oracle_2k, gpu_2k):
    o, g = oracle_2k, gpu_2k
    R = ol.Oracle.rodrigues(EPIPOLE)
    Rg = mm.rodrigues(EPIPOLE)
    assert np.array_equal(R.view(np.uint32), Rg.view(np.uint32)), "host-side setEpipole differs from oracle"
    o.set_rotation(ol.GED_CAMPOSE, R)
    g.set_rotation(0, Rg)
    ref = wl.pad_frame(wl.synth_frame(W, H, seed=11), 288, 1)
    cur = wl.synth_frame(W, H, seed=12, shift=(3.0, 1.0))
    g.set_ref(0, 0, [p.origin for p in ref], [p.stride for p in ref])
    # the oracle uses its own PaddedPlane type; wrap the same memory
    class P:  # noqa
        pass
    planes = []
    for p in ref:
        q = P(); q.origin = p.origin; q.stride = p.stride; planes.append(q)
    return o, g, planes, ref, cur


def rand_block(rng, w, h):
    x = int(rng.integers(0, (W - w) // 4 + 1)) * 4
    y = int(rng.integers(0, (H - h) // 4 + 1)) * 4
    return x, y


def test_mpa_cache_matches(env):
    """MotionPlaneAdaptiveMotionModel::fillCache on the device == oracle: checked through reproject of
    every 128x128 block of the picture with mv = 0 (reads the whole cache)."""
    o, g, *_ = env
    for model in (ol.MPA_FB, ol.MPA_LR, ol.MPA_TB):
        for y in range(0, H, 128):
            for x in range(0, W, 256):
                ox, oy = o.reproject((x, y), (128, 128), (0, 0), model)
                gx, gy = g.reproject((x, y), (128, 128), (0, 0), model)
                assert np.array_equal(ox, gx) and np.array_equal(oy, gy), (model, x, y)


@pytest.mark.parametrize("model", list(range(1, 11)))
def test_reproject_luma_chroma(env, model):
    o, g, *_ = env
    rng = np.random.default_rng(100 + model)
    n = 0
    for it in range(120):
        w, h = SIZES[it % len(SIZES)]
        x, y = rand_block(rng, w, h)
        if it % 9 == 0:
            y = 0 if it % 2 else H - h          # poles
        if it % 10 == 0:
            x = 0 if it % 4 else W - w          # ERP seam
        mv = (int(rng.integers(-800, 800)), int(rng.integers(-400, 400)))
        if it % 11 == 0: mv = (0, 0)
        if it % 13 == 0: mv = (mv[0], 0)
        if it % 17 == 0: mv = (0, mv[1])
        ox, oy = o.reproject((x, y), (w, h), mv, model)
        gx, gy = g.reproject((x, y), (w, h), mv, model)
        assert np.array_equal(ox, gx) and np.array_equal(oy, gy), ("luma", model, x, y, w, h, mv)
        if w >= 8 or h >= 8:
            cw, ch = max(w // 2, 2), max(h // 2, 2)
            ox, oy = o.reproject((x // 2, y // 2), (cw, ch), mv, model, comp=1, chroma_fmt=1)
            gx, gy = g.reproject((x // 2, y // 2), (cw, ch), mv, model, comp=1, chroma_fmt=1)
            assert np.array_equal(ox, gx) and np.array_equal(oy, gy), ("chroma", model, x, y, w, h, mv)
        n += len(ox)
    assert n > 0


def test_reproject_zero_epipole_nan_path(env):
    """Quirk Q4: default (0,0,0) epipole -> NaN rotation -> 'no motion' via isNaN().select."""
    o, g, *_ = env
    Rn = ol.Oracle.rodrigues([0, 0, 0])
    o.set_rotation(ol.GED_CAMPOSE, Rn)
    try:
        ox, oy = o.reproject((512, 256), (32, 32), (37, -21), ol.GED_CAMPOSE)
        gx, gy = g.reproject((512, 256), (32, 32), (37, -21), mm.GED_CAMPOSE, rot_slot=5)  # slot 5 still default
        assert np.array_equal(ox, gx) and np.array_equal(oy, gy)
        exp_x = (np.arange(8) * 4 + 512) * 16
        assert np.array_equal(gx.reshape(8, 8, order="F")[0], exp_x)
    finally:
        o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues(EPIPOLE))


@pytest.mark.parametrize("model", [0, 1, 2, 3, 4, 5, 6, 8, 10])
@pytest.mark.parametrize("bi", [False, True])
def test_mc_block_all_components(env, model, bi):
    """xPredInterBlkMM / xPredInterBlk: Y, Cb, Cr; uni (clipped) and bi (14-bit intermediate)."""
    o, g, planes, ref, cur = env
    rng = np.random.default_rng(200 + model + (50 if bi else 0))
    for it in range(24):
        w, h = SIZES[it % len(SIZES)]
        x, y = rand_block(rng, w, h)
        mv = (int(rng.integers(-500, 500)), int(rng.integers(-300, 300)))
        if it % 6 == 0: mv = ((mv[0] // 16) * 16, (mv[1] // 16) * 16)   # integer MV -> copy / single-pass cases
        if it % 7 == 0: mv = ((mv[0] // 16) * 16, mv[1])
        if it % 8 == 0: mv = (mv[0], (mv[1] // 16) * 16)
        if model == 0:
            x = min(max(x, 64), W - w - 64); y = min(max(y, 64), H - h - 64)
            a = o.mc_block_classic(planes[0], 0, 1, (x, y), (w, h), mv, bi)
        else:
            a = o.mc_block(planes[0], 0, 1, (x, y), (w, h), mv, model, bi)
        b = g.mc(0, 0, 1, (x, y), (w, h), mv, model, bi=bi)
        assert np.array_equal(a, b), ("Y", model, bi, x, y, w, h, mv)
        if w >= 8 and h >= 8:
            for comp in (1, 2):
                cp, cs = (x // 2, y // 2), (w // 2, h // 2)
                if model == 0:
                    a = o.mc_block_classic(planes[comp], comp, 1, cp, cs, mv, bi)
                else:
                    a = o.mc_block(planes[comp], comp, 1, cp, cs, mv, model, bi)
                b = g.mc(0, comp, 1, cp, cs, mv, model, bi=bi)
                assert np.array_equal(a, b), ("C", comp, model, bi, x, y, w, h, mv)


def test_mc_oob_zero_fill_both_rules(env):
    """Quirk Q2: MC uses +-maxCU slack, ME uses none; OOB sub-blocks are zero-filled."""
    o, g, planes, ref, cur = env
    seen_zero = False
    for oob in (ol.OOB_PRED, ol.OOB_ME):
        for (x, y, mv) in [(0, 0, (-40 * 16, -30 * 16)), (W - 64, H - 64, (90 * 16, 70 * 16)), (0, 512, (-200 * 16, 0)),
                           (W - 32, 0, (150 * 16 + 5, -140 * 16 - 3))]:
            for model in (ol.T3D, ol.MPA_LR, ol.ROT):
                a = o.mc_block(planes[0], 0, 1, (x, y), (32, 32), mv, model, False, oob=oob)
                b = g.mc(0, 0, 1, (x, y), (32, 32), mv, model, oob=oob)
                assert np.array_equal(a, b), (oob, x, y, mv, model)
                seen_zero |= bool((a == 0).all(axis=None) or (a[:4, :4] == 0).all())
    assert seen_zero


def _blocks_for(sizes, rng, model_cycle, cands, ref_slot=0):
    blocks = np.zeros(len(sizes), dtype=mm.api.BLOCK_DTYPE)
    mvs = np.zeros(len(sizes) * cands, dtype=mm.api.MV_DTYPE)
    for i, (w, h) in enumerate(sizes):
        x, y = rand_block(rng, w, h)
        blocks[i] = (x, y, w, h, model_cycle[i % len(model_cycle)], ref_slot, 0, i * cands, cands, x, y, 0)
    return blocks, mvs


@pytest.mark.parametrize("dist,sub_shift", [(ol.DIST_SAD, 0), (ol.DIST_SAD, 1), (ol.DIST_HAD, 0)])
@pytest.mark.parametrize("prec", [4, 2])
def test_me_batch_matches_oracle(env, dist, sub_shift, prec):
    """xMVReprojectionInterpolation + distFunc for batches of candidates, every block size and model."""
    o, g, planes, ref, cur = env
    rng = np.random.default_rng(300 + dist * 10 + sub_shift + prec)
    g.set_org(cur[0])
    sizes = SIZES * 2
    models = [1, 2, 3, 4, 5, 6, 10, 7, 9, 0]
    cands = 9
    blocks, mvs = _blocks_for(sizes, rng, models, cands)
    scale = 1 if prec == 4 else 4
    mvs["hor"] = rng.integers(-48 * scale, 48 * scale, size=len(mvs))
    mvs["ver"] = rng.integers(-32 * scale, 32 * scale, size=len(mvs))
    mvs["hor"][::9] = 0; mvs["ver"][::9] = 0
    for b in blocks:  # classic candidates must stay inside the padded plane
        if b["model"] == 0:
            b["x"] = b["org_x"] = min(max(b["x"], 64), W - b["w"] - 64)
            b["y"] = b["org_y"] = min(max(b["y"], 64), H - b["h"] - 64)
    got = g.me_batch(blocks, mvs, mv_prec_shift=prec, dist=dist, sub_shift=sub_shift)
    for b in blocks:
        sl = slice(b["cand_first"], b["cand_first"] + b["cand_count"])
        org = cur[0][b["y"]:b["y"] + b["h"], b["x"]:b["x"] + b["w"]]
        m = np.stack([mvs["hor"][sl], mvs["ver"][sl]], axis=1)
        exp = o.me_block(planes[0], org, None, (int(b["x"]), int(b["y"])), (int(b["w"]), int(b["h"])), int(b["model"]), m,
                         mv_prec_shift=prec, dist=dist, sub_shift=sub_shift, had_mode=ol.HAD_SIMD)
        assert np.array_equal(exp, got[sl]), (dict(zip(b.dtype.names, b.tolist())), exp, got[sl])


def test_me_batch_candidate_counts_of_the_search_loops(env):
    """Candidate-set sizes the reference's loops produce: 1-2 (AMVP template), 9 (half/quarter
    refinement), 16 (diamond round), 25 (DMVR), 64, 81 (bi-pred full search) -- mixed in ONE batch with ragged counts per
    block, small and large blocks, so that work items of different blocks and candidates share warps and the k_map item
    walk crosses block / chunk boundaries at every offset."""
    o, g, planes, ref, cur = env
    rng = np.random.default_rng(4242)
    g.set_org(cur[0])
    counts = [1, 2, 9, 16, 25, 64, 81, 3, 7, 33]
    sizes = [(8, 4), (4, 8), (8, 8), (16, 16), (32, 32), (64, 64), (128, 128), (16, 8), (64, 16), (4, 128)]
    models = [1, 6, 10, 2, 4, 5, 3, 9, 1, 6]
    n = len(sizes)
    blocks = np.zeros(n, dtype=mm.api.BLOCK_DTYPE)
    first = 0
    for i, (w, h) in enumerate(sizes):
        x, y = rand_block(rng, w, h)
        blocks[i] = (x, y, w, h, models[i], 0, 0, first, counts[i], x, y, 0)
        first += counts[i]
    mvs = np.zeros(first, dtype=mm.api.MV_DTYPE)
    mvs["hor"] = rng.integers(-160, 160, size=first)
    mvs["ver"] = rng.integers(-120, 120, size=first)
    for dist in (ol.DIST_SAD, ol.DIST_HAD):
        got = g.me_batch(blocks, mvs, mv_prec_shift=2, dist=dist)
        for b in blocks:
            sl = slice(b["cand_first"], b["cand_first"] + b["cand_count"])
            org = cur[0][b["y"]:b["y"] + b["h"], b["x"]:b["x"] + b["w"]]
            m = np.stack([mvs["hor"][sl], mvs["ver"][sl]], axis=1)
            exp = o.me_block(planes[0], org, None, (int(b["x"]), int(b["y"])), (int(b["w"]), int(b["h"])), int(b["model"]), m,
                             mv_prec_shift=2, dist=dist, had_mode=ol.HAD_SIMD)
            assert np.array_equal(exp, got[sl]), (dict(zip(b.dtype.names, b.tolist())), dist)


def test_me_batch_scalar_had_mode(oracle_2k):
    """--SIMD=SCALAR hosts use (int)(sad / sqrt(16*8) * 2); selectable per context."""
    o = oracle_2k
    g = mm.MmGpu(W, H, max_refs=1, had_mode=mm.HAD_SCALAR)
    try:
        ref = wl.pad_frame(wl.synth_frame(W, H, seed=21), 288, 1)
        cur = wl.synth_frame(W, H, seed=22)
        g.set_ref(0, 0, [p.origin for p in ref], [p.stride for p in ref])
        g.set_org(cur[0])
        rng = np.random.default_rng(5)
        sizes = [(16, 8), (8, 16), (32, 8), (8, 4), (4, 8), (64, 16), (16, 4)]
        blocks, mvs = _blocks_for(sizes, rng, [ol.T3D, ol.MPA_FB], 4)
        mvs["hor"] = rng.integers(-30, 30, size=len(mvs)); mvs["ver"] = rng.integers(-30, 30, size=len(mvs))
        got = g.me_batch(blocks, mvs, dist=mm.DIST_HAD)
        class P: pass
        pl = P(); pl.origin = ref[0].origin; pl.stride = ref[0].stride
        for b in blocks:
            sl = slice(b["cand_first"], b["cand_first"] + b["cand_count"])
            org = cur[0][b["y"]:b["y"] + b["h"], b["x"]:b["x"] + b["w"]]
            m = np.stack([mvs["hor"][sl], mvs["ver"][sl]], axis=1)
            exp = o.me_block(pl, org, None, (int(b["x"]), int(b["y"])), (int(b["w"]), int(b["h"])), int(b["model"]), m,
                             dist=ol.DIST_HAD, had_mode=ol.HAD_SCALAR)
            assert np.array_equal(exp, got[sl])
    finally:
        g.close()


def test_unpadded_upload_extends_border_like_reference(env):
    """Picture::extendPicBorder on the device: MC through slot 1 (device-padded) == slot 0 (host-padded)."""
    o, g, planes, ref, cur = env
    g.set_ref_unpadded(1, 0, [p.inner.copy() for p in ref])
    rng = np.random.default_rng(9)
    for (x, y) in [(0, 0), (W - 64, 0), (0, H - 64), (W - 64, H - 64), (1024, 512)]:
        mv = (int(rng.integers(-1500, 1500)), int(rng.integers(-1500, 1500)))
        for comp, pos, size in ((0, (x, y), (64, 64)), (1, (x // 2, y // 2), (32, 32)), (2, (x // 2, y // 2), (32, 32))):
            a = g.mc(0, comp, 1, pos, size, mv, mm.T3D)
            b = g.mc(1, comp, 1, pos, size, mv, mm.T3D)
            assert np.array_equal(a, b)


def test_unpadded_wrap_upload_equals_extend_wrap_border(env):
    """Picture::extendWrapBorder on the device (PIC_RECON_WRAP of a 360 ERP sequence): MC through a device-wrapped slot ==
    oracle MC on a plane padded by the oracle's restatement, for blocks whose windows cross the left / right picture
    border (where the two paddings differ) and for one in the middle; full and partial wrap offsets."""
    o, g, planes, ref, cur = env
    inner = [p.inner.copy() for p in ref]
    rng = np.random.default_rng(19)
    for xoff in (W, 96):
        g.set_ref_unpadded_wrap(1, 0, inner, xoff)
        wrapped = []
        for c, p in enumerate(ref):
            q = ol.PaddedPlane(p.w, p.h, p.mx, p.my)
            q.inner[:] = inner[c]
            q.extend_wrap(xoff >> (1 if c else 0))
            wrapped.append(q)
        differs = False
        for (x, y) in [(0, 0), (W - 64, 0), (0, H - 64), (W - 64, H - 64), (1024, 512), (0, 384), (W - 64, 640)]:
            for mv in ((int(rng.integers(-1800, 1800)), int(rng.integers(-600, 600))), (-2000, 33), (1999, -70)):
                for comp, pos, size in ((0, (x, y), (64, 64)), (1, (x // 2, y // 2), (32, 32)), (2, (x // 2, y // 2), (32, 32))):
                    exp = o.mc_block_classic(wrapped[comp], comp, 1, pos, size, mv, False)
                    got = g.mc(1, comp, 1, pos, size, mv, 0)
                    assert np.array_equal(exp, got), (xoff, x, y, mv, comp)
                    differs |= not np.array_equal(got, g.mc(0, comp, 1, pos, size, mv, 0))
        assert differs      # the wrapped margins are not the replicated ones


def test_add_avg(env):
    o, g, *_ = env
    rng = np.random.default_rng(3)
    for (w, h) in [(8, 4), (16, 16), (128, 64)]:
        a = rng.integers(-8192, 8191, size=(h, w)).astype(np.int16)
        b = rng.integers(-8192, 8191, size=(h, w)).astype(np.int16)
        exp = np.zeros((h, w), dtype=np.int16)
        ol.lib().orc_add_avg(a.ctypes.data, w, b.ctypes.data, w, exp.ctypes.data, w, w, h, 10, 0, 1023)
        assert np.array_equal(exp, g.add_avg(a, b))


def test_bi_pred_pipeline(env):
    """xPredInterBi shape: two MM MCs at bi=true + xWeightedAverage == oracle."""
    o, g, planes, ref, cur = env
    for model, mv0, mv1 in [(ol.MPA_TB, (35, -18), (-77, 40)), (ol.ROT, (100, 3), (-9, -250))]:
        a0 = o.mc_block(planes[0], 0, 1, (640, 384), (64, 32), mv0, model, True)
        a1 = o.mc_block(planes[0], 0, 1, (640, 384), (64, 32), mv1, model, True)
        exp = np.zeros((32, 64), dtype=np.int16)
        ol.lib().orc_add_avg(a0.ctypes.data, 64, a1.ctypes.data, 64, exp.ctypes.data, 64, 64, 32, 10, 0, 1023)
        b0 = g.mc(0, 0, 1, (640, 384), (64, 32), mv0, model, bi=True)
        b1 = g.mc(0, 0, 1, (640, 384), (64, 32), mv1, model, bi=True)
        assert np.array_equal(exp, g.add_avg(b0, b1))


def test_full_frame_properties(env):
    """BASELINE-size batch (whole 2048x1024 frame): size-independent properties.
    (1) distortion of MC(frame, mv=0, identity-at-zero models) against the reference frame itself is 0;
    (2) SAD(subShift=0) of the fast batch path == SAD recomputed from predictions of the MC batch path;
    (3) candidate order does not matter (checksum of distortions invariant under permutation)."""
    import torch
    o, g, planes, ref, cur = env
    blocks, mvs, prec = wl.frame_batch(W, H, seed=5, cands_per_block=4, mv_range=32, models=(mm.T3D, mm.ROT, mm.TAN))
    # (1) org := reference picture, all-zero MVs -> identity models give exact copies -> distortion 0
    g.set_org(np.ascontiguousarray(ref[0].inner))
    zero = np.zeros(len(mvs), dtype=mm.api.MV_DTYPE)
    for dist in (mm.DIST_SAD, mm.DIST_HAD):
        d = g.me_batch(blocks, zero, mv_prec_shift=prec, dist=dist)
        # ME rule (InterSearch.cpp:6332): sub-blocks with xPos >= W-4 or yPos >= H-4 are zero-filled (Q2)
        inner = (blocks["x"] + blocks["w"] < W) & (blocks["y"] + blocks["h"] < H)
        per_block = d.reshape(len(blocks), -1)
        assert int(per_block[inner].sum()) == 0
        assert int(per_block[~inner].sum()) > 0
    # (2)+(3)
    g.set_org(cur[0])
    d1 = g.me_batch(blocks, mvs, mv_prec_shift=prec, dist=mm.DIST_SAD)
    one = blocks.copy(); one["cand_count"] = 1; one["cand_first"] = np.arange(len(one))
    mv1 = mvs[blocks["cand_first"]]
    plan = g.plan(one, len(one))
    dmv = torch.from_numpy(mv1.view(np.int32).reshape(-1, 2).copy()).cuda()
    dst = torch.zeros((H, W), dtype=torch.int16, device="cuda")
    g.mc_batch_device(plan, dmv.data_ptr(), dst.data_ptr(), W, mv_prec_shift=prec)
    g.sync()
    pred = dst.cpu().numpy()
    # the MC batch uses the +-maxCU rule, the ME batch the strict rule: compare only where both agree (inside)
    sad = np.abs(cur[0].astype(np.int64) - pred.astype(np.int64))
    chk = 0
    for b in blocks[:400]:
        s = int(sad[b["y"]:b["y"] + b["h"], b["x"]:b["x"] + b["w"]].sum())
        if 160 <= b["x"] <= W - 300 and 160 <= b["y"] <= H - 300:
            assert s == int(d1[b["cand_first"]]), b
            chk += 1
    assert chk > 20
    perm = np.random.default_rng(1).permutation(len(blocks))
    pb = blocks[perm]
    d2 = g.me_batch(pb, mvs, mv_prec_shift=prec, dist=mm.DIST_SAD)
    assert np.array_equal(d1, d2)
    plan.close()


def test_pipelined_submit_wait_equals_synchronous_call(env):
    """mmgpu_me_batch_submit / _wait (two batches in flight on two slots, copies on their own stream) return exactly
    what the synchronous mmgpu_set_org + mmgpu_me_batch pair returns; slot misuse is reported, not ignored."""
    import torch
    o, g, planes, ref, cur = env
    orgs, batches, expect = [], [], []
    for i in range(4):
        blocks, mvs, prec = wl.frame_batch(W, H, seed=40 + i, cands_per_block=4, mv_range=24,
                                           models=(mm.MPA_FB, mm.ROT, mm.GED_CAMPOSE, mm.TAN, mm.T3D))
        org = wl.synth_plane(W, H, 900 + i, shift=(1.0 + i, 0.5))
        g.set_org(org)
        expect.append(g.me_batch(blocks, mvs, mv_prec_shift=prec, dist=mm.DIST_HAD).copy())
        pin = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).pin_memory().numpy()
        orgs.append(pin(org).view(np.int16).reshape(H, W))
        batches.append((pin(blocks).view(mm.api.BLOCK_DTYPE), pin(mvs).view(mm.api.MV_DTYPE),
                        pin(np.zeros(len(mvs), dtype=np.uint64)).view(np.uint64), prec))
    for i in range(4):
        b, m, out, prec = batches[i]
        if i >= 2:
            got = g.me_batch_wait(i % 2)
            assert np.array_equal(got, expect[i - 2]), i - 2
        g.me_batch_submit(i % 2, orgs[i], b, m, out, mv_prec_shift=prec, dist=mm.DIST_HAD)
    with pytest.raises(mm.MmGpuError):
        g.me_batch_submit(0, orgs[0], *batches[0][:3])      # slot 0 is still in flight
    for i in (2, 3):
        assert np.array_equal(g.me_batch_wait(i % 2), expect[i]), i
    with pytest.raises(mm.MmGpuError):
        g.me_batch_wait(0)                                    # nothing pending


def test_tma_boxes_equal_gather_path(env):
    """k_fir reads the reference through TMA-filled shared-memory boxes when the slot has descriptors and gathers from
    global memory otherwise (MMGPU_NO_TMA at registration time; also what unaligned device planes get).  Both paths,
    SAD and HAD, all models incl. poles / ERP seam / picture borders, give identical distortions; a sample is checked
    against the oracle."""
    import os
    o, g, planes, ref, cur = env
    g.set_ref(1, 0, [p.origin for p in ref], [p.stride for p in ref])           # descriptors
    os.environ["MMGPU_NO_TMA"] = "1"
    try:
        g.set_ref(2, 0, [p.origin for p in ref], [p.stride for p in ref])       # gather only
    finally:
        del os.environ["MMGPU_NO_TMA"]
    g.set_org(cur[0])
    blocks, mvs, prec = wl.frame_batch(W, H, seed=61, cands_per_block=3, mv_range=40,
                                       models=(mm.MPA_FB, mm.MPA_LR, mm.MPA_TB, mm.TAN, mm.T3D, mm.ROT, mm.GED_CAMPOSE, 0))
    res = {}
    for slot in (1, 2):
        b = blocks.copy(); b["ref_slot"] = slot
        for dist in (mm.DIST_SAD, mm.DIST_HAD):
            res[(slot, dist)] = g.me_batch(b, mvs, mv_prec_shift=prec, dist=dist).copy()
    for dist in (mm.DIST_SAD, mm.DIST_HAD):
        assert np.array_equal(res[(1, dist)], res[(2, dist)]), dist
    rng = np.random.default_rng(3)
    for bi in rng.choice(len(blocks), 60, replace=False):
        b = blocks[bi]
        if b["model"] == 0:
            continue
        sl = slice(b["cand_first"], b["cand_first"] + b["cand_count"])
        org = cur[0][b["y"]:b["y"] + b["h"], b["x"]:b["x"] + b["w"]]
        m = np.stack([mvs["hor"][sl], mvs["ver"][sl]], axis=1)
        exp = o.me_block(planes[0], org, None, (int(b["x"]), int(b["y"])), (int(b["w"]), int(b["h"])), int(b["model"]), m,
                         mv_prec_shift=prec, dist=mm.DIST_HAD, had_mode=ol.HAD_SIMD)
        assert np.array_equal(exp, res[(1, mm.DIST_HAD)][sl]), b


def test_reregistered_device_plane_gets_fresh_tma_descriptor(env):
    """A slot re-registered with another device buffer (what an encoder does for every new reference picture) must be
    read through its NEW descriptor: evaluate picture A, re-point the slot at picture B in other memory, evaluate again
    and compare with the gather path on B; then back to A."""
    import os
    import torch
    o, g, planes, ref, cur = env
    pics = []
    for seed in (71, 72):
        pl = wl.pad_frame(wl.synth_frame(W, H, seed=seed), 288, 1)
        dev = [torch.from_numpy(p.buf.copy()).cuda() for p in pl]
        pics.append((pl, dev))
    def register(slot, k, tma):
        pl, dev = pics[k]
        if not tma:
            os.environ["MMGPU_NO_TMA"] = "1"
        try:
            g.set_ref_device(slot, k, [d.data_ptr() + 2 * (p.my * p.stride + p.mx) for d, p in zip(dev, pl)],
                             [p.stride for p in pl], keep=dev)
        finally:
            os.environ.pop("MMGPU_NO_TMA", None)
    g.set_org(cur[0])
    blocks, mvs, prec = wl.frame_batch(W, H, seed=62, cands_per_block=2, mv_range=32, models=(mm.MPA_LR, mm.ROT, mm.GED_CAMPOSE))
    b1 = blocks.copy(); b1["ref_slot"] = 1
    b2 = blocks.copy(); b2["ref_slot"] = 2
    res = []
    for k in (0, 1, 0):
        register(1, k, True)
        register(2, k, False)
        a = g.me_batch(b1, mvs, mv_prec_shift=prec, dist=mm.DIST_HAD).copy()
        b = g.me_batch(b2, mvs, mv_prec_shift=prec, dist=mm.DIST_HAD).copy()
        assert np.array_equal(a, b), k
        res.append(a)
    assert np.array_equal(res[0], res[2]) and not np.array_equal(res[0], res[1])


def test_unclipped_classic_mv_stays_inside_the_padded_plane(env):
    """Classic candidates have no out-of-picture rule (the reference clips MVs with clipMv before xPredInterBlk); an
    unclipped MV far outside must not make the batch kernels read outside the padded plane: the call returns, and the
    valid candidates of the same batch are still exact."""
    o, g, planes, ref, cur = env
    g.set_org(cur[0])
    blocks = np.zeros(2, dtype=mm.api.BLOCK_DTYPE)
    blocks[0] = (64, 64, 64, 64, 0, 0, 0, 0, 3, 64, 64, 0)
    blocks[1] = (W - 128, H - 128, 128, 128, 0, 0, 0, 3, 3, W - 128, H - 128, 0)
    mvs = np.zeros(6, dtype=mm.api.MV_DTYPE)
    mvs["hor"] = [5, -30000, 30000, -7, 20000, -20000]
    mvs["ver"] = [-3, 30000, -30000, 9, -20000, 20000]
    got = g.me_batch(blocks, mvs, mv_prec_shift=4, dist=mm.DIST_HAD)
    for bi, ci in ((0, 0), (1, 3)):
        b = blocks[bi]
        org = cur[0][b["y"]:b["y"] + b["h"], b["x"]:b["x"] + b["w"]]
        exp = o.me_block(planes[0], org, None, (int(b["x"]), int(b["y"])), (int(b["w"]), int(b["h"])), 0,
                         np.array([[mvs["hor"][ci], mvs["ver"][ci]]]), mv_prec_shift=4, dist=mm.DIST_HAD, had_mode=ol.HAD_SIMD)
        assert int(exp[0]) == int(got[ci])
    g.sync()


def test_cpp_host_adapter_matches_oracle():
    """The C++ adapter (reference-shaped argument lists) end to end: MC of a GED_CAMPOSE block and
    xPatternRefinementProjected's 9-point HAD refinement equal the oracle."""
    import subprocess
    from test_native_cpu import _build_adapter
    exe = _build_adapter()
    out = subprocess.run([exe, "need-gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = dict((l.split()[0], l.split()[1:]) for l in out.stdout.strip().splitlines())
    Wd, Hd, M = 512, 256, 288
    o = ol.Oracle(Wd, Hd)
    o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues([1, 0, 0]))
    planes = []
    for c in range(3):
        w, h, m = (Wd, Hd, M) if c == 0 else (Wd // 2, Hd // 2, M // 2)
        xx, yy = np.meshgrid(np.arange(w), np.arange(h))
        p = ((xx * 7 + yy * 13 + (xx * yy) % 31 + c * 101) & 1023).astype(np.int16)
        planes.append(wl.Padded(p, m, m, align=1))
    class P: pass
    pl = P(); pl.origin = planes[0].origin; pl.stride = planes[0].stride
    a = o.mc_block(pl, 0, 1, (128, 64), (64, 64), (37, -21), ol.GED_CAMPOSE)
    s = 0
    for v in a.reshape(-1).astype(np.uint16).tolist():
        s = (s * 1000003 + v) & 0xFFFFFFFFFFFFFFFF
    assert int(lines["mc"][0]) == s
    org = ((np.arange(32 * 32) * 37) & 1023).astype(np.int16).reshape(32, 32)
    refine = [(0, 0), (0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (1, -1), (-1, 1), (1, 1)]
    cand = np.array([(6 + dx, -4 + dy) for dx, dy in refine])
    d = o.me_block(pl, org, None, (64, 96), (32, 32), ol.MPA_LR, cand, mv_prec_shift=3, dist=ol.DIST_HAD)
    best, bmv = None, None
    for (mh, mv), dd in zip(cand, d):
        u = int(dd) + abs(int(mh)) + abs(int(mv))
        if best is None or u < best:
            best, bmv = u, (int(mh), int(mv))
    assert [int(x) for x in lines["refine"]] == [best, bmv[0], bmv[1]]
    o.close()


def test_dmvr_projected_costs_and_decision(env):
    """a12: the 25 xDMVRCost values of xProcessDMVRProjected per sub-PU (two bi MCs + SAD subShift 1, >>1) and the
    resulting mvdL0SubPu == oracle; list-specific geodesic rotations and reference pictures."""
    o, g, planes, ref, cur = env
    ref1 = wl.pad_frame(wl.synth_frame(W, H, seed=31, shift=(1.5, -0.5)), 288, 1)
    g.set_ref(2, 8, [p.origin for p in ref1], [p.stride for p in ref1])
    class P: pass
    p1 = P(); p1.origin = ref1[0].origin; p1.stride = ref1[0].stride
    R0 = ol.Oracle.rodrigues(EPIPOLE); R1 = ol.Oracle.rodrigues([-0.2, 0.9, 0.1])
    g.set_rotation(0, mm.rodrigues(EPIPOLE)); g.set_rotation(1, mm.rodrigues([-0.2, 0.9, 0.1]))
    rng = np.random.default_rng(77)
    sub = np.zeros(30, dtype=mm.DMVR_DTYPE)
    sizes = [(16, 16), (8, 16), (16, 8), (8, 8)]
    models = [1, 2, 3, 4, 5, 6, 10, 8]
    for i in range(len(sub)):
        w, h = sizes[i % 4]
        x = int(rng.integers(4, (W - w) // 4 - 4)) * 4; y = int(rng.integers(4, (H - h) // 4 - 4)) * 4
        if i % 7 == 0: y = 0
        if i % 9 == 0: x = W - w
        mv0 = (int(rng.integers(-200, 200)), int(rng.integers(-120, 120)))
        sub[i] = (x, y, w, h, models[i % len(models)], 0, 2, 0, 1, mv0[0], mv0[1], -mv0[0] + int(rng.integers(-8, 8)), -mv0[1])
    got = g.dmvr_costs(sub)
    for i, s in enumerate(sub):
        campose = int(s["model"]) == ol.GED_CAMPOSE     # only the camera-pose model takes per-list epipoles
        exp = ol.dmvr_costs(o, planes[0], p1, (int(s["x"]), int(s["y"])), (int(s["w"]), int(s["h"])),
                            (int(s["mv0_hor"]), int(s["mv0_ver"])), (int(s["mv1_hor"]), int(s["mv1_ver"])), int(s["model"]),
                            R0 if campose else None, R1 if campose else None)
        assert np.array_equal(exp, got[i]), (i, s, exp, got[i])
        assert mm.dmvr_decide(got[i], int(s["w"]), int(s["h"])) == ol.dmvr_decide(exp, int(s["w"]), int(s["h"]))
    o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues(EPIPOLE))


def test_gpu_positions_equal_reference_source_directly(env):
    """The sm_100a mapping against the reference's OWN MVReprojection/MotionModels source (oracle/_ref/libvtmmap.so,
    compiled against the Eigen stand-in) without the oracle in between."""
    import ref_lib as rl
    if not rl.map_available():
        pytest.skip("oracle/_ref/libvtmmap.so not present")
    o, g, *_ = env
    ref = rl.RefMap(W, H, 1.0, 1)
    ref.set_epipole(EPIPOLE)
    g.set_rotation(0, mm.rodrigues(EPIPOLE))
    rng = np.random.default_rng(4242)
    n = 0
    for it in range(600):
        w, h = SIZES[it % len(SIZES)]
        x, y = rand_block(rng, w, h)
        if it % 8 == 0: y = 0 if it % 16 else H - h
        mv = (int(rng.integers(-1000, 1000)), int(rng.integers(-500, 500)))
        if it % 10 == 0: mv = (0, 0)
        m = 1 + it % 10
        ax, ay = ref.reproject((x, y), (w, h), mv, m)
        bx, by = g.reproject((x, y), (w, h), mv, m)
        assert np.array_equal(ax, bx) and np.array_equal(ay, by), (m, x, y, w, h, mv)
        if w >= 8 and h >= 8:
            ax, ay = ref.reproject((x // 2, y // 2), (w // 2, h // 2), mv, m, comp=1, chroma_fmt=1)
            bx, by = g.reproject((x // 2, y // 2), (w // 2, h // 2), mv, m, comp=1, chroma_fmt=1)
            assert np.array_equal(ax, bx) and np.array_equal(ay, by), ("chroma", m, x, y, w, h, mv)
        n += len(ax)
    assert n > 10000


def test_decoder_shape_full_picture_md5(env):
    """BASELINE config 4 shape: reconstruct a whole 2048x1024 4:2:0 picture from per-PU (model, MVs, uni/bi) decisions --
    DecCu::xReconInter -> motionCompensation -> xPredInterUni/xPredInterBi -> xPredInterBlkMM/xPredInterBlk + addAvg --
    on the GPU in one mmgpu_mc_jobs call and on the oracle; the decoded-picture MD5s (PicYuvMD5 style: per plane) match."""
    import hashlib
    o, g, planes, ref, cur = env
    ref1 = wl.pad_frame(wl.synth_frame(W, H, seed=41, shift=(-1.25, 0.75)), 288, 1)
    g.set_ref(3, 16, [p.origin for p in ref1], [p.stride for p in ref1])
    class P: pass
    planes1 = []
    for p in ref1:
        q = P(); q.origin = p.origin; q.stride = p.stride; planes1.append(q)
    Ra = ol.Oracle.rodrigues(EPIPOLE); Rb = ol.Oracle.rodrigues([0.1, 0.2, -0.97])
    g.set_rotation(0, mm.rodrigues(EPIPOLE)); g.set_rotation(2, mm.rodrigues([0.1, 0.2, -0.97]))
    rng = np.random.default_rng(2026)
    part = wl.frame_partition(W, H, seed=77)
    jobs = []
    exp = [np.zeros((H, W), np.int16), np.zeros((H // 2, W // 2), np.int16), np.zeros((H // 2, W // 2), np.int16)]
    for (x, y, w, h) in part:
        model = int(rng.integers(0, 11))
        bi = bool(rng.random() < 0.3) and (w * h > 32)          # 8x4 / 4x8 PUs are uni-predicted in VVC
        mv0 = (int(rng.integers(-700, 700)), int(rng.integers(-400, 400)))
        mv1 = (int(rng.integers(-700, 700)), int(rng.integers(-400, 400)))
        if rng.random() < 0.1: mv0 = ((mv0[0] >> 4) << 4, (mv0[1] >> 4) << 4)
        for comp in range(3):
            cx, cy, cw, ch = (x, y, w, h) if comp == 0 else (x // 2, y // 2, w // 2, h // 2)
            jobs.append((comp, cx, cy, cw, ch, model, int(bi), 0, 0, mv0[0], mv0[1], 3, 2, mv1[0], mv1[1]))
            def mc(pl, mv, R, want_bi):
                if model == 0:
                    return o.mc_block_classic(pl[comp], comp, 1, (cx, cy), (cw, ch), mv, want_bi)
                if model == ol.GED_CAMPOSE:
                    o.set_rotation(ol.GED_CAMPOSE, R)
                return o.mc_block(pl[comp], comp, 1, (cx, cy), (cw, ch), mv, model, want_bi)
            a = mc(planes, mv0, Ra, bi)
            if bi:
                b = mc(planes1, mv1, Rb, True)
                out = np.zeros_like(a)
                ol.lib().orc_add_avg(a.ctypes.data, cw, b.ctypes.data, cw, out.ctypes.data, cw, cw, ch, 10, 0, 1023)
                a = out
            exp[comp][cy:cy + ch, cx:cx + cw] = a
    o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues(EPIPOLE))
    got = [np.full((H, W), -1, np.int16), np.full((H // 2, W // 2), -1, np.int16), np.full((H // 2, W // 2), -1, np.int16)]
    g.mc_jobs(np.array(jobs, dtype=mm.JOB_DTYPE), got)
    for c in range(3):
        assert hashlib.md5(got[c].tobytes()).hexdigest() == hashlib.md5(exp[c].tobytes()).hexdigest(), \
            ("plane", c, int((got[c] != exp[c]).sum()))
    assert len(jobs) > 5000
    # the same reconstruction with the picture resident on the device (mmgpu_mc_jobs_device), and a partial job list on top
    # of a host picture that must keep the samples the jobs do not cover
    import torch
    dev = [torch.full((H, W), -1, dtype=torch.int16, device="cuda"), torch.full((H // 2, W // 2), -1, dtype=torch.int16, device="cuda"),
           torch.full((H // 2, W // 2), -1, dtype=torch.int16, device="cuda")]
    g.mc_jobs_device(np.array(jobs, dtype=mm.JOB_DTYPE), [t.data_ptr() for t in dev], [W, W // 2, W // 2])
    g.sync()
    for c in range(3):
        assert np.array_equal(dev[c].cpu().numpy(), exp[c]), c
    part_jobs = np.array(jobs[:300], dtype=mm.JOB_DTYPE)
    keep = [np.full((H, W), 77, np.int16), np.full((H // 2, W // 2), 78, np.int16), np.full((H // 2, W // 2), 79, np.int16)]
    g.mc_jobs(part_jobs, keep)
    for c in range(3):
        covered = np.zeros(keep[c].shape, dtype=bool)
        for j in part_jobs[part_jobs["comp"] == c]:
            covered[j["y"]:j["y"] + j["h"], j["x"]:j["x"] + j["w"]] = True
        assert np.array_equal(keep[c][covered], exp[c][covered]) and (keep[c][~covered] == 77 + c).all(), c


def test_config3_all_five_models_4096x2048():
    """BASELINE config 3 shape: 4096x2048 10-bit ERP, low-delay-P search range (+-64), all five model families
    (MPA x3, GED, ROT, TAN, 3DT) in one frame batch: SAD (integer search) and HAD (quarter-pel refinement) == oracle on a
    sample of blocks, and the whole-batch checksum is block-order invariant."""
    Wd, Hd = 4096, 2048
    o = ol.Oracle(Wd, Hd)
    g = mm.MmGpu(Wd, Hd, max_refs=1)
    try:
        o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues([1, 0, 0]))
        g.set_rotation(0, mm.rodrigues([1, 0, 0]))
        ref = wl.pad_frame(wl.synth_frame(Wd, Hd, seed=51), 288, 1)
        cur = wl.synth_frame(Wd, Hd, seed=52, shift=(4.0, -2.0))
        g.set_ref(0, 0, [p.origin for p in ref], [p.stride for p in ref])
        g.set_org(cur[0])
        class P: pass
        pl = P(); pl.origin = ref[0].origin; pl.stride = ref[0].stride
        models = (mm.MPA_FB, mm.MPA_LR, mm.MPA_TB, mm.GED_CAMPOSE, mm.ROT, mm.TAN, mm.T3D)
        for frac, dist in ((False, mm.DIST_SAD), (True, mm.DIST_HAD)):
            blocks, mvs, prec = wl.frame_batch(Wd, Hd, seed=9, cands_per_block=4, mv_range=64, models=models, frac=frac)
            got = g.me_batch(blocks, mvs, mv_prec_shift=prec, dist=dist)
            rng = np.random.default_rng(1)
            for bi in rng.choice(len(blocks), size=160, replace=False):
                b = blocks[bi]
                sl = slice(b["cand_first"], b["cand_first"] + b["cand_count"])
                org = cur[0][b["y"]:b["y"] + b["h"], b["x"]:b["x"] + b["w"]]
                m = np.stack([mvs["hor"][sl], mvs["ver"][sl]], axis=1)
                exp = o.me_block(pl, org, None, (int(b["x"]), int(b["y"])), (int(b["w"]), int(b["h"])), int(b["model"]), m,
                                 mv_prec_shift=prec, dist=dist, had_mode=ol.HAD_SIMD)
                assert np.array_equal(exp, got[sl]), (dict(zip(b.dtype.names, b.tolist())), exp, got[sl])
            perm = np.random.default_rng(2).permutation(len(blocks))
            assert np.array_equal(got, g.me_batch(blocks[perm], mvs, mv_prec_shift=prec, dist=dist))
    finally:
        g.close()
        o.close()


def test_tan_double_sincos_exhaustive(env):
    """TangentialMotionModel.cpp:27-28,38,40 narrow a DOUBLE libm sin/cos to float.  The device code
    (mm_libm.cuh: mm_sin_d2f / mm_cos_d2f) must give glibc's narrowed result for EVERY float of the domain |x| <= pi/2:
    its checksum over all 2^31 inputs equals the host sweep of oracle/tools/tan_exceptions.cpp, committed in
    csrc/mm_tan_exceptions.h (MM_TAN_SWEEP_CHECKSUM); a strided sweep is re-done against this box's glibc here."""
    import os, re, subprocess
    o, g, *_ = env
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "vvc-extension-mm_b200", "csrc", "mm_tan_exceptions.h")).read()
    full = int(re.search(r"MM_TAN_SWEEP_CHECKSUM 0x([0-9a-f]+)ull", hdr).group(1), 16)
    assert g.selftest_tan_sincos(1) == full
    tool = os.path.join(root, "oracle", "_build", "tan_exceptions")
    if not os.path.exists(tool):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(root, "oracle", "tools", "tan_exceptions.cpp"),
                               "-o", tool, "-lm"])
    host = int(subprocess.check_output([tool, "sum", "97"]).decode().strip(), 16)
    assert g.selftest_tan_sincos(97) == host


def test_multi_reference_batch_equals_single_reference_calls(env):
    """mmgpu_me_batch_device_multiref: one float mapping (per distinct GEODESIC_CAMPOSE rotation), one interpolation +
    distortion pass per reference picture == the plain call repeated with the blocks' ref_slot / rot_slot rewritten."""
    import torch
    o, g, planes, ref, cur = env
    refs = []
    for f in range(3):
        pp = wl.pad_frame(wl.synth_frame(W, H, seed=40 + f, shift=(0.5 * f, 0.25 * f)), 288, 1)
        g.set_ref(1 + f, 10 + f, [p.origin for p in pp], [p.stride for p in pp])
        refs.append(pp)
    g.set_rotation(1, mm.rodrigues([0.1, 0.9, -0.2]))
    g.set_org(cur[0])
    blocks, mvs, prec = wl.frame_batch(W, H, seed=9, cands_per_block=5, mv_range=32, models=(1, 2, 3, 4, 5, 6, 10), ctu_rows=[1, 5])
    plan = g.plan(blocks, len(mvs))
    dm = torch.from_numpy(mvs.view(np.int32).reshape(-1, 2).copy()).cuda()
    ref_slots, rot_slots = [1, 2, 3, 0], [0, 1, 0, 1]
    dd = torch.zeros(len(ref_slots) * len(mvs), dtype=torch.int64, device="cuda")
    for dist in (mm.DIST_SAD, mm.DIST_HAD):
        g.me_batch_device_multiref(plan, dm.data_ptr(), dd.data_ptr(), ref_slots, rot_slots, mv_prec_shift=prec, dist=dist)
        g.sync()
        got = dd.cpu().numpy().view(np.uint64).reshape(len(ref_slots), len(mvs))
        for r, (rs, ro) in enumerate(zip(ref_slots, rot_slots)):
            b = blocks.copy()
            b["ref_slot"] = rs
            b["rot_slot"] = ro
            exp = g.me_batch(b, mvs, mv_prec_shift=prec, dist=dist)
            assert np.array_equal(got[r], exp), (dist, r)
    assert len({tuple(x) for x in got}) == len(ref_slots)   # the four references really give different distortions


def test_merge_satd_uses_the_prediction_oob_rule(env):
    """merge-candidate SATD / xGetInterPredictionError evaluate predictions of motionCompensation, i.e. with
    xPredInterBlkMM's out-of-picture rule (InterPrediction.cpp:776-780), not the search's: mmgpu_set_dist_oob_mode(PRED)
    + the batch call == HAD(org, oracle MC with the PRED rule); near the picture border the two rules differ."""
    o, g, planes, ref, cur = env
    g.set_org(cur[0])
    rng = np.random.default_rng(5)
    blocks = np.zeros(24, dtype=mm.api.BLOCK_DTYPE)
    mvs = np.zeros(24, dtype=mm.api.MV_DTYPE)
    shapes = [(16, 16), (32, 8), (8, 32), (64, 64)]
    for i in range(24):
        w, h = shapes[i % 4]
        x = (W - w) if i % 3 == 0 else int(rng.integers(0, (W - w) // 4)) * 4      # right border: the +maxCU slack matters
        y = (H - h) if i % 3 == 1 else int(rng.integers(0, (H - h) // 4)) * 4      # bottom border
        blocks[i] = (x, y, w, h, [1, 2, 3, 4, 5, 6, 10][i % 7], 0, 0, i, 1, x, y, 0)
        mvs[i] = (int(rng.integers(0, 900)), int(rng.integers(0, 700)))
    try:
        g.set_dist_oob_mode(mm.OOB_PRED)
        got = g.me_batch(blocks, mvs, mv_prec_shift=0, dist=mm.DIST_HAD)
    finally:
        g.set_dist_oob_mode(mm.OOB_ME)
    me = g.me_batch(blocks, mvs, mv_prec_shift=0, dist=mm.DIST_HAD)
    for i, b in enumerate(blocks):
        pred = o.mc_block(planes[0], 0, 1, (int(b["x"]), int(b["y"])), (int(b["w"]), int(b["h"])), (int(mvs["hor"][i]), int(mvs["ver"][i])),
                          int(b["model"]), oob=ol.OOB_PRED)
        org = cur[0][b["y"]:b["y"] + b["h"], b["x"]:b["x"] + b["w"]]
        assert int(got[i]) == ol.had(np.ascontiguousarray(org), pred, ol.HAD_SIMD), i
    assert (got != me).any()


def test_device_side_unpadded_registration_single_and_batched(env):
    """mmgpu_set_ref_unpadded_device / mmgpu_set_refs_unpadded_device (the multi-GPU exchange moves un-padded pictures and
    rebuilds the margins on arrival): distortions against the device-padded slot == against the host-padded slot, for the
    replicated and the wrapped border, one picture per call and two pictures in one launch."""
    import torch
    o, g, planes, ref, cur = env
    g.set_org(cur[0])
    blocks, mvs, prec = wl.frame_batch(W, H, seed=21, cands_per_block=3, mv_range=48, models=(0, 1, 4, 6, 10), ctu_rows=[0, 7])
    blocks["ref_slot"] = 0
    exp = g.me_batch(blocks, mvs, mv_prec_shift=prec, dist=mm.DIST_SAD)
    un = [torch.from_numpy(np.ascontiguousarray(p.inner)).cuda() for p in ref]
    b1 = blocks.copy(); b1["ref_slot"] = 1
    b2 = blocks.copy(); b2["ref_slot"] = 2
    g.set_ref_unpadded_device(1, 5, [t.data_ptr() for t in un], [t.shape[1] for t in un], keep=un)
    assert np.array_equal(g.me_batch(b1, mvs, mv_prec_shift=prec, dist=mm.DIST_SAD), exp)
    other = [torch.from_numpy(np.ascontiguousarray(p.inner)).cuda() for p in wl.pad_frame(wl.synth_frame(W, H, seed=77), 288, 1)]
    g.set_refs_unpadded_device([2, 1], [6, 7], [[t.data_ptr() for t in un], [t.data_ptr() for t in other]],
                               [[t.shape[1] for t in un]] * 2, keep=(un, other))
    assert np.array_equal(g.me_batch(b2, mvs, mv_prec_shift=prec, dist=mm.DIST_SAD), exp)
    assert not np.array_equal(g.me_batch(b1, mvs, mv_prec_shift=prec, dist=mm.DIST_SAD), exp)      # slot 1 now holds the other picture
    # the multi-GPU form (mmgpu_pull_refs_unpadded: persistent pull kernels, here from local addresses): same planes, margins
    # included -- distortions at the picture border and MC deep inside every margin
    for ctas in (3, 24):
        g.pull_refs_unpadded([2, 1], [6, 7], [[t.data_ptr() for t in other], [t.data_ptr() for t in un]], [[t.shape[1] for t in un]] * 2,
                             max_ctas=ctas, keep=(un, other))
        assert np.array_equal(g.me_batch(b1, mvs, mv_prec_shift=prec, dist=mm.DIST_SAD), exp)
        assert not np.array_equal(g.me_batch(b2, mvs, mv_prec_shift=prec, dist=mm.DIST_SAD), exp)
        for (x, y, mv) in [(0, 0, (-4000, -4100)), (W - 64, 0, (4000, -4000)), (0, H - 64, (-3900, 4000)), (W - 64, H - 64, (4000, 4000)),
                           (1024, 0, (5, -4200)), (0, 512, (-4200, 7))]:
            for comp, pos, size in ((0, (x, y), (64, 64)), (1, (x // 2, y // 2), (32, 32)), (2, (x // 2, y // 2), (32, 32))):
                assert np.array_equal(g.mc(0, comp, 1, pos, size, mv, 0), g.mc(1, comp, 1, pos, size, mv, 0)), (ctas, x, y, mv, comp)
        g.set_refs_unpadded_device([2, 1], [6, 7], [[t.data_ptr() for t in un], [t.data_ptr() for t in other]], [[t.shape[1] for t in un]] * 2,
                                   keep=(un, other))
    # wrapped border: batched device padding == host-side extendWrapBorder of the oracle
    wr = []
    for p in ref:
        q = ol.PaddedPlane(p.w, p.h, p.mx, p.my)
        q.inner[:] = p.inner
        q.extend_wrap(W if p.w == W else W // 2)
        wr.append(q)
    g.set_ref(3, 9, [q.origin for q in wr], [q.stride for q in wr])
    b3 = blocks.copy(); b3["ref_slot"] = 3
    expw = g.me_batch(b3, mvs, mv_prec_shift=prec, dist=mm.DIST_SAD)
    g.set_refs_unpadded_device([2], [8], [[t.data_ptr() for t in un]], [[t.shape[1] for t in un]], wrap_offset=W, keep=un)
    assert np.array_equal(g.me_batch(b2, mvs, mv_prec_shift=prec, dist=mm.DIST_SAD), expw)
    g.set_ref(0, 0, [p.origin for p in ref], [p.stride for p in ref])


def test_mvp_convert_through_the_abi_equals_reference_source(env):
    """mmgpu_mvp_convert (8f-1, evaluated on the host inside the C-ABI library) == the reference's own
    MVReprojection::motionVectorInDesiredMotionModel (oracle/_ref/libvtmmap.so) for all model pairs."""
    import ref_lib as rl
    if not rl.map_available():
        pytest.skip("oracle/_ref/libvtmmap.so not built")
    o, g, *_ = env
    r = rl.RefMap(W, H, 1.0, 1)
    e2 = [0.1, 0.9, -0.2]
    r.set_epipole(EPIPOLE)
    r.set_epipole2(e2)
    g.set_rotation(0, mm.rodrigues(EPIPOLE))
    g.set_rotation(1, mm.rodrigues(e2))
    rng = np.random.default_rng(81)
    n = 0
    for mo in range(11):
        for md in range(11):
            for k in range(12):
                cw, chh, uw, uh = [int(2 ** rng.integers(2, 8)) for _ in range(4)]
                cx, cy = int(rng.integers(0, (W - cw) // 4 + 1)) * 4, int(rng.integers(0, (H - chh) // 4 + 1)) * 4
                ux, uy = int(rng.integers(0, (W - uw) // 4 + 1)) * 4, int(rng.integers(0, (H - uh) // 4 + 1)) * 4
                pos = (ux + int(rng.integers(0, uw)), uy + int(rng.integers(0, uh)))
                mv = (int(rng.integers(-600, 600)), int(rng.integers(-300, 300)))
                po, pd = (7 if k % 2 else 0), (7 if k % 4 < 2 else 0)
                exp = r.mvp(pos, mv, mo, md, (cx, cy, cw, chh), (ux, uy, uw, uh), (4, 4), po, pd)
                got = g.mvp_convert(pos, mv, mo, md, (cx, cy, cw, chh), (ux, uy, uw, uh), (4, 4), (1 if po == 7 else 0, 1 if pd == 7 else 0))
                assert got == exp, (mo, md, pos, mv, exp, got)
                n += 1
    assert n == 11 * 11 * 12
    g.set_rotation(1, mm.rodrigues([0.1, 0.9, -0.2]))


def test_classic_bi_mc_with_bdof_border_fetch(env):
    """a8: xPredInterBlk with bioApplied -- the 14-bit prediction inside the BDOF work buffer plus the ring of
    integer-position samples around it (InterPrediction.cpp:1007-1015, 1054-1093): mmgpu_mc_bdof == oracle, for fractions
    below / at / above half a sample (the ring fetch moves by one sample), integer MVs and blocks at the picture border."""
    o, g, planes, ref, cur = env
    rng = np.random.default_rng(88)
    for i, (w, h) in enumerate([(8, 8), (16, 8), (8, 16), (16, 16), (32, 32), (64, 16), (128, 128), (64, 64)] * 3):
        x = [0, W - w, int(rng.integers(0, (W - w) // 4)) * 4][i % 3]
        y = [int(rng.integers(0, (H - h) // 4)) * 4, 0, H - h][i % 3]
        mv = [(0, 0), (7, 8), (8, 7), (-9, 23), (16, -32), (int(rng.integers(-900, 900)), int(rng.integers(-500, 500)))][i % 6]
        exp = o.mc_bdof(planes[0], (x, y), (w, h), mv)
        got = g.mc_bdof(0, (x, y), (w, h), mv)
        assert np.array_equal(got[1:h + 3, 1:w + 3], exp[1:h + 3, 1:w + 3]), (x, y, w, h, mv)


def test_classic_dmvr_search_and_final_mc(env):
    """f4 (second half): InterPrediction::xProcessDMVR (InterPrediction.cpp:2239-2439).  The bilinear integer search of batched
    sub-PUs (mmgpu_dmvr_classic_costs == oracle's literal prefetch-buffer restatement, incl. windows in the picture margin), the
    decision (mmgpu_dmvr_decide == oracle) and the final MC from the window padded by replication (mmgpu_dmvr_final_mc):
    luma / chroma, every integer displacement, the alternative half-pel filter and the BDOF ring."""
    o, g, planes, ref, cur = env
    ref1 = wl.pad_frame(wl.synth_frame(W, H, seed=13, shift=(-2.0, 0.5)), 288, 1)
    g.set_ref(1, 1, [p.origin for p in ref1], [p.stride for p in ref1])
    rng = np.random.default_rng(2239)
    subs = np.zeros(96, dtype=mm.api.DMVR_CLASSIC_DTYPE)
    for i in range(len(subs)):
        w, h = [(16, 16), (8, 16), (16, 8), (8, 8)][i % 4]
        s = subs[i]
        s["w"], s["h"], s["ref_slot0"], s["ref_slot1"] = w, h, i % 2, 1 - i % 2
        for l in "01":
            if i % 5 == 0:   # windows reaching into the margin
                s["org_x" + l] = [-250, W + 200, int(rng.integers(0, W))][i % 3]
                s["org_y" + l] = [int(rng.integers(0, H)), -240, H + 190][i % 3]
            else:
                s["org_x" + l], s["org_y" + l] = int(rng.integers(-3, W - w - 4)), int(rng.integers(-3, H - h - 4))
            s["frac_x" + l], s["frac_y" + l] = [(0, 0), (int(rng.integers(0, 16)), 0), (0, int(rng.integers(0, 16))),
                                                 (int(rng.integers(1, 16)), int(rng.integers(1, 16)))][(i // 4 + int(l)) % 4]
    got = g.dmvr_classic_costs(subs)
    pl = [planes, ref1]
    for i, s in enumerate(subs):
        p0, p1 = pl[int(s["ref_slot0"])][0], pl[int(s["ref_slot1"])][0]
        exp = ol.dmvr_classic_costs(p0, p1, (int(s["org_x0"]), int(s["org_y0"])), (int(s["frac_x0"]), int(s["frac_y0"])),
                                    (int(s["org_x1"]), int(s["org_y1"])), (int(s["frac_x1"]), int(s["frac_y1"])), (int(s["w"]), int(s["h"])))
        assert np.array_equal(exp, got[i]), (i, s)
        assert ol.dmvr_decide(exp, int(s["w"]), int(s["h"])) == mm.dmvr_decide(got[i], int(s["w"]), int(s["h"]))
    assert len({mm.dmvr_decide(c, 16, 16)[1] for c in got}) > 3     # the search does move
    # final MC
    jobs, exps, off = [], [], 0
    for i in range(240):
        comp = i % 3
        luma = comp == 0
        w, h = ([(16, 16), (8, 16), (16, 8), (8, 8)] if luma else [(8, 8), (4, 8), (8, 4), (4, 4)])[(i // 3) % 4]
        pad = (2 if luma else 1) if i % 7 else 0
        dint = (int(rng.integers(-pad, pad + 1)), int(rng.integers(-pad, pad + 1)))
        nph = 16 if luma else 32
        frac = [(0, 0), (8 if luma else 16, 0), (0, 8 if luma else 16), (8 if luma else 16,) * 2, (int(rng.integers(1, nph)), int(rng.integers(1, nph))),
                (int(rng.integers(1, 8)), 0)][i % 6]
        alt, bio = int(luma and i % 2 == 0), int(luma and i % 4 < 2)
        pw, ph = (W, H) if luma else (W // 2, H // 2)
        org = (int(rng.integers(-8, pw - w)), int(rng.integers(-8, ph - h))) if i % 11 else (-(100 >> (not luma)), ph + 20)
        slot = i % 2
        n = (w + 4) * (h + 4) if bio else w * h
        jobs.append((comp, slot, org[0], org[1], w, h, pad, dint[0], dint[1], frac[0], frac[1], alt, bio, off, w))
        exps.append((off, n, ol.dmvr_final_mc(pl[slot][comp], luma, org, (w, h), pad, dint, frac, alt_hpel=bool(alt), bio=bool(bio))))
        off += n
    out = g.dmvr_final_mc(np.array(jobs, dtype=mm.api.DMVR_MC_JOB_DTYPE), off)
    for k, (o0, n, exp) in enumerate(exps):
        gotk = out[o0:o0 + n].reshape(exp.shape)
        if jobs[k][12]:   # bio: rows 1..h+2, columns 1..w+2 are defined
            assert np.array_equal(gotk[1:-1, 1:-1], exp[1:-1, 1:-1]), jobs[k]
        else:
            assert np.array_equal(gotk, exp), jobs[k]
    # plain xPredInterBlk with pu.cu->imv == IMV_HPEL
    for (w, h) in [(8, 8), (16, 32), (64, 64), (128, 128), (48, 8), (4, 16)]:
        for mv in [(8, 0), (0, 8), (8, 8), (-24, 40), (8, 3), (5, 8), (16, -16), (int(rng.integers(-500, 500)) * 8, int(rng.integers(-300, 300)) * 8)]:
            x, y = rand_block(rng, w, h)
            for bi in (False, True):
                assert np.array_equal(ol.mc_block_classic_alt(planes[0], (x, y), (w, h), mv, bi=bi), g.mc_alt_hpel(0, (x, y), (w, h), mv, bi=bi)), (w, h, mv, bi)
    g.set_ref(1, 0, [p.origin for p in ref], [p.stride for p in ref])


def test_mc_non_power_of_two_blocks(env):
    """The encoder's RD search also predicts 12 / 24 / 48-wide blocks (found by the random-access end-to-end run: a 48x8
    xPredInterBlkMM call): mmgpu_mc takes every size that is a multiple of the sub-block."""
    o, g, planes, ref, cur = env
    rng = np.random.default_rng(48)
    for (w, h) in [(48, 8), (24, 16), (12, 8), (8, 12), (96, 32), (48, 48)]:
        for model in (0, 1, 4, 5, 6, 10):
            x, y = int(rng.integers(0, (W - w) // 4)) * 4, int(rng.integers(0, (H - h) // 4)) * 4
            mv = (int(rng.integers(-300, 300)), int(rng.integers(-200, 200)))
            for comp in (0, 1):
                cx, cy, cw, ch = (x, y, w, h) if comp == 0 else (x // 2, y // 2, w // 2, h // 2)
                for bi in (False, True):
                    exp = (o.mc_block_classic(planes[comp], comp, 1, (cx, cy), (cw, ch), mv, bi) if model == 0 else
                           o.mc_block(planes[comp], comp, 1, (cx, cy), (cw, ch), mv, model, bi))
                    got = g.mc(0, comp, 1, (cx, cy), (cw, ch), mv, model, bi=bi)
                    assert np.array_equal(exp, got), (w, h, model, comp, bi)


@pytest.mark.parametrize("name,kw", [
    ("8bit", dict(bit_depth=8)), ("9bit", dict(bit_depth=9)), ("400", dict(chroma_fmt=0)), ("422", dict(chroma_fmt=2)),
    ("444", dict(chroma_fmt=3)), ("ged_original", dict(ged_flavor=0)), ("offset1.5", dict(offset4x4=1.5)),
    ("maxcu64", dict(max_cu=64, margin=160)),
])
def test_other_configurations_accepted_by_create(name, kw):
    """Everything mmgpu_create accepts is checked against the oracle, not only 10-bit 4:2:0 modulated-GED offset 1.0: other
    bit depths (FIR shifts / offsets, clip range, the int16 intermediates), chroma formats (sub-block sizes, position scaling,
    classic chroma phase), the original GED flavor, MMOffset4x4 = 4 (1.5), a smaller CTU; luma 4x4 blocks included."""
    w, h = 512, 256
    bd = kw.get("bit_depth", 10)
    fmt = kw.get("chroma_fmt", 1)
    o = ol.Oracle(w, h, chroma_fmt=fmt, max_cu=kw.get("max_cu", 128), offset4x4=kw.get("offset4x4", 1.0), ged_flavor=kw.get("ged_flavor", 1),
                  bit_depth=bd)
    g = mm.MmGpu(w, h, max_refs=1, **kw)
    try:
        R = ol.Oracle.rodrigues(EPIPOLE)
        o.set_rotation(ol.GED_CAMPOSE, R)
        g.set_rotation(0, mm.rodrigues(EPIPOLE))
        margin = kw.get("margin", 288)
        planes = wl.synth_frame(w, h, seed=5, chroma_fmt=fmt, bit_depth=bd)
        ref = wl.pad_frame(planes, margin, fmt)
        g.set_ref(0, 0, [p.origin for p in ref], [p.stride for p in ref])
        cur = wl.synth_frame(w, h, seed=6, chroma_fmt=fmt, bit_depth=bd, shift=(1.5, 0.5))
        hi = (1 << bd) - 1
        rng = np.random.default_rng(len(name))
        csx = 1 if fmt in (1, 2) else 0
        csy = 1 if fmt == 1 else 0
        for i, (bw, bh) in enumerate([(4, 4), (8, 4), (16, 16), (32, 8), (64, 64), (4, 16)] * 3):
            x, y = int(rng.integers(0, (w - bw) // 8)) * 8, int(rng.integers(0, (h - bh) // 8)) * 8
            mv = (int(rng.integers(-200, 200)), int(rng.integers(-120, 120)))
            for model in (0, 1, 3, 4, 5, 6, 7, 10):
                ncomp = 1 if fmt == 0 else 3
                for comp in range(ncomp):
                    cx, cy, cw, ch = (x, y, bw, bh) if comp == 0 else (x >> csx, y >> csy, bw >> csx, bh >> csy)
                    if cw < (4 >> (csx if comp else 0)) or ch < (4 >> (csy if comp else 0)):
                        continue
                    for bi in (False, True):
                        exp = (o.mc_block_classic(ref[comp], comp, fmt, (cx, cy), (cw, ch), mv, bi, clp=(0, hi)) if model == 0 else
                               o.mc_block(ref[comp], comp, fmt, (cx, cy), (cw, ch), mv, model, bi, clp=(0, hi)))
                        got = g.mc(0, comp, fmt, (cx, cy), (cw, ch), mv, model, bi=bi, clp=(0, hi))
                        assert np.array_equal(exp, got), (name, bw, bh, model, comp, bi)
        # motion search batches (luma): SAD and Hadamard, 4x4 blocks included
        g.set_org(cur[0])
        blocks, mvs, prec = wl.frame_batch(w, h, seed=3, cands_per_block=3, mv_range=24, models=(1, 2, 3, 4, 5, 6, 10, 7), frac=True)
        extra = np.zeros(8, dtype=mm.api.BLOCK_DTYPE)
        emv = np.zeros(8 * 3, dtype=mm.api.MV_DTYPE)
        for i in range(8):
            extra[i] = (8 * i + 64, 32, 4, 4, [1, 4, 5, 6, 10, 2, 3, 7][i], 0, 0, len(mvs) + 3 * i, 3, 8 * i + 64, 32, 0)
            emv["hor"][3 * i:3 * i + 3] = rng.integers(-60, 60, 3)
            emv["ver"][3 * i:3 * i + 3] = rng.integers(-60, 60, 3)
        blocks = np.concatenate([blocks, extra])
        mvs = np.concatenate([mvs, emv])
        for dist in (mm.DIST_SAD, mm.DIST_HAD):
            got = g.me_batch(blocks, mvs, mv_prec_shift=prec, dist=dist, clp=(0, hi))
            for b in blocks[::5].tolist() + extra.tolist():
                b = dict(zip(mm.api.BLOCK_DTYPE.names, b))
                sl = slice(b["cand_first"], b["cand_first"] + b["cand_count"])
                org = cur[0][b["y"]:b["y"] + b["h"], b["x"]:b["x"] + b["w"]]
                m = np.stack([mvs["hor"][sl], mvs["ver"][sl]], axis=1)
                exp = o.me_block(ref[0], org, None, (b["x"], b["y"]), (b["w"], b["h"]), b["model"], m, mv_prec_shift=prec, dist=dist,
                                 had_mode=ol.HAD_SIMD, clp=(0, hi))
                assert np.array_equal(exp, got[sl]), (name, dist, b)
    finally:
        g.close()
        o.close()


def test_bench_batch_equals_oracle(env):
    """The exact batch bench.py times (2048x1024, 1978 blocks x 16 candidates, MPA_FB/LR/TB + GED_CAMPOSE(1,0,0) + ROT, Hadamard)
    against the oracle: all 31 648 distortions."""
    import bench
    o, g, planes, ref, cur = env
    g.set_rotation(0, mm.rodrigues(bench.EPIPOLE))
    try:
        blocks, mvs, prec = wl.frame_batch(W, H, seed=1, cands_per_block=bench.CANDS, mv_range=bench.MV_RANGE, models=bench.models())
        g.set_org(cur[0])
        got = g.me_batch(blocks, mvs, mv_prec_shift=prec, dist=mm.DIST_HAD)
        exp = bench.oracle_distortions(ref[0], cur[0], blocks, mvs, prec, os.cpu_count() or 4)
        assert len(got) == 31648 and np.array_equal(got, exp)
    finally:
        g.set_rotation(0, mm.rodrigues(EPIPOLE))
