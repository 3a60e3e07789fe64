"""Synthetic ERP frames and frame-wide block/candidate batches.

numpy only; shared by tests/ and bench.py.  Frames: smooth field + uniform noise, 10-bit internal;
reference planes padded like Picture::extendPicBorder (replicated margins, Picture.cpp:988-1048).
"""
import numpy as np
from .api import BLOCK_DTYPE, MV_DTYPE, MPA_FB, MPA_LR, MPA_TB, TAN, T3D, ROT, GED_CAMPOSE

DEFAULT_MODELS = (MPA_FB, MPA_LR, MPA_TB, T3D, TAN, ROT, GED_CAMPOSE)


def synth_plane(w, h, seed, k=3, m=2, bit_depth=10, shift=(0, 0)):
    """512 + 400 sin(2 pi x/W k) cos(pi y/H m) + U(-32,32), clipped to the sample range (int16)."""
    rng = np.random.default_rng(seed)
    x = (np.arange(w, dtype=np.float64) + shift[0])[None, :]
    y = (np.arange(h, dtype=np.float64) + shift[1])[:, None]
    half = 1 << (bit_depth - 1)
    f = half + 0.78 * half * np.sin(2 * np.pi * x / w * k) * np.cos(np.pi * y / h * m)
    f = f + rng.integers(-32, 33, size=(h, w))
    return np.clip(np.rint(f), 0, (1 << bit_depth) - 1).astype(np.int16)


def synth_frame(w, h, seed, chroma_fmt=1, bit_depth=10, shift=(0, 0)):
    """Returns [Y, Cb, Cr] (or [Y] for 4:0:0) un-padded int16 planes."""
    planes = [synth_plane(w, h, seed * 3 + 1, 3, 2, bit_depth, shift)]
    if chroma_fmt != 0:
        cw = w >> (1 if chroma_fmt in (1, 2) else 0)
        ch = h >> (1 if chroma_fmt == 1 else 0)
        planes.append(synth_plane(cw, ch, seed * 3 + 2, 2, 1, bit_depth, (shift[0] / 2, shift[1] / 2)))
        planes.append(synth_plane(cw, ch, seed * 3 + 3, 1, 3, bit_depth, (shift[0] / 2, shift[1] / 2)))
    return planes


class Padded:
    """Plane with replicated margins; .origin = host address of sample (0,0); .stride in samples."""

    def __init__(self, plane, mx, my, align=8):
        h, w = plane.shape
        stride = (w + 2 * mx + align - 1) // align * align
        buf = np.zeros((h + 2 * my, stride), dtype=np.int16)
        buf[:, :w + 2 * mx] = np.pad(plane, ((my, my), (mx, mx)), mode="edge")
        self.buf, self.w, self.h, self.mx, self.my, self.stride = buf, w, h, mx, my, stride

    @property
    def origin(self):
        return self.buf.ctypes.data + 2 * (self.my * self.stride + self.mx)

    @property
    def inner(self):
        return self.buf[self.my:self.my + self.h, self.mx:self.mx + self.w]


def pad_frame(planes, margin=288, chroma_fmt=1):
    out = []
    for c, p in enumerate(planes):
        sx = 1 if (c and chroma_fmt in (1, 2)) else 0
        sy = 1 if (c and chroma_fmt == 1) else 0
        out.append(Padded(p, margin >> sx, margin >> sy))
    return out


def _split(x, y, w, h, rng, out, min_area=32, depth=0):
    """Random quad/binary/ternary-free partition of a CTU into VVC inter block sizes (powers of two)."""
    can_h = h >= 8 and (w * h // 2) >= min_area
    can_v = w >= 8 and (w * h // 2) >= min_area
    can_q = w >= 16 and h >= 16 and w == h
    p_stop = 0.0 if max(w, h) > 64 and depth == 0 else (0.15 + 0.12 * depth)
    if not (can_h or can_v or can_q) or rng.random() < p_stop:
        out.append((x, y, w, h))
        return
    choices = [c for c, ok in (("q", can_q), ("h", can_h), ("v", can_v)) if ok]
    c = choices[rng.integers(len(choices))]
    if c == "q":
        for dy in (0, h // 2):
            for dx in (0, w // 2):
                _split(x + dx, y + dy, w // 2, h // 2, rng, out, min_area, depth + 1)
    elif c == "h":
        _split(x, y, w, h // 2, rng, out, min_area, depth + 1)
        _split(x, y + h // 2, w, h // 2, rng, out, min_area, depth + 1)
    else:
        _split(x, y, w // 2, h, rng, out, min_area, depth + 1)
        _split(x + w // 2, y, w // 2, h, rng, out, min_area, depth + 1)


def frame_partition(w, h, seed, ctu=128, fixed=None):
    """Tile every CTU of the frame into blocks.  fixed=(bw,bh) gives a uniform tiling instead."""
    rng = np.random.default_rng(seed)
    blocks = []
    for cy in range(0, h, ctu):
        for cx in range(0, w, ctu):
            if fixed:
                bw, bh = fixed
                for y in range(cy, cy + ctu, bh):
                    for x in range(cx, cx + ctu, bw):
                        blocks.append((x, y, bw, bh))
            else:
                k = rng.integers(0, 8)
                if k == 0:
                    blocks.append((cx, cy, ctu, ctu))  # keep some 128x128 blocks
                else:
                    _split(cx, cy, ctu, ctu, rng, blocks)
    return blocks


def frame_batch(w, h, seed, cands_per_block=16, mv_range=64, models=DEFAULT_MODELS, ref_slot=0, rot_slot=0, ctu=128,
                fixed=None, ctu_rows=None, frac=False):
    """Blocks + candidate MVs of one frame (MV units: integer pel unless frac=True -> quarter pel).

    ctu_rows: optional iterable of CTU-row indices to keep (multi-GPU sharding by CTU rows).
    Returns (blocks[BLOCK_DTYPE], mvs[MV_DTYPE], mv_prec_shift).
    """
    rng = np.random.default_rng(seed + 7919)
    part = frame_partition(w, h, seed, ctu, fixed)
    if ctu_rows is not None:
        keep = set(ctu_rows)
        part = [b for b in part if (b[1] // ctu) in keep]
    nb = len(part)
    blocks = np.zeros(nb, dtype=BLOCK_DTYPE)
    mvs = np.zeros(nb * cands_per_block, dtype=MV_DTYPE)
    for i, (x, y, bw, bh) in enumerate(part):
        blocks[i] = (x, y, bw, bh, models[i % len(models)], ref_slot, rot_slot, i * cands_per_block, cands_per_block, x, y, 0)
    scale = 4 if frac else 1
    mvs["hor"] = rng.integers(-mv_range * scale, mv_range * scale + 1, size=len(mvs))
    mvs["ver"] = rng.integers(-mv_range * scale, mv_range * scale + 1, size=len(mvs))
    return blocks, mvs, (2 if frac else 4)


def shard_ctu_rows(n_rows, rank, world):
    """CTU rows of a frame handled by `rank`: round-robin."""
    return [r for r in range(n_rows) if r % world == rank]
