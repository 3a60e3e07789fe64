#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the CPU oracle (run in the build container, commit the output).

The reference ships no tests or golden vectors and cannot be built here (Eigen missing), so
these fixtures pin the ORACLE's behaviour at the time of generation; hand-derived known answers for the
integer steps live in tests/test_oracle_cpu.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as ol  # noqa: E402
from vvc_extension_mm_b200 import workload as wl  # noqa: E402


def main():
    W, H = 2048, 1024
    o = ol.Oracle(W, H)
    o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues([1, 0, 0]))
    rng = np.random.default_rng(20261018)
    sizes = [(8, 4), (4, 8), (8, 8), (16, 8), (8, 16), (16, 16), (32, 32), (64, 16), (128, 128), (4, 64)]
    cases, outs = [], []
    for model in range(1, 11):
        for it in range(24):
            w, h = sizes[it % len(sizes)]
            x = int(rng.integers(0, (W - w) // 4 + 1)) * 4
            y = int(rng.integers(0, (H - h) // 4 + 1)) * 4
            if it % 6 == 0:
                y = 0 if it % 12 else H - h
            mv = (int(rng.integers(-700, 700)), int(rng.integers(-350, 350)))
            if it % 8 == 0:
                mv = (0, 0)
            for comp, fmt in ((0, 0), (1, 1)):
                px, py, pw, ph = (x, y, w, h) if comp == 0 else (x // 2, y // 2, w // 2, h // 2)
                ox, oy = o.reproject((px, py), (pw, ph), mv, model, comp, fmt)
                cases.append((model, comp, fmt, px, py, pw, ph, mv[0], mv[1], len(ox)))
                outs.append(np.stack([ox, oy]))
    np.savez_compressed(os.path.join(HERE, "positions_2048x1024.npz"), cases=np.array(cases, dtype=np.int32),
                        out=np.concatenate(outs, axis=1))
    # MC + distortion on a small synthetic frame region
    ref = wl.pad_frame(wl.synth_frame(W, H, seed=1), 288, 1)
    cur = wl.synth_frame(W, H, seed=2, shift=(2.0, 1.0))
    class P: pass
    pl = P(); pl.origin = ref[0].origin; pl.stride = ref[0].stride
    dcases, dists = [], []
    for model in (1, 2, 3, 4, 5, 6, 10):
        for (w, h) in sizes:
            x = int(rng.integers(0, (W - w) // 4 + 1)) * 4
            y = int(rng.integers(0, (H - h) // 4 + 1)) * 4
            mvs = np.stack([rng.integers(-40, 40, size=4), rng.integers(-24, 24, size=4)], axis=1)
            org = cur[0][y:y + h, x:x + w]
            sad = o.me_block(pl, org, None, (x, y), (w, h), model, mvs, dist=ol.DIST_SAD)
            had = o.me_block(pl, org, None, (x, y), (w, h), model, mvs, dist=ol.DIST_HAD, had_mode=ol.HAD_SIMD)
            for i in range(4):
                dcases.append((model, x, y, w, h, mvs[i, 0], mvs[i, 1]))
                dists.append((sad[i], had[i]))
    np.savez_compressed(os.path.join(HERE, "distortion_2048x1024.npz"), cases=np.array(dcases, dtype=np.int32),
                        dist=np.array(dists, dtype=np.uint64))
    print("wrote", len(cases), "position cases and", len(dcases), "distortion cases")


if __name__ == "__main__":
    main()
