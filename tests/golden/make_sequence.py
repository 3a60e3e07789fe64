"""Synthetic 8-bit 4:2:0 ERP sequence with real motion on the sphere (test input for the end-to-end
EncoderApp/DecoderApp runs, BASELINE configs 1-4 at a small size): a seam-free texture defined on the unit
sphere, viewed by a camera that rotates a little and whose scene drifts along the x axis between frames
(so the rotational, geodesic and plane-adaptive models all find signal), plus a little noise.

usage: python tests/golden/make_sequence.py W H FRAMES out.yuv [seed]"""
import sys
import numpy as np


def texture(v):
    x, y, z = v
    t = (np.sin(7 * x + 3 * y) + np.sin(9 * y - 4 * z + 1.0) + np.sin(11 * z + 5 * x + 2.0) + 0.6 * np.sin(23 * x * y + 17 * z)
         + 0.5 * np.sin(31 * (x + y + z)))
    return t / 4.1


def rot(axis, ang):
    axis = np.asarray(axis, dtype=np.float64)
    axis /= np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def plane(w, h, t, gain, offs, rng, freq):
    xs = (np.arange(w) + 0.5) / w
    ys = (np.arange(h) + 0.5) / h
    phi = -xs[None, :] * 2 * np.pi
    th = ys[:, None] * np.pi
    v = np.stack([np.sin(th) * np.cos(phi), np.sin(th) * np.sin(phi), np.cos(th) * np.ones_like(phi)])
    R = rot([0.3, 0.2, 1.0], 0.012 * t) @ rot([1, 0, 0], 0.008 * t)
    v = np.tensordot(R, v, axes=1)
    # translation along x: points at distance ~ 1/(0.4 + 0.3*texture-dependent depth) shift towards -x
    d = 0.02 * t
    v = v + np.array([d, 0, 0])[:, None, None] * (0.6 + 0.4 * np.sin(3 * v[1] + 2 * v[2]))
    v /= np.linalg.norm(v, axis=0, keepdims=True)
    f = offs + gain * texture(v * freq) + rng.normal(0, 1.2, size=(h, w))
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def main():
    w, h, n, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    seed = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    rng = np.random.default_rng(seed)
    with open(out, "wb") as f:
        for t in range(n):
            f.write(plane(w, h, t, 90, 128, rng, 1.0).tobytes())
            f.write(plane(w // 2, h // 2, t, 40, 128, rng, 0.5).tobytes())
            f.write(plane(w // 2, h // 2, t, 40, 120, rng, 0.7).tobytes())


if __name__ == "__main__":
    main()
