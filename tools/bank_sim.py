"""Shared-memory wavefront model of k_fir's horizontal pass on the bench batch (CPU only): for every work item whose 32
windows fit a TMA box, counts the wavefronts of its 33 LDS.64 + 11 LDS.32 under different lane <-> sub-block mappings.
python tools/bank_sim.py [max_blocks]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from vvc_extension_mm_b200 import workload as wl
W, H = 2048, 1024
DIMS = [(104, 18), (56, 32), (40, 46), (104, 24), (56, 44), (40, 62)]


def tile_shape(w, h):
    if w > h and h % 8 == 0 and w % 16 == 0: return 2, 1
    if w < h and w % 8 == 0 and h % 16 == 0: return 1, 2
    if w > h and h % 4 == 0 and w % 8 == 0: return 1, 0
    if w < h and w % 4 == 0 and h % 8 == 0: return 0, 1
    if h % 8 == 0 and w % 8 == 0: return 1, 1
    return 0, 0


def sub_from_index(w, h, s):
    ltw, lth = tile_shape(w, h)
    cols = w // 4
    t = s >> (ltw + lth); q = s & ((1 << (ltw + lth)) - 1)
    ltpr = int(np.log2(cols)) - ltw
    ty, tx = t >> ltpr, t & ((1 << ltpr) - 1)
    qy, qx = q >> ltw, q & ((1 << ltw) - 1)
    return (ty << lth) + qy, (tx << ltw) + qx


def wavefronts64(addr8):     # addr8: 8-byte word index per lane (32,), -1 = inactive
    n = 0
    for h in range(2):
        a = addr8[h * 16:(h + 1) * 16]
        a = np.unique(a[a >= 0])
        n += max(1, np.bincount(a % 16, minlength=16).max()) if len(a) else 0
    return n


def wavefronts32(addr4):
    a = np.unique(addr4[addr4 >= 0])
    return max(1, np.bincount(a % 32, minlength=32).max()) if len(a) else 0


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    o = ol.Oracle(W, H)
    o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues([1, 0, 0]))
    blocks, mvs, prec = wl.frame_batch(W, H, seed=1, cands_per_block=16, mv_range=64, models=(1, 2, 3, 10, 6))
    rng = np.random.default_rng(0)
    pick = rng.permutation(len(blocks))[:nb]
    tot = {"tile": 0, "row16": 0, "row16p32": 0}
    big = {"tile": 0, "row16": 0, "row16p32": 0}; nbig = 0
    items = boxed = 0
    ideal = 0
    for bi in pick:
        b = blocks[bi]
        w, h = int(b["w"]), int(b["h"])
        n = (w // 4) * (h // 4)
        if n < 32:
            continue
        rows, cols = h // 4, w // 4
        rc = [sub_from_index(w, h, s) for s in range(n)]
        ltw, lth = tile_shape(w, h)
        for ci in range(0, int(b["cand_count"]), 4):
            mv = mvs[int(b["cand_first"]) + ci]
            ox, oy = o.reproject((int(b["x"]), int(b["y"])), (w, h), (int(mv["hor"]) << prec, int(mv["ver"]) << prec), int(b["model"]))
            ox = ox.reshape(cols, rows).T; oy = oy.reshape(cols, rows).T
            for ch in range(0, n, 32):
                subs = rc[ch:ch + 32]
                fx = np.array([ox[r, c] for r, c in subs]); fy = np.array([oy[r, c] for r, c in subs])
                xP, yP = fx >> 4, fy >> 4
                oob = (xP < 0) | (yP < 0) | (xP >= W - 4) | (yP >= H - 4)
                items += 1
                if oob.all():
                    continue
                ax = ((xP - 3) & ~3); ay = yP - 3
                x0 = ax[~oob].min() & ~7; y0 = ay[~oob].min()
                wb = (ax[~oob] + 16).max() - x0; hb = (ay[~oob] + 11).max() - y0
                bt = next((t for t, (bw_, bh_) in enumerate(DIMS) if wb <= bw_ and hb <= bh_), -1)
                if bt < 0:
                    continue
                boxed += 1
                pitchQ = DIMS[bt][0] // 4
                base8 = (ay - y0) * pitchQ + (ax - x0) // 4
                tail = ((xP - 3) & 3) >= 2
                # lane mappings: identity (tile-major) and row-major-16 (only when the chunk is 2 rows x 16 columns)
                r0 = np.array([r for r, c in subs]); c0 = np.array([c for r, c in subs])
                maps = {"tile": np.arange(32)}
                if lth == 1 and cols >= 16:
                    lane_of = (r0 - r0.min()) * 16 + (c0 - c0.min())
                    inv = np.empty(32, dtype=int); inv[lane_of] = np.arange(32)
                    maps["row16"] = inv
                else:
                    maps["row16"] = np.arange(32)
                if lth == 1 and cols >= 16:
                    maps["row16p32"] = maps["row16"]
                else:
                    maps["row16p32"] = np.arange(32)
                for name, perm in maps.items():
                    if name == "row16p32" and lth == 1 and cols >= 16 and wb <= 128:
                        pq = 32
                        b8x = (ay - y0) * pq + (ax - x0) // 4
                    else:
                        pq = pitchQ
                        b8x = base8
                    pitch_used = pq
                    b8 = np.where(oob, -1, b8x)[perm]
                    tl = (tail & ~oob)[perm]
                    wf = 0
                    for r in range(11):
                        for k in range(3):
                            wf += wavefronts64(np.where(b8 >= 0, b8 + r * pitch_used + k, -1))
                        wf += wavefronts32(np.where(tl, (b8 + r * pitch_used + 3) * 2, -1))
                    tot[name] += wf
                    if lth == 1 and cols >= 16: big[name] += wf
                if lth == 1 and cols >= 16: nbig += 1
                ideal += 11 * 7
    print("items %d, boxed %d (%.1f%%)" % (items, boxed, 100.0 * boxed / max(1, items)))
    for k, v in tot.items():
        print("%-9s wavefronts per boxed item: %.1f (ideal %.1f); items of >=16-column 2-row-tile chunks (%d): %.1f" % (k, v / max(1, boxed), ideal / max(1, boxed), nbig, big[k] / max(1, nbig)))


if __name__ == "__main__":
    main()
