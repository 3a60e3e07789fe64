"""Reader/replayer of the MC traces written by the patched reference apps (oracle/patches/MmGpuGlue.cpp, MMGPU_DUMP):
every xPredInterBlkMM call of a real decode/encode -- parameters, reference picture, predicted samples.
Test infrastructure: replays the calls through the CPU oracle (and, with a device, through libmmgpu.so)."""
import struct
import numpy as np
import oracle_lib as ol


def read_trace(path):
    """Yields ('R', poc, wrap, [planes]) and ('M', dict, dst) records."""
    with open(path, "rb") as f:
        data = f.read()
    off = 0
    while off < len(data):
        (tag,) = struct.unpack_from("<i", data, off)
        off += 4
        if tag == ord("R"):
            poc, wrap, nc = struct.unpack_from("<3i", data, off)
            off += 12
            planes = []
            for _ in range(nc):
                w, h = struct.unpack_from("<2i", data, off)
                off += 8
                planes.append(np.frombuffer(data, dtype=np.int16, count=w * h, offset=off).reshape(h, w).copy())
                off += 2 * w * h
            yield ("R", poc, wrap, planes)
        elif tag == ord("M"):
            v = struct.unpack_from("<15i3f", data, off)
            off += 15 * 4 + 12
            keys = ("comp", "fmt", "x", "y", "w", "h", "ref_poc", "wrap", "mv_hor", "mv_ver", "model", "bi", "clp_min", "clp_max",
                    "cur_poc")
            m = dict(zip(keys, v[:15]))
            m["epipole"] = v[15:18]
            dst = np.frombuffer(data, dtype=np.int16, count=m["w"] * m["h"], offset=off).reshape(m["h"], m["w"]).copy()
            off += 2 * m["w"] * m["h"]
            yield ("M", m, dst)
        else:
            raise ValueError("bad record tag %r at %d" % (tag, off - 4))


def pad_planes(planes, margin, fmt, wrap_offset=0):
    out = []
    for c, p in enumerate(planes):
        sx = 1 if (c and fmt in (1, 2)) else 0
        sy = 1 if (c and fmt == 1) else 0
        pp = ol.PaddedPlane(p.shape[1], p.shape[0], margin >> sx, margin >> sy)
        pp.inner[:] = p
        if wrap_offset:
            pp.extend_wrap(wrap_offset >> sx)
        else:
            pp.extend()
        out.append(pp)
    return out


def replay_oracle(path, pic_w, pic_h, margin=288, max_cu=128, limit=None, verbose=False):
    """Returns (calls, mismatches, first mismatch record or None)."""
    o = ol.Oracle(pic_w, pic_h, max_cu=max_cu)
    refs = {}
    n = bad = 0
    first = None
    last_e = None
    for rec in read_trace(path):
        if rec[0] == "R":
            _, poc, wrap, planes = rec
            refs[(poc, wrap)] = pad_planes(planes, margin, 1, pic_w if wrap else 0)
            continue
        _, m, dst = rec
        if m["model"] == ol.GED_CAMPOSE and m["epipole"] != last_e:
            o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues(m["epipole"]))
            last_e = m["epipole"]
        pl = refs[(m["ref_poc"], m["wrap"])][m["comp"]]
        got = o.mc_block(pl, m["comp"], m["fmt"], (m["x"], m["y"]), (m["w"], m["h"]), (m["mv_hor"], m["mv_ver"]), m["model"],
                         bi=bool(m["bi"]), clp=(m["clp_min"], m["clp_max"]))
        n += 1
        if not np.array_equal(got, dst):
            bad += 1
            if first is None:
                first = (m, dst, got)
            if verbose:
                print("MISMATCH", m, int(np.abs(got.astype(int) - dst).max()))
        if limit and n >= limit:
            break
    return n, bad, first


if __name__ == "__main__":
    import sys
    n, bad, first = replay_oracle(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), verbose=True)
    print("calls", n, "mismatches", bad)
    if first:
        m, dst, got = first
        print(m)
        print(dst[:8, :8])
        print(got[:8, :8])
