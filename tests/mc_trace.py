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
        elif tag == ord("O"):
            v = struct.unpack_from("<8i3f4i", data, off)
            off += 8 * 4 + 12 + 16
            keys = ("x", "y", "w", "h", "ref_poc", "wrap", "model", "cur_poc")
            m = dict(zip(keys, v[:8]))
            m["epipole"] = v[8:11]
            m["had"], m["sub_shift"], m["clp_min"], m["clp_max"] = v[11:15]
            org = np.frombuffer(data, dtype=np.int16, count=m["w"] * m["h"], offset=off).reshape(m["h"], m["w"]).copy()
            off += 2 * m["w"] * m["h"]
            yield ("O", m, org)
        elif tag == ord("E"):
            hor, ver, shift = struct.unpack_from("<3i", data, off)
            (dist,) = struct.unpack_from("<Q", data, off + 12)
            off += 20
            yield ("E", hor, ver, shift, dist)
        else:
            raise ValueError("bad record tag %r at %d" % (tag, off - 4))


def searches(path):
    """Groups the ME records: yields (header dict, org block, [(hor, ver, shift, dist), ...]) per dumped search, and the
    reference pictures seen so far as a dict {(poc, wrap): [planes]} (shared, updated while iterating)."""
    refs = {}
    cur = None
    for rec in read_trace(path):
        if rec[0] == "R":
            refs[(rec[1], rec[2])] = rec[3]
        elif rec[0] == "O":
            if cur:
                yield cur + (refs,)
            cur = (rec[1], rec[2], [])
        elif rec[0] == "E" and cur:
            cur[2].append(rec[1:])
    if cur:
        yield cur + (refs,)


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
        if rec[0] != "M":
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


def replay_me_oracle(path, pic_w, pic_h, margin=288, verbose=False):
    """Distortions of every dumped ME candidate through the oracle; returns (searches, candidates, mismatches)."""
    o = ol.Oracle(pic_w, pic_h)
    padded = {}
    ns = nc = bad = 0
    last_e = None
    for hdr, org, cands, refs in searches(path):
        key = (hdr["ref_poc"], hdr["wrap"])
        if key not in padded or padded[key][0] is not refs[key]:
            padded[key] = (refs[key], pad_planes(refs[key], margin, 1, pic_w if hdr["wrap"] else 0))
        pl = padded[key][1][0]
        if hdr["model"] == ol.GED_CAMPOSE and hdr["epipole"] != last_e:
            o.set_rotation(ol.GED_CAMPOSE, ol.Oracle.rodrigues(hdr["epipole"]))
            last_e = hdr["epipole"]
        ns += 1
        by_shift = {}
        for hor, ver, shift, dist in cands:
            by_shift.setdefault(shift, []).append((hor, ver, dist))
        for shift, lst in by_shift.items():
            mvs = np.array([(a, b) for a, b, _ in lst], dtype=np.int32)
            exp = np.array([d for _, _, d in lst], dtype=np.uint64)
            got = o.me_block(pl, org, None, (hdr["x"], hdr["y"]), (hdr["w"], hdr["h"]), hdr["model"], mvs, mv_prec_shift=shift,
                             dist=ol.DIST_HAD if hdr["had"] else ol.DIST_SAD, sub_shift=hdr["sub_shift"], had_mode=ol.HAD_SIMD,
                             clp=(hdr["clp_min"], hdr["clp_max"]))
            nc += len(lst)
            nb = int((got != exp).sum())
            bad += nb
            if nb and verbose:
                i = int(np.nonzero(got != exp)[0][0])
                print("ME MISMATCH", hdr, "mv", mvs[i], "shift", shift, "ref", int(exp[i]), "oracle", int(got[i]))
    return ns, nc, bad


if __name__ == "__main__":
    import sys
    print("ME searches, candidates, mismatches:", replay_me_oracle(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), verbose=True))
    n, bad, first = replay_oracle(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), verbose=True)
    print("calls", n, "mismatches", bad)
    if first:
        m, dst, got = first
        print(m)
        print(dst[:8, :8])
        print(got[:8, :8])
