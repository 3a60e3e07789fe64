"""Oracle restatement (TEST INFRASTRUCTURE) of the reference's integer-search point enumerators, written line by line after
EncoderLib/InterSearch.cpp: xTZ2PointSearch (:456-481), xTZ8PointDiamondSearch (:535-737), the raster scan of xTZSearch
(:5611-5647) and xTZSearchHelp's compare chain (:363-453).  `visit` collects (x, y, pointNr, distance) like xTZSearchHelp's
arguments.  Parity note: pinned only through the end-to-end runs (tests/test_e2e_reference_apps.py drives the REFERENCE'S OWN
enumerators with GPU distortions and gets a byte-identical bitstream); this file checks the product's host-side copies."""


def diamond(sr, sx, sy, dist, corners, visit):
    left_, right_, top_, bottom_ = sr
    top, bottom, left, right = sy - dist, sy + dist, sx - dist, sx + dist
    if dist == 1:
        if top >= top_:
            if corners:
                if left >= left_:
                    visit(left, top, 1, dist)
                visit(sx, top, 2, dist)
                if right <= right_:
                    visit(right, top, 3, dist)
            else:
                visit(sx, top, 2, dist)
        if left >= left_:
            visit(left, sy, 4, dist)
        if right <= right_:
            visit(right, sy, 5, dist)
        if bottom <= bottom_:
            if corners:
                if left >= left_:
                    visit(left, bottom, 6, dist)
                visit(sx, bottom, 7, dist)
                if right <= right_:
                    visit(right, bottom, 8, dist)
            else:
                visit(sx, bottom, 7, dist)
    elif dist <= 8:
        t2, b2, l2, r2 = sy - (dist >> 1), sy + (dist >> 1), sx - (dist >> 1), sx + (dist >> 1)
        if top >= top_ and left >= left_ and right <= right_ and bottom <= bottom_:
            visit(sx, top, 2, dist)
            visit(l2, t2, 1, dist >> 1)
            visit(r2, t2, 3, dist >> 1)
            visit(left, sy, 4, dist)
            visit(right, sy, 5, dist)
            visit(l2, b2, 6, dist >> 1)
            visit(r2, b2, 8, dist >> 1)
            visit(sx, bottom, 7, dist)
        else:
            if top >= top_:
                visit(sx, top, 2, dist)
            if t2 >= top_:
                if l2 >= left_:
                    visit(l2, t2, 1, dist >> 1)
                if r2 <= right_:
                    visit(r2, t2, 3, dist >> 1)
            if left >= left_:
                visit(left, sy, 4, dist)
            if right <= right_:
                visit(right, sy, 5, dist)
            if b2 <= bottom_:
                if l2 >= left_:
                    visit(l2, b2, 6, dist >> 1)
                if r2 <= right_:
                    visit(r2, b2, 8, dist >> 1)
            if bottom <= bottom_:
                visit(sx, bottom, 7, dist)
    else:
        if top >= top_ and left >= left_ and right <= right_ and bottom <= bottom_:
            visit(sx, top, 0, dist)
            visit(left, sy, 0, dist)
            visit(right, sy, 0, dist)
            visit(sx, bottom, 0, dist)
            for index in range(1, 4):
                yt, yb = top + (dist >> 2) * index, bottom - (dist >> 2) * index
                xl, xr = sx - (dist >> 2) * index, sx + (dist >> 2) * index
                visit(xl, yt, 0, dist)
                visit(xr, yt, 0, dist)
                visit(xl, yb, 0, dist)
                visit(xr, yb, 0, dist)
        else:
            if top >= top_:
                visit(sx, top, 0, dist)
            if left >= left_:
                visit(left, sy, 0, dist)
            if right <= right_:
                visit(right, sy, 0, dist)
            if bottom <= bottom_:
                visit(sx, bottom, 0, dist)
            for index in range(1, 4):
                yt, yb = top + (dist >> 2) * index, bottom - (dist >> 2) * index
                xl, xr = sx - (dist >> 2) * index, sx + (dist >> 2) * index
                if yt >= top_:
                    if xl >= left_:
                        visit(xl, yt, 0, dist)
                    if xr <= right_:
                        visit(xr, yt, 0, dist)
                if yb <= bottom_:
                    if xl >= left_:
                        visit(xl, yb, 0, dist)
                    if xr <= right_:
                        visit(xr, yb, 0, dist)


X_OFF = ((0, -1, -1, 0, -1, +1, -1, -1, +1), (0, 0, +1, +1, -1, +1, 0, +1, 0))
Y_OFF = ((0, 0, -1, -1, +1, -1, 0, +1, 0), (0, -1, -1, 0, -1, +1, +1, +1, +1))


def two_point(sr, bx, by, nr, visit):
    left_, right_, top_, bottom_ = sr
    for k in range(2):
        x, y = bx + X_OFF[k][nr], by + Y_OFF[k][nr]
        if left_ <= x <= right_ and top_ <= y <= bottom_:
            visit(x, y, 0, 2)


def raster(sr, step, visit):
    left_, right_, top_, bottom_ = sr
    y = top_
    while y <= bottom_:
        x = left_
        while x <= right_:
            visit(x, y, 0, step)
            x += step
        y += step


def replay(points, dist, mv_cost, best):
    """xTZSearchHelp, non-subShiftMode-1 branch; best = dict(sad, x, y, distance, round, nr)."""
    for (x, y, nr, d), sad in zip(points, dist):
        if sad < best["sad"]:
            sad += mv_cost(x, y)
            if sad < best["sad"]:
                best.update(sad=sad, x=x, y=y, distance=d, round=0, nr=nr)
    return best
