"""Pure-Python mirror of the integer algorithm used by csrc/postproc.cu (contour_hull_disk_kernel).

Test infrastructure only: it lets the CPU suite check, without a GPU, that "trace contour 0 ->
row spans -> monotone-chain hull -> inclusive interval raster" is equivalent to the oracle's
find_contours -> scipy ConvexHull -> polygon2mask on arbitrary masks. The CUDA kernel is then checked
against the oracle itself on the GPU (tests/test_gpu_metrics.py).
"""
import numpy as np

E_T, E_B, E_L, E_R, E_NONE = 0, 1, 2, 3, 4
CODE_BINS = (5, 7, 13, 15, 17, 21, 23, 25, 27, 33)


def paired_edge(cs, e):
    ul, ur, ll, lr = cs & 1, (cs >> 1) & 1, (cs >> 2) & 1, (cs >> 3) & 1
    if cs == 6:
        return {E_T: E_R, E_R: E_T, E_L: E_B, E_B: E_L}[e]
    if cs == 9:
        return {E_T: E_L, E_L: E_T, E_B: E_R, E_R: E_B}[e]
    ct, cb, cl, cr = ul != ur, ll != lr, ul != ll, ur != lr
    if ct and e != E_T:
        return E_T
    if cb and e != E_B:
        return E_B
    if cl and e != E_L:
        return E_L
    if cr and e != E_R:
        return E_R
    return E_NONE


FIRST_SEG = {1: (E_T, E_L), 2: (E_R, E_T), 3: (E_R, E_L), 4: (E_L, E_B), 5: (E_T, E_B), 6: (E_R, E_T),
             7: (E_R, E_B), 8: (E_B, E_R), 9: (E_T, E_L), 10: (E_B, E_T), 11: (E_B, E_L), 12: (E_L, E_R),
             13: (E_T, E_R), 14: (E_L, E_T)}


def edge_point(r0, c0, e):
    return (2 * r0 + (0 if e == E_T else 2 if e == E_B else 1),
            2 * c0 + (0 if e == E_L else 2 if e == E_R else 1))


def hull_stats(mask):
    """Returns dict(hull_area, hull_perim_hist[10], degenerate, contour_points) the way the kernel does."""
    m = mask.astype(np.uint8)
    H, W = m.shape
    case = None
    first = None
    if H >= 2 and W >= 2:
        case = m[:-1, :-1] + 2 * m[:-1, 1:] + 4 * m[1:, :-1] + 8 * m[1:, 1:]
        rs, cs = np.nonzero((case != 0) & (case != 15))
        if rs.size:
            first = (int(rs[0]), int(cs[0]))
    span = {}
    npts = 0

    def add(p):
        nonlocal npts
        y, x = p
        lo, hi = span.get(y, (10 ** 9, -10 ** 9))
        span[y] = (min(lo, x), max(hi, x))
        npts += 1

    def walk(r0, c0, e, stop):
        while True:
            nr, nc = r0, c0
            if e == E_T:
                nr, ne = r0 - 1, E_B
            elif e == E_B:
                nr, ne = r0 + 1, E_T
            elif e == E_L:
                nc, ne = c0 - 1, E_R
            else:
                nc, ne = c0 + 1, E_L
            if nr < 0 or nc < 0 or nr > H - 2 or nc > W - 2:
                return False
            ex = paired_edge(int(case[nr, nc]), ne)
            p = edge_point(nr, nc, ex)
            if p == stop:
                return True
            add(p)
            r0, c0, e = nr, nc, ex

    if first is not None:
        r0, c0 = first
        ef, et = FIRST_SEG[int(case[r0, c0])]
        pf, pt = edge_point(r0, c0, ef), edge_point(r0, c0, et)
        add(pf)
        add(pt)
        if not walk(r0, c0, et, pf):
            walk(r0, c0, ef, pt)
    out = dict(hull_area=0, hull_perim_hist=[0] * 10, degenerate=True, contour_points=npts)
    if npts == 0:
        return out
    ys = sorted(span)
    lch, rch = [], []
    for y in ys:
        lo, hi = span[y]
        while len(lch) >= 2:
            (ay, ax), (by, bx) = lch[-2], lch[-1]
            if (bx - ax) * (y - ay) - (lo - ax) * (by - ay) >= 0:
                lch.pop()
            else:
                break
        lch.append((y, lo))
        while len(rch) >= 2:
            (ay, ax), (by, bx) = rch[-2], rch[-1]
            if (bx - ax) * (y - ay) - (hi - ax) * (by - ay) <= 0:
                rch.pop()
            else:
                break
        rch.append((y, hi))
    poly = lch + rch[::-1]
    area2 = 0
    for i in range(len(poly)):
        (y0, x0), (y1, x1) = poly[i], poly[(i + 1) % len(poly)]
        area2 += x0 * y1 - x1 * y0
    if npts < 3 or area2 == 0:
        return out
    out["degenerate"] = False
    ymin, ymax = ys[0], ys[-1]
    pr0, pr1 = (ymin + 1) >> 1, ymax >> 1
    hull = np.zeros((H, W), bool)
    for r in range(pr0, pr1 + 1):
        y = 2 * r
        k = 0
        while k + 1 < len(lch) and lch[k + 1][0] < y:
            k += 1
        ay, ax = lch[k]
        if k + 1 < len(lch) and ay != y:
            by, bx = lch[k + 1]
            den, num = 2 * (by - ay), ax * (by - ay) + (bx - ax) * (y - ay)
            lo = -((-num) // den)
        else:
            lo = (ax + 1) >> 1
        k = 0
        while k + 1 < len(rch) and rch[k + 1][0] < y:
            k += 1
        ay, ax = rch[k]
        if k + 1 < len(rch) and ay != y:
            by, bx = rch[k + 1]
            den, num = 2 * (by - ay), ax * (by - ay) + (bx - ax) * (y - ay)
            hi = num // den
        else:
            hi = ax >> 1
        lo, hi = max(lo, 0), min(hi, W - 1)
        if hi >= lo:
            hull[r, lo:hi + 1] = True
    out["hull_area"] = int(hull.sum())
    if out["hull_area"] == 0:      # regionprops(hull_mask)[0] raises -> except branch, metrics.py:52-56
        out["degenerate"] = True
    out["hull_perim_hist"] = perim_hist(hull)
    out["hull_mask"] = hull
    return out


def perim_hist(mask):
    """Per-pixel formulation used by both CUDA kernels."""
    m = np.pad(mask.astype(np.uint8), 2)
    er = m[1:-1, 1:-1] & m[:-2, 1:-1] & m[2:, 1:-1] & m[1:-1, :-2] & m[1:-1, 2:]
    b = np.pad(m[1:-1, 1:-1] & (1 - er), 1)          # back to pad 2
    n4 = b[:-2, 1:-1] + b[2:, 1:-1] + b[1:-1, :-2] + b[1:-1, 2:]
    nd = b[:-2, :-2] + b[:-2, 2:] + b[2:, :-2] + b[2:, 2:]
    code = (1 + 2 * n4 + 10 * nd)[b[1:-1, 1:-1] == 1]
    return [int((code == c).sum()) for c in CODE_BINS]
