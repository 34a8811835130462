"""CPU ORACLE (test infrastructure, never the product path) -- morphometrics half of the hot path.

Line-by-line restatement of ``calculate_metrics`` in
/root/reference/src/yolo_sam_inference/utils/metrics.py:9-119.  The arithmetic of that function lives
in two third-party libraries:

* scikit-image (``scikit-image>=0.19.0``, requirements.txt:7; unpinned, NOT installed in this image and
  not installable offline): ``measure.regionprops`` (area, bbox, centroid, perimeter),
  ``measure.find_contours`` and ``draw.polygon2mask``.  Their published algorithms are restated below
  with numpy + the very ``scipy.ndimage`` calls skimage itself makes:
    - ``perimeter``      skimage/measure/_regionprops_utils.py ``perimeter(image, neighborhood=4)``
    - ``find_contours``  skimage/measure/_find_contours.py + _find_contours_cy.pyx (marching squares,
                         level 0.5, fully_connected='low', positive_orientation='low') and
                         ``_assemble_contours``
    - ``polygon2mask``   skimage/draw/_polygon2mask.py -> draw.polygon -> _draw.pyx ``_polygon`` ->
                         _shared/geometry.pxd ``point_in_polygon`` (O'Rourke; vertex/edge inclusive)
* SciPy/Qhull (scipy 1.18.1 IS installed): the real ``scipy.spatial.ConvexHull`` is called.

PARITY UNPINNED: the reference has no tests, fixtures or golden vectors, and skimage
cannot be imported here to cross-check the restated parts; the known-answer table in
tests/golden/metrics_kat.json is self-generated (first five rows derivable by hand).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.
"""
from __future__ import annotations

from collections import deque
from math import ceil, sqrt
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
from scipy import ndimage as ndi
from scipy.spatial import ConvexHull

# codes with non-zero weight in skimage's perimeter(), in increasing order
PERIM_CODES = (5, 7, 13, 15, 17, 21, 23, 25, 27, 33)
_STREL_4 = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=np.uint8)


def _perimeter_weights() -> np.ndarray:
    w = np.zeros(50, dtype=np.float64)
    w[[5, 7, 15, 17, 25, 27]] = 1
    w[[21, 33]] = sqrt(2)
    w[[13, 23]] = (1 + sqrt(2)) / 2
    return w


def perimeter_histogram(region_image: np.ndarray) -> np.ndarray:
    """skimage.measure.perimeter(image, neighborhood=4) up to (not including) the final dot product:
    returns the 50-bin histogram of border-pixel neighbourhood codes."""
    image = region_image.astype(np.uint8)
    eroded = ndi.binary_erosion(image, _STREL_4, border_value=0)
    border = image - eroded
    code = ndi.convolve(border, np.array([[10, 2, 10], [2, 1, 2], [10, 2, 10]]), mode="constant", cval=0)
    return np.bincount(code.ravel(), minlength=50)


def perimeter_from_histogram(hist50: np.ndarray) -> float:
    return float(np.asarray(hist50) @ _perimeter_weights())


def perimeter(region_image: np.ndarray) -> float:
    return perimeter_from_histogram(perimeter_histogram(region_image))


# ---------------------------------------------------------------------------------------------------
# find_contours restatement
# ---------------------------------------------------------------------------------------------------

def _contour_segments(mask: np.ndarray) -> List[Tuple[Tuple[float, float], Tuple[float, float]]]:
    """_get_contour_segments for a binary image at level 0.5, vertex_connect_high=False.
    Squares are visited in raster order (r0 outer, c0 inner); only mixed squares emit segments."""
    m = mask.astype(np.uint8)
    H, W = m.shape
    if H < 2 or W < 2:
        return []
    ul, ur, ll, lr = m[:-1, :-1], m[:-1, 1:], m[1:, :-1], m[1:, 1:]
    case = (ul * 1 + ur * 2 + ll * 4 + lr * 8).astype(np.uint8)
    rs, cs = np.nonzero((case != 0) & (case != 15))      # raster order
    segs = []
    for r0, c0, sc in zip(rs.tolist(), cs.tolist(), case[rs, cs].tolist()):
        r1, c1 = r0 + 1, c0 + 1
        # binary image, level 0.5 => every crossing is the exact edge midpoint
        top = (float(r0), c0 + 0.5)
        bottom = (float(r1), c0 + 0.5)
        left = (r0 + 0.5, float(c0))
        right = (r0 + 0.5, float(c1))
        if sc == 1:
            segs.append((top, left))
        elif sc == 2:
            segs.append((right, top))
        elif sc == 3:
            segs.append((right, left))
        elif sc == 4:
            segs.append((left, bottom))
        elif sc == 5:
            segs.append((top, bottom))
        elif sc == 6:
            segs.append((right, top))
            segs.append((left, bottom))
        elif sc == 7:
            segs.append((right, bottom))
        elif sc == 8:
            segs.append((bottom, right))
        elif sc == 9:
            segs.append((top, left))
            segs.append((bottom, right))
        elif sc == 10:
            segs.append((bottom, top))
        elif sc == 11:
            segs.append((bottom, left))
        elif sc == 12:
            segs.append((left, right))
        elif sc == 13:
            segs.append((top, right))
        elif sc == 14:
            segs.append((left, top))
    return segs


def _assemble_contours(segments) -> List[np.ndarray]:
    current_index = 0
    contours: Dict[int, deque] = {}
    starts: Dict[Tuple[float, float], Tuple[deque, int]] = {}
    ends: Dict[Tuple[float, float], Tuple[deque, int]] = {}
    for from_point, to_point in segments:
        if from_point == to_point:
            continue
        tail, tail_num = starts.pop(to_point, (None, None))
        head, head_num = ends.pop(from_point, (None, None))
        if tail is not None and head is not None:
            if tail is head:
                head.append(to_point)
            else:
                if tail_num > head_num:
                    head.extend(tail)
                    contours.pop(tail_num, None)
                    starts[head[0]] = (head, head_num)
                    ends[head[-1]] = (head, head_num)
                else:
                    tail.extendleft(reversed(head))
                    starts.pop(head[0], None)
                    contours.pop(head_num, None)
                    starts[tail[0]] = (tail, tail_num)
                    ends[tail[-1]] = (tail, tail_num)
        elif tail is None and head is None:
            new_contour = deque((from_point, to_point))
            contours[current_index] = new_contour
            starts[from_point] = (new_contour, current_index)
            ends[to_point] = (new_contour, current_index)
            current_index += 1
        elif head is None:
            tail.appendleft(from_point)
            starts[from_point] = (tail, tail_num)
        else:
            head.append(to_point)
            ends[to_point] = (head, head_num)
    return [np.array(c) for _, c in sorted(contours.items())]


def find_contours(mask: np.ndarray) -> List[np.ndarray]:
    """skimage.measure.find_contours(mask.astype(int), 0.5) (metrics.py:31)."""
    return _assemble_contours(_contour_segments(mask))


def first_contour(mask: np.ndarray) -> Optional[np.ndarray]:
    """contours[0] (metrics.py:34) without assembling every contour of a noisy mask: the surviving
    index-0 contour is the chain that contains the first emitted segment, so it is enough to assemble
    the segments of the connected chain reachable from it.  Equivalent to find_contours(mask)[0]
    (asserted in tests/test_oracle_metrics.py)."""
    segs = _contour_segments(mask)
    if not segs:
        return None
    by_point: Dict[Tuple[float, float], List[int]] = {}
    for i, (a, b) in enumerate(segs):
        by_point.setdefault(a, []).append(i)
        by_point.setdefault(b, []).append(i)
    seen = {0}
    stack = [0]
    while stack:
        i = stack.pop()
        for p in segs[i]:
            for j in by_point[p]:
                if j not in seen:
                    seen.add(j)
                    stack.append(j)
    chain = [segs[i] for i in sorted(seen)]
    return _assemble_contours(chain)[0]


# ---------------------------------------------------------------------------------------------------
# polygon2mask restatement
# ---------------------------------------------------------------------------------------------------

def _point_in_polygon(xp: np.ndarray, yp: np.ndarray, x: float, y: float) -> int:
    """skimage/_shared/geometry.pxd point_in_polygon: 0 outside, 1 inside, 2 vertex, 3 edge."""
    n = len(xp)
    eps = 1e-12
    l_cross = r_cross = 0
    x1 = xp[n - 1] - x
    y1 = yp[n - 1] - y
    for i in range(n):
        x0 = xp[i] - x
        y0 = yp[i] - y
        if (-eps < x0 < eps) and (-eps < y0 < eps):
            return 2
        if (y0 > 0) != (y1 > 0):
            if ((x0 * y1 - x1 * y0) / (y1 - y0)) > 0:
                r_cross += 1
        if (y0 < 0) != (y1 < 0):
            if ((x0 * y1 - x1 * y0) / (y1 - y0)) < 0:
                l_cross += 1
        x1, y1 = x0, y0
    if (r_cross & 1) != (l_cross & 1):
        return 3
    if r_cross & 1:
        return 1
    return 0


def polygon2mask(shape: Tuple[int, int], polygon: np.ndarray, literal: bool = False) -> np.ndarray:
    """skimage.draw.polygon2mask(shape, polygon) for a CONVEX polygon given as (row, col) vertices.

    literal=True runs the restated per-pixel point_in_polygon loop of _draw.pyx::_polygon (slow; used
    by the tests on small cases).  The default evaluates the same predicate vectorised: for a convex
    polygon "inside, on an edge or on a vertex" == all edge cross products have one sign or are zero,
    and with half-integer vertices and integer pixel centres every product is exact in float64."""
    r = np.asarray(polygon[:, 0], np.float64)
    c = np.asarray(polygon[:, 1], np.float64)
    minr = int(max(0, r.min()))
    maxr = int(ceil(r.max()))
    minc = int(max(0, c.min()))
    maxc = int(ceil(c.max()))
    maxr = min(shape[0] - 1, maxr)
    maxc = min(shape[1] - 1, maxc)
    out = np.zeros(shape, dtype=bool)
    if maxr < minr or maxc < minc:
        return out
    if literal:
        for ri in range(minr, maxr + 1):
            for ci in range(minc, maxc + 1):
                if _point_in_polygon(c, r, float(ci), float(ri)):
                    out[ri, ci] = True
        return out
    rr, cc = np.mgrid[minr:maxr + 1, minc:maxc + 1]
    rr = rr.astype(np.float64)
    cc = cc.astype(np.float64)
    pos = np.ones(rr.shape, bool)
    neg = np.ones(rr.shape, bool)
    n = len(r)
    for i in range(n):
        j = (i + 1) % n
        er, ec = r[j] - r[i], c[j] - c[i]
        if er == 0 and ec == 0:
            continue
        cross = er * (cc - c[i]) - ec * (rr - r[i])
        pos &= cross >= 0
        neg &= cross <= 0
    out[minr:maxr + 1, minc:maxc + 1] = pos | neg
    return out


# ---------------------------------------------------------------------------------------------------
# regionprops restatement (single label: every True pixel is one region, metrics.py:28)
# ---------------------------------------------------------------------------------------------------

class _Props:
    def __init__(self, mask: np.ndarray):
        rows, cols = np.nonzero(mask)
        if rows.size == 0:
            raise IndexError("list index out of range")    # regionprops(...)[0] on an empty label image
        self.area = int(rows.size)
        self.bbox = (int(rows.min()), int(cols.min()), int(rows.max()) + 1, int(cols.max()) + 1)
        self.centroid = (float(rows.mean()), float(cols.mean()))
        r0, c0, r1, c1 = self.bbox
        self.image = mask[r0:r1, c0:c1]

    @property
    def perimeter_hist(self) -> np.ndarray:
        return perimeter_histogram(self.image)

    @property
    def perimeter(self) -> float:
        return perimeter_from_histogram(self.perimeter_hist)


def calculate_metrics(image: np.ndarray, mask: np.ndarray, extras: bool = False) -> Dict[str, Any]:
    """metrics.py:9-119.  With extras=True the dict additionally carries the integer intermediates the
    CUDA kernels are compared with bit-for-bit (prefixed '_')."""
    if mask.ndim > 2:
        mask = mask.squeeze()
    mask = mask.astype(bool)
    assert mask.shape == image.shape[:2], \
        f"Mask shape {mask.shape} does not match image shape {image.shape[:2]}"

    props = _Props(mask)                                                         # :28
    contour = first_contour(mask)                                                # :31-34
    convex_props = None
    hull_vertices = np.zeros((0, 2))
    if contour is not None:
        try:
            hull = ConvexHull(contour)                                           # :36
            convex_hull_coords = contour[hull.vertices]                          # :38
            convex_hull_coords = np.vstack((convex_hull_coords, convex_hull_coords[0]))   # :40
            polygon_coords = np.column_stack((convex_hull_coords[:, 0], convex_hull_coords[:, 1]))
            convex_hull_mask = polygon2mask(mask.shape, polygon_coords)          # :46
            convex_props = _Props(convex_hull_mask)                              # :48
            hull_vertices = contour[hull.vertices]
        except Exception:                                                        # :52-56
            convex_props = None

    area = props.area                                                            # :62
    perimeter_ = props.perimeter                                                 # :65
    convex_hull_area = convex_props.area if convex_props else 0                  # :68
    convex_hull_perimeter = convex_props.perimeter if convex_props else 0        # :69
    area_ratio = convex_hull_area / area if area > 0 else 0                      # :72
    circularity = (2 * np.sqrt(np.pi * convex_hull_area)) / convex_hull_perimeter \
        if convex_hull_perimeter > 0 else 0                                      # :75
    deformability = 1 - circularity                                              # :78

    brightness_image = np.mean(image, axis=2)                                    # :81
    center_radius = int(min(mask.shape) * 0.1)                                   # :84-85
    center_x, center_y = props.centroid                                          # :87
    rr, cc = np.ogrid[:mask.shape[0], :mask.shape[1]]
    center_region_mask = (rr - center_x) ** 2 + (cc - center_y) ** 2 <= center_radius ** 2   # :89
    center_brightness = brightness_image[center_region_mask]                     # :92
    mean_brightness = np.mean(center_brightness) if center_brightness.size > 0 else 0
    brightness_std = np.std(center_brightness) if center_brightness.size > 0 else 0

    min_x, min_y, max_x, max_y = props.bbox                                      # :97
    aspect_ratio = (max_x - min_x) / (max_y - min_y) if (max_x - min_x) > 0 and (max_y - min_y) > 0 else 0
    mask_x_length = max_x - min_x
    mask_y_length = max_y - min_y

    out = {
        "deformability": float(deformability),
        "area": int(area),
        "area_ratio": float(area_ratio),
        "circularity": float(circularity),
        "convex_hull_area": int(convex_hull_area),
        "mask_x_length": int(mask_x_length),
        "mask_y_length": int(mask_y_length),
        "min_x": int(min_x),
        "min_y": int(min_y),
        "max_x": int(max_x),
        "max_y": int(max_y),
        "mean_brightness": float(mean_brightness),
        "brightness_std": float(brightness_std),
        "perimeter": float(perimeter_),
        "aspect_ratio": float(aspect_ratio),
        "convex_hull_perimeter": float(convex_hull_perimeter),
    }
    if extras:
        rows, cols = np.nonzero(mask)
        s = image.astype(np.int64).sum(axis=2)
        disk = s[center_region_mask]
        gray_floor = (s // 3)[mask]
        out["_perim_hist"] = [int(props.perimeter_hist[c]) for c in PERIM_CODES]
        out["_hull_perim_hist"] = ([int(convex_props.perimeter_hist[c]) for c in PERIM_CODES]
                                   if convex_props else [0] * len(PERIM_CODES))
        out["_sum_r"] = int(rows.sum())
        out["_sum_c"] = int(cols.sum())
        out["_disk_n"] = int(disk.size)
        out["_disk_sum"] = int(disk.sum())
        out["_disk_sumsq"] = int((disk * disk).sum())
        out["_hull_degenerate"] = convex_props is None
        out["_hull_vertices"] = hull_vertices.tolist()
        out["_contour_len"] = 0 if contour is None else int(len(contour))
        out["_mask_hist"] = np.bincount(gray_floor, minlength=256).astype(np.int64).tolist()
    return out
