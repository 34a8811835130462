"""CPU: the metrics oracle against the committed known-answer table and its own internal cross-checks."""
import json
import os

import numpy as np
import pytest

import algo_mirror as am
from oracle import metrics_oracle as mo
from test_gpu_metrics import FLOAT_KEYS, INT_KEYS, kat_masks

HERE = os.path.dirname(os.path.abspath(__file__))


def test_known_answer_table_matches_golden():
    with open(os.path.join(HERE, "golden", "metrics_kat.json")) as f:
        golden = json.load(f)
    rng = np.random.RandomState(golden["image_seed"])
    image = rng.randint(0, 256, size=(64, 64, 3)).astype(np.uint8)
    for m, ref in zip(kat_masks(), golden["rows"]):
        got = mo.calculate_metrics(image, m)
        assert list(got.keys()) == list(ref.keys())
        for k in INT_KEYS:
            assert got[k] == ref[k]
        for k in FLOAT_KEYS:
            assert got[k] == pytest.approx(ref[k], rel=1e-12, abs=1e-15)


def test_hand_derivable_answers():
    """rows derivable by hand from the library rules."""
    img = np.full((64, 64, 3), 100, np.uint8)
    m = np.zeros((64, 64), bool); m[5:15, 7:17] = 1
    r = mo.calculate_metrics(img, m)
    assert (r["area"], r["perimeter"], r["convex_hull_area"], r["convex_hull_perimeter"]) == (100, 36.0, 100, 36.0)
    assert (r["min_x"], r["min_y"], r["max_x"], r["max_y"]) == (5, 7, 15, 17)
    assert r["circularity"] == pytest.approx(2 * np.sqrt(np.pi * 100) / 36)
    assert r["mean_brightness"] == 100.0 and r["brightness_std"] == 0.0
    m = np.zeros((64, 64), bool); m[10, 10] = 1
    r = mo.calculate_metrics(img, m)
    assert (r["area"], r["perimeter"], r["convex_hull_area"], r["deformability"]) == (1, 0.0, 1, 1.0)
    m = np.zeros((64, 64), bool); m[0, 0] = 1           # 2 contour points -> QhullError branch
    r = mo.calculate_metrics(img, m)
    assert (r["convex_hull_area"], r["convex_hull_perimeter"], r["circularity"]) == (0, 0.0, 0.0)
    m = np.zeros((64, 64), bool); m[2, 3] = 1; m[20:40, 20:40] = 1   # contours[0] is the raster-first speck
    r = mo.calculate_metrics(img, m)
    assert (r["area"], r["convex_hull_area"]) == (401, 1)
    with pytest.raises(IndexError):
        mo.calculate_metrics(img, np.zeros((64, 64), bool))


def test_first_contour_equals_full_assembly_and_literal_raster():
    rng = np.random.RandomState(1)
    for trial in range(60):
        H, W = rng.randint(3, 30), rng.randint(3, 30)
        mask = rng.rand(H, W) < rng.choice([0.1, 0.5, 0.9])
        full = mo.find_contours(mask)
        first = mo.first_contour(mask)
        if not full:
            assert first is None
            continue
        assert np.array_equal(full[0], first)
        try:
            from scipy.spatial import ConvexHull
            hv = first[ConvexHull(first).vertices]
        except Exception:
            continue
        poly = np.vstack((hv, hv[0]))
        assert np.array_equal(mo.polygon2mask(mask.shape, poly), mo.polygon2mask(mask.shape, poly, literal=True))


def test_gpu_algorithm_mirror_equals_oracle():
    """The integer algorithm of csrc/postproc.cu (python mirror) == find_contours -> Qhull -> polygon2mask."""
    from scipy import ndimage as ndi
    rng = np.random.RandomState(0)
    img = np.full((1, 1, 3), 0, np.uint8)
    masks = []
    for trial in range(150):
        H, W = rng.randint(2, 40), rng.randint(2, 40)
        masks.append(rng.rand(H, W) < rng.choice([0.05, 0.2, 0.5, 0.8, 0.95]))
    for trial in range(80):
        H, W = rng.randint(8, 80), rng.randint(8, 80)
        rr, cc = np.ogrid[:H, :W]
        cy, cx = rng.uniform(-5, H + 5), rng.uniform(-5, W + 5)
        a, b = rng.uniform(1, H / 2), rng.uniform(1, W / 2)
        masks.append(((rr - cy) / a) ** 2 + ((cc - cx) / b) ** 2 <= 1)
        masks.append(ndi.gaussian_filter(rng.standard_normal((H, W)), rng.uniform(1, 5)) > 0)
    for mask in masks:
        if not mask.any():
            continue
        ref = mo.calculate_metrics(np.broadcast_to(img, mask.shape + (3,)), mask, extras=True)
        got = am.hull_stats(mask)
        assert got["hull_area"] == ref["convex_hull_area"]
        assert got["hull_perim_hist"] == ref["_hull_perim_hist"]
        assert got["degenerate"] == ref["_hull_degenerate"]
        assert am.perim_hist(mask) == ref["_perim_hist"]


def _upstream():
    with open(os.path.join(HERE, "golden", "skimage_upstream_vectors.json")) as f:
        return json.load(f)


def test_find_contours_matches_upstream_skimage_vector():
    """scikit-image's own test_find_contours.py::test_binary: one closed contour around an L-shaped hole, vertex by vertex
    (start point, direction and closing point included) -- pins the case table, the assembly order and the orientation of
    the restated marching squares against something this repo did not generate."""
    v = _upstream()["find_contours_binary"]
    a = np.ones((8, 8), dtype=np.float32)
    a[1:-1, 1] = 0
    a[1, 1:-1] = 0
    contours = mo.find_contours(a > v["level"])
    assert len(contours) == v["n_contours"]
    assert np.array_equal(contours[0], np.array(v["contour0_default_orientation"]))
    assert np.array_equal(mo.first_contour(a > v["level"]), np.array(v["contour0_default_orientation"]))


def test_polygon2mask_matches_upstream_skimage_vectors():
    """scikit-image's test_polygon2mask (pixel count of a concave 8-gon) and test_draw.py's rectangle cases (edges and
    vertices are inside; clipping to the shape) through the restated point_in_polygon."""
    v = _upstream()
    p = v["polygon2mask"]
    m = mo.polygon2mask(tuple(p["shape"]), np.array(p["polygon"], float), literal=True)
    assert m.shape == tuple(p["shape"]) and int(m.sum()) == p["mask_sum"]
    for key in ("polygon_rectangle", "polygon_exceed"):
        c = v[key]
        got = mo.polygon2mask(tuple(c["shape"]), np.array(c["polygon"], float), literal=True)
        exp = np.zeros(tuple(c["shape"]), bool)
        exp[c["filled_rows"][0]:c["filled_rows"][1], c["filled_cols"][0]:c["filled_cols"][1]] = True
        assert np.array_equal(got, exp), key
    # the vectorised convex form used for the hull raster agrees with the literal loop on the convex rectangle
    c = v["polygon_rectangle"]
    assert np.array_equal(mo.polygon2mask(tuple(c["shape"]), np.array(c["polygon"][:-1], float)),
                          mo.polygon2mask(tuple(c["shape"]), np.array(c["polygon"], float), literal=True))
