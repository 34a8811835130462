"""Synthetic inputs for BASELINE.json's configs.

Datasets and YOLO weights are not available offline, so every config is driven with generated
brightfield-like frames: a noisy mid-grey background (statistics of the reference's example PNGs
under examples/example_image/: mean ~139, range 44..240) with darker ellipses standing in for cells,
and the ellipse bounding boxes (jittered) standing in for the YOLO detections that
pipeline.py:84-87 would return (float32 xyxy in original-image pixels).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def synth_image(index: int, size: int = 1024, n_boxes: int = 1,
                bit_depth: int = 8) -> Tuple[np.ndarray, np.ndarray]:
    """Return (gray image [size,size] uint8|uint16, boxes float32 [n_boxes,4] xyxy).

    Image ``index`` is a pure function of (index, size, n_boxes): RandomState(1000 + index).
    n_boxes == 1 : one ellipse, semi-axes U[30,70]*(size/1024), centre U[200,824]*(size/1024).
    n_boxes  > 1 : non-overlapping ellipses on a jittered ceil(sqrt(n))^2 grid.
    """
    rng = np.random.RandomState(1000 + index)
    s = size / 1024.0
    bg = 139.0 + 8.0 * rng.standard_normal((size, size))
    yy, xx = np.mgrid[0:size, 0:size]
    boxes = np.zeros((n_boxes, 4), np.float32)
    if n_boxes == 1:
        cells = [(rng.uniform(200, 824) * s, rng.uniform(200, 824) * s,
                  rng.uniform(30, 70) * s, rng.uniform(30, 70) * s)]
    else:
        g = int(np.ceil(np.sqrt(n_boxes)))
        pitch = size / g
        cells = []
        for k in range(n_boxes):
            gy, gx = divmod(k, g)
            ry = rng.uniform(0.18, 0.32) * pitch
            rx = rng.uniform(0.18, 0.32) * pitch
            cy = (gy + 0.5) * pitch + rng.uniform(-0.12, 0.12) * pitch
            cx = (gx + 0.5) * pitch + rng.uniform(-0.12, 0.12) * pitch
            cells.append((cy, cx, ry, rx))
    for k, (cy, cx, ry, rx) in enumerate(cells):
        inside = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        bg[inside] -= 60.0
        j = rng.uniform(0, 4, size=4) * s
        boxes[k] = [max(cx - rx - j[0], 0), max(cy - ry - j[1], 0),
                    min(cx + rx + j[2], size - 1), min(cy + ry + j[3], size - 1)]
    img8 = np.clip(np.rint(bg), 0, 255).astype(np.uint8)
    if bit_depth == 16:
        return (img8.astype(np.uint16) * 257), boxes
    return img8, boxes


def gray_to_rgb_u8(gray: np.ndarray) -> np.ndarray:
    """What pipeline.py:206-210 hands to the SAM stage: cv2.imread default flags reduce 16-bit to
    8-bit (v >> 8) and replicate grey to three channels; BGR->RGB is a no-op on grey."""
    if gray.dtype == np.uint16:
        gray = (gray >> 8).astype(np.uint8)
    return np.ascontiguousarray(np.repeat(gray[:, :, None], 3, axis=2))


def synth_batch(start: int, count: int, size: int = 1024, n_boxes: int = 1
                ) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    imgs, boxes = [], []
    for i in range(start, start + count):
        g, b = synth_image(i, size, n_boxes)
        imgs.append(gray_to_rgb_u8(g))
        boxes.append(b)
    return imgs, boxes
