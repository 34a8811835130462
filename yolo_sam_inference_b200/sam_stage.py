"""SamStage: host-side mirror of the reference's SAM stage, executing on libysi.so (sm_100a CUDA).

Replaces, for one image, the loop body of CellSegmentationPipeline.process_single_image
(/root/reference/src/yolo_sam_inference/pipeline.py:161-175): SamProcessor preprocessing, per-box
SamModel forward (:89-124), post_process_masks + threshold, and calculate_metrics
(utils/metrics.py:9-119).  Inputs and outputs keep the reference's types:

    run(image uint8[H,W,3] RGB, boxes float32[N,4] xyxy)
        -> masks bool[N,H,W], metrics list[dict] (the 16 keys / python types of metrics.py:102-119),
           crops list[uint8[h,w,3]] (image[y1:y2, x1:x2] with box.astype(int), pipeline.py:379)

There is no CPU or PyTorch fallback: construction fails if the CUDA library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
from math import sqrt
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as nat
from .ingest import as_rgb_u8, pixel_format_of
from .weights import SamVariant, variant_of


def _perimeter_weights() -> np.ndarray:
    # skimage.measure.perimeter weights (see oracle/metrics_oracle.py for the derivation)
    w = np.zeros(50, dtype=np.float64)
    w[[5, 7, 15, 17, 25, 27]] = 1
    w[[21, 33]] = sqrt(2)
    w[[13, 23]] = (1 + sqrt(2)) / 2
    return w


_PERIM_W = _perimeter_weights()


def _perimeter_from_bins(bins: np.ndarray) -> float:
    hist = np.zeros(50, dtype=np.int64)
    hist[list(nat.PERIM_CODES)] = bins.astype(np.int64)
    return float(hist @ _PERIM_W)          # the very reduction skimage performs


_PERIM_W10 = np.array([_PERIM_W[c] for c in nat.PERIM_CODES], dtype=np.float64)
_ZERO_ROW = {"deformability": 1.0, "area": 0, "area_ratio": 0.0, "circularity": 0.0, "convex_hull_area": 0,
             "mask_x_length": 0, "mask_y_length": 0, "min_x": 0, "min_y": 0, "max_x": 0, "max_y": 0,
             "mean_brightness": 0.0, "brightness_std": 0.0, "perimeter": 0.0, "aspect_ratio": 0.0,
             "convex_hull_perimeter": 0.0}
_PI = float(np.pi)


def metrics_from_rows(rows: np.ndarray, on_empty: str = "raise") -> List[Dict[str, Any]]:
    """ysi_mask_metrics rows -> the dicts of utils/metrics.py:102-119 (same keys, order and python types).

    Every integer comes straight from the CUDA kernels; the float64 scalars are formed here with the reference's own
    formulas (:62-100).  An empty mask raises IndexError like metrics.py:28 does (``on_empty="raise"``, parity) or yields an
    all-zero row (``on_empty="zeros"``).  Works column-wise on the whole batch (one numpy -> python conversion per field);
    the arithmetic per row is exactly the scalar one:
      * perimeters: the 50-bin dot product skimage forms (``hist @ weights``, zero bins included) -- same call, same order
      * mean / std of the centre disk: exact integer sums from the device, one correctly rounded division each (python's
        int / int), which is what numpy's float64 mean / std agree with to 1e-9 (they sum pairwise)
    """
    n = len(rows)
    if n == 0:
        return []
    hist = np.zeros((n, 50), dtype=np.float64)
    hist[:, list(nat.PERIM_CODES)] = rows["perim_hist"]
    hull_hist = np.zeros((n, 50), dtype=np.float64)
    hull_hist[:, list(nat.PERIM_CODES)] = rows["hull_perim_hist"]
    area = rows["area"].tolist()
    flags = rows["flags"].tolist()
    hull_area = rows["hull_area"].tolist()
    disk_n, disk_s1, disk_s2 = rows["disk_n"].tolist(), rows["disk_sum"].tolist(), rows["disk_sumsq"].tolist()
    min_r, min_c, max_r, max_c = rows["min_r"].tolist(), rows["min_c"].tolist(), rows["max_r"].tolist(), rows["max_c"].tolist()
    out: List[Dict[str, Any]] = []
    for i in range(n):
        a = area[i]
        if a == 0:
            if on_empty == "raise":
                raise IndexError("list index out of range")     # regionprops(mask)[0] on an empty mask
            out.append(dict(_ZERO_ROW))
            continue
        degenerate = bool(flags[i] & nat.FLAG_HULL_DEGENERATE)
        perimeter = float(hist[i] @ _PERIM_W)                                            # :65
        convex_hull_area = 0 if degenerate else hull_area[i]                              # :68
        convex_hull_perimeter = 0.0 if degenerate else float(hull_hist[i] @ _PERIM_W)     # :69
        area_ratio = convex_hull_area / a                                                 # :72
        circularity = (2 * sqrt(_PI * convex_hull_area)) / convex_hull_perimeter if convex_hull_perimeter > 0 else 0.0   # :75
        nn, s1, s2 = disk_n[i], disk_s1[i], disk_s2[i]
        if nn > 0:                                                                        # :92-94
            mean_brightness = s1 / (3 * nn)
            brightness_std = sqrt((nn * s2 - s1 * s1) / (9 * nn * nn))
        else:
            mean_brightness = brightness_std = 0.0
        dx, dy = max_r[i] - min_r[i], max_c[i] - min_c[i]                                 # :97-100 (rows are "x")
        out.append({
            "deformability": float(1 - circularity),                                      # :78
            "area": a,
            "area_ratio": float(area_ratio),
            "circularity": float(circularity),
            "convex_hull_area": convex_hull_area,
            "mask_x_length": dx,
            "mask_y_length": dy,
            "min_x": min_r[i],
            "min_y": min_c[i],
            "max_x": max_r[i],
            "max_y": max_c[i],
            "mean_brightness": float(mean_brightness),
            "brightness_std": float(brightness_std),
            "perimeter": perimeter,
            "aspect_ratio": float(dx / dy) if dx > 0 and dy > 0 else 0.0,
            "convex_hull_perimeter": float(convex_hull_perimeter),
        })
    return out


def metrics_from_raw(raw: np.void, on_empty: str = "raise") -> Dict[str, Any]:
    """One ysi_mask_metrics row -> the dict of utils/metrics.py:102-119 (see ``metrics_from_rows``)."""
    rows = np.zeros(1, dtype=nat.METRICS_DTYPE)
    rows[0] = raw
    return metrics_from_rows(rows, on_empty)[0]


def box_crop(image: np.ndarray, box: np.ndarray) -> np.ndarray:
    """image[y1:y2, x1:x2] with ``box.astype(int)`` -- the reference's box-to-pixel convention (pipeline.py:379) -- as uint8
    RGB (a zero-copy view for RGB inputs; raw grey samples are converted like ``_load_image`` does)."""
    x1, y1, x2, y2 = np.asarray(box, np.float32).astype(int)
    return as_rgb_u8(image[max(y1, 0):max(y2, 0), max(x1, 0):max(x2, 0)])


def expanded_crop_window(min_x: int, min_y: int, max_x: int, max_y: int, H: int, W: int) -> Tuple[int, int, int, int]:
    """The 2x-expanded mask-bbox window the reference's CSV consumers recompute from the metric row
    (examples/plot_scatter_example.py:107-147, deformability_training_data.py:97-153): the CSV calls rows "x" and columns
    "y", so they are swapped first; the box is doubled about its integer centre and clipped to the image.
    Returns (row0, row1, col0, col1) with image[row0:row1, col0:col1] the crop."""
    c0, c1, r0, r1 = int(min_y), int(max_y), int(min_x), int(max_x)
    cc, rc = (c0 + c1) // 2, (r0 + r1) // 2
    nw, nh = int((c1 - c0) * 2.0), int((r1 - r0) * 2.0)
    c0, c1 = cc - nw // 2, cc + nw // 2
    r0, r1 = rc - nh // 2, rc + nh // 2
    c0 = max(0, min(c0, W - 1))
    c1 = max(c0 + 1, min(c1, W))
    r0 = max(0, min(r0, H - 1))
    r1 = max(r0 + 1, min(r1, H))
    return r0, r1, c0, c1


def expanded_crop(image: np.ndarray, raw: np.void) -> np.ndarray:
    """the expanded mask-bbox crop, cut from the image the stage was given using the bounding box
    the CUDA kernels measured (a zero-copy view for RGB inputs), so downstream tools need not re-read the TIFF."""
    H, W = image.shape[:2]
    r0, r1, c0, c1 = expanded_crop_window(int(raw["min_r"]), int(raw["min_c"]), int(raw["max_r"]), int(raw["max_c"]), H, W)
    return as_rgb_u8(image[r0:r1, c0:c1])


def parse_device(device: str) -> int:
    """'cuda' / 'cuda:3' -> CUDA device index. 'cpu' is refused: this path has no CPU fallback."""
    d = str(device)
    if d == "cuda":
        return 0
    if d.startswith("cuda:"):
        return int(d.split(":", 1)[1])
    raise ValueError(f"SamStage runs on a B200 only (device={device!r}); the CPU path is the reference's")


class SamStage:
    """One context per (GPU, worker); not thread-safe; distinct instances are independent."""

    def __init__(self, sam_model_type: str = "facebook/sam-vit-huge", device: str = "cuda",
                 state_dict: Optional[Dict[str, Any]] = None, max_batch: int = 8, max_boxes: int = 64,
                 max_image_hw: Tuple[int, int] = (1024, 1024), on_empty: str = "raise",
                 precision: Optional[str] = None):
        """``precision``: 16-bit encoding of the tensor-core operands, "fp16" or "bf16" (default: $YSI_PRECISION, else
        fp16); accumulation, residual stream and statistics are fp32 either way (csrc/common.h)."""
        self.variant: SamVariant = variant_of(sam_model_type)
        self.precision = precision or nat.default_precision()
        self.device_index = parse_device(device)
        self.on_empty = on_empty
        self.max_batch, self.max_boxes = int(max_batch), int(max_boxes)
        self._lib = nat.load(precision=self.precision)
        cfg = nat.YsiConfig()
        v = self.variant
        cfg.hidden_size, cfg.num_layers, cfg.num_heads, cfg.mlp_dim = v.hidden_size, v.num_layers, v.num_heads, v.mlp_dim
        cfg.num_global = len(v.global_attn_indexes)
        for i, g in enumerate(v.global_attn_indexes):
            cfg.global_attn_indexes[i] = g
        cfg.max_batch, cfg.max_boxes = self.max_batch, self.max_boxes
        cfg.max_image_h, cfg.max_image_w = int(max_image_hw[0]), int(max_image_hw[1])
        self._ctx = nat._ctx()
        rc = self._lib.ysi_create(self.device_index, C.byref(cfg), C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.ysi_last_error(None)
            self._ctx = None
            raise RuntimeError(f"ysi_create failed ({rc}): {msg.decode() if msg else '?'}")
        self.last_timing: Dict[str, float] = {}
        self._pinned: Dict[Any, Any] = {}
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # ------------------------------------------------------------------ lifecycle
    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self._lib.ysi_last_error(self._ctx)
            raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        """Upload SamModel.state_dict() (torch tensors or numpy arrays, any float dtype)."""
        keep: List[np.ndarray] = []
        descs = (nat.YsiTensorDesc * len(state_dict))()
        for i, (name, t) in enumerate(state_dict.items()):
            if hasattr(t, "detach"):
                t = t.detach().float().cpu().numpy()
            a = np.ascontiguousarray(t, dtype=np.float32)
            keep.append(a)
            descs[i].name = name.encode()
            descs[i].data = nat.as_f32p(a)
            descs[i].ndim = a.ndim
            for k in range(a.ndim):
                descs[i].shape[k] = a.shape[k]
        self._check(self._lib.ysi_load_weights(self._ctx, descs, len(state_dict)), "ysi_load_weights")

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.ysi_destroy(self._ctx)       # drains every stream before the pinned result buffers go away
            self._ctx = None
        for m, r in getattr(self, "_pinned", {}).values():
            m.close(); r.close()
        self._pinned = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self._lib.ysi_launch_count(self._ctx))

    # ------------------------------------------------------------------ the hot path
    def run(self, image: np.ndarray, boxes: np.ndarray, want_masks: bool = True, masks: Optional[str] = "auto",
            expanded_crops: bool = False):
        out = self.run_batch([image], [boxes], want_masks=want_masks, masks=masks, expanded_crops=expanded_crops)
        return out[0]

    def run_packed(self, image: np.ndarray, boxes: np.ndarray) -> Tuple[np.ndarray, List[Dict[str, Any]]]:
        """Masks in the reference's wire format instead of one byte per pixel: row k is ``np.packbits(mask_k)`` (MSB
        first over the row-major mask), i.e. exactly what ``utils/mask_encoding.encode_binary_mask`` feeds to zlib
        (mask_encoding.py:24) -- 8x less device-to-host traffic.  Returns (uint8 [N, ceil(H*W/8)], metrics)."""
        m, mets, _ = self.run(image, boxes, masks="packed")
        return m, mets

    @staticmethod
    def _mask_mode(want_masks: bool, masks: Optional[str]) -> Optional[str]:
        if masks == "auto":
            return "bool" if want_masks else None
        if masks not in (None, "bool", "packed"):
            raise ValueError(f"masks must be 'bool', 'packed' or None, got {masks!r}")
        return masks

    def _submit(self, slot: int, images: Sequence[np.ndarray], boxes: Sequence[np.ndarray], mode: Optional[str], pinned: bool):
        """Validate one batch, allocate its result buffers and enqueue it into ``slot`` (ysi_submit). Returns the
        bookkeeping record ``_finish`` needs, or None for a batch without boxes (nothing is launched, pipeline.py:176-179)."""
        n = len(images)
        assert n == len(boxes) and n >= 1
        fmt = pixel_format_of(images[0])
        H, W = images[0].shape[:2]
        imgs = []
        for im in images:
            if pixel_format_of(im) != fmt or im.shape[:2] != (H, W):
                raise ValueError("images of one batch must share size and pixel format")
            inner_ok = (im.strides[-1] == im.itemsize) and (im.ndim == 2 or im.strides[1] == 3)
            imgs.append(im if inner_ok else np.ascontiguousarray(im))
        row_stride = imgs[0].strides[0]
        if any(im.strides[0] != row_stride for im in imgs):
            imgs = [np.ascontiguousarray(im) for im in imgs]
            row_stride = imgs[0].strides[0]
        blist = [np.asarray(b, np.float32).reshape(-1, 4) for b in boxes]
        counts = np.array([len(b) for b in blist], dtype=np.int32)
        nb = int(counts.sum())
        if nb == 0:
            return None
        allb = np.ascontiguousarray(np.concatenate(blist, 0))
        PB = (H * W + 7) // 8
        if pinned:
            mbuf, rows = self._out_buffers(slot, nb, H, W, mode)
        else:
            mbuf = np.empty((nb, H, W), np.uint8) if mode == "bool" else (np.empty((nb, PB), np.uint8) if mode == "packed" else None)
            rows = np.zeros(nb, dtype=nat.METRICS_DTYPE)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        bt = nat.YsiBatch()
        bt.n_images, bt.height, bt.width, bt.row_stride, bt.pixel_format = n, H, W, row_stride, fmt
        bt.images = ptrs
        bt.boxes_xyxy = nat.as_f32p(allb)
        bt.box_counts = nat.as_i32p(counts)
        bt.masks_out = nat.as_u8p(mbuf) if mode == "bool" else None
        bt.packed_out = nat.as_u8p(mbuf) if mode == "packed" else None
        bt.metrics_out = rows.ctypes.data_as(C.c_void_p)
        self._check(self._lib.ysi_submit(self._ctx, slot, C.byref(bt)), "ysi_submit")
        return dict(slot=slot, images=images, boxes=blist, counts=counts, masks=mbuf, rows=rows, mode=mode, hw=(H, W),
                    keep=(imgs, allb, ptrs, bt))

    def _finish(self, rec, raw: bool, copy: bool, expanded_crops: bool, want_crops: bool = True):
        tm = nat.YsiTiming()
        self._check(self._lib.ysi_wait_batch(self._ctx, rec["slot"], C.byref(tm)), "ysi_wait_batch")
        self.last_timing = tm.as_dict()
        out = []
        k = 0
        mbuf, rows, mode = rec["masks"], rec["rows"], rec["mode"]
        H, W = rec["hw"]
        for i, image in enumerate(rec["images"]):
            c = int(rec["counts"][i])
            if mode is None:
                m = None
            else:
                m = mbuf[k:k + c].view(bool) if mode == "bool" else mbuf[k:k + c]
                if copy:
                    m = m.copy()
            r = rows[k:k + c].copy() if copy else rows[k:k + c]
            mets = r if raw else metrics_from_rows(r, self.on_empty)
            crops = [box_crop(image, bx) for bx in rec["boxes"][i]] if want_crops else None
            if expanded_crops:
                out.append((m, mets, crops, [expanded_crop(image, r[j]) for j in range(c)]))
            else:
                out.append((m, mets, crops))
            k += c
        return out

    def _empty_result(self, images, mode, expanded_crops):
        res = []
        for im in images:
            H, W = im.shape[:2]
            m = None if mode is None else (np.zeros((0, H, W), bool) if mode == "bool" else np.zeros((0, (H * W + 7) // 8), np.uint8))
            res.append((m, [], [], []) if expanded_crops else (m, [], []))
        return res

    def run_batch(self, images: Sequence[np.ndarray], boxes: Sequence[np.ndarray], want_masks: bool = True,
                  raw: bool = False, masks: Optional[str] = "auto", expanded_crops: bool = False):
        """Process several same-sized images in one call. Returns a list of (masks, metrics, crops[, expanded crops]).

        Images are uint8 [H,W,3] RGB (what ``_load_image`` returns) or the raw single-channel samples of the file (uint8 /
        uint16 [H,W]; the device then performs ``_load_image``'s 16 -> 8 bit reduction and grey -> RGB replication).
        ``masks``: "bool" (bool [N,H,W]), "packed" (uint8 [N, ceil(H*W/8)], np.packbits rows) or None (metrics only).
        Any number of boxes is accepted; more than ``max_boxes`` are decoded in chunks that share the image embeddings."""
        mode = self._mask_mode(want_masks, masks)
        rec = self._submit(0, images, boxes, mode, pinned=False)
        if rec is None:
            return self._empty_result(images, mode, expanded_crops)
        return self._finish(rec, raw, copy=False, expanded_crops=expanded_crops)

    # ------------------------------------------------------------------ pipelined form
    def run_stream(self, batches, want_masks: bool = True, raw: bool = False, copy_masks: bool = True,
                   masks: Optional[str] = "auto", expanded_crops: bool = False, crops: bool = True):
        """Generator over ``batches`` (an iterable of (images, boxes) with same-sized images per batch) that keeps two
        batches in flight: the H2D copy and the encoder of batch i+1 overlap the decoder, metrics and D2H copy of
        batch i (ysi_submit / ysi_wait_batch). Yields, per batch, what ``run_batch`` returns.

        ``copy_masks=False`` yields masks and raw rows as views into the slot's pinned host buffer instead of copies (at 32
        boxes per image the copies are 268 MB of host memcpy per batch and dominate); such a view is valid only until
        the generator is advanced again (the slot's buffer is then handed to the next batch). ``crops=False`` skips cutting
        the box crops (the third element of every result is then None)."""
        mode = self._mask_mode(want_masks, masks)
        pending: List[Any] = []
        slot = 0
        for images, boxes in batches:
            if len(pending) == 2:
                yield self._finish(pending.pop(0), raw, copy_masks, expanded_crops, crops)
            rec = self._submit(slot, images, boxes, mode, pinned=True)
            if rec is None:
                while pending:
                    yield self._finish(pending.pop(0), raw, copy_masks, expanded_crops, crops)
                yield self._empty_result(images, mode, expanded_crops)
                continue
            pending.append(rec)
            slot ^= 1
        while pending:
            yield self._finish(pending.pop(0), raw, copy_masks, expanded_crops, crops)

    def _out_buffers(self, slot: int, nb: int, H: int, W: int, mode: Optional[str]):
        """Pinned host buffers for a slot's results (true copy/compute overlap needs page-locked memory); they grow when a
        batch brings more boxes or larger images than any before."""
        per = H * W if mode == "bool" else ((H * W + 7) // 8 if mode == "packed" else 0)
        need_m, need_r = nb * per, nb * nat.METRICS_DTYPE.itemsize
        m, r = self._pinned.get(slot, (None, None))
        if m is None or m.nbytes < need_m:
            m = nat.PinnedBuffer(max(need_m, self.max_boxes * per), self.precision, self.device_index)
        if r is None or r.nbytes < need_r:
            r = nat.PinnedBuffer(max(need_r, self.max_boxes * nat.METRICS_DTYPE.itemsize), self.precision, self.device_index)
        self._pinned[slot] = (m, r)
        masks = m.array[:need_m].reshape(nb, H, W) if mode == "bool" else (m.array[:need_m].reshape(nb, per) if mode == "packed" else None)
        rows = r.array[:need_r].view(nat.METRICS_DTYPE)
        return masks, rows

    # ------------------------------------------------------------------ measurement support (bench.py)
    def pool_upload(self, images: Sequence[np.ndarray]) -> None:
        """Make `images` resident in HBM (device-resident leg of the bench)."""
        n = len(images)
        for i, im in enumerate(images):
            im = np.ascontiguousarray(im)
            H, W = im.shape[:2]
            self._check(self._lib.ysi_pool_upload(self._ctx, n, i, nat.as_u8p(im), H, W, im.strides[0]), "ysi_pool_upload")
        self._pool_hw = (H, W)

    def compute_pool(self, first: int, n: int, boxes: Sequence[np.ndarray], sync: bool = True) -> Dict[str, float]:
        counts = np.array([len(b) for b in boxes], dtype=np.int32)
        allb = np.ascontiguousarray(np.concatenate([np.asarray(b, np.float32).reshape(-1, 4) for b in boxes], 0))
        tm = nat.YsiTiming()
        self._check(self._lib.ysi_compute_pool(self._ctx, first, n, nat.as_f32p(allb), nat.as_i32p(counts), 1 if sync else 0,
                                               C.byref(tm)), "ysi_compute_pool")
        self._staged = (int(counts.sum()), *self._pool_hw)
        return tm.as_dict() if sync else {}

    def timer_record(self, slot: int) -> None:
        self._check(self._lib.ysi_timer_record(self._ctx, slot), "ysi_timer_record")

    def timer_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float(0)
        self._check(self._lib.ysi_timer_elapsed_ms(self._ctx, a, b, C.byref(ms)), "ysi_timer_elapsed_ms")
        return float(ms.value)

    def sync(self) -> None:
        self._check(self._lib.ysi_sync(self._ctx), "ysi_sync")

    def profile(self, enable: bool) -> None:
        self._check(self._lib.ysi_profile(self._ctx, 1 if enable else 0), "ysi_profile")

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        n = 32
        names = (C.c_char_p * n)()
        ms = np.zeros(n, np.float64)
        rec = np.zeros(n, np.int64)
        fl = np.zeros(n, np.float64)
        by = np.zeros(n, np.float64)
        k = self._lib.ysi_profile_read(self._ctx, n, names, nat.as_f64p(ms), rec.ctypes.data_as(C.POINTER(C.c_int64)), nat.as_f64p(fl),
                                       nat.as_f64p(by))
        if k < 0:
            self._check(k, "ysi_profile_read")
        return {names[i].decode(): {"ms": float(ms[i]), "records": int(rec[i]), "flops": float(fl[i]), "bytes": float(by[i])}
                for i in range(k)}

    # ------------------------------------------------------------------ stage-level API (parity tests)
    def preprocess(self, images: Sequence[np.ndarray]) -> np.ndarray:
        n = len(images)
        H, W = images[0].shape[:2]
        imgs = [np.ascontiguousarray(im) for im in images]
        ptrs = (nat._u8p * n)(*[nat.as_u8p(im) for im in imgs])
        out = np.empty((n, 3, 1024, 1024), np.float32)
        self._check(self._lib.ysi_preprocess(self._ctx, n, ptrs, H, W, imgs[0].strides[0], nat.as_f32p(out)), "ysi_preprocess")
        return out

    def encode(self, pixel_values: np.ndarray, want_hidden: bool = False):
        pv = np.ascontiguousarray(pixel_values, np.float32)
        n = pv.shape[0]
        emb = np.empty((n, 256, 64, 64), np.float32)
        hid = np.empty((self.variant.num_layers + 1, n, 64, 64, self.variant.hidden_size), np.float32) if want_hidden else None
        self._check(self._lib.ysi_encode(self._ctx, n, nat.as_f32p(pv), nat.as_f32p(emb), nat.as_f32p(hid)), "ysi_encode")
        return (emb, hid) if want_hidden else emb

    def decode(self, image_embeddings: np.ndarray, boxes_1024: np.ndarray, want_sparse: bool = False):
        emb = np.ascontiguousarray(image_embeddings, np.float32).reshape(256, 64, 64)
        b = np.ascontiguousarray(boxes_1024, np.float64).reshape(-1, 4)
        nb = b.shape[0]
        low = np.empty((nb, 256, 256), np.float32)
        sp = np.empty((nb, 2, 256), np.float32) if want_sparse else None
        self._check(self._lib.ysi_decode(self._ctx, nat.as_f32p(emb), nat.as_f64p(b), nb, nat.as_f32p(low), nat.as_f32p(sp)), "ysi_decode")
        return (low, sp) if want_sparse else low

    def postprocess(self, low_res: np.ndarray, H: int, W: int, want_logits: bool = False):
        low = np.ascontiguousarray(low_res, np.float32).reshape(-1, 256, 256)
        nb = low.shape[0]
        masks = np.empty((nb, H, W), np.uint8)
        up = np.empty((nb, H, W), np.float32) if want_logits else None
        self._check(self._lib.ysi_postprocess(self._ctx, nat.as_f32p(low), nb, H, W, nat.as_u8p(masks), nat.as_f32p(up)), "ysi_postprocess")
        return (masks.view(bool), up) if want_logits else masks.view(bool)

    def metrics(self, image: np.ndarray, masks: np.ndarray, raw: bool = False):
        img = np.ascontiguousarray(image, np.uint8)
        H, W = img.shape[:2]
        m = np.ascontiguousarray(masks).astype(np.uint8).reshape(-1, H, W)
        rows = np.zeros(m.shape[0], dtype=nat.METRICS_DTYPE)
        self._check(self._lib.ysi_metrics(self._ctx, nat.as_u8p(img), H, W, img.strides[0], nat.as_u8p(m), m.shape[0],
                                          rows.ctypes.data_as(C.c_void_p)), "ysi_metrics")
        return rows if raw else metrics_from_rows(rows, self.on_empty)

    def gemm(self, A: np.ndarray, W: np.ndarray, bias: Optional[np.ndarray] = None, act: int = 0) -> np.ndarray:
        A = np.ascontiguousarray(A, np.float32)
        W = np.ascontiguousarray(W, np.float32)
        b = None if bias is None else np.ascontiguousarray(bias, np.float32)
        M, K = A.shape
        N = W.shape[0]
        out = np.empty((M, N), np.float32)
        self._check(self._lib.ysi_gemm(self._ctx, nat.as_f32p(A), nat.as_f32p(W), nat.as_f32p(b), M, N, K, act, nat.as_f32p(out)), "ysi_gemm")
        return out

    def attention(self, qkv: np.ndarray, rel_h: np.ndarray, rel_w: np.ndarray, heads: int, is_global: bool) -> np.ndarray:
        q = np.ascontiguousarray(qkv, np.float32)
        n_seq, T = q.shape[0], q.shape[1]
        head_dim = q.shape[2] // (3 * heads)
        out = np.empty((n_seq, T, heads * head_dim), np.float32)
        self._check(self._lib.ysi_attention(self._ctx, nat.as_f32p(q), nat.as_f32p(np.ascontiguousarray(rel_h, np.float32)),
                                            nat.as_f32p(np.ascontiguousarray(rel_w, np.float32)), n_seq, heads, head_dim,
                                            1 if is_global else 0, nat.as_f32p(out)), "ysi_attention")
        return out

    def gemm_ex(self, A: np.ndarray, W: np.ndarray, bias: Optional[np.ndarray], act: int, out_kind: int,
                C0: Optional[np.ndarray] = None) -> np.ndarray:
        """GEMM through the production dispatcher: out_kind 0 fp32, 1 bf16-rounded, 2 C0 += result."""
        A = np.ascontiguousarray(A, np.float32)
        W = np.ascontiguousarray(W, np.float32)
        b = None if bias is None else np.ascontiguousarray(bias, np.float32)
        M, K = A.shape
        N = W.shape[0]
        out = np.zeros((M, N), np.float32) if C0 is None else np.array(C0, dtype=np.float32, order="C", copy=True)
        self._check(self._lib.ysi_gemm_ex(self._ctx, nat.as_f32p(A), nat.as_f32p(W), nat.as_f32p(b), M, N, K, act, out_kind,
                                          nat.as_f32p(out)), "ysi_gemm_ex")
        return out

    def gemm_bench(self, M: int, N: int, K: int, pair: bool, mode: int, iters: int = 20) -> float:
        """ms per launch of one GEMM shape on device-resident operands (measurement support)."""
        ms = C.c_float(0)
        self._check(self._lib.ysi_gemm_bench(self._ctx, M, N, K, int(pair), mode, iters, C.byref(ms)), "ysi_gemm_bench")
        return float(ms.value)

    def attention_bench(self, n_seq: int, heads: int, head_dim: int, is_global: bool, iters: int = 20) -> float:
        """ms per launch of one attention shape on device-resident operands (measurement support)."""
        ms = C.c_float(0)
        self._check(self._lib.ysi_attention_bench(self._ctx, n_seq, heads, head_dim, int(is_global), iters, C.byref(ms)),
                    "ysi_attention_bench")
        return float(ms.value)

    def image_pe(self) -> np.ndarray:
        out = np.empty((256, 64, 64), np.float32)
        self._check(self._lib.ysi_get_image_pe(self._ctx, nat.as_f32p(out)), "ysi_get_image_pe")
        return out
