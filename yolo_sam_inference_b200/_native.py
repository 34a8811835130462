"""ctypes binding of libysi.so (include/ysi.h).  No torch types cross this boundary.

The library is the product path: if it cannot be loaded or a call fails, we raise -- there is no CPU
or PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

YSI_MAX_GLOBAL_LAYERS = 8
YSI_PERIM_BINS = 10
PERIM_CODES = (5, 7, 13, 15, 17, 21, 23, 25, 27, 33)
FLAG_EMPTY_MASK = 1
FLAG_HULL_DEGENERATE = 2
FLAG_CONTOUR_TRUNCATED = 4


class YsiConfig(C.Structure):
    _fields_ = [("hidden_size", C.c_int32), ("num_layers", C.c_int32), ("num_heads", C.c_int32),
                ("mlp_dim", C.c_int32), ("num_global", C.c_int32),
                ("global_attn_indexes", C.c_int32 * YSI_MAX_GLOBAL_LAYERS),
                ("max_batch", C.c_int32), ("max_boxes", C.c_int32),
                ("max_image_h", C.c_int32), ("max_image_w", C.c_int32)]


class YsiTensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4)]


class YsiTiming(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "preprocess_ms", "encoder_ms", "decoder_ms",
                                         "postprocess_ms", "metrics_ms", "d2h_ms", "total_ms")]

    def as_dict(self):
        return {n: float(getattr(self, n)) for n, _ in self._fields_}


PIX_RGB8, PIX_GRAY8, PIX_GRAY16 = 0, 1, 2


class YsiBatch(C.Structure):
    _fields_ = [("n_images", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("row_stride", C.c_int32),
                ("pixel_format", C.c_int32), ("images", C.POINTER(C.c_void_p)), ("boxes_xyxy", C.POINTER(C.c_float)),
                ("box_counts", C.POINTER(C.c_int32)), ("masks_out", C.POINTER(C.c_uint8)),
                ("packed_out", C.POINTER(C.c_uint8)), ("metrics_out", C.c_void_p)]


# numpy mirror of ysi_mask_metrics (same field order / sizes; no padding needed: all naturally aligned)
METRICS_DTYPE = np.dtype([
    ("area", "<i8"), ("sum_r", "<i8"), ("sum_c", "<i8"),
    ("min_r", "<i4"), ("min_c", "<i4"), ("max_r", "<i4"), ("max_c", "<i4"),
    ("perim_hist", "<u4", (YSI_PERIM_BINS,)),
    ("hull_area", "<i8"),
    ("hull_perim_hist", "<u4", (YSI_PERIM_BINS,)),
    ("disk_n", "<i8"), ("disk_sum", "<u8"), ("disk_sumsq", "<u8"),
    ("flags", "<u4"), ("contour_points", "<i4"), ("hull_vertices", "<i4"), ("reserved", "<i4"),
    ("mask_hist", "<u4", (256,)),
], align=True)

_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_ctx = C.c_void_p

EXPORTS = {
    "ysi_version": (C.c_int, []),
    "ysi_operand_dtype": (C.c_char_p, []),
    "ysi_create": (C.c_int, [C.c_int, C.POINTER(YsiConfig), C.POINTER(_ctx)]),
    "ysi_load_weights": (C.c_int, [_ctx, C.POINTER(YsiTensorDesc), C.c_size_t]),
    "ysi_destroy": (None, [_ctx]),
    "ysi_last_error": (C.c_char_p, [_ctx]),
    "ysi_run": (C.c_int, [_ctx, _u8p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, _u8p, _u8p, C.c_void_p,
                          C.POINTER(YsiTiming)]),
    "ysi_run_batch": (C.c_int, [_ctx, C.c_int, C.POINTER(_u8p), C.c_int, C.c_int, C.c_int, _f32p, _i32p, _u8p, _u8p,
                                C.c_void_p, C.POINTER(YsiTiming)]),
    "ysi_submit_batch": (C.c_int, [_ctx, C.c_int, C.c_int, C.POINTER(_u8p), C.c_int, C.c_int, C.c_int, _f32p, _i32p,
                                   _u8p, _u8p, C.c_void_p]),
    "ysi_wait_batch": (C.c_int, [_ctx, C.c_int, C.POINTER(YsiTiming)]),
    "ysi_submit": (C.c_int, [_ctx, C.c_int, C.POINTER(YsiBatch)]),
    "ysi_alloc_pinned": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "ysi_free_pinned": (None, [C.c_void_p]),
    "ysi_pool_upload": (C.c_int, [_ctx, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, C.c_int]),
    "ysi_compute_pool": (C.c_int, [_ctx, C.c_int, C.c_int, _f32p, _i32p, C.c_int, C.POINTER(YsiTiming)]),
    "ysi_timer_record": (C.c_int, [_ctx, C.c_int]),
    "ysi_timer_elapsed_ms": (C.c_int, [_ctx, C.c_int, C.c_int, _f32p]),
    "ysi_sync": (C.c_int, [_ctx]),
    "ysi_profile": (C.c_int, [_ctx, C.c_int]),
    "ysi_profile_read": (C.c_int, [_ctx, C.c_int, C.POINTER(C.c_char_p), _f64p, C.POINTER(C.c_int64), _f64p, _f64p]),
    "ysi_preprocess": (C.c_int, [_ctx, C.c_int, C.POINTER(_u8p), C.c_int, C.c_int, C.c_int, _f32p]),
    "ysi_encode": (C.c_int, [_ctx, C.c_int, _f32p, _f32p, _f32p]),
    "ysi_decode": (C.c_int, [_ctx, _f32p, _f64p, C.c_int, _f32p, _f32p]),
    "ysi_postprocess": (C.c_int, [_ctx, _f32p, C.c_int, C.c_int, C.c_int, _u8p, _f32p]),
    "ysi_metrics": (C.c_int, [_ctx, _u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_void_p]),
    "ysi_gemm": (C.c_int, [_ctx, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]),
    "ysi_attention": (C.c_int, [_ctx, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]),
    "ysi_gemm_ex": (C.c_int, [_ctx, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]),
    "ysi_gemm_bench": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "ysi_attention_bench": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "ysi_get_image_pe": (C.c_int, [_ctx, _f32p]),
    "ysi_launch_count": (C.c_int64, [_ctx]),
}

PRECISIONS = ("bf16", "fp16")      # 16-bit operand encoding of the tensor-core contractions (csrc/common.h)
_LIBS: dict = {}


def default_precision() -> str:
    return os.environ.get("YSI_PRECISION", "fp16")


def lib_path(precision: str = "bf16") -> str:
    from .build import lib_path as _lp
    return _lp(precision)


def load(build_if_missing: bool = True, precision: Optional[str] = None) -> C.CDLL:
    """Load the library of one operand precision and bind every symbol of include/ysi.h (raises if any is missing)."""
    precision = precision or default_precision()
    if precision not in PRECISIONS:
        raise ValueError(f"precision must be one of {PRECISIONS}, got {precision!r}")
    if precision in _LIBS:
        return _LIBS[precision]
    path = lib_path(precision)
    if not os.path.exists(path):
        if not build_if_missing:
            raise RuntimeError(f"{path} is missing: run `python -m yolo_sam_inference_b200.build`")
        from .build import build
        build(precision=precision)
    lib = C.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.ysi_operand_dtype().decode() != precision:
        raise RuntimeError(f"{path} was built for {lib.ysi_operand_dtype().decode()} operands, expected {precision}")
    _LIBS[precision] = lib
    return lib


class PinnedBuffer:
    """Page-locked host memory (ysi_alloc_pinned) exposed as a uint8 numpy array; freed with the object."""

    def __init__(self, nbytes: int, precision: Optional[str] = None, device: int = 0):
        self._lib = load(precision=precision)
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        rc = self._lib.ysi_alloc_pinned(int(device), self.nbytes, C.byref(p))
        if rc != 0 or not p.value:
            raise MemoryError(f"ysi_alloc_pinned({nbytes}) failed ({rc})")
        self._ptr = p
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(self.nbytes, 1)).from_address(p.value))[:self.nbytes]

    def close(self) -> None:
        if getattr(self, "_ptr", None):
            self.array = None
            self._lib.ysi_free_pinned(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def as_u8p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_u8p)


def as_f32p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_f32p)


def as_f64p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_f64p)


def as_i32p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_i32p)
