"""Image ingest for the folder entry point: the host half of SURVEY.md section 8 row f3.

The reference loads every file with ``cv2.imread`` + BGR->RGB (``_load_image``,
/root/reference/src/yolo_sam_inference/pipeline.py:206-210): 16-bit samples are reduced to 8 bits (``v >> 8``) and grey is
replicated to three channels on the host, one core at ~18 files/s for 2048 x 2048 16-bit LZW TIFFs.

Here a *baseline* TIFF (uncompressed strips, 8- or 16-bit single channel, or 8-bit RGB) is not decoded at all: its strips
are read straight into page-locked memory (``readinto``, no intermediate copy, GIL released) and handed to the device as
raw samples -- 1 or 2 bytes per pixel over PCIe instead of 3 -- where ``gray_ingest_kernel`` (csrc/postproc.cu) performs
the 16 -> 8 bit reduction and the grey -> RGB replication.  Anything else (compressed TIFF, PNG, JPEG, tiles, big-endian
16-bit, palettes ...) takes the reference's own ``cv2.imread`` path, in a thread pool.  Both routes yield bit-identical
RGB images (tests/test_ingest.py compares them with cv2.imread file by file).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _native as nat

_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8}


@dataclass(frozen=True)
class RawTiff:
    """A baseline TIFF whose samples can be handed to the device as they sit in the file."""
    path: str
    height: int
    width: int
    pixel_format: int                 # nat.PIX_GRAY8 / PIX_GRAY16 / PIX_RGB8
    strips: Tuple[Tuple[int, int], ...]   # (file offset, byte count) in row order

    @property
    def bytes_per_pixel(self) -> int:
        return {nat.PIX_GRAY8: 1, nat.PIX_GRAY16: 2, nat.PIX_RGB8: 3}[self.pixel_format]

    @property
    def nbytes(self) -> int:
        return self.height * self.width * self.bytes_per_pixel

    @property
    def shape(self) -> Tuple[int, ...]:
        return (self.height, self.width, 3) if self.pixel_format == nat.PIX_RGB8 else (self.height, self.width)

    @property
    def dtype(self):
        return np.uint16 if self.pixel_format == nat.PIX_GRAY16 else np.uint8


def probe_tiff(path: str) -> Optional[RawTiff]:
    """Parse the first IFD; return a RawTiff if the file is a baseline image the raw route supports, else None."""
    try:
        with open(path, "rb") as f:
            head = f.read(8)
            if len(head) < 8 or head[:2] not in (b"II", b"MM"):
                return None
            bo = "<" if head[:2] == b"II" else ">"
            magic, ifd = struct.unpack(bo + "HI", head[2:8])
            if magic != 42:                    # BigTIFF (43) and anything else: decode route
                return None
            f.seek(ifd)
            raw = f.read(2)
            if len(raw) < 2:
                return None
            (n,) = struct.unpack(bo + "H", raw)
            entries = f.read(12 * n)
            if len(entries) < 12 * n:
                return None
            tags = {}
            for i in range(n):
                tag, typ, cnt = struct.unpack(bo + "HHI", entries[12 * i:12 * i + 8])
                size = _TYPE_SIZE.get(typ)
                if size is None:
                    continue
                val = entries[12 * i + 8:12 * i + 12]
                tags[tag] = (typ, cnt, val, size)

            def values(tag) -> Optional[List[int]]:
                if tag not in tags:
                    return None
                typ, cnt, val, size = tags[tag]
                if typ not in (1, 3, 4):
                    return None
                nbytes = size * cnt
                if nbytes <= 4:
                    data = val[:nbytes]
                else:
                    (off,) = struct.unpack(bo + "I", val)
                    f.seek(off)
                    data = f.read(nbytes)
                    if len(data) < nbytes:
                        return None
                return list(struct.unpack(bo + {1: "B", 3: "H", 4: "I"}[typ] * cnt, data))

            width, height = values(256), values(257)
            if not width or not height:
                return None
            W, H = width[0], height[0]
            if (values(259) or [1])[0] != 1:           # Compression: uncompressed only
                return None
            if 322 in tags or 324 in tags:             # tiles
                return None
            if (values(284) or [1])[0] != 1:           # PlanarConfiguration: chunky
                return None
            if (values(274) or [1])[0] != 1:           # Orientation: top-left only
                return None
            if (values(317) or [1])[0] != 1:           # Predictor
                return None
            if 338 in tags:                            # ExtraSamples (alpha)
                return None
            fmt = values(339)                          # SampleFormat: unsigned integer only
            if fmt and any(v != 1 for v in fmt):
                return None
            spp = (values(277) or [1])[0]
            bits = values(258) or [1]
            photo = (values(262) or [None])[0]
            if spp == 1 and photo == 1 and bits == [8]:
                pf = nat.PIX_GRAY8
            elif spp == 1 and photo == 1 and bits == [16] and bo == "<":
                pf = nat.PIX_GRAY16
            elif spp == 3 and photo == 2 and bits == [8, 8, 8]:
                pf = nat.PIX_RGB8
            else:
                return None
            offs, cnts = values(273), values(279)
            if not offs or not cnts or len(offs) != len(cnts):
                return None
            bpp = {nat.PIX_GRAY8: 1, nat.PIX_GRAY16: 2, nat.PIX_RGB8: 3}[pf]
            if sum(cnts) != H * W * bpp:
                return None
            rps = (values(278) or [H])[0]
            if len(offs) != (H + rps - 1) // rps:
                return None
            # merge strips that are contiguous in the file (the usual case: one read for the whole image)
            strips: List[Tuple[int, int]] = []
            for o, c in zip(offs, cnts):
                if strips and strips[-1][0] + strips[-1][1] == o:
                    strips[-1] = (strips[-1][0], strips[-1][1] + c)
                else:
                    strips.append((o, c))
            return RawTiff(path, H, W, pf, tuple(strips))
    except (OSError, struct.error):
        return None


def read_raw_into(info: RawTiff, out: np.ndarray) -> None:
    """Read the strips of ``info`` into ``out`` (a writable C-contiguous uint8 view of info.nbytes bytes, typically pinned)."""
    mv = memoryview(out).cast("B")
    assert len(mv) == info.nbytes
    pos = 0
    with open(info.path, "rb", buffering=0) as f:
        for off, cnt in info.strips:
            f.seek(off)
            got = 0
            while got < cnt:
                k = f.readinto(mv[pos + got:pos + cnt])
                if not k:
                    raise OSError(f"{info.path}: unexpected end of file in a TIFF strip")
                got += k
            pos += cnt


def decode_rgb(path: str) -> np.ndarray:
    """The reference's loader, verbatim (pipeline.py:206-210): cv2.imread (default flags) + BGR->RGB -> uint8 [H,W,3]."""
    import cv2
    image = cv2.imread(path)
    return cv2.cvtColor(image, cv2.COLOR_BGR2RGB)


def as_rgb_u8(samples: np.ndarray) -> np.ndarray:
    """Raw samples (grey uint8 / uint16, or RGB) -> the uint8 RGB image _load_image would have returned (host copy; used
    for detectors that need pixels, crops and visualisations -- the device performs the same conversion for the SAM path)."""
    if samples.ndim == 3:
        return samples
    g = (samples >> 8).astype(np.uint8) if samples.dtype == np.uint16 else samples
    return np.repeat(g[:, :, None], 3, axis=2)


def pixel_format_of(image: np.ndarray) -> int:
    if image.ndim == 3 and image.shape[2] == 3 and image.dtype == np.uint8:
        return nat.PIX_RGB8
    if image.ndim == 2 and image.dtype == np.uint8:
        return nat.PIX_GRAY8
    if image.ndim == 2 and image.dtype == np.uint16:
        return nat.PIX_GRAY16
    raise ValueError(f"unsupported image array: shape {image.shape}, dtype {image.dtype} "
                     "(uint8 [H,W,3] RGB, or raw grey uint8 / uint16 [H,W])")
