"""CPU: the raw-TIFF ingest route (host half) yields exactly what the reference's loader
(cv2.imread + BGR->RGB, pipeline.py:206-210) yields; everything it cannot take falls back to that loader."""
import struct

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _roundtrip(path):
    from yolo_sam_inference_b200 import ingest
    info = ingest.probe_tiff(str(path))
    ref = ingest.decode_rgb(str(path))
    if info is None:
        return None, ref
    buf = np.zeros(info.nbytes, np.uint8)
    ingest.read_raw_into(info, buf)
    return ingest.as_rgb_u8(buf.view(info.dtype).reshape(info.shape)), ref


@pytest.mark.parametrize("kind", ["gray8", "gray16", "rgb8"])
def test_baseline_tiff_raw_route_equals_cv2(tmp_path, kind):
    from yolo_sam_inference_b200 import _native as nat
    from yolo_sam_inference_b200 import ingest
    rng = np.random.RandomState(7)
    if kind == "gray8":
        a = rng.randint(0, 256, (301, 417)).astype(np.uint8)
    elif kind == "gray16":
        a = rng.randint(0, 65536, (260, 333)).astype(np.uint16)        # every high/low byte combination: v >> 8, not v / 256
    else:
        a = rng.randint(0, 256, (128, 200, 3)).astype(np.uint8)
    p = tmp_path / f"{kind}.tiff"
    assert cv2.imwrite(str(p), a if a.ndim == 2 else cv2.cvtColor(a, cv2.COLOR_RGB2BGR), [cv2.IMWRITE_TIFF_COMPRESSION, 1])
    info = ingest.probe_tiff(str(p))
    assert info is not None and (info.height, info.width) == a.shape[:2]
    assert info.pixel_format == {"gray8": nat.PIX_GRAY8, "gray16": nat.PIX_GRAY16, "rgb8": nat.PIX_RGB8}[kind]
    got, ref = _roundtrip(p)
    assert got.dtype == np.uint8 and np.array_equal(got, ref)
    if kind == "gray16":
        assert np.array_equal(got[..., 0], (a >> 8).astype(np.uint8))


def test_multi_strip_tiff(tmp_path):
    """Strips that are not contiguous in the file (hand-written baseline TIFF: 3 strips of 2 rows with gaps between them)."""
    from yolo_sam_inference_b200 import ingest
    H, W, rps = 6, 5, 2
    img = np.arange(H * W, dtype=np.uint8).reshape(H, W) * 3
    strips = [img[r:r + rps].tobytes() for r in range(0, H, rps)]
    body = b""
    offs = []
    pos = 8
    for s in strips:
        body += b"\xEE" * 3                      # gap
        pos += 3
        offs.append(pos)
        body += s
        pos += len(s)
    n = len(strips)
    off_arr, cnt_arr = pos, pos + 4 * n
    extra = struct.pack("<%dI" % n, *offs) + struct.pack("<%dI" % n, *[len(s) for s in strips])
    ifd_off = cnt_arr + 4 * n
    entries = [(256, 3, 1, W), (257, 3, 1, H), (258, 3, 1, 8), (259, 3, 1, 1), (262, 3, 1, 1), (273, 4, n, off_arr),
               (277, 3, 1, 1), (278, 3, 1, rps), (279, 4, n, cnt_arr)]
    ifd = struct.pack("<H", len(entries))
    for tag, typ, cnt, val in entries:
        ifd += struct.pack("<HHI", tag, typ, cnt) + (struct.pack("<HH", val, 0) if typ == 3 and cnt == 1 else struct.pack("<I", val))
    ifd += struct.pack("<I", 0)
    p = tmp_path / "strips.tiff"
    p.write_bytes(b"II" + struct.pack("<HI", 42, ifd_off) + body + extra + ifd)
    info = ingest.probe_tiff(str(p))
    assert info is not None and len(info.strips) == 3
    got, ref = _roundtrip(p)
    assert np.array_equal(got, ref) and np.array_equal(got[..., 0], img)


def test_unsupported_files_take_the_reference_loader(tmp_path):
    from yolo_sam_inference_b200 import ingest
    rng = np.random.RandomState(3)
    g = rng.randint(0, 256, (64, 80)).astype(np.uint8)
    cv2.imwrite(str(tmp_path / "lzw.tiff"), g)                                   # OpenCV's default: LZW
    cv2.imwrite(str(tmp_path / "deflate.tiff"), g, [cv2.IMWRITE_TIFF_COMPRESSION, 8])
    cv2.imwrite(str(tmp_path / "plain.png"), g)
    (tmp_path / "garbage.tiff").write_bytes(b"II*\x00\xff\xff\xff\x7f")
    (tmp_path / "short.tiff").write_bytes(b"II")
    for name in ("lzw.tiff", "deflate.tiff", "plain.png", "garbage.tiff", "short.tiff"):
        assert ingest.probe_tiff(str(tmp_path / name)) is None, name
    assert np.array_equal(ingest.decode_rgb(str(tmp_path / "lzw.tiff"))[..., 1], g)
    # big-endian 16-bit samples would need a byte swap: not taken raw
    be = tmp_path / "be16.tiff"
    H, W = 4, 4
    data = (np.arange(16, dtype=">u2") * 1000).tobytes()
    entries = [(256, 3, 1, W), (257, 3, 1, H), (258, 3, 1, 16), (259, 3, 1, 1), (262, 3, 1, 1), (273, 4, 1, 8),
               (277, 3, 1, 1), (278, 3, 1, H), (279, 4, 1, len(data))]
    ifd = struct.pack(">H", len(entries))
    for tag, typ, cnt, val in entries:
        ifd += struct.pack(">HHI", tag, typ, cnt) + (struct.pack(">HH", val, 0) if typ == 3 else struct.pack(">I", val))
    ifd += struct.pack(">I", 0)
    be.write_bytes(b"MM" + struct.pack(">HI", 42, 8 + len(data)) + data + ifd)
    assert ingest.probe_tiff(str(be)) is None


def test_pixel_format_of_arrays():
    from yolo_sam_inference_b200 import _native as nat
    from yolo_sam_inference_b200.ingest import pixel_format_of
    assert pixel_format_of(np.zeros((4, 4, 3), np.uint8)) == nat.PIX_RGB8
    assert pixel_format_of(np.zeros((4, 4), np.uint8)) == nat.PIX_GRAY8
    assert pixel_format_of(np.zeros((4, 4), np.uint16)) == nat.PIX_GRAY16
    with pytest.raises(ValueError):
        pixel_format_of(np.zeros((4, 4), np.float32))
