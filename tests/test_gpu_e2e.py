"""End-to-end SAM stage (image + boxes -> masks + metrics) through ysi_run_batch vs the oracle."""
import numpy as np
import pytest

from conftest import iou_gate, rel_l2

pytestmark = pytest.mark.gpu


def test_zero_boxes_short_circuit(tiny_stage):
    img = np.zeros((1024, 1024, 3), np.uint8)
    before = tiny_stage.launch_count
    masks, mets, crops = tiny_stage.run(img, np.zeros((0, 4), np.float32))
    assert masks.shape == (0, 1024, 1024) and mets == [] and crops == []
    assert tiny_stage.launch_count == before       # pipeline.py:176-179: SAM skipped, nothing launched


def test_stage_end_to_end(tiny_stage, tiny_oracle):
    from oracle import metrics_oracle as mo
    from oracle import sam_oracle
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    imgs, boxes, ref_masks, ref_up = [], [], [], []
    for idx, nb in ((10, 1), (11, 3)):
        g, b = synth_image(idx, 1024, nb)
        im = gray_to_rgb_u8(g)
        m, d = sam_oracle.run_stage(tiny_oracle, im, b, dump=True)
        imgs.append(im); boxes.append(b); ref_masks.append(m); ref_up.append(d["upsampled_logits"])
    tiny_stage.on_empty = "zeros"
    out = tiny_stage.run_batch(imgs, boxes)
    ious = []
    for (masks, mets, crops), rm, im, b in zip(out, ref_masks, imgs, boxes):
        assert masks.shape == rm.shape and len(mets) == len(rm) == len(crops)
        for k in range(len(rm)):
            inter = np.logical_and(masks[k], rm[k]).sum()
            union = np.logical_or(masks[k], rm[k]).sum()
            ious.append(inter / max(union, 1))
            # metrics of OUR mask must equal the oracle's metrics of OUR mask exactly (a7 on identical masks)
            if masks[k].any():
                ref = mo.calculate_metrics(im, masks[k])
                for key, val in ref.items():
                    if isinstance(val, int):
                        assert mets[k][key] == val, (key, mets[k][key], val)
                    else:
                        assert mets[k][key] == pytest.approx(val, rel=1e-9, abs=1e-12), key
            x1, y1, x2, y2 = b[k].astype(int)
            assert np.array_equal(crops[k], im[y1:y2, x1:x2])
    print("end-to-end mask IoU vs fp32 oracle (random-init noise-field logits):", ["%.4f" % i for i in ious])
    # random-init logits are a zero-mean noise field, the hardest case for a thresholded-mask
    # IoU: north-star gate 0.999 with fp16 operands; bf16 operands have their own measured floor (conftest.iou_gate)
    assert min(ious) >= iou_gate(tiny_stage.precision)


def test_batch_equals_single(tiny_stage):
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    imgs, boxes = [], []
    for idx in (20, 21):
        g, b = synth_image(idx, 1024, 2)
        imgs.append(gray_to_rgb_u8(g)); boxes.append(b)
    tiny_stage.on_empty = "zeros"
    both = tiny_stage.run_batch(imgs, boxes)
    for i in range(2):
        single = tiny_stage.run(imgs[i], boxes[i])
        assert np.array_equal(both[i][0], single[0])
        assert both[i][1] == single[1]


@pytest.mark.parametrize("H,W", [(2048, 2048), (348, 704)])
def test_stage_end_to_end_other_sizes(tiny_stage, tiny_oracle, H, W):
    """Config-5 sized frames (2x antialias downscale in, 1024->2048 second upsample out) and the aspect ratio of the
    reference's example PNGs (upscale in, unpadded 506x1024 -> 348x704 downscale out)."""
    from oracle import metrics_oracle as mo
    from oracle import sam_oracle
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    g, b = synth_image(31, max(H, W), 2)
    im = np.ascontiguousarray(gray_to_rgb_u8(g)[:H, :W])
    b = np.clip(b, 0, [W - 1, H - 1, W - 1, H - 1]).astype(np.float32)
    ref_masks, d = sam_oracle.run_stage(tiny_oracle, im, b, dump=True)
    tiny_stage.on_empty = "zeros"
    masks, mets, crops = tiny_stage.run(im, b)
    assert masks.shape == ref_masks.shape == (2, H, W)
    # a6 alone on the oracle's logits is exact at this geometry too
    m6 = tiny_stage.postprocess(d["low_res_logits"], H, W)
    assert np.array_equal(m6, ref_masks)
    for k in range(2):
        iou = np.logical_and(masks[k], ref_masks[k]).sum() / max(np.logical_or(masks[k], ref_masks[k]).sum(), 1)
        assert iou >= iou_gate(tiny_stage.precision), iou
        if masks[k].any():
            ref = mo.calculate_metrics(im, masks[k])
            for key, val in ref.items():
                if isinstance(val, int):
                    assert mets[k][key] == val, (key, mets[k][key], val)
                else:
                    assert mets[k][key] == pytest.approx(val, rel=1e-9, abs=1e-12), key


def test_run_stream_equals_run_batch(tiny_stage):
    """The two-slot pipelined API (copies / encoder / decoder of neighbouring batches overlap on four streams) must
    return exactly what the synchronous call returns, batch by batch, including a zero-box batch in the middle."""
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    tiny_stage.on_empty = "zeros"
    batches = []
    for i, nbx in enumerate((1, 2, 0, 3, 1)):
        imgs, boxes = [], []
        for k in range(2):
            g, b = synth_image(40 + 2 * i + k, 1024, max(nbx, 1))
            imgs.append(gray_to_rgb_u8(g)); boxes.append(b if nbx else np.zeros((0, 4), np.float32))
        batches.append((imgs, boxes))
    ref = [tiny_stage.run_batch(im, bx) for im, bx in batches]
    got = list(tiny_stage.run_stream(iter(batches)))
    assert len(got) == len(ref)
    for r, g in zip(ref, got):
        for (rm, rmet, rc), (gm, gmet, gc) in zip(r, g):
            assert np.array_equal(rm, gm) and rmet == gmet and len(rc) == len(gc)


@pytest.mark.parametrize("H,W", [(1024, 1024), (348, 701)])
def test_packed_mask_wire_format(tiny_stage, H, W):
    """f2 of SURVEY section 8f: masks as np.packbits rows (utils/mask_encoding.py:24), byte-identical to packing the
    byte masks on the host, with the same metrics; 348 x 701 has a pixel count that is not a multiple of 8."""
    import base64
    import zlib
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    g, b = synth_image(90, max(H, W), 2)
    img = np.ascontiguousarray(gray_to_rgb_u8(g)[:H, :W])
    b = np.clip(b, 0, [W - 1, H - 1, W - 1, H - 1]).astype(np.float32)
    tiny_stage.on_empty = "zeros"
    masks, mets, _ = tiny_stage.run(img, b)
    packed, mets2 = tiny_stage.run_packed(img, b)
    assert mets2 == mets
    for k in range(len(masks)):
        assert np.array_equal(packed[k], np.packbits(masks[k]))
        # the reference's encode / decode round trip on the packed row (mask_encoding.py:10-58)
        data = base64.b64encode(zlib.compress(packed[k].tobytes())).decode("ascii")
        back = np.unpackbits(np.frombuffer(zlib.decompress(base64.b64decode(data)), np.uint8))[:H * W].reshape(H, W)
        assert np.array_equal(back.astype(bool), masks[k])
