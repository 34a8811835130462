"""End-to-end SAM stage (image + boxes -> masks + metrics) through ysi_run_batch vs the oracle."""
import numpy as np
import pytest

from conftest import check_mask_parity, rel_l2

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
    for (masks, mets, crops), rm, ru, im, b in zip(out, ref_masks, ref_up, imgs, boxes):
        assert masks.shape == rm.shape and len(mets) == len(rm) == len(crops)
        for k in range(len(rm)):
            ious.append(check_mask_parity(masks[k], rm[k], ru[k], tiny_stage.precision))
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
    print("end-to-end mask IoU vs fp32 oracle (random-init noise-field logits, %s operands):" % tiny_stage.precision,
          ["%.4f" % i for i in ious])


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
        check_mask_parity(masks[k], ref_masks[k], d["upsampled_logits"][k], tiny_stage.precision, (H, W, k))
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


def test_raw_grey_ingest_equals_rgb_path(tiny_stage):
    """f3: raw 8-/16-bit single-channel samples handed to ysi_submit (device-side v >> 8 and grey -> RGB) give bit-identical
    masks, metric rows and crops to the RGB image _load_image (cv2.imread, pipeline.py:206-210) would have produced."""
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    tiny_stage.on_empty = "zeros"
    rng = np.random.RandomState(4)
    for H, W, bits in ((1024, 1024, 8), (1024, 1024, 16), (600, 812, 16), (2048, 2048, 16)):
        g, b = synth_image(70 + bits + H, max(H, W), 2, bit_depth=bits)
        g = np.ascontiguousarray(g[:H, :W])
        if bits == 16:
            g = (g + rng.randint(0, 256, g.shape).astype(np.uint16)).astype(np.uint16)    # exercise the low byte: v >> 8 drops it
        b = np.clip(b, 0, [W - 1, H - 1, W - 1, H - 1]).astype(np.float32)
        rgb = gray_to_rgb_u8(g)
        (m0, r0, c0), = tiny_stage.run_batch([rgb], [b], raw=True)
        (m1, r1, c1), = tiny_stage.run_batch([g], [b], raw=True)
        assert np.array_equal(m0, m1)
        assert r0.tobytes() == r1.tobytes()
        assert all(np.array_equal(x, y) for x, y in zip(c0, c1))
    # through the pipelined API as well (raw samples straight from pinned memory)
    g, b = synth_image(75, 1024, 2, bit_depth=16)
    ref = tiny_stage.run_batch([gray_to_rgb_u8(g)], [b], raw=True)[0]
    got = list(tiny_stage.run_stream(iter([([g], [b])] * 3), raw=True))
    for batch in got:
        assert np.array_equal(batch[0][0], ref[0]) and batch[0][1].tobytes() == ref[1].tobytes()


def test_more_boxes_than_max_boxes_are_chunked(tiny_weights, tiny_stage):
    """The reference loops over any number of boxes (pipeline.py:170): a context with max_boxes = 2 must return for 5 + 3
    boxes exactly what a context that holds them all returns (chunks share the batch's image embeddings)."""
    from yolo_sam_inference_b200.sam_stage import SamStage
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    imgs, boxes = [], []
    for idx, nb in ((80, 5), (81, 3)):
        g, b = synth_image(idx, 1024, nb)
        imgs.append(gray_to_rgb_u8(g)); boxes.append(b)
    tiny_stage.on_empty = "zeros"
    ref = tiny_stage.run_batch(imgs, boxes, raw=True)
    small = SamStage("vit_t", device="cuda:0", state_dict=tiny_weights, max_batch=2, max_boxes=2, on_empty="zeros",
                     precision=tiny_stage.precision)
    try:
        got = small.run_batch(imgs, boxes, raw=True)
        packed = small.run_batch(imgs, boxes, raw=True, masks="packed")
        streamed = list(small.run_stream(iter([(imgs, boxes), (imgs[:1], boxes[:1])]), raw=True))
    finally:
        small.close()
    for (m0, r0, _), (m1, r1, _), (p1, _, _) in zip(ref, got, packed):
        assert np.array_equal(m0, m1) and r0.tobytes() == r1.tobytes()
        assert all(np.array_equal(p1[k], np.packbits(m0[k])) for k in range(len(m0)))
    assert np.array_equal(streamed[0][1][0], ref[1][0]) and np.array_equal(streamed[1][0][0], ref[0][0])


def test_image_buffers_grow_on_demand(tiny_weights, tiny_stage):
    """max_image_hw is only the initial capacity: a larger image re-sizes the slot buffers instead of failing."""
    from yolo_sam_inference_b200.sam_stage import SamStage
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    g, b = synth_image(85, 1500, 2)
    big = gray_to_rgb_u8(g)
    g2, b2 = synth_image(86, 1024, 1)
    small_img = gray_to_rgb_u8(g2)
    tiny_stage.on_empty = "zeros"
    ref_big = tiny_stage.run_batch([big], [b], raw=True)[0]
    ref_small = tiny_stage.run_batch([small_img], [b2], raw=True)[0]
    st = SamStage("vit_t", device="cuda:0", state_dict=tiny_weights, max_batch=1, max_boxes=4, max_image_hw=(512, 512),
                  on_empty="zeros", precision=tiny_stage.precision)
    try:
        for img, bx, ref in ((small_img, b2, ref_small), (big, b, ref_big), (small_img, b2, ref_small)):
            m, r, _ = st.run_batch([img], [bx], raw=True)[0]
            assert np.array_equal(m, ref[0]) and r.tobytes() == ref[1].tobytes()
    finally:
        st.close()


def test_expanded_crops_come_with_the_metric_rows(tiny_stage):
    """f1: box crop (pipeline.py:379 convention) and the 2x-expanded mask-bbox crop the CSV consumers recompute."""
    from yolo_sam_inference_b200.sam_stage import expanded_crop_window
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    g, b = synth_image(88, 1024, 3)
    img = gray_to_rgb_u8(g)
    tiny_stage.on_empty = "zeros"
    masks, mets, crops, big = tiny_stage.run(img, b, expanded_crops=True)
    assert len(crops) == len(big) == 3
    for k in range(3):
        x1, y1, x2, y2 = b[k].astype(int)
        assert np.array_equal(crops[k], img[y1:y2, x1:x2])
        if mets[k]["area"] > 0:
            r0, r1, c0, c1 = expanded_crop_window(mets[k]["min_x"], mets[k]["min_y"], mets[k]["max_x"], mets[k]["max_y"], 1024, 1024)
            assert np.array_equal(big[k], img[r0:r1, c0:c1])


def test_two_contexts_on_two_devices_in_one_process(tiny_weights):
    """Distinct contexts are independent, also across GPUs of one process (the shared-memory opt-in of every kernel is
    per device). Needs two visible GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from yolo_sam_inference_b200.sam_stage import SamStage
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    g, b = synth_image(89, 1024, 2)
    img = gray_to_rgb_u8(g)
    a = SamStage("vit_t", device="cuda:0", state_dict=tiny_weights, max_batch=1, max_boxes=4, on_empty="zeros")
    c = SamStage("vit_t", device="cuda:1", state_dict=tiny_weights, max_batch=1, max_boxes=4, on_empty="zeros")
    try:
        ra = a.run_batch([img], [b], raw=True)[0]
        rc = c.run_batch([img], [b], raw=True)[0]
        assert np.array_equal(ra[0], rc[0]) and ra[1].tobytes() == rc[1].tobytes()
    finally:
        a.close(); c.close()
