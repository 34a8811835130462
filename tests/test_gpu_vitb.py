"""The benchmarked models end to end vs the fp32 oracle: ViT-B on the 8 images of BASELINE configs[0] (one batch of 8, as
the bench runs it), ViT-B at 32 boxes per image (configs[3]: the many-box decoder kernels), a head_dim-80 tower and ViT-H."""
import numpy as np
import pytest

from conftest import LOGIT_TOL, check_mask_parity, rel_l2

pytestmark = pytest.mark.gpu


def _metrics_equal(mo, img, mask, got):
    ref = mo.calculate_metrics(img, mask)
    for key, val in ref.items():
        if isinstance(val, int):
            assert got[key] == val, (key, got[key], val)
        else:
            assert got[key] == pytest.approx(val, rel=1e-9, abs=1e-12), key


@pytest.fixture(scope="module")
def vit_b():
    from oracle import sam_oracle
    from yolo_sam_inference_b200.weights import seeded_state_dict
    sd = seeded_state_dict("vit_b", 1234)
    return sd, sam_oracle.build_model("vit_b", state_dict=sd)


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_vit_b_config0_eight_images(vit_b, precision):
    """configs[0]: ViT-B, the 8 synthetic 1024x1024 images (RandomState(1000 + i)), 1 box each -- processed as ONE batch of
    8 through ysi_run_batch (the unit the bench times): per-layer hidden states, embeddings, logits, masks and the metrics
    of every mask against the reference path on the same inputs."""
    from oracle import metrics_oracle as mo
    from oracle import sam_oracle
    from yolo_sam_inference_b200.sam_stage import SamStage
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    sd, model = vit_b
    imgs, boxes, refs = [], [], []
    for i in range(8):
        g, b = synth_image(i, 1024, 1)
        im = gray_to_rgb_u8(g)
        rm, d = sam_oracle.run_stage(model, im, b, dump=(i == 0) or True)
        imgs.append(im); boxes.append(b); refs.append((rm, d))
    stage = SamStage("vit_b", device="cuda:0", state_dict=sd, max_batch=8, max_boxes=8, on_empty="zeros", precision=precision)
    try:
        pv = np.stack([d["pixel_values"] for _, d in refs])
        emb, hid = stage.encode(pv, want_hidden=True)
        errs = [max(rel_l2(hid[li + 1, i], refs[i][1][f"hidden_{li}"]) for i in range(8)) for li in range(12)]
        e_emb = max(rel_l2(emb[i], refs[i][1]["image_embeddings"]) for i in range(8))
        print("vit_b/%s per-layer rel-L2 (max over 8 images):" % precision, ["%.1e" % e for e in errs], "embeddings %.2e" % e_emb)
        assert max(errs) < LOGIT_TOL and e_emb < LOGIT_TOL
        e_dec, e_e2e = [], []
        for i in range(8):
            b1024 = sam_oracle.rescale_boxes(imgs[i], boxes[i])[0].numpy()
            e_dec.append(rel_l2(stage.decode(refs[i][1]["image_embeddings"], b1024), refs[i][1]["low_res_logits"]))
            e_e2e.append(rel_l2(stage.decode(emb[i], b1024), refs[i][1]["low_res_logits"]))
        print("vit_b/%s logits rel-L2: decoder alone max %.2e, end to end max %.2e" % (precision, max(e_dec), max(e_e2e)))
        assert max(e_dec) < LOGIT_TOL and max(e_e2e) < LOGIT_TOL
        out = stage.run_batch(imgs, boxes)
        ious = []
        for i, (masks, mets, _) in enumerate(out):
            rm, d = refs[i]
            ious.append(check_mask_parity(masks[0], rm[0], d["upsampled_logits"][0], precision, i))
            if masks[0].any():
                _metrics_equal(mo, imgs[i], masks[0], mets[0])
        print("vit_b/%s IoU of the 8 masks:" % precision, ["%.4f" % v for v in ious], "positive fraction %.3f" % refs[0][0][0].mean())
    finally:
        stage.close()


def test_vit_b_thirty_two_boxes_end_to_end(vit_b):
    """configs[3]: ViT-B with 32 boxes on one image -- the streaming token->image attention and the tiled token GEMM (the
    kernels the 32-box bench runs) compared end to end on the benchmarked model: logits, masks, metrics of our masks."""
    from oracle import metrics_oracle as mo
    from oracle import sam_oracle
    from yolo_sam_inference_b200.sam_stage import SamStage
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    sd, model = vit_b
    g, b = synth_image(300, 1024, 32)
    im = gray_to_rgb_u8(g)
    rm, d = sam_oracle.run_stage(model, im, b, dump=True)
    stage = SamStage("vit_b", device="cuda:0", state_dict=sd, max_batch=1, max_boxes=32, on_empty="zeros")
    try:
        emb = stage.encode(d["pixel_values"][None])
        b1024 = sam_oracle.rescale_boxes(im, b)[0].numpy()
        low = stage.decode(emb[0], b1024)
        errs = [rel_l2(low[k], d["low_res_logits"][k]) for k in range(32)]
        print("vit_b 32 boxes: logits rel-L2 max %.2e mean %.2e" % (max(errs), float(np.mean(errs))))
        assert max(errs) < LOGIT_TOL
        masks, mets, _ = stage.run(im, b)
        ious = [check_mask_parity(masks[k], rm[k], d["upsampled_logits"][k], stage.precision, k) for k in range(32)]
        print("vit_b 32 boxes: IoU min %.4f mean %.4f" % (min(ious), float(np.mean(ious))))
        for k in (0, 13, 31):
            if masks[k].any():
                _metrics_equal(mo, im, masks[k], mets[k])
    finally:
        stage.close()


def test_head_dim_80_tower_parity():
    """ViT-H's head_dim (80): small 4-layer tower (2 windowed + 2 global layers) end to end vs the fp32 oracle."""
    from oracle import sam_oracle
    from yolo_sam_inference_b200.sam_stage import SamStage
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    from yolo_sam_inference_b200.weights import seeded_state_dict
    sd = seeded_state_dict("vit_t80", 1234)
    model = sam_oracle.build_model("vit_t80", state_dict=sd)
    g, boxes = synth_image(3, 1024, 2)
    img = gray_to_rgb_u8(g)
    ref_masks, d = sam_oracle.run_stage(model, img, boxes, dump=True)
    stage = SamStage("vit_t80", device="cuda:0", state_dict=sd, max_batch=1, max_boxes=2, on_empty="zeros")
    try:
        emb, hid = stage.encode(d["pixel_values"][None], want_hidden=True)
        errs = [rel_l2(hid[i + 1, 0], d[f"hidden_{i}"]) for i in range(4)]
        e_emb = rel_l2(emb[0], d["image_embeddings"])
        print("vit_t80 per-layer rel-L2:", ["%.1e" % e for e in errs], "embeddings %.2e" % e_emb)
        assert max(errs) < LOGIT_TOL and e_emb < LOGIT_TOL
        masks, mets, _ = stage.run(img, boxes)
        for k in range(2):
            check_mask_parity(masks[k], ref_masks[k], d["upsampled_logits"][k], stage.precision, k)
    finally:
        stage.close()


def test_vit_h_stage_parity():
    """BASELINE configs[2] model (the reference's default sam_model_type, pipeline.py:51): ViT-H, D=1280, 32 layers,
    16 heads of 80, global layers 7/15/23/31 -- embeddings, logits and masks vs the fp32 oracle on one image."""
    from oracle import sam_oracle
    from yolo_sam_inference_b200.sam_stage import SamStage
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    from yolo_sam_inference_b200.weights import seeded_state_dict
    sd = seeded_state_dict("vit_h", 1234)
    model = sam_oracle.build_model("vit_h", state_dict=sd)
    g, boxes = synth_image(0, 1024, 1)
    img = gray_to_rgb_u8(g)
    ref_masks, d = sam_oracle.run_stage(model, img, boxes, dump=True)
    del model
    stage = SamStage("vit_h", device="cuda:0", state_dict=sd, max_batch=1, max_boxes=2, on_empty="zeros")
    try:
        emb, hid = stage.encode(d["pixel_values"][None], want_hidden=True)
        errs = [rel_l2(hid[i + 1, 0], d[f"hidden_{i}"]) for i in range(32)]
        e_emb = rel_l2(emb[0], d["image_embeddings"])
        masks, mets, _ = stage.run(img, boxes)
        v = check_mask_parity(masks[0], ref_masks[0], d["upsampled_logits"][0], stage.precision)
        print("vit_h per-layer rel-L2 (every 4th):", ["%.1e" % e for e in errs[::4]], "last %.1e" % errs[-1])
        print("vit_h embeddings %.2e | IoU %.4f" % (e_emb, v))
        assert max(errs) < LOGIT_TOL and e_emb < LOGIT_TOL
    finally:
        stage.close()
