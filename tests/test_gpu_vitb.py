"""BASELINE configs[0]/[1] model: ViT-B end to end on one synthetic image vs the fp32 oracle."""
import numpy as np
import pytest

from conftest import iou_gate, rel_l2

pytestmark = pytest.mark.gpu


def test_vit_b_stage_parity():
    from oracle import sam_oracle
    from yolo_sam_inference_b200.sam_stage import SamStage
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    from yolo_sam_inference_b200.weights import seeded_state_dict
    sd = seeded_state_dict("vit_b", 1234)
    model = sam_oracle.build_model("vit_b", state_dict=sd)
    g, boxes = synth_image(0, 1024, 1)
    img = gray_to_rgb_u8(g)
    ref_masks, d = sam_oracle.run_stage(model, img, boxes, dump=True)
    stage = SamStage("vit_b", device="cuda:0", state_dict=sd, max_batch=1, max_boxes=2, on_empty="zeros")
    try:
        emb, hid = stage.encode(d["pixel_values"][None], want_hidden=True)
        errs = [rel_l2(hid[i + 1, 0], d[f"hidden_{i}"]) for i in range(12)]
        e_emb = rel_l2(emb[0], d["image_embeddings"])
        b1024 = sam_oracle.rescale_boxes(img, boxes)[0].numpy()
        low_dec = stage.decode(d["image_embeddings"], b1024)          # decoder alone (oracle embeddings)
        low_e2e = stage.decode(emb[0], b1024)                         # our embeddings
        e_dec, e_e2e = rel_l2(low_dec, d["low_res_logits"]), rel_l2(low_e2e, d["low_res_logits"])
        masks, mets, _ = stage.run(img, boxes)
        iou = np.logical_and(masks[0], ref_masks[0]).sum() / max(np.logical_or(masks[0], ref_masks[0]).sum(), 1)
        frac_pos = ref_masks[0].mean()
        print("vit_b per-layer rel-L2:", ["%.1e" % e for e in errs])
        print("vit_b embeddings %.2e | logits dec-only %.2e e2e %.2e | IoU %.4f (ref positive frac %.3f)"
              % (e_emb, e_dec, e_e2e, iou, frac_pos))
        assert max(errs) < 2e-2 and e_emb < 2e-2 and e_dec < 2e-2 and e_e2e < 2e-2
        assert iou >= iou_gate(stage.precision)
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
        assert max(errs) < 2e-2 and e_emb < 2e-2
        masks, mets, _ = stage.run(img, boxes)
        for k in range(2):
            iou = np.logical_and(masks[k], ref_masks[k]).sum() / max(np.logical_or(masks[k], ref_masks[k]).sum(), 1)
            assert iou >= iou_gate(stage.precision)
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
        iou = np.logical_and(masks[0], ref_masks[0]).sum() / max(np.logical_or(masks[0], ref_masks[0]).sum(), 1)
        print("vit_h per-layer rel-L2 (every 4th):", ["%.1e" % e for e in errs[::4]], "last %.1e" % errs[-1])
        print("vit_h embeddings %.2e | IoU %.4f" % (e_emb, iou))
        assert max(errs) < 2e-2 and e_emb < 2e-2
        assert iou >= iou_gate(stage.precision)
    finally:
        stage.close()
