"""CPU: the SAM oracle (transformers) against the committed golden fixture and the reference's call pattern."""
import os

import numpy as np
import pytest

from conftest import rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def case(tiny_oracle):
    from oracle import sam_oracle
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    g, boxes = synth_image(1, 1024, 2)
    img = gray_to_rgb_u8(g)
    masks, d = sam_oracle.run_stage(tiny_oracle, img, boxes, dump=True)
    return img, boxes, masks, d


def test_weight_inventory_loads_strict(tiny_weights):
    from oracle import sam_oracle
    from yolo_sam_inference_b200.weights import VARIANTS, state_dict_shapes
    sam_oracle.build_model("vit_t", state_dict=tiny_weights)      # load_state_dict(strict=True) inside
    for t in tiny_weights.values():                               # bf16-representable by construction
        assert (t.to(dtype=t.dtype).bfloat16().float() == t).all()
    assert len(state_dict_shapes(VARIANTS["vit_b"])) == len(state_dict_shapes(VARIANTS["vit_t"])) + 8 * 14


def test_oracle_matches_golden_fixture(case):
    img, boxes, masks, d = case
    g = np.load(os.path.join(HERE, "golden", "sam_vit_t.npz"))
    assert np.array_equal(g["boxes"], boxes)
    assert float(d["pixel_values"].astype(np.float64).sum()) == pytest.approx(float(g["pixel_values_sum"]), rel=1e-9)
    assert rel_l2(d["hidden_3"][::8, ::8, ::4], g["hidden_last"]) < 1e-4
    assert rel_l2(d["image_embeddings"][:, ::8, ::8], g["image_embeddings"]) < 1e-4
    assert np.abs(d["sparse_embeddings"] - g["sparse_embeddings"]).max() < 1e-5
    assert rel_l2(d["low_res_logits"][:, ::4, ::4], g["low_res_logits"]) < 1e-3
    area = masks.reshape(len(masks), -1).sum(1)
    assert np.abs(area - g["mask_area"]).max() <= 0.002 * 1024 * 1024


def test_batched_boxes_equal_reference_per_box_loop(case, tiny_oracle):
    """pipeline.py:170 runs one box at a time; the oracle may batch boxes (result-neutral)."""
    from oracle import sam_oracle
    img, boxes, masks, d = case
    masks2, d2 = sam_oracle.run_stage(tiny_oracle, img, boxes, dump=True, per_box=True)
    assert np.abs(d["low_res_logits"] - d2["low_res_logits"]).max() < 1e-6
    assert (masks != masks2).mean() < 1e-5


def test_box_rescale_matches_processor(case):
    """processing_sam.py:215-234 via the real SamProcessor vs oracle.rescale_boxes (used by the C ABI shim too)."""
    from oracle import sam_oracle
    img = np.zeros((348, 704, 3), np.uint8)
    boxes = np.array([[10.5, 20.25, 100.0, 200.0], [0, 0, 703, 347]], np.float32)
    out = sam_oracle.processor()(img, input_boxes=[[b.tolist() for b in boxes]], return_tensors="pt")
    assert np.array_equal(out["input_boxes"][0].numpy(), sam_oracle.rescale_boxes(img, boxes)[0].numpy())


def test_postprocess_threshold_is_logit_gt_zero(case):
    img, boxes, masks, d = case
    assert np.array_equal(masks, d["upsampled_logits"] > 0.0)     # pipeline.py:123 `> 0.5` on bool is identity
