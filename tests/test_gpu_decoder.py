"""a4/a5: prompt encoder + mask decoder + upscaler vs transformers, fed with the ORACLE's image embeddings."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _case(tiny_oracle, idx, n_boxes):
    from oracle import sam_oracle
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    g, boxes = synth_image(idx, 1024, n_boxes)
    img = gray_to_rgb_u8(g)
    _, dumps = sam_oracle.run_stage(tiny_oracle, img, boxes, dump=True)
    b1024 = sam_oracle.rescale_boxes(img, boxes)[0].numpy()
    return img, boxes, dumps, b1024


def test_prompt_encoder_and_image_pe(tiny_stage, tiny_oracle):
    img, boxes, dumps, b1024 = _case(tiny_oracle, 4, 3)
    low, sparse = tiny_stage.decode(dumps["image_embeddings"], b1024, want_sparse=True)
    assert np.abs(sparse - dumps["sparse_embeddings"]).max() < 1e-5
    assert np.abs(tiny_stage.image_pe() - dumps["image_pe"]).max() < 1e-5


@pytest.mark.parametrize("n_boxes", [1, 5, 40])     # 40 boxes: 280 token rows -> tiled fp32 token GEMM path
def test_decoder_logits_parity(tiny_stage, tiny_oracle, n_boxes):
    img, boxes, dumps, b1024 = _case(tiny_oracle, 5 + n_boxes, n_boxes)
    low = tiny_stage.decode(dumps["image_embeddings"], b1024)
    ref = dumps["low_res_logits"]
    assert low.shape == ref.shape and np.isfinite(low).all()
    err = rel_l2(low, ref)
    print("decoder low-res logits rel-L2 %.2e, max-abs/max %.2e" % (err, np.abs(low - ref).max() / np.abs(ref).max()))
    assert err < 2e-2
