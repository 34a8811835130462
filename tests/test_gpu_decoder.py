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


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_decoder_many_boxes_kernels(tiny_weights, tiny_oracle, precision):
    """72 boxes in ONE launch (configs[3] regime: >= 64 boxes): register-resident token->image attention + merge, four-token
    image->token attention, pipelined fp32 token GEMM (504 token rows). Same gate as the few-box paths, and the result must
    not depend on how the boxes are chunked."""
    from yolo_sam_inference_b200.sam_stage import SamStage
    img, boxes, dumps, b1024 = _case(tiny_oracle, 77, 72)
    st = SamStage("vit_t", device="cuda:0", state_dict=tiny_weights, max_batch=1, max_boxes=72, precision=precision)
    try:
        low = st.decode(dumps["image_embeddings"], b1024)
        low8 = st.decode(dumps["image_embeddings"], b1024[:8])
    finally:
        st.close()
    ref = dumps["low_res_logits"]
    err = rel_l2(low, ref)
    print("decoder low-res logits at 72 boxes (%s) rel-L2 %.2e" % (precision, err))
    assert np.isfinite(low).all() and err < 2e-2
    # few-box kernels vs many-box kernels on the same boxes: same arithmetic up to summation order
    assert rel_l2(low[:8], low8) < 1e-4
