"""a6: two-stage bilinear upsample + threshold vs transformers' post_process_masks (bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W", [(1024, 1024), (348, 704), (300, 704), (2048, 2048), (1000, 1500), (64, 2000)])
def test_postprocess_bit_exact(tiny_stage, H, W):
    from oracle import sam_oracle
    rng = np.random.RandomState(H * 7 + W)
    nb = 3
    low = (rng.standard_normal((nb, 256, 256)) * 4e-3).astype(np.float32)
    low[2] = np.abs(low[2]) * np.sign(rng.standard_normal((256, 256))).astype(np.float32)
    scale = 1024.0 / max(H, W)
    reshaped = (int(H * scale + 0.5), int(W * scale + 0.5))
    ref_up, ref_mask = sam_oracle.postprocess_logits(low, (H, W), reshaped)
    masks, up = tiny_stage.postprocess(low, H, W, want_logits=True)
    assert masks.shape == (nb, H, W)
    assert np.array_equal(up, ref_up), f"max diff {np.abs(up - ref_up).max()}"
    assert np.array_equal(masks, ref_mask)


def test_postprocess_edge_logits(tiny_stage):
    """zeros (not > 0), exact +-0 crossings and saturated values"""
    from oracle import sam_oracle
    low = np.zeros((2, 256, 256), np.float32)
    low[0, 100:140, 90:160] = 1.0
    low[0, 120, 120] = -1.0
    low[1] = -1.0
    low[1, 0, 0] = 5.0
    low[1, 255, 255] = 5.0
    ref_up, ref_mask = sam_oracle.postprocess_logits(low, (1024, 1024), (1024, 1024))
    masks, up = tiny_stage.postprocess(low, 1024, 1024, want_logits=True)
    assert np.array_equal(up, ref_up)
    assert np.array_equal(masks, ref_mask)


def test_postprocess_fast_path_1024(tiny_stage):
    """H = W = 1024 without the logit dump runs the bit-parallel kernel (csrc/postproc.cu, upsample_stats_fast_kernel):
    noise logits, exact zeros / sign changes at tile borders, the image corners and both signs of zero."""
    from oracle import sam_oracle
    rng = np.random.RandomState(11)
    low = (rng.standard_normal((5, 256, 256)) * 4e-3).astype(np.float32)
    low[1] = 0.0
    low[1, 100:140, 90:160] = 1.0
    low[1, 120, 120] = -1.0
    low[2] = -1.0
    low[2, 0, 0] = 5.0; low[2, 255, 255] = 5.0; low[2, 0, 255] = 5.0; low[2, 255, 0] = 5.0
    low[3] = np.where(rng.rand(256, 256) < 0.5, -0.0, 0.0).astype(np.float32)      # +-0 everywhere: nothing is > 0
    low[3, 7::16, :] = 1e-30
    low[4, :, 31:33] *= -1.0                                                        # sign flips across a lane boundary
    _, ref_mask = sam_oracle.postprocess_logits(low, (1024, 1024), (1024, 1024))
    masks = tiny_stage.postprocess(low, 1024, 1024)
    assert np.array_equal(masks, ref_mask)


def test_fast_path_stats_equal_generic_stats(tiny_stage):
    """The fused statistics of the fast kernel (area, centroid sums, bbox, perimeter codes, first contour cell -- seen
    through the metric rows) equal those the generic mask-statistics kernel derives from the same mask bytes."""
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    g, b = synth_image(77, 1024, 3)
    img = gray_to_rgb_u8(g)
    tiny_stage.on_empty = "zeros"
    (masks, rows, _), = tiny_stage.run_batch([img], [b], raw=True)     # ysi_run_batch: fast kernel
    again = tiny_stage.metrics(img, masks, raw=True)                    # ysi_metrics: generic kernel on the same bytes
    gray = (img.astype(np.int64).sum(2) // 3)
    for k in range(len(rows)):
        for name in rows.dtype.names:
            assert np.array_equal(rows[k][name], again[k][name]), (k, name, rows[k][name], again[k][name])
        assert np.array_equal(rows[k]["mask_hist"], np.bincount(gray[masks[k]], minlength=256))
        assert int(rows[k]["area"]) == int(masks[k].sum())
