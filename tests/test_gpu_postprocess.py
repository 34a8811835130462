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
