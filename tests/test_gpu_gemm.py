"""tcgen05 GEMM core vs a float64 reference on bf16-rounded operands (through the C ABI, ysi_gemm)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _bf16(a):
    return torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("M,N,K", [
    (128, 256, 64),        # one tile, one k-block
    (128, 64, 64),         # BN=64 path
    (256, 128, 192),       # BN=128 path, 3 k-blocks
    (4096, 768, 768),      # ViT-B proj
    (4900, 2304, 768),     # ViT-B qkv incl. window padding rows, ragged M tail
    (1000, 192, 200),      # ragged M, N=192 (vit_t), K not a multiple of 64
    (333, 3072, 768),      # fc1 shape, small M
    (4096, 256, 2304),     # neck 3x3 as GEMM
    (16384, 128, 64),      # ConvT2 shape
])
def test_gemm_matches_reference(tiny_stage, M, N, K):
    rng = np.random.RandomState(M + N + K)
    A = _bf16(rng.standard_normal((M, K)).astype(np.float32))
    W = _bf16((rng.standard_normal((N, K)) * 0.05).astype(np.float32))
    bias = rng.standard_normal(N).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64).T + bias
    out = tiny_stage.gemm(A, W, bias)
    assert out.shape == (M, N)
    assert np.isfinite(out).all()
    # fp32 accumulation of exact bf16 products: tolerance is fp32 round-off, not bf16
    assert rel_l2(out, ref) < 1e-5
    assert np.abs(out - ref).max() < 1e-3 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("act", [1, 2])
def test_gemm_activation_epilogues(tiny_stage, act):
    rng = np.random.RandomState(act)
    M, N, K = 512, 768, 192
    A = _bf16(rng.standard_normal((M, K)).astype(np.float32))
    W = _bf16((rng.standard_normal((N, K)) * 0.1).astype(np.float32))
    bias = rng.standard_normal(N).astype(np.float32)
    pre = torch.from_numpy(A.astype(np.float64) @ W.astype(np.float64).T + bias)
    ref = (torch.nn.functional.gelu(pre) if act == 1 else torch.relu(pre)).numpy()
    out = tiny_stage.gemm(A, W, bias, act=act)
    assert np.abs(out - ref).max() < 2e-5
