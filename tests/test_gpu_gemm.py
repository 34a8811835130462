"""tcgen05 GEMM core vs a float64 reference on bf16-rounded operands (through the C ABI, ysi_gemm)."""
import numpy as np
import pytest
import torch

from conftest import op16_round, rel_l2

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


@pytest.mark.parametrize("M,N,K,act,out_kind", [
    (4900, 2304, 768, 0, 1),     # qkv: bf16 output through the staged TMA-store epilogue, ragged M tail
    (4096, 3072, 768, 1, 1),     # fc1: GELU + bf16 output
    (4100, 768, 3072, 0, 2),     # fc2: residual add through cp.reduce.async.bulk, ragged M tail
    (2048, 768, 768, 0, 2),      # proj
    (512, 768, 768, 0, 2),       # small M: single-CTA kernel, red.global.add epilogue
    (512, 192, 192, 1, 1),       # small N: single-CTA kernel, direct bf16 stores
])
def test_gemm_production_epilogues(tiny_stage, M, N, K, act, out_kind):
    rng = np.random.RandomState(M + N + K + act)
    A = _bf16(rng.standard_normal((M, K)).astype(np.float32))
    W = _bf16((rng.standard_normal((N, K)) * 0.05).astype(np.float32))
    bias = rng.standard_normal(N).astype(np.float32)
    pre = torch.from_numpy(A.astype(np.float64) @ W.astype(np.float64).T + bias)
    ref = (torch.nn.functional.gelu(pre) if act == 1 else pre).numpy()
    if out_kind == 2:
        C0 = rng.standard_normal((M, N)).astype(np.float32)
        out = tiny_stage.gemm_ex(A, W, bias, act, 2, C0)
        assert np.abs(out - (C0.astype(np.float64) + ref)).max() < 2e-5 * max(1.0, np.abs(ref).max())
    else:
        out = tiny_stage.gemm_ex(A, W, bias, act, 1)
        ref_bf = op16_round(ref, tiny_stage.precision)
        # 16-bit rounding of a value that is itself only fp32-accurate: allow one ulp of the operand encoding
        ulp = 2.0 ** -7 if tiny_stage.precision == "bf16" else 2.0 ** -10
        assert np.abs(out - ref_bf).max() <= ulp * max(1.0, np.abs(ref).max())
        assert rel_l2(out, ref) < (4e-3 if tiny_stage.precision == "bf16" else 5e-4)
