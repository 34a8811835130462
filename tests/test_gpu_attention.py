"""Fused windowed / global attention with decomposed rel-pos bias vs a plain torch fp32 restatement of
modeling_sam.py:761-801, 843-882 (through the C ABI, ysi_attention)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def reference_attention(qkv, rel_h, rel_w, heads, S):
    """qkv [B, S*S, 3*heads*hd] (already bf16-representable) -> [B, S*S, heads*hd], fp32 math."""
    B, T, _ = qkv.shape
    hd = qkv.shape[2] // (3 * heads)
    x = qkv.reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4).reshape(3, B * heads, T, hd)
    q, k, v = x[0], x[1], x[2]
    idx = torch.arange(S)[:, None] - torch.arange(S)[None, :] + (S - 1)
    Rh, Rw = rel_h[idx], rel_w[idx]                       # [S, S, hd]
    rq = q.reshape(B * heads, S, S, hd)
    bias = (torch.einsum("bhwc,hkc->bhwk", rq, Rh)[:, :, :, :, None] +
            torch.einsum("bhwc,wkc->bhwk", rq, Rw)[:, :, :, None, :]).reshape(B * heads, T, T)
    attn = torch.softmax((q * hd ** -0.5) @ k.transpose(-2, -1) + bias, dim=-1)
    out = (attn @ v).reshape(B, heads, T, hd).permute(0, 2, 1, 3).reshape(B, T, heads * hd)
    return out


@pytest.mark.parametrize("is_global,n_seq,heads,hd", [(False, 3, 3, 64), (False, 50, 2, 64), (True, 1, 2, 64), (True, 2, 3, 64),
                                                       (False, 27, 2, 80), (True, 1, 3, 80)])
def test_attention_matches_reference(tiny_stage, is_global, n_seq, heads, hd):
    g = torch.Generator().manual_seed(7 + n_seq + heads)
    S = 64 if is_global else 14
    T = S * S
    qkv = _bf16(torch.randn(n_seq, T, 3 * heads * hd, generator=g) * 1.2)
    rel_h = _bf16(torch.randn(2 * S - 1, hd, generator=g) * 0.15)
    rel_w = _bf16(torch.randn(2 * S - 1, hd, generator=g) * 0.15)
    ref = reference_attention(qkv, rel_h, rel_w, heads, S).numpy()
    out = tiny_stage.attention(qkv.numpy(), rel_h.numpy(), rel_w.numpy(), heads, is_global)
    assert np.isfinite(out).all()
    # P and the output are rounded to bf16 (2^-9); everything else is fp32
    assert rel_l2(out, ref) < 6e-3
    assert np.abs(out - ref).max() < 0.03 * np.abs(ref).max()


@pytest.mark.parametrize("is_global", [False, True])
def test_attention_growing_logits_exercise_lazy_rescale(tiny_stage, is_global):
    """Keys whose logits grow steadily along the sequence force the running maximum to move by far more
    than the lazy-rescale threshold (2^8), tile after tile; some query rows see it shrink instead."""
    g = torch.Generator().manual_seed(99)
    S = 64 if is_global else 14
    T, heads, n_seq = S * S, 2, 2
    qkv = torch.randn(n_seq, T, 3 * heads * 64, generator=g) * 0.5
    u = torch.randn(64, generator=g)
    u = u / u.norm()
    ramp = torch.linspace(0.0, 60.0, T)                      # up to ~ +60*|q.u|*0.125 in the exponent
    for h in range(heads):
        k0 = heads * 64 + h * 64
        qkv[:, :, k0:k0 + 64] += ramp[None, :, None] * u[None, None, :]
        qkv[:, :, h * 64:h * 64 + 64] += 2.0 * u[None, None, :] * torch.sign(torch.randn(n_seq, T, 1, generator=g))
    qkv = _bf16(qkv)
    rel_h = _bf16(torch.randn(2 * S - 1, 64, generator=g) * 0.15)
    rel_w = _bf16(torch.randn(2 * S - 1, 64, generator=g) * 0.15)
    ref = reference_attention(qkv, rel_h, rel_w, heads, S).numpy()
    out = tiny_stage.attention(qkv.numpy(), rel_h.numpy(), rel_w.numpy(), heads, is_global)
    assert np.isfinite(out).all()
    assert rel_l2(out, ref) < 8e-3
