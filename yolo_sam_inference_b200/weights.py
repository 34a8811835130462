"""SAM weight inventory and the seeded test-weight recipe.

The C-ABI takes weights as a flat list of (name, fp32 host pointer, shape) with the names of
``transformers.SamModel.state_dict()`` (the object the reference builds at
/root/reference/src/yolo_sam_inference/pipeline.py:76).  This module knows that inventory without
importing transformers, so the product path has no dependency on it.

``seeded_state_dict`` is the shared test-weight recipe of SURVEY.md §8(c): checkpoints are not
available offline and the library default init is degenerate (encoder std 1e-10, zero positional
tables), so both the oracle and the CUDA path are driven with the same seeded, bf16-representable
weights.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np
import torch


@dataclass(frozen=True)
class SamVariant:
    name: str
    hidden_size: int
    num_layers: int
    num_heads: int
    global_attn_indexes: Tuple[int, ...]
    mlp_dim: int
    hf_id: str = ""

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads


# facebook/sam-vit-{base,large,huge} vision towers (configuration_sam.py:116-134 defaults are ViT-B);
# "vit_t" is a small test-only tower with the same head_dim so every kernel is exercised quickly.
VARIANTS: Dict[str, SamVariant] = {
    "vit_t": SamVariant("vit_t", 192, 4, 3, (1, 3), 768),
    "vit_t80": SamVariant("vit_t80", 160, 4, 2, (1, 3), 640),      # test-only tower with ViT-H's head_dim 80
    "vit_b": SamVariant("vit_b", 768, 12, 12, (2, 5, 8, 11), 3072, "facebook/sam-vit-base"),
    "vit_l": SamVariant("vit_l", 1024, 24, 16, (5, 11, 17, 23), 4096, "facebook/sam-vit-large"),
    "vit_h": SamVariant("vit_h", 1280, 32, 16, (7, 15, 23, 31), 5120, "facebook/sam-vit-huge"),
}

HF_ID_TO_VARIANT = {v.hf_id: k for k, v in VARIANTS.items() if v.hf_id}

WINDOW = 14
GRID = 64          # 1024 / 16 patch grid
DEC_C = 256        # prompt/decoder hidden size


def variant_of(sam_model_type: str) -> SamVariant:
    """Map the reference's ``sam_model_type`` string (pipeline.py:51) or a short name to a variant."""
    if sam_model_type in VARIANTS:
        return VARIANTS[sam_model_type]
    if sam_model_type in HF_ID_TO_VARIANT:
        return VARIANTS[HF_ID_TO_VARIANT[sam_model_type]]
    raise ValueError(f"unknown SAM model type {sam_model_type!r}")


def state_dict_shapes(v: SamVariant) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every tensor in SamModel.state_dict(), in its order."""
    D, hd = v.hidden_size, v.head_dim
    out: List[Tuple[str, Tuple[int, ...]]] = []
    a = out.append
    a(("shared_image_embedding.positional_embedding", (2, 128)))
    a(("vision_encoder.pos_embed", (1, GRID, GRID, D)))
    a(("vision_encoder.patch_embed.projection.weight", (D, 3, 16, 16)))
    a(("vision_encoder.patch_embed.projection.bias", (D,)))
    for i in range(v.num_layers):
        p = f"vision_encoder.layers.{i}."
        S = GRID if i in v.global_attn_indexes else WINDOW
        a((p + "layer_norm1.weight", (D,)))
        a((p + "layer_norm1.bias", (D,)))
        a((p + "attn.rel_pos_h", (2 * S - 1, hd)))
        a((p + "attn.rel_pos_w", (2 * S - 1, hd)))
        a((p + "attn.qkv.weight", (3 * D, D)))
        a((p + "attn.qkv.bias", (3 * D,)))
        a((p + "attn.proj.weight", (D, D)))
        a((p + "attn.proj.bias", (D,)))
        a((p + "layer_norm2.weight", (D,)))
        a((p + "layer_norm2.bias", (D,)))
        a((p + "mlp.lin1.weight", (v.mlp_dim, D)))
        a((p + "mlp.lin1.bias", (v.mlp_dim,)))
        a((p + "mlp.lin2.weight", (D, v.mlp_dim)))
        a((p + "mlp.lin2.bias", (D,)))
    a(("vision_encoder.neck.conv1.weight", (DEC_C, D, 1, 1)))
    a(("vision_encoder.neck.layer_norm1.weight", (DEC_C,)))
    a(("vision_encoder.neck.layer_norm1.bias", (DEC_C,)))
    a(("vision_encoder.neck.conv2.weight", (DEC_C, DEC_C, 3, 3)))
    a(("vision_encoder.neck.layer_norm2.weight", (DEC_C,)))
    a(("vision_encoder.neck.layer_norm2.bias", (DEC_C,)))
    a(("prompt_encoder.shared_embedding.positional_embedding", (2, 128)))
    a(("prompt_encoder.mask_embed.conv1.weight", (4, 1, 2, 2)))
    a(("prompt_encoder.mask_embed.conv1.bias", (4,)))
    a(("prompt_encoder.mask_embed.conv2.weight", (16, 4, 2, 2)))
    a(("prompt_encoder.mask_embed.conv2.bias", (16,)))
    a(("prompt_encoder.mask_embed.conv3.weight", (DEC_C, 16, 1, 1)))
    a(("prompt_encoder.mask_embed.conv3.bias", (DEC_C,)))
    a(("prompt_encoder.mask_embed.layer_norm1.weight", (4,)))
    a(("prompt_encoder.mask_embed.layer_norm1.bias", (4,)))
    a(("prompt_encoder.mask_embed.layer_norm2.weight", (16,)))
    a(("prompt_encoder.mask_embed.layer_norm2.bias", (16,)))
    a(("prompt_encoder.no_mask_embed.weight", (1, DEC_C)))
    for i in range(4):
        a((f"prompt_encoder.point_embed.{i}.weight", (1, DEC_C)))
    a(("prompt_encoder.not_a_point_embed.weight", (1, DEC_C)))
    a(("mask_decoder.iou_token.weight", (1, DEC_C)))
    a(("mask_decoder.mask_tokens.weight", (4, DEC_C)))

    def attn(prefix: str, internal: int) -> None:
        for n in ("q_proj", "k_proj", "v_proj"):
            a((f"{prefix}.{n}.weight", (internal, DEC_C)))
            a((f"{prefix}.{n}.bias", (internal,)))
        a((f"{prefix}.out_proj.weight", (DEC_C, internal)))
        a((f"{prefix}.out_proj.bias", (DEC_C,)))

    for i in range(2):
        p = f"mask_decoder.transformer.layers.{i}"
        attn(p + ".self_attn", 256)
        a((p + ".layer_norm1.weight", (DEC_C,)))
        a((p + ".layer_norm1.bias", (DEC_C,)))
        attn(p + ".cross_attn_token_to_image", 128)
        a((p + ".layer_norm2.weight", (DEC_C,)))
        a((p + ".layer_norm2.bias", (DEC_C,)))
        a((p + ".mlp.lin1.weight", (2048, DEC_C)))
        a((p + ".mlp.lin1.bias", (2048,)))
        a((p + ".mlp.lin2.weight", (DEC_C, 2048)))
        a((p + ".mlp.lin2.bias", (DEC_C,)))
        a((p + ".layer_norm3.weight", (DEC_C,)))
        a((p + ".layer_norm3.bias", (DEC_C,)))
        a((p + ".layer_norm4.weight", (DEC_C,)))
        a((p + ".layer_norm4.bias", (DEC_C,)))
        attn(p + ".cross_attn_image_to_token", 128)
    attn("mask_decoder.transformer.final_attn_token_to_image", 128)
    a(("mask_decoder.transformer.layer_norm_final_attn.weight", (DEC_C,)))
    a(("mask_decoder.transformer.layer_norm_final_attn.bias", (DEC_C,)))
    a(("mask_decoder.upscale_conv1.weight", (DEC_C, 64, 2, 2)))
    a(("mask_decoder.upscale_conv1.bias", (64,)))
    a(("mask_decoder.upscale_conv2.weight", (64, 32, 2, 2)))
    a(("mask_decoder.upscale_conv2.bias", (32,)))
    a(("mask_decoder.upscale_layer_norm.weight", (64,)))
    a(("mask_decoder.upscale_layer_norm.bias", (64,)))

    def ff(prefix: str, out_dim: int) -> None:
        a((prefix + ".proj_in.weight", (256, 256)))
        a((prefix + ".proj_in.bias", (256,)))
        a((prefix + ".proj_out.weight", (out_dim, 256)))
        a((prefix + ".proj_out.bias", (out_dim,)))
        a((prefix + ".layers.0.weight", (256, 256)))
        a((prefix + ".layers.0.bias", (256,)))

    for i in range(4):
        ff(f"mask_decoder.output_hypernetworks_mlps.{i}", 32)
    ff("mask_decoder.iou_prediction_head", 4)
    return out


def _bf16_representable(t: torch.Tensor) -> torch.Tensor:
    """Round to a value that BOTH 16-bit operand encodings hold exactly: bf16's 8 significand bits, and magnitudes
    below fp16's smallest normal (2^-14) flushed to zero (0.24 % of N(0, 0.02) draws)."""
    t = t.to(torch.bfloat16).to(torch.float32)
    return torch.where(t.abs() < 2.0 ** -14, torch.zeros_like(t), t)


def seeded_state_dict(variant: str | SamVariant, seed: int = 1234,
                      logit_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded, non-degenerate, bf16-representable SAM weights.

    * matrices / conv kernels ~ N(0, 0.02); biases ~ N(0, 0.02)
    * LayerNorm gamma ~ N(1, 0.05), beta ~ N(0, 0.05)
    * pos_embed ~ N(0, 0.02); rel_pos_h/w ~ N(0, 0.1) so the decomposed bias matters
    * the random-Fourier matrix ~ N(0, 1) (SAM's scale; the HF default of hidden_size//2 is a bug)
    * embeddings (tokens, point/no-mask embeds) ~ N(0, 0.5)
    Every value is rounded to an fp32 that is exactly representable in bf16 AND fp16, so the oracle and either
    operand encoding of the device path hold identical numbers.
    ``logit_gain`` scales the two upscaler convs (test knob only).
    """
    v = VARIANTS[variant] if isinstance(variant, str) else variant
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in state_dict_shapes(v):
        if name == "prompt_encoder.shared_embedding.positional_embedding":
            sd[name] = sd["shared_image_embedding.positional_embedding"]  # tied (modeling_sam.py:1110-1112)
            continue
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        if name.endswith("positional_embedding"):
            t = r
        elif "rel_pos" in name:
            t = 0.1 * r
        elif "layer_norm" in name and name.endswith(".weight"):
            t = 1.0 + 0.05 * r
        elif "layer_norm" in name and name.endswith(".bias"):
            t = 0.05 * r
        elif name in ("mask_decoder.iou_token.weight", "mask_decoder.mask_tokens.weight",
                      "prompt_encoder.no_mask_embed.weight", "prompt_encoder.not_a_point_embed.weight") \
                or name.startswith("prompt_encoder.point_embed."):
            t = 0.5 * r
        else:
            t = 0.02 * r
        if logit_gain != 1.0 and name.startswith("mask_decoder.upscale_conv"):
            t = t * logit_gain
        sd[name] = _bf16_representable(t).contiguous()
    return sd


def to_numpy_state_dict(sd: Dict[str, "torch.Tensor"]) -> Dict[str, np.ndarray]:
    """fp32 C-contiguous numpy views for the ctypes upload (ysi_load_weights)."""
    out = {}
    for k, t in sd.items():
        out[k] = np.ascontiguousarray(t.detach().to(torch.float32).cpu().numpy())
    return out
