"""CPU ORACLE (test infrastructure, never the product path) -- SAM half of the hot path.

This is the reference's own arithmetic for /root/reference/src/yolo_sam_inference/pipeline.py:89-124
and :161-175, i.e. the third-party ``transformers`` (5.5.0 in this image; the reference pins only
``transformers>=4.30.0`` in requirements.txt:5) ``SamModel`` + ``SamProcessor`` executed in fp32 on
the CPU, driven exactly the way the reference drives them:

* ``preprocess``      -> pipeline.py:165-166   sam_processor(image, return_tensors="pt")
* ``rescale_boxes``   -> pipeline.py:97-102    sam_processor(image, input_boxes=[[box]])
* ``run_stage``       -> pipeline.py:161-175 / 105-124  (forward, post_process_masks, > 0.5)

Deviations, all result-neutral and verified in tests/test_oracle_sam.py:
  - ``return_tensors="pt"`` is NOT passed to ``post_process_masks`` (transformers 5.5.0 rejects that
    4.x-era kwarg with TypeError).
  - the image encoder runs once per image and its ``image_embeddings`` are reused for every box
    (the reference re-runs it per box with identical results, bit for bit).
  - boxes of one image may be batched through the decoder (agrees with box-at-a-time to ~1e-8).

Parity status: the reference ships no tests or golden vectors for this path, so the
oracle is pinned only by being the very library code the reference calls.  Golden fixtures under
tests/golden/ are generated HERE by tests/golden/make_golden.py from this module.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.
"""
from __future__ import annotations

import os
import sys
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from yolo_sam_inference_b200.weights import VARIANTS, SamVariant, seeded_state_dict  # noqa: E402


def build_model(variant: str = "vit_b", seed: int = 1234, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                logit_gain: float = 1.0):
    """SamModel(SamConfig(...)) in eval/fp32 with the shared seeded weights loaded (strict)."""
    from transformers import SamConfig, SamModel, SamVisionConfig

    v: SamVariant = VARIANTS[variant]
    vc = SamVisionConfig(hidden_size=v.hidden_size, num_hidden_layers=v.num_layers,
                         num_attention_heads=v.num_heads, global_attn_indexes=list(v.global_attn_indexes),
                         mlp_dim=v.mlp_dim)
    model = SamModel(SamConfig(vision_config=vc))
    sd = state_dict if state_dict is not None else seeded_state_dict(v, seed, logit_gain)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return model.eval().float()


_PROCESSOR = None


def processor():
    """SamProcessor(SamImageProcessor()) -- defaults equal facebook/sam-vit-*'s preprocessor config."""
    global _PROCESSOR
    if _PROCESSOR is None:
        from transformers import SamImageProcessor, SamProcessor
        _PROCESSOR = SamProcessor(SamImageProcessor())
    return _PROCESSOR


def preprocess(image_rgb_u8: np.ndarray):
    """pipeline.py:165-166.  Returns (pixel_values fp32[1,3,1024,1024], original (H,W), reshaped (h',w'))."""
    out = processor()(image_rgb_u8, return_tensors="pt")
    return (out["pixel_values"], tuple(int(x) for x in out["original_sizes"][0]),
            tuple(int(x) for x in out["reshaped_input_sizes"][0]))


def rescale_boxes(image_rgb_u8: np.ndarray, boxes_xyxy: np.ndarray) -> torch.Tensor:
    """pipeline.py:97-102 -> processing_sam.py:215-234: float64 boxes in the 1024-frame, [1,nb,4]."""
    H, W = image_rgb_u8.shape[:2]
    scale = 1024.0 / max(H, W)
    newh, neww = int(H * scale + 0.5), int(W * scale + 0.5)
    b = np.asarray(boxes_xyxy, dtype=np.float32).astype(np.float64).reshape(-1, 2, 2).copy()
    b[..., 0] = b[..., 0] * (neww / W)
    b[..., 1] = b[..., 1] * (newh / H)
    return torch.from_numpy(b.reshape(1, -1, 4))


@torch.no_grad()
def run_stage(model, image_rgb_u8: np.ndarray, boxes_xyxy: np.ndarray, dump: bool = False,
              per_box: bool = False) -> Tuple[np.ndarray, Dict[str, np.ndarray]]:
    """The SAM stage of pipeline.py:161-175 for one image.

    Returns (masks bool[nb,H,W], dumps).  ``dumps`` (when dump=True) holds the stage tensors the CUDA
    path is compared with: pixel_values, hidden_<i> after every encoder layer (NHWC), image_embeddings,
    sparse_embeddings, image_pe, low_res_logits [nb,256,256], upsampled_logits [nb,H,W].
    ``per_box`` runs the decoder one box at a time exactly like the reference loop (pipeline.py:170).
    """
    dumps: Dict[str, np.ndarray] = {}
    nb = int(len(boxes_xyxy))
    H, W = image_rgb_u8.shape[:2]
    if nb == 0:
        return np.zeros((0, H, W), bool), dumps
    pixel_values, orig, reshaped = preprocess(image_rgb_u8)
    hooks = []
    if dump:
        dumps["pixel_values"] = pixel_values[0].numpy().copy()
        enc = model.vision_encoder
        hooks.append(enc.patch_embed.register_forward_hook(
            lambda m, i, o: dumps.__setitem__("patch_embed", o[0].numpy().copy())))
        for li, layer in enumerate(enc.layers):
            hooks.append(layer.register_forward_hook(
                lambda m, i, o, li=li: dumps.__setitem__(f"hidden_{li}", o[0].numpy().copy())))
    emb = model.get_image_embeddings(pixel_values)
    for h in hooks:
        h.remove()
    boxes = rescale_boxes(image_rgb_u8, boxes_xyxy)
    if per_box:
        lows = []
        for k in range(nb):
            o = model(image_embeddings=emb, input_boxes=boxes[:, k:k + 1], multimask_output=False)
            lows.append(o.pred_masks[0, 0])
        low = torch.stack(lows, 0)                      # [nb,1,256,256]
    else:
        o = model(image_embeddings=emb, input_boxes=boxes, multimask_output=False)
        low = o.pred_masks[0]                           # [nb,1,256,256]
    ip = processor().image_processor
    up = ip.post_process_masks([low], [orig], [reshaped], binarize=False)[0]   # [nb,1,H,W] fp32
    masks = (up > 0.0)[:, 0].numpy()
    masks = masks > 0.5          # pipeline.py:123 (identity on bool)
    if dump:
        dumps["image_embeddings"] = emb[0].numpy().copy()
        sp, _ = model.prompt_encoder(input_points=None, input_labels=None, input_boxes=boxes, input_masks=None)
        dumps["sparse_embeddings"] = sp[0].numpy().copy()
        dumps["image_pe"] = model.get_image_wide_positional_embeddings()[0].numpy().copy()
        dumps["low_res_logits"] = low[:, 0].numpy().copy()
        dumps["upsampled_logits"] = up[:, 0].numpy().copy()
    return masks, dumps


@torch.no_grad()
def postprocess_logits(low_res: np.ndarray, orig: Tuple[int, int], reshaped: Tuple[int, int]
                       ) -> Tuple[np.ndarray, np.ndarray]:
    """image_processing_sam.py:410-430 on given low-res logits [nb,256,256] -> (upsampled fp32, masks)."""
    ip = processor().image_processor
    low = torch.from_numpy(np.ascontiguousarray(low_res, np.float32))[:, None]
    up = ip.post_process_masks([low], [tuple(orig)], [tuple(reshaped)], binarize=False)[0][:, 0]
    return up.numpy(), (up > 0.0).numpy()
