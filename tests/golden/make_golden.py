"""Generates the committed golden fixtures from the CPU oracle (run in the dev container):

    python tests/golden/make_golden.py

* metrics_kat.json  -- calculate_metrics rows for the known-answer masks of tests/test_gpu_metrics.kat_masks
* sam_vit_t.npz     -- subsampled stage tensors of the transformers SamModel ("vit_t" test tower, seed 1234)
                       on synthetic image 1 with 2 boxes

The reference ships no golden vectors; these are self-generated from the third-party
library code the reference calls (transformers 5.5.0) and from the restated skimage semantics.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import metrics_oracle as mo  # noqa: E402
from oracle import sam_oracle  # noqa: E402
from test_gpu_metrics import kat_masks  # noqa: E402
from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image  # noqa: E402


def main():
    rng = np.random.RandomState(3)
    image = rng.randint(0, 256, size=(64, 64, 3)).astype(np.uint8)
    rows = []
    for m in kat_masks():
        rows.append(mo.calculate_metrics(image, m))
    with open(os.path.join(HERE, "metrics_kat.json"), "w") as f:
        json.dump({"image_seed": 3, "rows": rows}, f, indent=1)

    model = sam_oracle.build_model("vit_t", 1234)
    g, boxes = synth_image(1, 1024, 2)
    img = gray_to_rgb_u8(g)
    masks, d = sam_oracle.run_stage(model, img, boxes, dump=True)
    np.savez_compressed(
        os.path.join(HERE, "sam_vit_t.npz"),
        boxes=boxes,
        pixel_values_sum=np.float64(d["pixel_values"].astype(np.float64).sum()),
        hidden_last=d["hidden_3"][::8, ::8, ::4].astype(np.float32),
        image_embeddings=d["image_embeddings"][:, ::8, ::8].astype(np.float32),
        sparse_embeddings=d["sparse_embeddings"].astype(np.float32),
        low_res_logits=d["low_res_logits"][:, ::4, ::4].astype(np.float32),
        mask_area=masks.reshape(len(masks), -1).sum(1).astype(np.int64),
    )
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
