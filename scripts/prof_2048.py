"""Per-class device times of the folder route's device side for a batch of 2048x2048 16-bit grey images (configs[4] shape):
raw samples -> ingest -> resize -> encoder -> decoder -> upsample/metrics. Usage: python scripts/prof_2048.py [size] [boxes]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_sam_inference_b200.sam_stage import SamStage          # noqa: E402
from yolo_sam_inference_b200.synth import synth_image            # noqa: E402
from yolo_sam_inference_b200.weights import seeded_state_dict    # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nbox = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B = 8
imgs, boxes = [], []
for i in range(B):
    g, b = synth_image(9000 + i, size, nbox, bit_depth=16) if size != 1024 else synth_image(9000 + i, size, nbox)
    imgs.append(g)
    boxes.append(b)
print("image dtype", imgs[0].dtype, imgs[0].shape)
st = SamStage("vit_b", device="cuda:0", state_dict=seeded_state_dict("vit_b", 1234), max_batch=B, max_boxes=B * nbox,
              max_image_hw=(size, size), on_empty="zeros")
for _ in range(3):
    st.run_batch(imgs, boxes, masks="packed", raw=True)
st.profile(True)
N = 4
for _ in range(N):
    st.run_batch(imgs, boxes, masks="packed", raw=True)
prof = st.profile_read()
st.profile(False)
tot = 0.0
for k, v in prof.items():
    if v["ms"] > 0:
        print("%-14s %8.3f ms/batch  (%d records)" % (k, v["ms"] / N, v["records"] // N))
        tot += v["ms"] / N
print("total %.3f ms/batch -> %.1f images/s device-side" % (tot, B / tot * 1e3))
print("last timing", st.last_timing)
st.close()
