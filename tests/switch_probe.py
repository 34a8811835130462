"""Helper of test_gpu_switches.py: one ViT-B encode + decode with whatever YSI_* switches the environment carries, results to an
.npz file.  The switches are read once per process by the library, so every configuration needs its own process.

    python tests/switch_probe.py <out.npz>
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from yolo_sam_inference_b200.sam_stage import SamStage
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image
    from yolo_sam_inference_b200.weights import seeded_state_dict
    g, boxes = synth_image(11, 1024, 4)
    img = gray_to_rgb_u8(g)
    st = SamStage("vit_b", device="cuda:0", state_dict=seeded_state_dict("vit_b", 1234), max_batch=1, max_boxes=8,
                  on_empty="zeros")
    pv = st.preprocess([img])
    emb = st.encode(pv)
    low = st.decode(emb[0], np.asarray(boxes, np.float64))          # 1024 x 1024 image: the boxes already are in model coordinates
    np.savez(sys.argv[1], emb=np.asarray(emb), low=np.asarray(low))
    st.close()


if __name__ == "__main__":
    main()
