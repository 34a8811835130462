"""One realistic launch of each encoder attention kernel (ViT-B batch 8 shapes) for ncu captures."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolo_sam_inference_b200.sam_stage import SamStage
from yolo_sam_inference_b200.weights import seeded_state_dict
which = sys.argv[1] if len(sys.argv) > 1 else "global"
hd = int(sys.argv[2]) if len(sys.argv) > 2 else 64
heads = 12 if hd == 64 else 16
st = SamStage("vit_t", device="cuda:0", state_dict=seeded_state_dict("vit_t", 1234), max_batch=1, max_boxes=2)
rng = np.random.RandomState(0)
is_global = which == "global"
S = 64 if is_global else 14
n_seq = 8 if is_global else 200
qkv = rng.standard_normal((n_seq, S * S, 3 * heads * hd)).astype(np.float32)
rh = (0.1 * rng.standard_normal((2 * S - 1, hd))).astype(np.float32)
rw = (0.1 * rng.standard_normal((2 * S - 1, hd))).astype(np.float32)
for _ in range(3):
    out = st.attention(qkv, rh, rw, heads, is_global)
print("ok", float(np.abs(out).mean()))
st.close()
