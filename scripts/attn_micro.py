"""Device-resident timing of the encoder attention kernels at the bench shapes (ViT-B batch 8; ViT-H batch 8)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolo_sam_inference_b200.sam_stage import SamStage
from yolo_sam_inference_b200.weights import seeded_state_dict
st = SamStage("vit_t", device="cuda:0", state_dict=seeded_state_dict("vit_t", 1234), max_batch=1, max_boxes=2)
tag = sys.argv[1] if len(sys.argv) > 1 else ""
res = []
for name, n_seq, heads, hd, glob in (("global64", 8, 12, 64, True), ("window64", 200, 12, 64, False), ("global80", 8, 16, 80, True), ("window80", 200, 16, 80, False)):
    ms = st.attention_bench(n_seq, heads, hd, glob, 20)
    T, S = (4096, 64) if glob else (196, 14)
    fl = n_seq * heads * (4.0 * T * T * hd + 4.0 * T * S * hd)
    res.append("%s %.1f us %.0f TF/s" % (name, ms * 1e3, fl / ms / 1e9))
print("[%s]" % tag, " | ".join(res), flush=True)
st.close()
