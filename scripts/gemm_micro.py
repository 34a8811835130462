"""GEMM micro-benchmark on the ViT-B shapes (8 images): mainloop-only vs full epilogues, single CTA vs CTA pair."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_sam_inference_b200.sam_stage import SamStage
st = SamStage("vit_t", device="cuda:0", max_batch=1, max_boxes=1)
only = sys.argv[1:]   # optional: shape pair mode
shapes = {"qkv": (32768, 2304, 768), "proj": (32768, 768, 768), "fc1": (32768, 3072, 768), "fc2": (32768, 768, 3072)}
for name, (M, N, K) in shapes.items():
    for pair in (0, 1, 2):
        for mode in (2, 0, 1):
            if pair == 2 and mode == 2:
                continue
            if only and (name != only[0] or pair != int(only[1]) or mode != int(only[2])):
                continue
            ms = st.gemm_bench(M, N, K, pair, mode, 20)
            print(f"{name:5s} pair={pair} mode={mode} {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:7.0f} TF")
st.close()
