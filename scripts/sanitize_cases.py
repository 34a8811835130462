"""Small invocations of every hand-written kernel family for compute-sanitizer (racecheck / synccheck / memcheck).

    compute-sanitizer --tool racecheck python scripts/sanitize_cases.py attn
    compute-sanitizer --tool synccheck python scripts/sanitize_cases.py tower
    ... python scripts/sanitize_cases.py vit_h         (one full ViT-H encode, batch 1)

Sizes are the smallest that still reach every code path (both head_dims, windowed and global tiles, CTA-pair GEMM,
staged epilogues, decoder at few / many boxes, fast and generic upsample paths)."""
import sys
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolo_sam_inference_b200.sam_stage import SamStage          # noqa: E402
from yolo_sam_inference_b200.synth import gray_to_rgb_u8, synth_image   # noqa: E402
from yolo_sam_inference_b200.weights import seeded_state_dict    # noqa: E402


def attn_cases(stage):
    rng = np.random.RandomState(0)
    for is_global, n_seq, heads, hd in ((False, 2, 2, 64), (False, 2, 1, 80), (True, 1, 1, 64), (True, 1, 1, 80)):
        S = 64 if is_global else 14
        qkv = rng.standard_normal((n_seq, S * S, 3 * heads * hd)).astype(np.float32)
        rh = (0.1 * rng.standard_normal((2 * S - 1, hd))).astype(np.float32)
        rw = (0.1 * rng.standard_normal((2 * S - 1, hd))).astype(np.float32)
        out = stage.attention(qkv, rh, rw, heads, is_global)
        print("attention", "global" if is_global else "window", hd, float(np.abs(out).mean()), flush=True)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "attn"
    if what == "attn":
        st = SamStage("vit_t", device="cuda:0", state_dict=seeded_state_dict("vit_t", 1234), max_batch=1, max_boxes=2)
        attn_cases(st)
        st.close()
    elif what in ("tower", "tower80"):
        name = "vit_t" if what == "tower" else "vit_t80"
        st = SamStage(name, device="cuda:0", state_dict=seeded_state_dict(name, 1234), max_batch=2, max_boxes=40,
                      max_image_hw=(1024, 1024), on_empty="zeros")
        imgs, boxes = [], []
        for i, nb in ((0, 1), (1, 3)):
            g, b = synth_image(i, 1024, nb)
            imgs.append(gray_to_rgb_u8(g)); boxes.append(b)
        out = st.run_batch(imgs, boxes)
        print("run_batch", [len(o[1]) for o in out], flush=True)
        g, b = synth_image(2, 1024, 20)                       # many boxes: streaming t2i kernel + tiled token GEMM
        out = st.run_batch([gray_to_rgb_u8(g)], [b])
        print("run_batch 20 boxes", len(out[0][1]), flush=True)
        g, b = synth_image(3, 700, 2)                         # generic (non-1024) pre/post-processing path
        im = np.ascontiguousarray(gray_to_rgb_u8(g)[:348, :700])
        b = np.clip(b, 0, [699, 347, 699, 347]).astype(np.float32)
        out = st.run(im, b)
        print("run 348x700", len(out[1]), flush=True)
        st.close()
    elif what == "vit_h":
        st = SamStage("vit_h", device="cuda:0", state_dict=seeded_state_dict("vit_h", 1234), max_batch=1, max_boxes=2, on_empty="zeros")
        g, b = synth_image(0, 1024, 1)
        out = st.run(gray_to_rgb_u8(g), b)
        print("vit_h run", len(out[1]), flush=True)
        st.close()
    print("done", what, flush=True)


if __name__ == "__main__":
    main()
