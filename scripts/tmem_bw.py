"""Tensor-memory read rate seen by a GEMM epilogue: drain-only GEMMs (EpiDrain: tcgen05.ld of the whole accumulator, no
global stores) with a short K, so that the TMEM -> register readout, not the MMA, bounds the tile time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolo_sam_inference_b200.sam_stage import SamStage
from yolo_sam_inference_b200.weights import seeded_state_dict
st = SamStage("vit_t", device="cuda:0", state_dict=seeded_state_dict("vit_t", 1234), max_batch=1, max_boxes=2)
for pair in (0, 1):
    for K in (64, 128, 256, 768):
        M, N = 148 * 128 * 8 * (2 if pair else 1), 256
        ms = st.gemm_bench(M, N, K, bool(pair), 2, 10)
        tiles_per_sm = M / 128 / 148 * (N / 256)
        clk = ms * 1e-3 * 1.7e9 / tiles_per_sm
        print("pair %d K %4d: %.3f ms, %.0f clks (at 1.7 GHz) per 128x256 fp32 accumulator tile and SM -> %.0f B/clk/SM" % (pair, K, ms, clk, 131072 / clk), flush=True)
st.close()
