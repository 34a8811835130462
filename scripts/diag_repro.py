"""Localise run-to-run non-determinism (compute-sanitizer is closed on this pool): run a kernel family several times on
the same input and report where the outputs differ bitwise."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolo_sam_inference_b200.sam_stage import SamStage          # noqa: E402
from yolo_sam_inference_b200.weights import seeded_state_dict    # noqa: E402


def diff_report(name, a, b, shape_names):
    d = a.view(np.uint32) != b.view(np.uint32) if a.dtype == np.float32 else a != b
    n = int(d.sum())
    print(f"{name}: {n} of {d.size} elements differ", flush=True)
    if n:
        idx = np.argwhere(d)
        for ax, nm in enumerate(shape_names):
            vals, cnt = np.unique(idx[:, ax], return_counts=True)
            print(f"   axis {nm}: {len(vals)} distinct values, first {vals[:12].tolist()} counts {cnt[:12].tolist()}")
        rel = np.abs(a[d].astype(np.float64) - b[d]) / np.maximum(np.abs(b[d]), 1e-30)
        print(f"   rel diff median {np.median(rel):.2e} max {rel.max():.2e}; abs max {np.abs(a[d] - b[d]).max():.3e}")
    return n


def main():
    what = sys.argv[1:] or ["attn", "gemm", "enc"]
    rng = np.random.RandomState(0)
    if "attn" in what:
        st = SamStage("vit_t", device="cuda:0", state_dict=seeded_state_dict("vit_t", 1234), max_batch=1, max_boxes=2)
        for is_global, n_seq, heads, hd in ((False, 200, 16, 80), (False, 200, 12, 64), (True, 8, 16, 80), (True, 8, 12, 64)):
            S = 64 if is_global else 14
            qkv = rng.standard_normal((n_seq, S * S, 3 * heads * hd)).astype(np.float32)
            rh = (0.1 * rng.standard_normal((2 * S - 1, hd))).astype(np.float32)
            rw = (0.1 * rng.standard_normal((2 * S - 1, hd))).astype(np.float32)
            ref = st.attention(qkv, rh, rw, heads, is_global)
            tot = 0
            for r in range(4):
                out = st.attention(qkv, rh, rw, heads, is_global)
                tot += diff_report(f"attention {'global' if is_global else 'window'} hd{hd} n_seq {n_seq} run {r}",
                                   out.reshape(n_seq, S * S, heads, hd), ref.reshape(n_seq, S * S, heads, hd), ["seq", "token", "head", "ch"])
            print("   => total", tot, flush=True)
        st.close()
    if "gemm" in what:
        st = SamStage("vit_t", device="cuda:0", state_dict=seeded_state_dict("vit_t", 1234), max_batch=1, max_boxes=2)
        for M, N, K, act, kind in ((39200, 3840, 1280, 0, 1), (32768, 1280, 1280, 0, 2), (32768, 5120, 1280, 1, 1), (32768, 1280, 5120, 0, 2)):
            A = rng.standard_normal((M, K)).astype(np.float32)
            W = (0.05 * rng.standard_normal((N, K))).astype(np.float32)
            bias = rng.standard_normal(N).astype(np.float32)
            C0 = rng.standard_normal((M, N)).astype(np.float32)
            ref = st.gemm_ex(A, W, bias, act, kind, C0)
            tot = 0
            for r in range(3):
                out = st.gemm_ex(A, W, bias, act, kind, C0)
                tot += diff_report(f"gemm M{M} N{N} K{K} act{act} kind{kind} run {r}", out, ref, ["row", "col"])
            print("   => total", tot, flush=True)
        st.close()
    if "enc" in what:
        variant = os.environ.get("DIAG_VARIANT", "vit_h")
        batch = int(os.environ.get("DIAG_BATCH", "8"))
        st = SamStage(variant, device="cuda:0", state_dict=seeded_state_dict(variant, 1234), max_batch=batch, max_boxes=2)
        pv = rng.standard_normal((batch, 3, 1024, 1024)).astype(np.float32)
        emb0, hid0 = st.encode(pv, want_hidden=True)
        D = hid0.shape[-1]
        hd = 80 if variant in ("vit_h", "vit_t80") else 64
        for r in range(3):
            emb, hid = st.encode(pv, want_hidden=True)
            for li in range(hid.shape[0]):
                if not np.array_equal(hid[li], hid0[li]):
                    a = hid[li].reshape(batch, 64, 64, D // hd, hd)
                    b = hid0[li].reshape(batch, 64, 64, D // hd, hd)
                    diff_report(f"{variant} hidden slot {li} (first differing) run {r}", a, b, ["image", "y", "x", "headcol", "ch"])
                    break
            else:
                print(f"{variant} run {r}: all hidden states identical", flush=True)
        st.close()


if __name__ == "__main__":
    main()
