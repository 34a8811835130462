"""CPU: host-side logic above the C ABI -- metric formulas, result contract, sharding (incl. 2-rank gloo)."""
import os
import sys

import numpy as np
import pytest

from oracle import metrics_oracle as mo
from test_gpu_metrics import FLOAT_KEYS, INT_KEYS, kat_masks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CSV_KEYS = ["deformability", "area", "area_ratio", "circularity", "convex_hull_area", "mask_x_length",
            "mask_y_length", "min_x", "min_y", "max_x", "max_y", "mean_brightness", "brightness_std", "perimeter",
            "aspect_ratio", "convex_hull_perimeter"]


def _raw_from_oracle(image, mask):
    """Build the ysi_mask_metrics row the kernels would produce, from the oracle's integer intermediates."""
    from yolo_sam_inference_b200 import _native as nat
    e = mo.calculate_metrics(image, mask, extras=True)
    r = np.zeros(1, dtype=nat.METRICS_DTYPE)[0]
    r["area"] = e["area"]; r["sum_r"] = e["_sum_r"]; r["sum_c"] = e["_sum_c"]
    r["min_r"], r["min_c"], r["max_r"], r["max_c"] = e["min_x"], e["min_y"], e["max_x"], e["max_y"]
    r["perim_hist"] = e["_perim_hist"]; r["hull_perim_hist"] = e["_hull_perim_hist"]
    r["hull_area"] = e["convex_hull_area"]
    r["disk_n"], r["disk_sum"], r["disk_sumsq"] = e["_disk_n"], e["_disk_sum"], e["_disk_sumsq"]
    r["flags"] = 2 if e["_hull_degenerate"] else 0
    return r, e


def test_metrics_from_raw_reproduces_reference_dict():
    from yolo_sam_inference_b200.sam_stage import metrics_from_raw
    rng = np.random.RandomState(3)
    image = rng.randint(0, 256, size=(64, 64, 3)).astype(np.uint8)
    for m in kat_masks():
        raw, ref = _raw_from_oracle(image, m)
        got = metrics_from_raw(raw)
        assert list(got.keys()) == CSV_KEYS                      # utils/metrics.py:102-119 order
        for k in INT_KEYS:
            assert type(got[k]) is int and got[k] == ref[k]
        for k in FLOAT_KEYS:
            assert type(got[k]) is float and got[k] == pytest.approx(ref[k], rel=1e-9, abs=1e-12)
        assert got["perimeter"] == ref["perimeter"]


def test_empty_mask_policy():
    from yolo_sam_inference_b200 import _native as nat
    from yolo_sam_inference_b200.sam_stage import metrics_from_raw
    r = np.zeros(1, dtype=nat.METRICS_DTYPE)[0]
    with pytest.raises(IndexError):
        metrics_from_raw(r, "raise")
    z = metrics_from_raw(r, "zeros")
    assert list(z.keys()) == CSV_KEYS and z["area"] == 0 and z["deformability"] == 1.0


def test_partition_matches_reference_rule():
    from yolo_sam_inference_b200.sharding import gather_in_order, partition_contiguous, shard_range
    import math
    for n in (0, 1, 7, 8, 9, 256, 4096, 10000):
        for w in (1, 2, 3, 4, 8):
            files = list(range(n))
            bs = math.ceil(n / w) if n else 0
            ref = [files[i:i + bs] for i in range(0, n, bs)] if n else []      # pipeline.py:540-541
            got = partition_contiguous(files, w)
            assert got == ref
            assert gather_in_order(got) == files
            assert [list(shard_range(n, r, w)) for r in range(len(got))] == got


def test_csv_row_contract(tmp_path):
    """_collect_metrics_data / _collect_timing_data rows feed reporting.save_results_to_csv unchanged."""
    import pandas as pd
    from yolo_sam_inference_b200.pipeline import CellSegmentationPipeline, ProcessingResult
    from yolo_sam_inference_b200.sam_stage import metrics_from_raw
    image = np.full((64, 64, 3), 90, np.uint8)
    raw, _ = _raw_from_oracle(image, kat_masks()[2])
    res = ProcessingResult("dir/img_0.tiff", [metrics_from_raw(raw)], 1,
                           {"image_load": 0.001, "yolo_detection": 0.0, "sam_preprocess": 0.002, "inference": 0.01,
                            "postprocess": 0.001, "total_time": 0.02, "cells_processed": 1})
    md, td = [], []
    CellSegmentationPipeline._collect_metrics_data(md, res)
    CellSegmentationPipeline._collect_timing_data(td, res)
    assert list(pd.DataFrame(md).columns) == ["image_name", "cell_id"] + CSV_KEYS
    assert list(td[0].keys())[:2] == ["image_name", "cells_processed"] and "inference_ms" in td[0]
    tot = CellSegmentationPipeline._initialize_timing_dict()
    CellSegmentationPipeline._update_total_timing(tot, res.timing)
    assert tot["total_cells"] == 1 and tot["image_load"] == 0.001


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        line = bench.host_dry_run(rank, world, n_images=10)
        if rank == 0:
            import json
            with open(os.path.join(out_dir, "line.json"), "w") as f:
                json.dump(line, f)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_on_gloo(tmp_path):
    """N>1 host path on CPU: 2 ranks shard the image list, no data-path collective, max-over-ranks timing."""
    import json
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    line = json.load(open(tmp_path / "line.json"))
    assert line["n_gpus"] == 2 and line["images_total"] == 10 and line["shards"] == [5, 5]
    assert line["scaling"] == "weak" or line["scaling"] == "strong"


def _reference_expanded_window(min_x, min_y, max_x, max_y, img_w, img_h):
    """examples/plot_scatter_example.py:114-140 restated (PIL box = (left, upper, right, lower))."""
    min_x_img, max_x_img, min_y_img, max_y_img = int(min_y), int(max_y), int(min_x), int(max_x)
    center_x = (min_x_img + max_x_img) // 2
    center_y = (min_y_img + max_y_img) // 2
    width, height = max_x_img - min_x_img, max_y_img - min_y_img
    new_width, new_height = int(width * 2.0), int(height * 2.0)
    min_x_img, max_x_img = center_x - (new_width // 2), center_x + (new_width // 2)
    min_y_img, max_y_img = center_y - (new_height // 2), center_y + (new_height // 2)
    min_x_img = max(0, min(min_x_img, img_w - 1))
    max_x_img = max(min_x_img + 1, min(max_x_img, img_w))
    min_y_img = max(0, min(min_y_img, img_h - 1))
    max_y_img = max(min_y_img + 1, min(max_y_img, img_h))
    return min_x_img, min_y_img, max_x_img, max_y_img


def test_expanded_crop_is_what_the_csv_consumers_cut():
    """f1: the 2x-expanded mask-bbox crop equals PIL's img.crop of the consumer's window, incl. clipping at every border."""
    from PIL import Image
    from yolo_sam_inference_b200 import _native as nat
    from yolo_sam_inference_b200.sam_stage import expanded_crop, expanded_crop_window
    rng = np.random.RandomState(11)
    H, W = 90, 140
    img = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    pil = Image.fromarray(img)
    for _ in range(200):
        r0, c0 = rng.randint(0, H - 1), rng.randint(0, W - 1)
        r1, c1 = rng.randint(r0 + 1, H + 1), rng.randint(c0 + 1, W + 1)
        box = _reference_expanded_window(r0, c0, r1, c1, W, H)
        ref = np.asarray(pil.crop(box))
        raw = np.zeros(1, dtype=nat.METRICS_DTYPE)[0]
        raw["min_r"], raw["min_c"], raw["max_r"], raw["max_c"] = r0, c0, r1, c1
        got = expanded_crop(img, raw)
        assert np.array_equal(got, ref)
        a, b, c, d = expanded_crop_window(r0, c0, r1, c1, H, W)
        assert (c, a, d, b) == box
    # raw grey samples give the same crop as the RGB image _load_image would have produced
    g16 = rng.randint(0, 65536, (H, W)).astype(np.uint16)
    rgb = np.repeat((g16 >> 8).astype(np.uint8)[:, :, None], 3, 2)
    raw = np.zeros(1, dtype=nat.METRICS_DTYPE)[0]
    raw["min_r"], raw["min_c"], raw["max_r"], raw["max_c"] = 10, 20, 40, 70
    assert np.array_equal(expanded_crop(g16, raw), expanded_crop(rgb, raw))


def test_visualisation_writer_layout(tmp_path):
    """f4: folders and file names of pipeline.py:352-432; the original image round-trips through the TIFF writer."""
    import cv2
    from yolo_sam_inference_b200.pipeline import CellSegmentationPipeline
    rng = np.random.RandomState(2)
    img = rng.randint(0, 256, (40, 50, 3)).astype(np.uint8)
    masks = np.zeros((2, 40, 50), bool)
    masks[0, 5:15, 5:20] = True
    masks[1, 20:30, 25:45] = True
    boxes = np.array([[4, 4, 21, 16], [24, 19, 46, 31]], np.float32)
    pipe = CellSegmentationPipeline.__new__(CellSegmentationPipeline)          # the writer needs no device
    pipe._save_visualizations(img, masks, boxes, [{}, {}], tmp_path / "run" / "cell_007.tiff")
    base = tmp_path / "run"
    expect = ["1_original_images/cell_007_original.tiff", "2_yolo_detections/cell_007_yolo.tiff",
              "3_processed_masks/masks/cell_007_mask_0.tiff", "3_processed_masks/masks/cell_007_mask_1.tiff",
              "3_processed_masks/overlay_images/cell_007_mask_0_overlay.tiff",
              "3_processed_masks/convex_hull_overlay/cell_007_mask_1_convex_hull.tiff",
              "4_combined_visualization/cell_007_combined.tiff"]
    for rel in expect:
        assert (base / rel).exists(), rel
    back = cv2.cvtColor(cv2.imread(str(base / expect[0])), cv2.COLOR_BGR2RGB)
    assert np.array_equal(back, img)
    m0 = cv2.imread(str(base / expect[2]), cv2.IMREAD_UNCHANGED)
    assert m0.dtype == np.uint8 and np.array_equal(m0 > 0, masks[0]) and set(np.unique(m0)) == {0, 255}
    comb = cv2.imread(str(base / expect[6]))
    assert comb.shape == (40, 100, 3)
    ov = cv2.cvtColor(cv2.imread(str(base / expect[4])), cv2.COLOR_BGR2RGB)
    exp = img.copy()
    exp[masks[0]] = exp[masks[0]] * 0.7 + np.array([255, 0, 0]) * 0.3
    assert np.array_equal(ov, exp)


def test_bench_folder_writer_repeats_and_names(tmp_path, monkeypatch):
    """bench.write_folder: `repeat` copies of the frames under distinct *.tiff names (what process_directory globs), one box
    table entry per file, every file a complete baseline TIFF that the raw-strip ingest accepts."""
    import shutil

    import bench
    from yolo_sam_inference_b200.ingest import probe_tiff
    frames = [np.full((32, 48), 10 * (i + 1), np.uint8) for i in range(3)]
    boxes = [np.array([[1.0, 2.0, 3.0 + i, 4.0]], np.float32) for i in range(3)]
    d, table = bench.write_folder(frames, boxes, repeat=2)
    try:
        names = sorted(os.listdir(d))
        assert names == sorted(table) and len(names) == 6 and all(n.endswith(".tiff") for n in names)
        for k, n in enumerate(names):
            assert np.array_equal(table[n], boxes[k % 3])
            info = probe_tiff(os.path.join(d, n))
            assert info is not None and tuple(info.shape) == (32, 48)
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_window_partition_maps_are_inverse():
    """The token -> window-row map the folded LayerNorm writes through (csrc/encoder.cu: build_tok_win_map_kernel) and the
    window-row -> token map of the windowed layers (window_row_to_token) restated in numpy: inverse on the 4096 real tokens,
    -1 exactly on the 804 pad rows of the 5 x 5 windows of 14 x 14."""
    tok = np.arange(4096)
    y, x = tok >> 6, tok & 63
    t2w = ((y // 14) * 5 + x // 14) * 196 + (y % 14) * 14 + x % 14
    w = np.arange(4900)
    win, l = w // 196, w % 196
    yy, xx = (win // 5) * 14 + l // 14, (win % 5) * 14 + l % 14
    w2t = np.where((yy < 64) & (xx < 64), yy * 64 + xx, -1)
    assert len(np.unique(t2w)) == 4096 and t2w.max() < 4900
    assert np.array_equal(w2t[t2w], tok)
    assert int((w2t < 0).sum()) == 804


def test_epilogue_slab_index_maps():
    """Index arithmetic of the GEMM epilogues' shared-memory slabs (csrc/gemm.cuh: slab_put / slab_flush / slab_load_*, slab64_*),
    restated: a lane's row written piece by piece comes back as whole rows in the flush order and vice versa, and every
    16-byte access of a quarter warp (8 lanes) goes to 8 different bank groups (128 B apart modulo 128 B)."""
    # 128-byte rows: put = lane L, piece j -> L*128 + ((j ^ (L & 7)) << 4); flush / load step j: lane -> row 4j + L//8, piece L%8
    put = {(L, j): L * 128 + ((j ^ (L & 7)) << 4) for L in range(32) for j in range(8)}
    assert sorted(put.values()) == list(range(0, 4096, 16))                         # a bijection onto the 4 KB slab
    for j in range(8):
        addrs = []
        for L in range(32):
            r, pc = 4 * j + L // 8, L % 8
            a = r * 128 + ((pc ^ (r & 7)) << 4)
            assert a == put[(r, pc)]                                                 # reads what row r put as its piece pc
            addrs.append(a)
        for q in range(4):                                                           # quarter warps: conflict-free wavefronts
            assert len({(a >> 4) & 7 for a in addrs[8 * q:8 * q + 8]}) == 8
        for q in range(4):                                                           # the per-row writes as well
            assert len({(put[(L, j)] >> 4) & 7 for L in range(8 * q, 8 * q + 8)}) == 8
    # 64-byte rows: put = L*64 + ((p ^ ((L >> 1) & 3)) << 4); flush step j: lane -> row 8j + L//4, piece L%4
    put64 = {(L, p): L * 64 + ((p ^ ((L >> 1) & 3)) << 4) for L in range(32) for p in range(4)}
    assert sorted(put64.values()) == list(range(0, 2048, 16))
    for j in range(4):
        addrs = []
        for L in range(32):
            r, pc = 8 * j + L // 4, L % 4
            a = r * 64 + ((pc ^ ((r >> 1) & 3)) << 4)
            assert a == put64[(r, pc)]
            addrs.append(a)
        for q in range(4):
            assert len({(a >> 4) & 7 for a in addrs[8 * q:8 * q + 8]}) == 8
    for p in range(4):
        for q in range(4):
            assert len({(put64[(L, p)] >> 4) & 7 for L in range(8 * q, 8 * q + 8)}) == 8


def test_folded_layernorm_statistics_merge():
    """The folded LayerNorm's row statistics (csrc/gemm.cuh: EpiResidLN writes (mean, sum of squared deviations) per group of
    columns, EpiStaged::prefetch merges them in slot order with Chan's update), restated in float32: equal to the two-pass
    mean / variance also for rows whose mean dwarfs their spread, where sum / sum-of-squares statistics lose every digit."""
    rng = np.random.RandomState(3)
    D, slots = 768, 8
    m = D // slots
    for offset in (0.0, 50.0, 3000.0):
        x = (rng.standard_normal(D) * 0.7 + offset).astype(np.float32)
        parts = []
        for s in range(slots):                       # producer: 32 columns at a time inside a slot
            cnt = np.float32(0); mean = np.float32(0); m2 = np.float32(0)
            for c in range(0, m, 32):
                v = x[s * m + c:s * m + c + 32]
                cm = np.float32(v.sum(dtype=np.float32) / np.float32(32))
                cM2 = np.float32(((v - cm) ** 2).sum(dtype=np.float32))
                d = cm - mean; tot = cnt + np.float32(32)
                mean = np.float32(mean + d * (np.float32(32) / tot))
                m2 = np.float32(m2 + cM2 + d * d * (cnt * np.float32(32) / tot))
                cnt = tot
            parts.append((mean, m2))
        mean = np.float32(0); m2 = np.float32(0); k = np.float32(0)
        for pm, pM2 in parts:                        # consumer
            d = pm - mean; r = np.float32(1) / (k + np.float32(1))
            mean = np.float32(mean + d * r)
            m2 = np.float32(m2 + pM2 + d * d * (np.float32(m) * k * r))
            k += np.float32(1)
        ref_mean, ref_var = float(x.astype(np.float64).mean()), float(x.astype(np.float64).var())
        assert abs(float(mean) - ref_mean) <= 1e-6 * max(1.0, abs(ref_mean))
        assert abs(float(m2) / D - ref_var) <= 1e-4 * ref_var
        naive = float(np.float32((x * x).sum(dtype=np.float32) / np.float32(D)) - np.float32(x.sum(dtype=np.float32) / np.float32(D)) ** 2)
        if offset >= 3000.0:
            assert abs(naive - ref_var) > 1e-2 * ref_var      # what the merge avoids
