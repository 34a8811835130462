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
