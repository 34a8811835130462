"""The drop-in entry points (pipeline.py mirror) on a real folder: files -> cv2.imread -> boxes -> SAM stage -> CSV rows,
single process and the multi-worker folder partition, vs the oracle run on the same files."""
import numpy as np
import pytest

from conftest import iou_gate

pytestmark = pytest.mark.gpu

CSV_COLUMNS = ["image_name", "cell_id", "deformability", "area", "area_ratio", "circularity", "convex_hull_area",
               "mask_x_length", "mask_y_length", "min_x", "min_y", "max_x", "max_y", "mean_brightness", "brightness_std",
               "perimeter", "aspect_ratio", "convex_hull_perimeter"]        # pipeline.py:293-305 + utils/metrics.py:102-119


def _make_folder(root):
    import cv2
    from yolo_sam_inference_b200.synth import synth_image
    d = root / "in"
    d.mkdir()
    boxes = {}
    g, b = synth_image(60, 1024, 2)
    cv2.imwrite(str(d / "a_000.tiff"), g)                                   # 8-bit single-channel TIFF (configs[0..3])
    boxes["a_000.tiff"] = b
    g, b = synth_image(61, 2048, 1)
    cv2.imwrite(str(d / "b_001.tiff"), g.astype(np.uint16) * 257)           # 16-bit TIFF, 2048x2048 (configs[4])
    boxes["b_001.tiff"] = b
    g, _ = synth_image(62, 1024, 1)
    cv2.imwrite(str(d / "c_002.png"), g)                                    # no detections: SAM is skipped
    boxes["c_002.png"] = np.zeros((0, 4), np.float32)
    g, b = synth_image(63, 1024, 3)
    cv2.imwrite(str(d / "d_003.tiff"), g)
    boxes["d_003.tiff"] = b
    return d, boxes


def test_process_directory_matches_oracle(tmp_path, tiny_weights, tiny_oracle):
    import cv2
    from oracle import metrics_oracle as mo
    from oracle import sam_oracle
    from yolo_sam_inference_b200.pipeline import (BatchProcessingResult, BoxTable, CellSegmentationPipeline,
                                                  ParallelCellSegmentationPipeline)
    d, boxes = _make_folder(tmp_path)
    kw = dict(detector=BoxTable(boxes), sam_state_dict=tiny_weights, max_boxes=8, max_image_hw=(2048, 2048), on_empty="zeros")
    pipe = CellSegmentationPipeline(None, "vit_t", device="cuda:0", **kw)
    res = pipe.process_directory(d, tmp_path / "out", save_visualizations=False)
    assert isinstance(res, BatchProcessingResult)
    assert [r.image_path.split("/")[-1] for r in res.results] == sorted(boxes)          # sorted file order
    assert [r.num_cells for r in res.results] == [2, 1, 0, 3]
    assert len(res.metrics_data) == 6 and all(list(row.keys()) == CSV_COLUMNS for row in res.metrics_data)
    assert res.total_timing["total_cells"] == 6
    for r in res.results:
        assert {"image_load", "yolo_detection", "sam_preprocess", "inference", "postprocess", "total_time",
                "cells_processed"} <= set(r.timing)
    # per image: masks vs the oracle on the same decoded file, metrics of OUR masks vs the oracle's metrics of them
    for r in res.results:
        name = r.image_path.split("/")[-1]
        if len(boxes[name]) == 0:
            assert r.cell_metrics == [] and r.timing["sam_preprocess"] == 0.0           # pipeline.py:176-179
            continue
        img = cv2.cvtColor(cv2.imread(r.image_path), cv2.COLOR_BGR2RGB)                 # pipeline.py:206-210
        assert img.dtype == np.uint8                                                    # 16-bit TIFF arrives as v >> 8
        ref_masks, _ = sam_oracle.run_stage(tiny_oracle, img, boxes[name], dump=True)
        masks, mets, _ = pipe.sam_stage.run(img, boxes[name])
        assert mets == r.cell_metrics
        for k in range(len(masks)):
            iou = np.logical_and(masks[k], ref_masks[k]).sum() / max(np.logical_or(masks[k], ref_masks[k]).sum(), 1)
            assert iou >= iou_gate(pipe.sam_stage.precision), (name, k, iou)
            if masks[k].any():
                ref = mo.calculate_metrics(img, masks[k])
                for key, val in ref.items():
                    if isinstance(val, int):
                        assert mets[k][key] == val, (key, mets[k][key], val)
                    else:
                        assert mets[k][key] == pytest.approx(val, rel=1e-9, abs=1e-12), key
    # folder partition over 2 worker processes (one ysi_ctx each, no collective): same rows in the same order
    par = ParallelCellSegmentationPipeline(None, "vit_t", device="cuda", num_pipelines=2, **kw)
    res2 = par.process_directory(d, tmp_path / "out2", save_visualizations=False)
    assert res2.metrics_data == res.metrics_data
    assert [r.num_cells for r in res2.results] == [2, 1, 0, 3]
