"""The drop-in entry points (pipeline.py mirror) on a real folder: files -> loader -> boxes -> SAM stage -> CSV rows,
single process (batched, pipelined, raw-TIFF ingest) and the multi-worker folder partition, vs the oracle run on the same
files."""
import numpy as np
import pytest

from conftest import check_mask_parity

pytestmark = pytest.mark.gpu

CSV_COLUMNS = ["image_name", "cell_id", "deformability", "area", "area_ratio", "circularity", "convex_hull_area",
               "mask_x_length", "mask_y_length", "min_x", "min_y", "max_x", "max_y", "mean_brightness", "brightness_std",
               "perimeter", "aspect_ratio", "convex_hull_perimeter"]        # pipeline.py:293-305 + utils/metrics.py:102-119


def _make_folder(root, uncompressed):
    import cv2
    from yolo_sam_inference_b200.synth import synth_image
    d = root / ("in_raw" if uncompressed else "in_lzw")
    d.mkdir()
    flags = [cv2.IMWRITE_TIFF_COMPRESSION, 1] if uncompressed else []
    boxes = {}
    g, b = synth_image(60, 1024, 2)
    cv2.imwrite(str(d / "a_000.tiff"), g, flags)                            # 8-bit single-channel TIFF (configs[0..3])
    boxes["a_000.tiff"] = b
    g, b = synth_image(64, 1024, 1)
    cv2.imwrite(str(d / "a_001.tiff"), g, flags)                            # same size: shares a batch with a_000
    boxes["a_001.tiff"] = b
    g, b = synth_image(61, 2048, 1)
    cv2.imwrite(str(d / "b_001.tiff"), g.astype(np.uint16) * 257 + 91, flags)   # 16-bit TIFF, 2048x2048 (configs[4])
    boxes["b_001.tiff"] = b
    g, _ = synth_image(62, 1024, 1)
    cv2.imwrite(str(d / "c_002.png"), g)                                    # no detections: SAM is skipped
    boxes["c_002.png"] = np.zeros((0, 4), np.float32)
    g, b = synth_image(63, 1024, 3)
    cv2.imwrite(str(d / "d_003.tiff"), g, flags)
    boxes["d_003.tiff"] = b
    g, b = synth_image(65, 1024, 11)                                        # more boxes than max_boxes: chunked
    cv2.imwrite(str(d / "e_004.png"), g)
    boxes["e_004.png"] = b
    return d, boxes


def test_process_directory_matches_oracle(tmp_path, tiny_weights, tiny_oracle):
    import cv2
    from oracle import metrics_oracle as mo
    from oracle import sam_oracle
    from yolo_sam_inference_b200.pipeline import (BatchProcessingResult, BoxTable, CellSegmentationPipeline,
                                                  ParallelCellSegmentationPipeline)
    d, boxes = _make_folder(tmp_path, uncompressed=True)
    d2, _ = _make_folder(tmp_path, uncompressed=False)
    kw = dict(detector=BoxTable(boxes), sam_state_dict=tiny_weights, max_boxes=8, max_image_hw=(1024, 1024), on_empty="zeros",
              batch_size=2)
    pipe = CellSegmentationPipeline(None, "vit_t", device="cuda:0", mask_output="packed", **kw)
    res = pipe.process_directory(d, tmp_path / "out", save_visualizations=False)
    assert isinstance(res, BatchProcessingResult)
    assert [r.image_path.split("/")[-1] for r in res.results] == sorted(boxes)          # sorted file order
    assert [r.num_cells for r in res.results] == [2, 1, 1, 0, 3, 11]
    assert len(res.metrics_data) == 18 and all(list(row.keys()) == CSV_COLUMNS for row in res.metrics_data)
    assert res.total_timing["total_cells"] == 18
    for r in res.results:
        assert {"image_load", "yolo_detection", "sam_preprocess", "inference", "postprocess", "total_time",
                "cells_processed"} <= set(r.timing)
    # the same folder as LZW TIFFs goes through cv2.imread instead of the raw-strip route: identical rows
    res_lzw = pipe.process_directory(d2, tmp_path / "out_lzw", save_visualizations=False)
    assert res_lzw.metrics_data == res.metrics_data
    # ... and so does the one-image-at-a-time entry point
    for r in res.results:
        single = pipe.process_single_image(r.image_path, tmp_path / "o1" / "x.tiff", save_visualizations=False)
        assert single.cell_metrics == r.cell_metrics
    # per image: masks vs the oracle on the same decoded file, metrics of OUR masks vs the oracle's metrics of them
    for r in res.results:
        name = r.image_path.split("/")[-1]
        if len(boxes[name]) == 0:
            assert r.cell_metrics == [] and r.timing["sam_preprocess"] == 0.0           # pipeline.py:176-179
            continue
        img = cv2.cvtColor(cv2.imread(r.image_path), cv2.COLOR_BGR2RGB)                 # pipeline.py:206-210
        assert img.dtype == np.uint8                                                    # 16-bit TIFF arrives as v >> 8
        ref_masks, dd = sam_oracle.run_stage(tiny_oracle, img, boxes[name], dump=True)
        packed = pipe.last_masks[name]
        H, W = img.shape[:2]
        for k in range(len(ref_masks)):
            mask = np.unpackbits(packed[k])[:H * W].reshape(H, W).astype(bool)
            check_mask_parity(mask, ref_masks[k], dd["upsampled_logits"][k], pipe.sam_stage.precision, (name, k))
            if mask.any():
                ref = mo.calculate_metrics(img, mask)
                for key, val in ref.items():
                    if isinstance(val, int):
                        assert r.cell_metrics[k][key] == val, (key, r.cell_metrics[k][key], val)
                    else:
                        assert r.cell_metrics[k][key] == pytest.approx(val, rel=1e-9, abs=1e-12), key
    # visualisations on: the files the CSV consumers read exist and hold the loader's image
    res_vis = pipe.process_directory(d, tmp_path / "out_vis", save_visualizations=True)
    assert res_vis.metrics_data == res.metrics_data
    run_dir = tmp_path / "out_vis" / pipe.run_id
    orig = cv2.cvtColor(cv2.imread(str(run_dir / "1_original_images" / "b_001_original.tiff")), cv2.COLOR_BGR2RGB)
    assert np.array_equal(orig, cv2.cvtColor(cv2.imread(str(d / "b_001.tiff")), cv2.COLOR_BGR2RGB))
    assert (run_dir / "3_processed_masks" / "masks" / "e_004_mask_10.tiff").exists()
    pipe.close()
    # folder partition over 2 persistent worker processes (one ysi_ctx each, no collective): same rows in the same order,
    # also on the second call (workers, contexts and weights are reused)
    par = ParallelCellSegmentationPipeline(None, "vit_t", device="cuda", num_pipelines=2, **kw)
    try:
        res2 = par.process_directory(d, tmp_path / "out2", save_visualizations=False)
        assert res2.metrics_data == res.metrics_data
        assert [r.num_cells for r in res2.results] == [2, 1, 1, 0, 3, 11]
        pids = [p.pid for p, _ in par._workers]
        res3 = par.process_directory(d2, tmp_path / "out3", save_visualizations=False)
        assert res3.metrics_data == res.metrics_data and [p.pid for p, _ in par._workers] == pids
        img = cv2.cvtColor(cv2.imread(str(d / "a_000.tiff")), cv2.COLOR_BGR2RGB)
        par._ctor_kwargs["detector"].current_name = None
        bx, masks, scores = par.process_image(img)                                      # pipeline.py:469-503
        assert len(bx) == len(masks) == len(scores)
    finally:
        par.close()


def test_dead_worker_is_reported_not_waited_for(tmp_path, tiny_weights):
    """A worker killed by a native crash must surface as an error in the parent, not hang it."""
    import os
    import signal
    from yolo_sam_inference_b200.pipeline import BoxTable, ParallelCellSegmentationPipeline
    d, boxes = _make_folder(tmp_path, uncompressed=True)
    par = ParallelCellSegmentationPipeline(None, "vit_t", device="cuda", num_pipelines=1, detector=BoxTable(boxes),
                                           sam_state_dict=tiny_weights, on_empty="zeros")
    try:
        par._ensure_workers()
        os.kill(par._workers[0][0].pid, signal.SIGKILL)
        with pytest.raises(RuntimeError, match="died"):
            par.process_directory(d, tmp_path / "out", save_visualizations=False)
    finally:
        par.close()
