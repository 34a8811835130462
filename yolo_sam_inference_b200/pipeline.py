"""Drop-in mirror of the reference pipeline entry points, with the SAM stage on libysi.so.

Mirrors /root/reference/src/yolo_sam_inference/pipeline.py: same class names, constructor arguments,
``process_single_image`` / ``process_directory`` signatures, ``ProcessingResult`` /
``BatchProcessingResult`` dataclasses (:31-45), timing keys (:143-194, :271-283) and CSV row builders
(:293-317).  What changes is only the body of ``if len(boxes) > 0:`` (:161-175), which becomes one
``SamStage.run`` call.  The YOLO detector stays on the reference's torch path (ultralytics, :72-73,
:84-87); because ultralytics is not installable offline, any callable ``image -> float32[N,4] xyxy``
can be injected as ``detector`` instead.

Out of scope here: visualisation / TIFF writers (:331-438), MLflow, ROI web UI.
"""
from __future__ import annotations

import logging
import math
import time
import uuid
from dataclasses import dataclass
from datetime import datetime
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .sam_stage import SamStage
from .sharding import partition_contiguous

logger = logging.getLogger(__name__)
logger.setLevel(logging.WARNING)


@dataclass
class ProcessingResult:
    """Per-image result (pipeline.py:31-37)."""
    image_path: str
    cell_metrics: List[Dict[str, Any]]
    num_cells: int
    timing: Dict[str, float]


@dataclass
class BatchProcessingResult:
    """Per-directory result (pipeline.py:39-45)."""
    results: List[ProcessingResult]
    total_timing: Dict[str, float]
    metrics_data: List[Dict[str, Any]]
    timing_data: List[Dict[str, Any]]


Detector = Callable[[np.ndarray], np.ndarray]


class BoxTable:
    """Detector stand-in: boxes looked up by image file name (synthetic 'YOLO boxes' of BASELINE configs)."""

    def __init__(self, boxes_by_name: Dict[str, np.ndarray]):
        self.boxes_by_name = boxes_by_name
        self.current_name: Optional[str] = None

    def __call__(self, image: np.ndarray) -> np.ndarray:
        return np.asarray(self.boxes_by_name.get(self.current_name, np.zeros((0, 4))), np.float32).reshape(-1, 4)


def _load_yolo(yolo_model_path) -> Detector:
    try:
        from ultralytics import YOLO          # reference path, pipeline.py:72-73
    except ImportError as e:
        raise RuntimeError("ultralytics is not installed: pass detector=<callable image->boxes> "
                           "(the YOLO detector is outside the accelerated path)") from e
    model = YOLO(yolo_model_path)
    model.args["verbose"] = False

    def detect(image: np.ndarray) -> np.ndarray:       # pipeline.py:84-87
        return model(image)[0].boxes.xyxy.cpu().numpy()
    return detect


def _load_sam_state_dict(sam_model_type: str):
    """pipeline.py:76 loads SamModel.from_pretrained(hub id); offline this only works from a local cache."""
    from transformers import SamModel
    return SamModel.from_pretrained(sam_model_type, local_files_only=True).state_dict()


class CellSegmentationPipeline:
    def __init__(self, yolo_model_path: Union[str, Path, None], sam_model_type: str = "facebook/sam-vit-huge",
                 device: str = "cuda", *, detector: Optional[Detector] = None,
                 sam_state_dict: Optional[Dict[str, Any]] = None, max_boxes: int = 64,
                 max_image_hw: Tuple[int, int] = (1024, 1024), on_empty: str = "raise"):
        self.device = device
        self.sam_model_type = sam_model_type
        self.detector: Detector = detector if detector is not None else _load_yolo(yolo_model_path)
        sd = sam_state_dict if sam_state_dict is not None else _load_sam_state_dict(sam_model_type)
        self.sam_stage = SamStage(sam_model_type, device=device, state_dict=sd, max_batch=1, max_boxes=max_boxes,
                                  max_image_hw=max_image_hw, on_empty=on_empty)
        self.run_id = self._generate_run_id()

    @staticmethod
    def _generate_run_id() -> str:
        return f"{datetime.now().strftime('%Y%m%d_%H%M%S')}_{uuid.uuid4().hex[:8]}"

    def _detect_cells(self, image: np.ndarray) -> np.ndarray:
        return np.asarray(self.detector(image), np.float32).reshape(-1, 4)

    def process_single_image(self, image_path: Union[str, Path], output_path: Union[str, Path],
                             save_visualizations: bool = True) -> ProcessingResult:
        timings: Dict[str, float] = {}
        start_time = time.time()
        image = self._load_image(str(image_path))
        timings["image_load"] = time.time() - start_time

        start_time = time.time()
        if isinstance(self.detector, BoxTable):
            self.detector.current_name = Path(image_path).name
        boxes = self._detect_cells(image)
        timings["yolo_detection"] = time.time() - start_time

        masks: Sequence[np.ndarray] = []
        cell_metrics: List[Dict[str, Any]] = []
        sam_times = {"inference": 0.0, "postprocess": 0.0}
        if len(boxes) > 0:
            # pipeline.py:161-175 -> one call; the stage reports honest per-phase device times
            start_time = time.time()
            masks, cell_metrics, _crops = self.sam_stage.run(image, boxes)
            t = self.sam_stage.last_timing
            timings["sam_preprocess"] = (t["h2d_ms"] + t["preprocess_ms"]) / 1e3
            sam_times["inference"] = (t["encoder_ms"] + t["decoder_ms"]) / 1e3
            sam_times["postprocess"] = (t["postprocess_ms"] + t["metrics_ms"] + t["d2h_ms"]) / 1e3
        else:
            timings["sam_preprocess"] = 0.0
        timings.update(sam_times)
        if save_visualizations:
            start_time = time.time()
            self._save_visualizations(image, masks, boxes, cell_metrics, output_path)
            timings["visualization"] = time.time() - start_time
        total_time = time.time() - start_time
        timings.update({"total_time": total_time, "cells_processed": len(boxes)})
        return ProcessingResult(image_path=str(image_path), cell_metrics=cell_metrics,
                                num_cells=len(cell_metrics), timing=timings)

    @staticmethod
    def _load_image(image_path: str) -> np.ndarray:
        import cv2                                   # pipeline.py:206-210
        image = cv2.imread(image_path)
        return cv2.cvtColor(image, cv2.COLOR_BGR2RGB)

    def _save_visualizations(self, image, masks, boxes, cell_metrics, output_path) -> None:
        """Visualisation / TIFF output (pipeline.py:331-438) is outside the accelerated path; nothing is written."""
        logger.debug("save_visualizations requested for %s: not part of the SAM stage", output_path)

    def process_directory(self, input_dir: Union[str, Path], output_dir: Union[str, Path],
                          save_visualizations: bool = True, pbar=None) -> BatchProcessingResult:
        input_dir = Path(input_dir)
        output_dir = Path(output_dir) / self.run_id
        output_dir.mkdir(parents=True, exist_ok=True)
        image_files = self._get_image_files(input_dir)
        results, metrics_data, timing_data = [], [], []
        total_timing = self._initialize_timing_dict()
        for image_path in image_files:
            result = self.process_single_image(image_path, output_dir / image_path.name, save_visualizations)
            results.append(result)
            self._update_progress(pbar, result)
            self._collect_metrics_data(metrics_data, result)
            self._collect_timing_data(timing_data, result)
            self._update_total_timing(total_timing, result.timing)
        return BatchProcessingResult(results=results, total_timing=total_timing, metrics_data=metrics_data,
                                     timing_data=timing_data)

    @staticmethod
    def _get_image_files(directory: Path) -> List[Path]:
        # same globs as pipeline.py:265-269, but sorted: the reference's order is filesystem dependent
        return sorted(list(directory.glob("*.png")) + list(directory.glob("*.jpg")) + list(directory.glob("*.tiff")))

    @staticmethod
    def _initialize_timing_dict() -> Dict[str, float]:
        return {"image_load": 0, "yolo_detection": 0, "sam_preprocess": 0, "sam_inference_total": 0,
                "sam_postprocess_total": 0, "metrics_total": 0, "visualization": 0, "total_time": 0, "total_cells": 0}

    @staticmethod
    def _update_progress(pbar, result: ProcessingResult) -> None:
        if pbar is not None:
            pbar.update(1)
            pbar.set_postfix({"cells": result.num_cells}, refresh=True)

    @staticmethod
    def _collect_metrics_data(metrics_data: List[Dict[str, Any]], result: ProcessingResult) -> None:
        for cell_idx, metrics in enumerate(result.cell_metrics):
            metrics_data.append({"image_name": Path(result.image_path).name, "cell_id": cell_idx, **metrics})

    @staticmethod
    def _collect_timing_data(timing_data: List[Dict[str, Any]], result: ProcessingResult) -> None:
        timing_data.append({"image_name": Path(result.image_path).name,
                            "cells_processed": result.timing["cells_processed"],
                            **{f"{k}_ms": v * 1000 for k, v in result.timing.items() if k != "cells_processed"}})

    @staticmethod
    def _update_total_timing(total_timing: Dict[str, float], timing: Dict[str, float]) -> None:
        for key in total_timing:
            if key == "total_cells":
                total_timing[key] += timing["cells_processed"]
            elif key in timing:
                total_timing[key] += timing[key]


def _worker_process_directory(rank: int, device: str, files: List[str], output_dir: str, save_visualizations: bool,
                              ctor_kwargs: Dict[str, Any], queue) -> None:
    try:
        pipe = CellSegmentationPipeline(**ctor_kwargs, device=device)
        out = [pipe.process_single_image(f, Path(output_dir) / Path(f).name, save_visualizations) for f in files]
        queue.put((rank, out, None))
    except Exception as e:   # noqa: BLE001 - forwarded to the parent
        queue.put((rank, None, repr(e)))


class ParallelCellSegmentationPipeline:
    """pipeline.py:440-584 with replicas spread over GPUs instead of stacked on one device.

    ``num_pipelines`` workers, worker k on ``cuda:k % n_gpus`` (one process per GPU, no NCCL: the path
    has no cross-image reduction).  Files are split into contiguous chunks of ceil(n / num_pipelines)
    (:540-541) and results are concatenated in chunk order (:569-577)."""

    def __init__(self, yolo_model_path, sam_model_type: str = "facebook/sam-vit-huge", device: str = "cuda",
                 num_pipelines: int = 2, **kwargs):
        self.device = device
        self.sam_model_type = sam_model_type
        self.num_pipelines = num_pipelines
        self._ctor_kwargs = dict(yolo_model_path=yolo_model_path, sam_model_type=sam_model_type, **kwargs)
        self.run_id = CellSegmentationPipeline._generate_run_id()

    _get_image_files = staticmethod(CellSegmentationPipeline._get_image_files)
    _initialize_timing_dict = staticmethod(CellSegmentationPipeline._initialize_timing_dict)

    def process_directory(self, input_dir, output_dir, save_visualizations: bool = True, pbar=None) -> BatchProcessingResult:
        import multiprocessing as mp
        input_dir = Path(input_dir)
        output_dir = Path(output_dir) / self.run_id
        output_dir.mkdir(parents=True, exist_ok=True)
        files = [str(p) for p in self._get_image_files(input_dir)]
        chunks = partition_contiguous(files, self.num_pipelines)
        n_gpus = _gpu_count()
        ctx = mp.get_context("spawn")
        queue = ctx.Queue()
        procs = []
        for k, chunk in enumerate(chunks):
            p = ctx.Process(target=_worker_process_directory,
                            args=(k, f"cuda:{k % max(n_gpus, 1)}", chunk, str(output_dir), save_visualizations,
                                  self._ctor_kwargs, queue))
            p.start()
            procs.append(p)
        got: Dict[int, List[ProcessingResult]] = {}
        for _ in procs:
            rank, out, err = queue.get()
            if err is not None:
                for p in procs:
                    p.terminate()
                raise RuntimeError(f"worker {rank} failed: {err}")
            got[rank] = out
        for p in procs:
            p.join()
        results, metrics_data, timing_data = [], [], []
        total_timing = self._initialize_timing_dict()
        for k in range(len(chunks)):
            for result in got[k]:
                results.append(result)
                if pbar:
                    pbar.update(1)
                CellSegmentationPipeline._collect_metrics_data(metrics_data, result)
                CellSegmentationPipeline._collect_timing_data(timing_data, result)
                CellSegmentationPipeline._update_total_timing(total_timing, result.timing)
        return BatchProcessingResult(results=results, total_timing=total_timing, metrics_data=metrics_data,
                                     timing_data=timing_data)


def _gpu_count() -> int:
    import ctypes
    try:
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import torch
            return torch.cuda.device_count()
        except Exception:
            return 1
    n = ctypes.c_int(0)
    return n.value if cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0 else 1
