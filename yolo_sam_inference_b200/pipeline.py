"""Drop-in mirror of the reference pipeline entry points, with the SAM stage on libysi.so.

Mirrors /root/reference/src/yolo_sam_inference/pipeline.py: same class names, constructor arguments,
``process_single_image`` / ``process_directory`` signatures, ``ProcessingResult`` /
``BatchProcessingResult`` dataclasses (:31-45), timing keys (:143-194, :271-283), CSV row builders
(:293-317) and visualisation file layout (:331-438).  What changes is the body of ``if len(boxes) > 0:``
(:161-175), which becomes SAM-stage calls on the device, and HOW a folder is driven through it:

* ``process_directory`` (:212-263) no longer handles one image at a time.  A thread pool reads files ahead
  (baseline TIFFs straight into page-locked memory as raw samples, everything else through the
  reference's own ``cv2.imread`` path -- ``ingest.py``), consecutive same-sized images are grouped into
  batches of ``batch_size`` and the batches flow through the two-slot pipeline of ``SamStage.run_stream``
  (copies, encoder and decoder/metrics of neighbouring batches overlap).  Per-image ``ProcessingResult``s,
  their order and the CSV rows are unchanged.
* ``ParallelCellSegmentationPipeline`` (:440-584) spreads its replicas over GPUs (one persistent worker
  process per replica, worker k on ``cuda:k % n_gpus``) instead of stacking them on one device; the folder
  is split by the reference's ``ceil(n / num_pipelines)`` rule (:540-541), no collective is involved.

The YOLO detector stays on the reference's torch path (ultralytics, :72-73, :84-87); because
ultralytics is not installable offline, any callable ``image -> float32[N,4] xyxy`` can be injected as
``detector`` instead.
"""
from __future__ import annotations

import logging
import os
import threading
import time
import uuid
from collections import deque
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from datetime import datetime
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _native as nat
from .ingest import as_rgb_u8, decode_rgb, probe_tiff, read_raw_into
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
    needs_pixels = False          # the pipeline may hand it the raw samples instead of building an RGB image

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


class _Prefetcher:
    """Reads the files of a folder ahead of the GPU, in order, on a thread pool (bounded window).

    Baseline TIFFs land as raw samples in recycled page-locked buffers (``readinto``: disk -> pinned memory, no copy in
    between, GIL released); every other file goes through ``decode_rgb`` (cv2.imread + BGR->RGB, pipeline.py:206-210).
    ``iterate(files)`` yields (index, path, image, load_seconds, token); ``release(token)`` returns a pinned buffer to the
    pool once the device has consumed it.  Threads and pinned buffers live as long as the pipeline (page-locking memory
    costs milliseconds per buffer: it must not be paid per folder)."""

    def __init__(self, workers: int, depth: int, raw_ingest: bool, precision: str, device: int):
        self.depth = max(1, depth)
        self.raw_ingest = raw_ingest
        self.precision, self.device = precision, device
        self.pool = ThreadPoolExecutor(max_workers=max(1, workers), thread_name_prefix="ysi-load")
        self.free: Dict[int, List[nat.PinnedBuffer]] = {}
        self.all: List[nat.PinnedBuffer] = []
        self.lock = threading.Lock()

    def _acquire(self, nbytes: int) -> nat.PinnedBuffer:
        with self.lock:
            lst = self.free.get(nbytes)
            if lst:
                return lst.pop()
        buf = nat.PinnedBuffer(nbytes, self.precision, self.device)
        with self.lock:
            self.all.append(buf)
        return buf

    def release(self, token) -> None:
        if token is not None:
            with self.lock:
                self.free.setdefault(token.nbytes, []).append(token)

    def _load(self, path: str):
        t0 = time.time()
        info = probe_tiff(path) if self.raw_ingest and path.lower().endswith((".tif", ".tiff")) else None
        if info is not None:
            buf = self._acquire(info.nbytes)
            read_raw_into(info, buf.array)
            image = buf.array.view(info.dtype).reshape(info.shape)
            return image, time.time() - t0, buf
        return decode_rgb(path), time.time() - t0, None

    def iterate(self, files: Sequence[Union[str, Path]]):
        files = [str(f) for f in files]
        window: deque = deque()
        nxt = 0
        n = len(files)
        try:
            while nxt < n or window:
                while nxt < n and len(window) < self.depth:
                    window.append((nxt, self.pool.submit(self._load, files[nxt])))
                    nxt += 1
                idx, fut = window.popleft()
                image, t_load, token = fut.result()
                yield idx, files[idx], image, t_load, token
        finally:
            for _, fut in window:             # abandoned mid-folder: let the reads finish, hand their buffers back
                try:
                    self.release(fut.result()[2])
                except Exception:
                    pass

    def close(self) -> None:
        self.pool.shutdown(wait=True, cancel_futures=True)
        for b in self.all:
            b.close()
        self.all, self.free = [], {}


class CellSegmentationPipeline:
    def __init__(self, yolo_model_path: Union[str, Path, None], sam_model_type: str = "facebook/sam-vit-huge",
                 device: str = "cuda", *, detector: Optional[Detector] = None,
                 sam_state_dict: Optional[Dict[str, Any]] = None, max_boxes: int = 64,
                 max_image_hw: Tuple[int, int] = (1024, 1024), on_empty: str = "raise", batch_size: int = 8,
                 decode_workers: Optional[int] = None, raw_ingest: bool = True, mask_output: Optional[str] = None,
                 mask_sink: Optional[Callable[[str, np.ndarray], None]] = None, precision: Optional[str] = None):
        """Beyond the reference's three arguments (keyword-only, all optional):
        ``batch_size`` images per device launch in ``process_directory``; ``max_boxes`` / ``max_image_hw`` are capacity
        hints, not limits (more boxes are decoded in chunks, larger images re-size the buffers); ``decode_workers`` threads
        read files ahead; ``raw_ingest`` hands baseline TIFFs to the device undecoded; ``mask_output`` ("packed" | "bool" |
        None) brings every image's masks to the host (None: masks stay on the device unless visualisations are written) and
        hands them to ``mask_sink(image_name, masks)`` -- a view valid during the call, e.g. for utils/mask_encoding -- or,
        without a sink, keeps copies in ``last_masks``."""
        self.device = device
        self.sam_model_type = sam_model_type
        self.detector: Detector = detector if detector is not None else _load_yolo(yolo_model_path)
        sd = sam_state_dict if sam_state_dict is not None else _load_sam_state_dict(sam_model_type)
        self.batch_size = max(1, int(batch_size))
        self.sam_stage = SamStage(sam_model_type, device=device, state_dict=sd, max_batch=self.batch_size,
                                  max_boxes=max_boxes, max_image_hw=max_image_hw, on_empty=on_empty, precision=precision)
        self.decode_workers = decode_workers if decode_workers else min(16, max(2, (os.cpu_count() or 4) // 2))
        self.raw_ingest = raw_ingest
        self.mask_output = mask_output
        self.mask_sink = mask_sink
        self.last_masks: Dict[str, np.ndarray] = {}
        self._prefetch: Optional[_Prefetcher] = None
        self.run_id = self._generate_run_id()

    def close(self) -> None:
        self.sam_stage.close()                 # drains the device before the pinned staging buffers are released
        if self._prefetch is not None:
            self._prefetch.close()
            self._prefetch = None

    @staticmethod
    def _generate_run_id() -> str:
        return f"{datetime.now().strftime('%Y%m%d_%H%M%S')}_{uuid.uuid4().hex[:8]}"

    def _detect_cells(self, image: np.ndarray) -> np.ndarray:
        return np.asarray(self.detector(image), np.float32).reshape(-1, 4)

    def _detect(self, image: np.ndarray, image_path: str) -> np.ndarray:
        if isinstance(self.detector, BoxTable):
            self.detector.current_name = Path(image_path).name
        if image.ndim == 2 and getattr(self.detector, "needs_pixels", True):
            image = as_rgb_u8(image)        # a real detector sees exactly what _load_image would have produced
        return self._detect_cells(image)

    def process_single_image(self, image_path: Union[str, Path], output_path: Union[str, Path],
                             save_visualizations: bool = True) -> ProcessingResult:
        timings: Dict[str, float] = {}
        start_time = time.time()
        image = self._load_image(str(image_path))
        timings["image_load"] = time.time() - start_time

        start_time = time.time()
        boxes = self._detect(image, str(image_path))
        timings["yolo_detection"] = time.time() - start_time

        masks: Sequence[np.ndarray] = []
        cell_metrics: List[Dict[str, Any]] = []
        sam_times = {"inference": 0.0, "postprocess": 0.0}
        if len(boxes) > 0:
            # pipeline.py:161-175 -> one call; the stage reports honest per-phase device times
            start_time = time.time()
            mode = "bool" if save_visualizations else self.mask_output
            masks, cell_metrics, _crops = self.sam_stage.run(image, boxes, masks=mode)
            if self.mask_output:
                if self.mask_sink is not None:
                    self.mask_sink(Path(image_path).name, masks)
                else:
                    self.last_masks[Path(image_path).name] = masks
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
        return decode_rgb(image_path)                # pipeline.py:206-210

    # ------------------------------------------------------------------ visualisation / TIFF writers (pipeline.py:331-438)
    @staticmethod
    def _write_tiff(path: Path, array: np.ndarray) -> None:
        """utils/image_utils.save_optimized_tiff / save_mask_as_tiff (:8-102): uint8 TIFF with deflate compression. The
        reference writes through tifffile (absent here); OpenCV's libtiff writer produces the same pixels."""
        import cv2
        if array.dtype == np.bool_:
            array = array.astype(np.uint8) * 255
        if array.ndim == 3:
            array = cv2.cvtColor(array, cv2.COLOR_RGB2BGR)
        if not cv2.imwrite(str(path), array, [cv2.IMWRITE_TIFF_COMPRESSION, 8]):      # 8 = Adobe deflate (zlib)
            raise IOError(f"Failed to save TIFF file: {path}")

    def _save_visualizations(self, image, masks, boxes, cell_metrics, output_path) -> None:
        """Same folders and file names as the reference (pipeline.py:352-432): 1_original_images/<stem>_original.tiff (the
        file the CSV consumers crop from), 2_yolo_detections, 3_processed_masks/{masks,overlay_images,convex_hull_overlay},
        4_combined_visualization.  Errors are reported and swallowed like the reference does (:436-438)."""
        try:
            import cv2
            image = as_rgb_u8(image)
            output_path = Path(output_path)
            base_dir = output_path.parent
            dirs = {"original": base_dir / "1_original_images", "yolo": base_dir / "2_yolo_detections",
                    "processed_masks": base_dir / "3_processed_masks/masks",
                    "processed_overlays": base_dir / "3_processed_masks/overlay_images",
                    "convex_hull": base_dir / "3_processed_masks/convex_hull_overlay",
                    "combined": base_dir / "4_combined_visualization"}
            for d in dirs.values():
                d.mkdir(parents=True, exist_ok=True)
            stem = output_path.stem
            self._write_tiff(dirs["original"] / f"{stem}_original.tiff", image)
            yolo_vis = image.copy()
            for box in boxes:
                x1, y1, x2, y2 = np.asarray(box).astype(int)
                cv2.rectangle(yolo_vis, (int(x1), int(y1)), (int(x2), int(y2)), (255, 0, 0), 2)
            self._write_tiff(dirs["yolo"] / f"{stem}_yolo.tiff", yolo_vis)
            red = np.array([255, 0, 0])
            for i, (mask, metrics) in enumerate(zip(masks, cell_metrics)):
                mask = np.asarray(mask, bool)
                self._write_tiff(dirs["processed_masks"] / f"{stem}_mask_{i}.tiff", mask)
                overlay = image.copy()
                overlay[mask] = overlay[mask] * 0.7 + red * 0.3
                self._write_tiff(dirs["processed_overlays"] / f"{stem}_mask_{i}_overlay.tiff", overlay)
                # metrics.py:102-119 never returns 'convex_hull_coords', so the reference's hull overlay is the plain image
                self._write_tiff(dirs["convex_hull"] / f"{stem}_mask_{i}_convex_hull.tiff", image)
            combined = np.zeros((image.shape[0], image.shape[1] * 2, 3), dtype=np.uint8)
            combined[:, :image.shape[1]] = yolo_vis
            overlay_vis = image.copy()
            for mask in masks:
                mask = np.asarray(mask, bool)
                overlay_vis[mask] = overlay_vis[mask] * 0.8 + red * 0.2
            combined[:, image.shape[1]:] = overlay_vis
            self._write_tiff(dirs["combined"] / f"{stem}_combined.tiff", combined)
        except Exception as e:                         # noqa: BLE001 - pipeline.py:436-438
            print(f"Warning: Error during visualization saving: {str(e)}")

    # ------------------------------------------------------------------ folder entry point
    def process_files(self, image_files: Sequence[Union[str, Path]], output_dir: Union[str, Path],
                      save_visualizations: bool = True, pbar=None) -> List[ProcessingResult]:
        """The batched, pipelined body of ``process_directory``: results come back in the order of ``image_files``."""
        files = [Path(f) for f in image_files]
        output_dir = Path(output_dir)
        n = len(files)
        results: List[Optional[ProcessingResult]] = [None] * n
        if n == 0:
            return []
        stage = self.sam_stage
        mode = "bool" if save_visualizations else self.mask_output
        if self._prefetch is None:
            self._prefetch = _Prefetcher(self.decode_workers, depth=4 * self.batch_size, raw_ingest=self.raw_ingest,
                                         precision=stage.precision, device=stage.device_index)
        pre = self._prefetch
        metas: deque = deque()        # bookkeeping of the batches handed to run_stream, in submission order

        def finish_image(idx, image, boxes, timing, masks, cell_metrics):
            t_vis0 = time.time()
            if save_visualizations:
                self._save_visualizations(image, masks if masks is not None else [], boxes, cell_metrics,
                                          output_dir / files[idx].name)
                timing["visualization"] = time.time() - t_vis0
            if self.mask_output and masks is not None and len(boxes) > 0:
                if self.mask_sink is not None:
                    self.mask_sink(files[idx].name, masks)
                else:
                    self.last_masks[files[idx].name] = np.array(masks, copy=True)
            timing["total_time"] = sum(v for k, v in timing.items() if k != "cells_processed")
            timing["cells_processed"] = len(boxes)
            results[idx] = ProcessingResult(image_path=str(files[idx]), cell_metrics=cell_metrics,
                                            num_cells=len(cell_metrics), timing=timing)
            self._update_progress(pbar, results[idx])

        def batches():
            cur: List[Any] = []
            key = None
            for idx, path, image, t_load, token in pre.iterate(files):
                t0 = time.time()
                boxes = self._detect(image, path)
                timing = {"image_load": t_load, "yolo_detection": time.time() - t0, "sam_preprocess": 0.0,
                          "inference": 0.0, "postprocess": 0.0}
                if len(boxes) == 0:                     # pipeline.py:176-179: SAM is skipped
                    finish_image(idx, image, boxes, timing, None, [])
                    pre.release(token)
                    continue
                k = (image.shape, image.dtype)
                if cur and (k != key or len(cur) == self.batch_size):
                    metas.append(cur)
                    yield [c[1] for c in cur], [c[2] for c in cur]
                    cur = []
                key = k
                cur.append((idx, image, boxes, timing, token))
            if cur:
                metas.append(cur)
                yield [c[1] for c in cur], [c[2] for c in cur]

        try:
            for out in stage.run_stream(batches(), masks=mode, copy_masks=False, crops=False):
                meta = metas.popleft()
                t = stage.last_timing
                share = 1.0 / len(meta)
                for (idx, image, boxes, timing, token), (masks, cell_metrics, _crops) in zip(meta, out):
                    timing["sam_preprocess"] = (t["h2d_ms"] + t["preprocess_ms"]) / 1e3 * share
                    timing["inference"] = (t["encoder_ms"] + t["decoder_ms"]) / 1e3 * share
                    timing["postprocess"] = (t["postprocess_ms"] + t["metrics_ms"] + t["d2h_ms"]) / 1e3 * share
                    finish_image(idx, image, boxes, timing, masks, cell_metrics)
                    pre.release(token)
        finally:
            stage.sync()
        return [r for r in results if r is not None]

    def process_directory(self, input_dir: Union[str, Path], output_dir: Union[str, Path],
                          save_visualizations: bool = True, pbar=None) -> BatchProcessingResult:
        input_dir = Path(input_dir)
        output_dir = Path(output_dir) / self.run_id
        output_dir.mkdir(parents=True, exist_ok=True)
        image_files = self._get_image_files(input_dir)
        results = self.process_files(image_files, output_dir, save_visualizations, pbar)
        return self._assemble(results)

    @classmethod
    def _assemble(cls, results: List[ProcessingResult]) -> BatchProcessingResult:
        metrics_data: List[Dict[str, Any]] = []
        timing_data: List[Dict[str, Any]] = []
        total_timing = cls._initialize_timing_dict()
        for result in results:
            cls._collect_metrics_data(metrics_data, result)
            cls._collect_timing_data(timing_data, result)
            cls._update_total_timing(total_timing, result.timing)
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


# ---------------------------------------------------------------------------------------------------------------------
# multi-GPU folder partition
# ---------------------------------------------------------------------------------------------------------------------
def _worker_main(rank: int, device: str, ctor_kwargs: Dict[str, Any], conn) -> None:
    """One persistent replica: builds its pipeline (context + weights) once, then serves requests until told to stop."""
    try:
        pipe = CellSegmentationPipeline(**ctor_kwargs, device=device)
        conn.send(("ready", rank, None))
    except Exception as e:   # noqa: BLE001 - forwarded to the parent
        conn.send(("error", rank, repr(e)))
        return
    while True:
        try:
            msg = conn.recv()
        except EOFError:
            break
        if msg[0] == "stop":
            break
        try:
            if msg[0] == "files":
                _, files, output_dir, save_visualizations = msg
                conn.send(("ok", rank, pipe.process_files(files, output_dir, save_visualizations)))
            elif msg[0] == "image":
                image = msg[1]
                boxes = pipe._detect_cells(image)
                masks = pipe.sam_stage.run(image, boxes)[0] if len(boxes) else np.zeros((0,) + image.shape[:2], bool)
                conn.send(("ok", rank, (boxes, masks)))
            else:
                conn.send(("error", rank, f"unknown request {msg[0]!r}"))
        except Exception as e:   # noqa: BLE001
            conn.send(("error", rank, repr(e)))
    pipe.close()


class ParallelCellSegmentationPipeline:
    """pipeline.py:440-584 with replicas spread over GPUs instead of stacked on one device.

    ``num_pipelines`` persistent worker processes (spawned on first use, kept until ``close()``), worker k on
    ``cuda:k % n_gpus`` (one process / one ysi_ctx per replica, no NCCL: the path has no cross-image reduction).  Files are
    split into contiguous chunks of ceil(n / num_pipelines) (:540-541), every worker runs the batched, pipelined
    ``CellSegmentationPipeline.process_files`` on its chunk, and results are concatenated in chunk order (:569-577)."""

    POLL_S = 0.2

    def __init__(self, yolo_model_path, sam_model_type: str = "facebook/sam-vit-huge", device: str = "cuda",
                 num_pipelines: int = 2, **kwargs):
        self.device = device
        self.sam_model_type = sam_model_type
        self.num_pipelines = num_pipelines
        self._ctor_kwargs = dict(yolo_model_path=yolo_model_path, sam_model_type=sam_model_type, **kwargs)
        self._workers: List[Tuple[Any, Any]] = []      # (process, parent end of the pipe)
        self.run_id = CellSegmentationPipeline._generate_run_id()

    _get_image_files = staticmethod(CellSegmentationPipeline._get_image_files)
    _initialize_timing_dict = staticmethod(CellSegmentationPipeline._initialize_timing_dict)
    _generate_run_id = staticmethod(CellSegmentationPipeline._generate_run_id)

    # -- worker management
    def _recv(self, k: int, what: str):
        """Next message of worker k; raises if the worker died (e.g. a native crash) instead of waiting forever."""
        proc, conn = self._workers[k]
        try:
            while not conn.poll(self.POLL_S):
                if not proc.is_alive():
                    raise EOFError
            kind, rank, payload = conn.recv()
        except (EOFError, OSError):
            proc.join(timeout=1)
            code = proc.exitcode
            self.close()
            raise RuntimeError(f"worker {k} died (exit code {code}) while {what}") from None
        if kind == "error":
            self.close()
            raise RuntimeError(f"worker {rank} failed while {what}: {payload}")
        return payload

    def _send(self, k: int, msg) -> None:
        proc, conn = self._workers[k]
        try:
            if not proc.is_alive():
                raise OSError
            conn.send(msg)
        except OSError:
            proc.join(timeout=1)
            code = proc.exitcode
            self.close()
            raise RuntimeError(f"worker {k} died (exit code {code}) before it could be given work") from None

    def _ensure_workers(self) -> None:
        if self._workers:
            return
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        n_gpus = _gpu_count()
        base = self.device.split(":")[0]
        for k in range(self.num_pipelines):
            parent, child = ctx.Pipe()
            dev = f"{base}:{k % max(n_gpus, 1)}"
            p = ctx.Process(target=_worker_main, args=(k, dev, self._ctor_kwargs, child), daemon=True)
            p.start()
            child.close()
            self._workers.append((p, parent))
        for k in range(self.num_pipelines):
            self._recv(k, "starting up")

    def close(self) -> None:
        for p, conn in self._workers:
            try:
                conn.send(("stop",))
            except Exception:
                pass
        for p, conn in self._workers:
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()
            conn.close()
        self._workers = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the reference's public methods
    def process_image(self, image: np.ndarray) -> Tuple[np.ndarray, List[np.ndarray], List[float]]:
        """pipeline.py:469-503: one RGB array through detector + SAM on the first replica -> (boxes, masks, scores)."""
        self._ensure_workers()
        self._send(0, ("image", np.ascontiguousarray(image)))
        boxes, masks = self._recv(0, "processing an image")
        return boxes, [m for m in masks], [1.0] * len(boxes)

    def process_directory(self, input_dir, output_dir, save_visualizations: bool = True, pbar=None) -> BatchProcessingResult:
        input_dir = Path(input_dir)
        output_dir = Path(output_dir) / self.run_id
        output_dir.mkdir(parents=True, exist_ok=True)
        files = [str(p) for p in self._get_image_files(input_dir)]
        chunks = partition_contiguous(files, self.num_pipelines)
        self._ensure_workers()
        for k, chunk in enumerate(chunks):
            self._send(k, ("files", chunk, str(output_dir), save_visualizations))
        results: List[ProcessingResult] = []
        for k in range(len(chunks)):                   # chunk order == input order (pipeline.py:569-577)
            out = self._recv(k, "processing its chunk")
            results.extend(out)
            if pbar:
                pbar.update(len(out))
        return CellSegmentationPipeline._assemble(results)


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
