"""B200-native SAM stage of the yolo-sam-inference pipeline (image + box prompts -> masks, crops, metrics).

Public surface mirrors the reference package (src/yolo_sam_inference/__init__.py:1-21):
CellSegmentationPipeline / ParallelCellSegmentationPipeline keep their entry points; SamStage is the
seam that replaces pipeline.py:161-175 and runs on libysi.so (hand-written sm_100a CUDA, C ABI in
include/ysi.h).
"""
from .sam_stage import SamStage, metrics_from_raw  # noqa: F401
from .pipeline import (BatchProcessingResult, BoxTable, CellSegmentationPipeline,  # noqa: F401
                       ParallelCellSegmentationPipeline, ProcessingResult)

__all__ = ["SamStage", "metrics_from_raw", "CellSegmentationPipeline", "ParallelCellSegmentationPipeline",
           "ProcessingResult", "BatchProcessingResult", "BoxTable"]
