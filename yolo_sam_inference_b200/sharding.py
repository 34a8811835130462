"""Folder partition of the image list across workers / GPUs (no collective: images are independent).

Same rule as ParallelCellSegmentationPipeline.process_directory
(/root/reference/src/yolo_sam_inference/pipeline.py:540-541): contiguous chunks of ceil(n / workers);
results are concatenated in chunk order (:569-577) so the output order equals the (sorted) input order.
"""
from __future__ import annotations

import math
from typing import List, Sequence, TypeVar

T = TypeVar("T")


def partition_contiguous(items: Sequence[T], workers: int) -> List[List[T]]:
    """image_batches = [files[i:i+batch] for i in range(0, n, batch)], batch = ceil(n / workers).
    Like the reference, this can yield fewer than `workers` chunks (and none for an empty list)."""
    n = len(items)
    if n == 0 or workers <= 0:
        return []
    batch = math.ceil(n / workers)
    return [list(items[i:i + batch]) for i in range(0, n, batch)]


def shard_range(n: int, rank: int, world: int) -> range:
    """Index range of rank `rank` under the same contiguous rule (empty if the chunks run out)."""
    if n == 0:
        return range(0)
    batch = math.ceil(n / world)
    lo = min(rank * batch, n)
    return range(lo, min(lo + batch, n))


def gather_in_order(per_rank_results: Sequence[Sequence[T]]) -> List[T]:
    out: List[T] = []
    for r in per_rank_results:
        out.extend(r)
    return out
