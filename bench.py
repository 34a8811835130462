#!/usr/bin/env python
"""bench.py -- SAM box-prompt stage throughput on B200 (BASELINE.json metric), plus the reference CPU arm.

    python bench.py --gpus N --steps K --warmup W [--workload b1|b32|vit_h|folder5]   # this repo (libysi, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...                            # the reference's CPU path (oracle)

Default workload ``b1`` = BASELINE.json configs[1]: SAM ViT-B, 256 synthetic 1024x1024 images, 1 box each.
A *step* is one pass of the hot path (ingest -> preprocess -> ViT encoder -> prompt encoder / mask decoder ->
upsample + threshold -> morphometrics) over the 256 images of the config (32 device batches of 8), so the default
``--steps 20`` times 5120 images (~7 s): the GPU is in its sustained, power-capped regime.  Rank r of N processes its own
contiguous shard of the image list (folder partition, no data-path collective; per-GPU work is fixed => weak scaling).

  value  images/s with the inputs resident in HBM before the timed region (every step enqueued back to back on the
         context's streams, CUDA events on them, max over ranks)
  e2e    images/s through the reference-facing entry point: CellSegmentationPipeline.process_directory on a FOLDER of TIFF
         files (read-ahead thread pool -> pinned memory -> H2D -> device path -> D2H of packed masks + metric rows ->
         16-key metric dicts, ProcessingResult per image), one call per step, wall clock, max over ranks
  roofline      tensor-pipe roofline of the dominant kernel class (the tcgen05 GEMM behind every ViT linear), from
                per-launch CUDA events in a separate profiled pass over the same batches
  roofline_hbm  HBM roofline of upsample + threshold + morphometrics (a6 + a7) at 32 boxes per image (256 masks per launch)
  extra         configs[3] (32 boxes / image) masks/s, configs[2]'s model (ViT-H) images/s, configs[4]-style folder of
                2048x2048 16-bit TIFFs through process_directory, and the bf16-operand build on configs[1]
  cpu_baseline  the reference path (transformers SamModel fp32 + restated metrics) on the host cores, bounded sample,
                rank 0 / N=1 only
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = int(os.environ.get("YSI_BENCH_BATCH", "8"))     # images per device launch
PRECISION = os.environ.get("YSI_PRECISION", "fp16")     # 16-bit operand encoding of the tensor-core contractions (fp16 | bf16)
STEP_IMAGES = 256                                        # BASELINE configs[1]: 256 synthetic 1024x1024 images = one step
ENC_FLOPS = {"vit_b": 937.6e9, "vit_l": 2837.0e9, "vit_h": 5641.8e9}   # algorithmic FLOPs / image
DEC_FLOPS_BOX = 3.61e9
MODEL_NAME = {"vit_b": "ViT-B", "vit_l": "ViT-L", "vit_h": "ViT-H"}
WORKLOADS = {
    # name: (model, boxes/image, image size, images per step, description)
    "b1": ("vit_b", 1, 1024, STEP_IMAGES, "configs[1]: SAM ViT-B, 256 synthetic 1024x1024 images, 1 box/image"),
    "b32": ("vit_b", 32, 1024, 64, "configs[3]: SAM ViT-B, 32 boxes/image (decoder + morphometrics dominated)"),
    "vit_h": ("vit_h", 1, 1024, 64, "configs[2]'s model: SAM ViT-H, synthetic 1024x1024 images, 1 box/image"),
    "folder5": ("vit_b", 1, 2048, 64, "configs[4]: folder of synthetic 2048x2048 16-bit TIFFs through process_directory"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "hbm_gbs": p["hbm_gbs"], "src": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
                "src": "fallback (B200_PROFILING.md)"}


def ncu_traffic(key: str):
    """DRAM bytes per launch of a kernel class from the committed ncu summary (profiles/ncu_traffic.json, written by
    scripts/ncu_traffic.py from a `ncu --set full` capture); None when no capture of this round exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            d = json.load(f)
        return d.get(key)
    except Exception:
        return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def init_dist(world: int, backend: str):
    if world <= 1:
        return None
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to the process's stdout (fd 1) when the first communicator is created; rank 0
        # must print ONE JSON line, so fd 1 points at /dev/null until the communicator exists.
        import torch
        sys.stdout.flush()
        saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        try:
            dist.init_process_group(backend)
            if backend == "nccl":
                local = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
                torch.cuda.set_device(local)
                t = torch.zeros(1, device=f"cuda:{local}")
                dist.all_reduce(t)                 # creates the communicator (and prints the banner into /dev/null)
                torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)
    return dist


def max_over_ranks(dist, value: float, device=None) -> float:
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class ClockSampler:
    """nvidia-smi clocks/throttle sampling DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       power_w_median=float(np.median(pw)))
        return out


def make_inputs(first_index: int, count: int, size: int = 1024, boxes: int = 1, bit_depth: int = 8):
    """Grey frames + box prompts of the synthetic generator, generated on a few threads."""
    from concurrent.futures import ThreadPoolExecutor
    from yolo_sam_inference_b200.synth import synth_image
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        out = list(ex.map(lambda i: synth_image(i, size, boxes, bit_depth), range(first_index, first_index + count)))
    return [o[0] for o in out], [o[1] for o in out]


def write_folder(grays, boxes, prefix="img", repeat=1):
    """The images as baseline (uncompressed) single-channel TIFFs named *.tiff (the reference globs only *.png, *.jpg and
    *.tiff, pipeline.py:265-269) in a fresh directory; returns (dir, {file name: boxes}). ``repeat`` writes the whole set that
    many times under different names: a longer folder from the same synthetic frames, so that one process_directory call is
    many batches long and the two-slot pipeline's fill / drain does not dominate the measurement."""
    import cv2
    need = repeat * sum(g.nbytes for g in grays) * 1.5
    bases = []
    if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) and shutil.disk_usage("/dev/shm").free > need:
        bases.append("/dev/shm")
    bases.append(None)                   # the default temporary directory: always tried last
    for base in bases:
        d = tempfile.mkdtemp(prefix="ysi_bench_", dir=base)
        table, ok = {}, True
        for r in range(repeat):
            for i, (g, b) in enumerate(zip(grays, boxes)):
                name = f"{prefix}_{r * len(grays) + i:05d}.tiff"
                path = os.path.join(d, name)
                if not cv2.imwrite(path, g, [cv2.IMWRITE_TIFF_COMPRESSION, 1]) or os.path.getsize(path) < g.nbytes:
                    ok = False           # e.g. another rank filled the shared-memory file system meanwhile
                    break
                table[name] = b
            if not ok:
                break
        if ok:
            return d, table
        shutil.rmtree(d, ignore_errors=True)
    raise IOError("could not write the benchmark folder (no space in /dev/shm nor in the temporary directory)")


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (transformers fp32 + restated skimage metrics) on host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_images_per_s(model_name: str, boxes: int, n_images: int, warmup: int, threads: int, budget_s: float = 150.0):
    """Times pipeline.py:161-175 as the reference executes it on the CPU (same images, same boxes, same weights)."""
    import torch
    from oracle import metrics_oracle, sam_oracle
    from yolo_sam_inference_b200.synth import gray_to_rgb_u8
    torch.set_num_threads(threads)
    model = sam_oracle.build_model(model_name, 1234)
    grays, bxs = make_inputs(0, max(min(n_images, 32), 1), 1024, boxes)
    imgs = [gray_to_rgb_u8(g) for g in grays]

    def one(i):
        im, bx = imgs[i % len(imgs)], bxs[i % len(imgs)]
        masks, _ = sam_oracle.run_stage(model, im, bx)
        for m in masks:
            if m.any():
                metrics_oracle.calculate_metrics(im, m)

    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    done = 0
    for i in range(n_images):
        one(i)
        done += 1
        if time.perf_counter() - t0 > budget_s:      # keep the whole run within a few minutes
            break
    dt = time.perf_counter() - t0
    return done / dt, dt, done


def metric_name(model: str, boxes: int) -> str:
    return "SAM box-prompt images/s (%s, 1024x1024, %d box%s/image%s)" % (
        MODEL_NAME.get(model, model), boxes, "" if boxes == 1 else "es", "; masks/s == images/s" if boxes == 1 else "")


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    model, boxes, size, _, desc = WORKLOADS[args.workload if args.workload != "folder5" else "b1"]
    threads = os.cpu_count() or 1
    warm = min(args.warmup, 1)
    ips, dt, done = cpu_reference_images_per_s(model, boxes, args.steps, warm, threads)
    line = {
        "impl": "reference", "metric": metric_name(model, boxes), "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": done, "warmup": warm, "ms_per_step": 1e3 * dt / max(done, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc + "; CPU arm: each step = 1 image of that workload (a bounded sample: the whole 256-image "
                               "step would take minutes per step) through transformers SamModel fp32 + restated skimage/scipy "
                               "metrics on all host cores; warm-up clamped to 1 step (a CPU has nothing to warm beyond the "
                               "first call)"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"{done} images, {boxes} box(es) each, seeded random-init {MODEL_NAME.get(model, model)}"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
def host_dry_run(rank: int, world: int, n_images: int = 10):
    """Host-side logic of the multi-rank bench without a GPU (used by the gloo test): shard, time, reduce."""
    import torch.distributed as dist
    from yolo_sam_inference_b200.sharding import shard_range
    import torch
    mine = shard_range(n_images * 1, rank, world)
    t0 = time.perf_counter()
    time.sleep(0.01 * (rank + 1))
    dt = time.perf_counter() - t0
    tmax = max_over_ranks(dist, dt)
    counts = torch.zeros(world, dtype=torch.int64)
    counts[rank] = len(mine)
    dist.all_reduce(counts)           # bookkeeping only (not on the data path)
    return {"n_gpus": world, "images_total": int(counts.sum()), "shards": counts.tolist(), "scaling": "weak",
            "ms_per_step": tmax * 1e3}


class Leg:
    """One model / box-count configuration on one GPU: resident pool, batches, timing helpers."""

    def __init__(self, model: str, boxes: int, local: int, images: int, first_index: int, precision: str = PRECISION):
        from yolo_sam_inference_b200.sam_stage import SamStage
        from yolo_sam_inference_b200.synth import gray_to_rgb_u8
        from yolo_sam_inference_b200.weights import seeded_state_dict
        self.model, self.boxes_per_image, self.n = model, boxes, images
        self.grays, self.boxes = make_inputs(first_index, images, 1024, boxes)
        self.stage = SamStage(model, device=f"cuda:{local}", state_dict=seeded_state_dict(model, 1234), max_batch=BATCH,
                              max_boxes=BATCH * boxes, max_image_hw=(1024, 1024), on_empty="zeros", precision=precision)
        self.stage.pool_upload([gray_to_rgb_u8(g) for g in self.grays])
        self.nbat = images // BATCH

    def step(self, sync_last: bool = False):
        """One step = every batch of the pool once, enqueued back to back."""
        for k in range(self.nbat):
            b = k * BATCH
            self.stage.compute_pool(b, BATCH, self.boxes[b:b + BATCH], sync=sync_last and k == self.nbat - 1)

    def timed(self, steps: int, warmup: int, dist=None, dev=None, sampler=None):
        for _ in range(warmup):
            self.step(sync_last=True)
        if dist is not None:
            dist.barrier()
        self.stage.sync()
        if sampler is not None:
            sampler.start()
        l0 = self.stage.launch_count
        self.stage.timer_record(0)
        for _ in range(steps):
            self.step()
        self.stage.timer_record(1)
        self.stage.sync()
        ms = self.stage.timer_elapsed_ms(0, 1)
        launches = self.stage.launch_count - l0
        clocks = sampler.stop() if sampler is not None else {}
        if dist is not None:
            dist.barrier()
        ms = max_over_ranks(dist, ms, dev)
        return ms, launches, clocks

    def profile(self, batches: int):
        self.stage.profile(True)
        for k in range(batches):
            b = (k % self.nbat) * BATCH
            self.stage.compute_pool(b, BATCH, self.boxes[b:b + BATCH], sync=True)
        prof = self.stage.profile_read()
        self.stage.profile(False)
        return prof

    def close(self):
        self.stage.close()


def folder_e2e(model: str, grays, boxes, local: int, steps: int, warmup: int, dist=None, dev=None, precision: str = PRECISION,
               repeat: int = 1):
    """images/s of CellSegmentationPipeline.process_directory over a folder of baseline TIFFs (one call per step)."""
    from yolo_sam_inference_b200.pipeline import BoxTable, CellSegmentationPipeline
    from yolo_sam_inference_b200.weights import seeded_state_dict
    folder, table = write_folder(grays, boxes, repeat=repeat)
    n_files = repeat * len(grays)
    out_dir = tempfile.mkdtemp(prefix="ysi_bench_out_")
    nb = max(len(b) for b in boxes)
    H, W = grays[0].shape
    try:
        pipe = CellSegmentationPipeline(None, model, device=f"cuda:{local}", detector=BoxTable(table),
                                        sam_state_dict=seeded_state_dict(model, 1234), max_boxes=BATCH * nb, max_image_hw=(H, W),
                                        on_empty="zeros", batch_size=BATCH, mask_output="packed",
                                        mask_sink=lambda name, packed: None, precision=precision)
        cells = 0
        for _ in range(max(warmup, 1)):
            pipe.process_directory(folder, out_dir, save_visualizations=False)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = pipe.process_directory(folder, out_dir, save_visualizations=False)
            cells += len(res.metrics_data)
        pipe.sam_stage.sync()
        dt = time.perf_counter() - t0
        assert len(res.results) == n_files and cells == steps * repeat * sum(len(b) for b in boxes)
        load_s = res.total_timing["image_load"]
        pipe.close()
    finally:
        shutil.rmtree(folder, ignore_errors=True)
        shutil.rmtree(out_dir, ignore_errors=True)
    dt = max_over_ranks(dist, dt, dev)
    bpp = grays[0].dtype.itemsize
    h2d = repeat * (len(grays) * H * W * bpp + sum(len(b) for b in boxes) * (4 * 8 + 8))
    d2h = repeat * sum(len(b) for b in boxes) * ((H * W + 7) // 8 + 1192)
    return {"images_per_s": n_files * steps / dt, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "loader_thread_seconds_per_step": load_s, "files_per_step": n_files}


def run_ours(args):
    import torch
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dist = init_dist(world, "nccl")
    dev = torch.device("cuda", local)
    model, boxes, size, step_images, desc = WORKLOADS[args.workload]
    K, Wm = args.steps, max(args.warmup, 3)
    pk = peaks()

    if args.workload == "folder5":
        grays, bxs = make_inputs(rank * step_images, step_images, size, boxes, bit_depth=16)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        f = folder_e2e(model, grays, bxs, local, K, Wm, dist, dev, repeat=2)
        step_images = f["files_per_step"]
        clocks = sampler.stop() if rank == 0 else {}
        if rank == 0:
            v = world * f["images_per_s"]
            print(json.dumps({
                "metric": "SAM box-prompt images/s (ViT-B, folder of 2048x2048 16-bit TIFFs, 1 box/image)", "value": v,
                "unit": "images/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": 1e3 * step_images / f["images_per_s"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": PRECISION, "data": "synthetic",
                "config": {"workload": desc + f"; {step_images} files per GPU and step, baseline (uncompressed) TIFF strips read "
                                              "into pinned memory, 16->8 bit + grey->RGB + 2x antialias downscale on the device"},
                "clocks": clocks, "gpu_launches": None,
                "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": f["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": f["d2h_bytes_per_step"]}}))
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- device-resident leg -------------------------------------------------------------------
    leg = Leg(model, boxes, local, step_images, rank * step_images)
    ms_total, launches, clocks = leg.timed(K, Wm, dist, dev, ClockSampler(local) if rank == 0 else None)
    value = world * K * step_images / (ms_total / 1e3)

    # ---- profiled pass: per-launch CUDA events by kernel class (roofline + breakdown) ---------------------
    roof, breakdown = None, None
    if rank == 0:
        pb = 4
        prof = leg.profile(pb)
        gemm_classes = ["gemm_patch", "gemm_qkv", "gemm_proj", "gemm_fc1", "gemm_fc2"]
        g_ms = sum(prof[c]["ms"] for c in gemm_classes)
        g_fl = sum(prof[c]["flops"] for c in gemm_classes)
        g_n = sum(prof[c]["records"] for c in gemm_classes)
        tot_ms = sum(v["ms"] for v in prof.values())
        achieved = g_fl / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "gemm2_op16_kernel / gemm_op16_kernel (tcgen05 GEMMs of the ViT linears)",
                "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_tflops_sustained"], "frac_of_burst": achieved / pk["bf16_tflops"],
                "peak_source": pk["src"] + ", sustained figure: the kernels are timed inside a seconds-long, power-capped run",
                "flops_per_launch": g_fl / max(g_n, 1), "avg_launch_ms": g_ms / max(g_n, 1),
                "share_of_step": g_ms / tot_ms if tot_ms else None,
                "traffic": ncu_traffic(f"gemm_class_{model}_b{BATCH}"),
                "traffic_source": "profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch, class average)",
                "how": f"CUDA events around every launch, {pb} profiled batches of {BATCH} images after the timed region",
                "note": "since round 2 this class also carries LayerNorm1 (folded into the patch-embed / fc2 / qkv epilogues: operand "
                        "copy, row statistics, mean / rstd): 0.4 ms of LayerNorm kernels per batch became 0.25 ms of extra GEMM "
                        "time (the class averaged 0.88 of the sustained peak with separate LayerNorm kernels, YSI_LN_FUSED=0)"}
        breakdown = {k: {"ms_per_batch": v["ms"] / pb,
                         "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["ms"] > 0 and v["flops"] > 0 else None}
                     for k, v in prof.items()}
        enc_ms = sum(prof[c]["ms"] for c in gemm_classes + ["attn_window", "attn_global", "layernorm", "neck"]) / pb
        breakdown["_encoder_alg_tflops"] = ENC_FLOPS[model] * BATCH / (enc_ms / 1e3) / 1e12 if enc_ms else None
        breakdown["_encoder_tensor_util_of_sustained"] = breakdown["_encoder_alg_tflops"] / pk["bf16_tflops_sustained"] if enc_ms else None
    grays_main, boxes_main = leg.grays, leg.boxes
    leg.close()

    # ---- end-to-end leg: folder -> process_directory -> metric dicts -----------------------------------------
    e2e = folder_e2e(model, grays_main, boxes_main, local, K, min(Wm, 3), dist, dev)
    e2e_value = world * e2e["images_per_s"]

    # ---- the other configs, short runs (rank 0 keeps the line; every rank runs them so GPUs stay in step) ----
    extra = {}
    roof_hbm = None
    if not args.no_extra and args.workload == "b1":
        # configs[3]: 32 boxes per image
        l32 = Leg("vit_b", 32, local, 32, 5000 + rank * 32)
        ms32, _, _ = l32.timed(6, 3, dist, dev)
        ips32 = world * 6 * 32 / (ms32 / 1e3)
        extra["configs3_b32_images_per_s"] = ips32
        extra["configs3_b32_masks_per_s"] = ips32 * 32
        if rank == 0:
            p32 = l32.profile(3)
            up, hull = p32["post_upsample"], p32["post_hull"]
            ms = up["ms"] + hull["ms"]
            by = up["bytes"] + hull["bytes"]
            gbs = by / (ms / 1e3) / 1e9 if ms > 0 else 0.0
            roof_hbm = {"bound": "hbm", "kernel": "upsample_stats_fast_kernel + contour_hull_disk_kernel (a6 + a7, 256 masks per launch)",
                        "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                        "peak_source": pk["src"],
                        "bytes_per_mask": by / max(up["records"], 1) / (BATCH * 32),
                        "bytes_note": "algorithmic (SURVEY section 8d): read 256x256 fp32 logits + write the packed-bit mask "
                                      "(131072 B at 1024x1024; the wire format of utils/mask_encoding.py) once, the contour kernel "
                                      "re-reads the packed mask + writes the 1192 B metric row",
                        "avg_launch_ms": {"upsample": up["ms"] / max(up["records"], 1), "hull": hull["ms"] / max(hull["records"], 1)},
                        "traffic": ncu_traffic("post_b32"), "workload": "configs[3]: 32 boxes/image, batch 8 images"}
            extra["configs3_breakdown_ms_per_batch"] = {k: v["ms"] / 3 for k, v in p32.items() if v["ms"] > 0}
        e32 = folder_e2e("vit_b", l32.grays, l32.boxes, local, 3, 1, dist, dev, repeat=8)       # 256 files = 32 batches per call
        extra["configs3_b32_e2e_images_per_s"] = world * e32["images_per_s"]
        l32.close()
        # configs[2]'s model
        lh = Leg("vit_h", 1, local, 32, 6000 + rank * 32)
        msh, _, _ = lh.timed(3, 3, dist, dev)
        extra["configs2_vit_h_images_per_s"] = world * 3 * 32 / (msh / 1e3)
        extra["configs2_vit_h_alg_tflops_per_gpu"] = (ENC_FLOPS["vit_h"] + DEC_FLOPS_BOX) * extra["configs2_vit_h_images_per_s"] / world / 1e12
        lh.close()
        # configs[4]-style folder: 2048x2048 16-bit TIFFs from files
        g5, b5 = make_inputs(7000 + rank * 32, 32, 2048, 1, bit_depth=16)
        # 128 files = 16 batches per call on one GPU; 64 per rank on several (1 GB of files per rank would crowd /dev/shm at 8 ranks)
        e5 = folder_e2e("vit_b", g5, b5, local, 3, 1, dist, dev, repeat=4 if world == 1 else 2)
        extra["configs4_folder_2048_16bit_e2e_images_per_s"] = world * e5["images_per_s"]
        del g5, b5
        # the bf16-operand build on configs[1] (not the default, does not meet the IoU gate)
        other = "bf16" if PRECISION == "fp16" else "fp16"
        lo = Leg("vit_b", 1, local, 64, rank * STEP_IMAGES, precision=other)
        mso, _, _ = lo.timed(4, 3, dist, dev)
        extra[f"configs1_{other}_operands_images_per_s"] = world * 4 * 64 / (mso / 1e3)
        lo.close()

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ips, dt, n = cpu_reference_images_per_s(model, boxes, 6 if model == "vit_b" else 2, 1, threads, 45.0)
        cpu = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"{n} images ({boxes} box(es) each) of the same workload, {dt:.1f} s: transformers SamModel fp32 + "
                         "restated skimage/scipy metrics"}

    if rank == 0:
        line = {
            "metric": metric_name(model, boxes), "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": PRECISION, "data": "synthetic",
            "config": {"workload": desc + f"; {PRECISION} operands (tcgen05 kind::f16, fp32 accumulate / residual / statistics), "
                                          f"one step = {step_images} images per GPU = {step_images // BATCH} device batches of {BATCH}",
                       "batch": BATCH, "images_per_step_per_gpu": step_images, "boxes_per_image": boxes,
                       "weights": "seeded random-init (no checkpoints offline)",
                       "l2": "every batch of a step is a different set of images and the per-batch working set (~0.9 GB of "
                             "activations at 8 ViT-B images) exceeds the 126 MB L2, so no L2 flush is needed between timed iterations",
                       "parallelism": f"image-sharded x{world}, no collective",
                       "e2e_path": "CellSegmentationPipeline.process_directory on a folder of baseline 8-bit TIFFs (one call per "
                                   "step): read-ahead pool -> pinned -> H2D (raw grey samples) -> device -> D2H (packed masks + "
                                   "metric rows) -> 16-key dicts"},
            "masks_per_s": value * boxes, "boxes_per_image": boxes,
            "alg_tflops": (ENC_FLOPS[model] + DEC_FLOPS_BOX * boxes) * value / world / 1e12,
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": e2e["d2h_bytes_per_step"],
                    "loader_thread_seconds_per_step": e2e["loader_thread_seconds_per_step"]},
            "roofline": roof, "roofline_hbm": roof_hbm, "cpu_baseline": cpu, "breakdown": breakdown, "extra": extra,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="b1", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other configs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
