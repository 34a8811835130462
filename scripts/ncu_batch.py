"""One device batch of the bench workload between cudaProfilerStart / Stop (for `ncu --profile-from-start off`).

    python scripts/ncu_batch.py [boxes_per_image=1] [model=vit_b] [profiled_batches=1]

Same objects as bench.py's device-resident leg (resident image pool, ysi_compute_pool); two warm-up batches first, so the
profiled batch replays the encoder's CUDA graph like the timed region does."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402  (cudaProfilerStart / Stop only)

import bench  # noqa: E402


def main():
    boxes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    model = sys.argv[2] if len(sys.argv) > 2 else "vit_b"
    nprof = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    leg = bench.Leg(model, boxes, 0, 3 * bench.BATCH, 0)
    for k in range(2):
        leg.stage.compute_pool(k * bench.BATCH, bench.BATCH, leg.boxes[k * bench.BATCH:(k + 1) * bench.BATCH], sync=True)
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(nprof):
        b = 2 * bench.BATCH
        leg.stage.compute_pool(b, bench.BATCH, leg.boxes[b:b + bench.BATCH], sync=True)
    torch.cuda.cudart().cudaProfilerStop()
    leg.close()
    print("ok", model, boxes)


if __name__ == "__main__":
    main()
