"""profiles/ncu_traffic.json + launch lists from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--clock-control none --csv` captures of one bench batch (gpurun_out/ncu_launches_b1.csv, ncu_launches_b32.csv).

    python scripts/ncu_traffic.py r02          (run in the dev container)

bench.py reads the JSON for `roofline.traffic` (bytes per launch, dram read + write)."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        d = launches.setdefault(r[ii], {"name": r[ki]})
        d[r[mi]] = float(r[vi].replace(",", ""))
    return list(launches.values())


def short(name):
    n = name.split("(")[0]
    for a, b in (("ysi::", ""), ("void ", "")):
        n = n.replace(a, b)
    return n[:90]


def summarize(launches, out_path, title):
    agg = collections.OrderedDict()
    for l in launches:
        a = agg.setdefault(short(l["name"]), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += l.get("gpu__time_duration.sum", 0.0)
        a[2] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    tot = sum(v[1] for v in agg.values())
    with open(out_path, "w") as fh:
        fh.write(f"# {title}\n# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                 "(cold-cache, serialised launches: compare SHARES, not absolutes)\n")
        fh.write("launches  total_us  share  dram_MB/launch  kernel\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            fh.write("%8d %9.1f %5.1f%% %12.2f  %s\n" % (v[0], v[1] / 1e3, 100 * v[1] / tot, v[2] / v[0] / 1e6, k))
        fh.write("total %.1f us over %d launches\n" % (tot / 1e3, sum(v[0] for v in agg.values())))
    return agg


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = {}
    p1 = os.path.join(ROOT, "gpurun_out", "ncu_launches_b1.csv")
    if os.path.exists(p1):
        agg = summarize(load(p1), os.path.join(ROOT, "profiles", f"{tag}_ncu_launch_list.txt"),
                        "launch list of one default bench batch (configs[1]: ViT-B, 8 images, 1 box each)")
        g = [(k, v) for k, v in agg.items() if "gemm2_op16_kernel" in k]
        n = sum(v[0] for _, v in g)
        if n:
            out["gemm_class_vit_b_b8"] = sum(v[2] for _, v in g) / n
            out["gemm_class_vit_b_b8_launches"] = n
    p32 = os.path.join(ROOT, "gpurun_out", "ncu_launches_b32.csv")
    if os.path.exists(p32):
        agg = summarize(load(p32), os.path.join(ROOT, "profiles", f"{tag}_ncu_launch_list_b32.txt"),
                        "launch list of one configs[3] batch (ViT-B, 8 images, 32 boxes each)")
        post = [(k, v) for k, v in agg.items() if "upsample_stats" in k or "contour_hull_disk" in k]
        batches = max(min(v[0] for _, v in post), 1) if post else 1
        if post:
            out["post_b32"] = sum(v[2] for _, v in post) / batches
    json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
