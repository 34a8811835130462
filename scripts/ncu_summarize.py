"""Turns gpurun_out/*.ncu-rep / ncu_launches.csv into the small text summaries committed under profiles/.

    python scripts/ncu_summarize.py <tag>      (run in the dev container; needs ncu on PATH)
"""
import collections
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEYS = [
    "gpu__time_duration.sum", "gpc__cycles_elapsed.max.per_second", "launch__grid_size", "launch__registers_per_thread",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
]


def summarize_rep(path, fh):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h]
    for r in rows[2:]:
        fh.write("kernel: %s\n" % r[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr:
                fh.write("  %-72s %-12s %s\n" % (k, units[hdr.index(k)], r[hdr.index(k)]))
        st = sorted(((float(r[hdr.index(h)]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for h in stall), reverse=True)[:6]
        fh.write("  top stalls (warps per issue): %s\n\n" % ", ".join("%s %.2f" % (n, v) for v, n in st))


def summarize_launches(path, fh):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki].split("(")[0][:80], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    fh.write("launches  total_us  share   kernel      (gpu__time_duration.sum, ncu --clock-control none; cold-cache, serialised)\n")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        fh.write("%8d %9.1f %6.1f%%  %s\n" % (v[0], v[1] / 1e3, 100 * v[1] / tot, k))
    fh.write("total %.1f us over %d launches\n" % (tot / 1e3, sum(v[0] for v in agg.values())))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "*.ncu-rep"))):
        name = os.path.splitext(os.path.basename(rep))[0]
        with open(os.path.join(OUT, f"{tag}_{name}.txt"), "w") as fh:
            fh.write(f"# ncu --set full --clock-control none summary of gpurun_out/{name}.ncu-rep (numbers under ncu are not bench values)\n")
            summarize_rep(rep, fh)
    lc = os.path.join(ROOT, "gpurun_out", "ncu_launches.csv")
    if os.path.exists(lc):
        with open(os.path.join(OUT, f"{tag}_ncu_launch_list.txt"), "w") as fh:
            summarize_launches(lc, fh)


if __name__ == "__main__":
    main()
