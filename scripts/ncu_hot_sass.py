"""Top stalled SASS instructions of one kernel from an ncu report's source page.

    python scripts/ncu_hot_sass.py <file.ncu-rep> <kernel-name substring> [top N]
"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# the page concatenates kernels: "Kernel Name",<name> line, header line, then instruction rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    if pat not in b["name"]:
        continue
    h = b["hdr"]
    si = h.index("# Samples")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    total = sum(int(r[si] or 0) for r in b["rows"])
    print("kernel:", b["name"][:100], "total samples", total)
    order = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][si] or 0))[:top]
    for i in sorted(order):
        r = b["rows"][i]
        st = sorted(((int(r[c] or 0), h[c][6:]) for c in stall_cols), reverse=True)[:3]
        print("%5d %5.1f%%  #%-5d %-70s %s" % (int(r[si]), 100.0 * int(r[si]) / max(total, 1), i, r[1].strip()[:70],
                                           " ".join("%s=%d" % (n, v) for v, n in st if v)))
    break
