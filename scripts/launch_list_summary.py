"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
python scripts/launch_list_summary.py gpurun_out/launches.csv "command line that was profiled" > profiles/xyz.txt"""
import csv
import sys
from collections import OrderedDict

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    name = r[ki]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
ours = sum(v[1] for k, v in agg.items() if "cgc::" in k)
print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 %s" % cmd)
print("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's CUDA-event shares (kernel_ms), not absolutes")
for k, (n, t) in agg.items():
    short = k.split("(")[0][:70]
    print("%-70s launches %4d  total %10.1f us  avg %9.1f us  share_of_cgc %.3f" % (short, n, t, t / n, (t / ours) if "cgc::" in k and ours else 0.0))
