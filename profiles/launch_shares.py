#!/usr/bin/env python3
"""Per-kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python profiles/launch_shares.py gpurun_out/r02_launches.csv > profiles/r02_launch_shares.csv"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
agg = OrderedDict()
for r in rows[1:]:
    name = r[ik].split("(")[0]
    t = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    n, tot = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, tot + t)
total = sum(t for _, t in agg.values())
print("kernel,launches,total_us,share (bench.py --steps 1 --warmup 3 --no-exhibits under ncu: set-up, warm-up, timed step, side ensembles, the "
      "two time-step variants; cold-cache serialised times - compare shares, not absolutes)")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name},{n},{t:.1f},{t / total:.4f}")
