#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list: one row per kernel name with launch count,
total and mean device time, and the share of the listed launches.  torch's data-generation kernels (bench.py
synthesises its input on the device before the timed region) are kept but marked."""
import csv
import sys
from collections import OrderedDict

src = sys.argv[1]
rows = [r for r in csv.reader(open(src, newline="")) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rows[1:]:
    name = r[ik]
    short = name.split("(")[0].split("::")[-1][:60] if "at::" not in name else "torch:" + name.split("<")[0].split("::")[-1][:40]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", ""))
ours = {k: v for k, v in agg.items() if not k.startswith("torch:")}
tot = sum(v[1] for v in ours.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_us", "mean_us", "share_of_own_kernels"])
for k, v in agg.items():
    w.writerow([k, v[0], f"{v[1] / 1e3:.1f}", f"{v[1] / 1e3 / v[0]:.2f}", f"{v[1] / tot:.4f}" if k in ours else "(input synthesis, outside the timed region)"])
