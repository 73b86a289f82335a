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
# kernels of ours that run outside bench.py's timed steps: once per bank (cycle detection at gpuchan_create) and during
# the first warm-up submits only (derotator prepass while some channel is still in its transient)
OUTSIDE = {"rot_cycle_detect_kernel": "bank creation", "rot_prepass_kernel": "warm-up submits only",
           "rot_prepass_table_kernel": "warm-up submits only", "carry_save_kernel": "warm-up submits only"}
timed = {k: v for k, v in ours.items() if k not in OUTSIDE}
tot_timed = sum(v[1] for v in timed.values())
w.writerow(["kernel", "launches", "total_us", "mean_us", "share_of_own_kernels", "share_of_timed_steps", "note"])
for k, v in agg.items():
    if k not in ours:
        row = ["", "", "input synthesis, outside the timed region"]
    elif k in OUTSIDE:
        row = [f"{v[1] / tot:.4f}", "", OUTSIDE[k] + ", outside the timed steps"]
    else:
        row = [f"{v[1] / tot:.4f}", f"{v[1] / tot_timed:.4f}", "timed steps (ncu serialises and adds launch gaps; cold cache)"]
    w.writerow([k, v[0], f"{v[1] / 1e3:.1f}", f"{v[1] / 1e3 / v[0]:.2f}"] + row)
