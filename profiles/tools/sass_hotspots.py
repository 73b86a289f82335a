#!/usr/bin/env python
"""Per-SASS-instruction execution counts and stall samples from an ncu report (needs -lineinfo / --import-source on).
Writes the annotated listing to argv[2] and prints segments of equal execution count (= loops / roles) to stdout."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[starts[0]]
end = starts[1] - 1 if len(starts) > 1 else len(rows)
body = [r for r in rows[starts[0] + 1:end] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in body)
tot_s = sum(int(r[ix["# Samples"]]) for r in body)
agg = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
print(f"sass lines {len(body)}  warp-instructions {tot_i}  samples {tot_s}")
print("stalls:", sorted(agg.items(), key=lambda x: -x[1])[:8])
lines = []
for r in body:
    top = max(stalls, key=lambda s: int(r[ix[s]]))
    lines.append((int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]), top[6:], r[ix["Source"]]))
with open(out, "w") as f:
    for n, (c, s, t, src) in enumerate(lines):
        f.write(f"{n:5} {c:>9} {s:>6} {t:<14} {src}\n")
seg = []
cur = None
for i, (c, s, t, src) in enumerate(lines):
    if cur and abs(c - cur[2]) <= 0.02 * max(c, cur[2]):
        cur[1] = i; cur[3] += c; cur[4] += s
    else:
        if cur: seg.append(cur)
        cur = [i, i, c, c, s]
seg.append(cur)
for s in seg:
    if s[3] > 0.003 * tot_i or s[4] > 0.01 * tot_s:
        print(f"lines {s[0]:5}-{s[1]:5} n={s[1]-s[0]+1:4} exec/inst={s[2]:>9} warp-inst={s[3]:>10} ({100*s[3]/tot_i:4.1f}%) samples={s[4]} ({100*s[4]/tot_s:4.1f}%)")
