#!/bin/bash
# 6 or 8 transform warps for the shapes that keep one channel group per CTA
mkdir -p gpurun_out
for cfg in c4 c5 headline; do
  for xf in 6 8; do
    extra=""; [ $cfg = headline ] && extra="GPUCHAN_TC_GPC=1"
    env $extra GPUCHAN_TC_XF=$xf timeout 200 python bench.py --config $cfg --steps 6 --submits 8 --no-cpu-baseline > gpurun_out/xfsel_$cfg.$xf.json 2> gpurun_out/xfsel_$cfg.$xf.err
    python - "$cfg" "$xf" <<'PY'
import json, sys
cfg, xf = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/xfsel_{cfg}.{xf}.json")); r = d["roofline"]
    print("%s XF=%s: kernel_ms %.4f value %.4g" % (cfg, xf, r["kernel_ms_per_launch"], d["value"]))
except Exception as e:
    print(cfg, xf, "FAILED", e, open(f"gpurun_out/xfsel_{cfg}.{xf}.err").read()[-600:])
PY
  done
done
