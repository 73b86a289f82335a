#!/bin/bash
# two channel groups per CTA (default when it fits) against GPUCHAN_TC_GPC=1 (one group per CTA)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_fullsize_chain.py -x -q 2>&1 | grep -E "passed|failed|^FAILED|Error|assert" | tail -6
for mode in 2 1; do
  for cfg in headline c3 c2; do
    GPUCHAN_TC_GPC=$mode timeout 200 python bench.py --config $cfg --steps 6 --submits 8 --no-cpu-baseline > gpurun_out/pair$mode.$cfg.json 2> gpurun_out/pair$mode.$cfg.err
    python - "$mode" "$cfg" <<'PY'
import json, sys
v, cfg = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/pair{v}.{cfg}.json")); r = d["roofline"]
    print("gpc<=%s %s: kernel_ms %.4f value %.4g frac %.3f" % (v, cfg, r["kernel_ms_per_launch"], d["value"], r["frac"]))
except Exception as e:
    print(v, cfg, "FAILED", e, open(f"gpurun_out/pair{v}.{cfg}.err").read()[-600:])
PY
  done
done
