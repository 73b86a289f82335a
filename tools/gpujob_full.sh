#!/bin/bash
# full GPU suite + smoke + the two bench shapes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for cfg in headline c2; do
  timeout 300 python bench.py --config $cfg --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  python - "$cfg" <<'PY'
import json, sys
cfg = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/bench_{cfg}.json")); r = d["roofline"]
    print("%s: kernel_ms %.4f value %.4g e2e %.4g frac %.3f clocks %s" % (cfg, r["kernel_ms_per_launch"], d["value"], d["e2e"]["value"], r["frac"], d["clocks"]))
except Exception as e:
    print(cfg, "FAILED", e, open(f"gpurun_out/bench_{cfg}.err").read()[-800:])
PY
done
