#!/bin/bash
# experiment: number of transform warps (4 / 8 / 12) -- is the raw -> byte-plane transform what paces the tile pipeline?
mkdir -p gpurun_out
for v in "" _xf4 _xf12; do
  lib=$PWD/tsl-sdr_b200/libtslb200$v.so
  TSLB200_LIB=$lib timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  for cfg in headline c2; do
    TSLB200_LIB=$lib timeout 200 python bench.py --config $cfg --steps 5 --submits 8 --no-cpu-baseline > gpurun_out/xf$v.$cfg.json 2> gpurun_out/xf$v.$cfg.err
    python - "$v" "$cfg" <<'PY'
import json, sys
v, cfg = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/xf{v}.{cfg}.json")); r = d["roofline"]
    print("variant '%s' %s: kernel_ms %.4f value %.4g clocks %s" % (v, cfg, r["kernel_ms_per_launch"], d["value"], d["clocks"]))
except Exception as e:
    print(v, cfg, "FAILED", e, open(f"gpurun_out/xf{v}.{cfg}.err").read()[-800:])
PY
  done
done
