#!/bin/bash
# experiment runner: bench (headline, c2) + parity for every libtslb200*.so variant in the package directory
mkdir -p gpurun_out
for lib in ${VARLIBS:-tsl-sdr_b200/libtslb200.so tsl-sdr_b200/libtslb200_*.so}; do
  case $lib in *compat*) continue;; esac
  v=$(basename $lib .so)
  [ -n "$VARNOTEST" ] || TSLB200_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_math.py -x -q 2>&1 | grep -E "passed|failed|^FAILED|Error|assert" | tail -6
  for cfg in ${VARCFGS:-headline c2}; do
    TSLB200_LIB=$PWD/$lib timeout 200 python bench.py --config $cfg --steps 8 --submits 8 --no-cpu-baseline > gpurun_out/var_$v.$cfg.json 2> gpurun_out/var_$v.$cfg.err
    python - "$v" "$cfg" <<'PY'
import json, sys
v, cfg = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/var_{v}.{cfg}.json")); r = d["roofline"]
    print("%s %s: kernel_ms %.4f value %.4g clocks %s" % (v, cfg, r["kernel_ms_per_launch"], d["value"], d["clocks"]["reasons"]))
except Exception as e:
    print(v, cfg, "FAILED", e, open(f"gpurun_out/var_{v}.{cfg}.err").read()[-800:])
PY
  done
done
