#!/bin/bash
# diagnostics builds (TC_DIAG bits: 1 no epilogue arithmetic, 2 no transform, 4 one MMA per warp and tile): kernel time only
mkdir -p gpurun_out
for lib in tsl-sdr_b200/libtslb200.so tsl-sdr_b200/libtslb200_diag*.so; do
  v=$(basename $lib .so)
  for cfg in headline c2; do
    TSLB200_LIB=$PWD/$lib timeout 200 python bench.py --config $cfg --steps 5 --submits 8 --no-cpu-baseline > gpurun_out/var_$v.$cfg.json 2> gpurun_out/var_$v.$cfg.err
    python - "$v" "$cfg" <<'PY'
import json, sys
v, cfg = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/var_{v}.{cfg}.json")); r = d["roofline"]
    print("%s %s: kernel_ms %.4f" % (v, cfg, r["kernel_ms_per_launch"]))
except Exception as e:
    print(v, cfg, "FAILED", e, open(f"gpurun_out/var_{v}.{cfg}.err").read()[-600:])
PY
  done
done
