#!/bin/bash
# round 2: packed-pair epilogue (v3) check: GPU tests (incl. the math selftests), bench headline + c2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_headline_v3.json 2> gpurun_out/bench_headline_v3.err; python - <<'PY'
import json
for f in ("bench_headline_v3",):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d["roofline"]
        print(f, "value %.4g e2e %.4g kernel_ms %.4f frac %.4f mac_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], r["kernel_ms_per_launch"], r["frac"], r["mac_frac"], d["clocks"]))
    except Exception as e: print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
timeout 300 python bench.py --config c2 --no-cpu-baseline > gpurun_out/bench_c2_v3.json 2> gpurun_out/bench_c2_v3.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c2_v3.json")); r=d["roofline"]
print("c2 value %.4g e2e %.4g kernel_ms %.4f frac %.4f" % (d["value"], d["e2e"]["value"], r["kernel_ms_per_launch"], r["frac"]))
PY
