#!/bin/bash
mkdir -p gpurun_out
NP=$(nvidia-smi --query-gpu=index --format=csv,noheader | wc -l)
run() { # name, nproc, args...
  name=$1; np=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print(n, "N=%d value %.4g e2e %.4g ms/step %.3f kernel_ms %.4f (max %.4f) share %.2f fanout %s h2d %d" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["kernel_ms_per_launch_max_over_ranks"], r["kernel_share_of_step"], d.get("fanout"), d["e2e"]["h2d_bytes_per_step"]))
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/{n}.err").read()[-2500:])
PY
}
run bench_n${NP}_headline_relay $NP --no-cpu-baseline
run bench_n${NP}_headline_e2ehost $NP --no-cpu-baseline --e2e-fanout host
