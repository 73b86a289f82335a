#!/bin/bash
# round 2, eight GPUs: weak scaling at the headline shape (N = 4, 8), c4 strong scaling (N = 4, 8), c2 at N = 8, 8-GPU host binary
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
run() { # name, nproc, args...
  name=$1; np=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print(n, "N=%d value %.4g e2e %.4g ms/step %.3f kernel_ms %.4f (max %.4f) share %.2f fanout %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["kernel_ms_per_launch_max_over_ranks"], r["kernel_share_of_step"], d.get("fanout")))
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/{n}.err").read()[-2500:])
PY
}
run bench_n8_headline_relay 8 --no-cpu-baseline
run bench_n4_headline_relay 4 --no-cpu-baseline
run bench_n8_c4 8 --no-cpu-baseline --config c4
run bench_n4_c4 4 --no-cpu-baseline --config c4
run bench_n8_c2_relay 8 --no-cpu-baseline --config c2
run bench_n8_headline_nccl 8 --no-cpu-baseline --fanout nccl
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
