#!/bin/bash
# two GPUs at HEAD: the 2-GPU tests, bench at N = 2 (relay chain) for the headline shape and c4
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_host_binary.py -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_2gpu.log; cat gpurun_out/pytest_gpu_2gpu.log
run() { name=$1; np=$2; shift 2
  if [ "$np" = 1 ]; then timeout 300 python bench.py --gpus 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; fi
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
run bench_n1_headline 1 --no-cpu-baseline --steps 10
run bench_n2_headline_relay 2 --no-cpu-baseline --steps 10
run bench_n2_c4 2 --no-cpu-baseline --config c4 --steps 10
