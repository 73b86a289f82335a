#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; head -12 gpurun_out/topo.txt
numactl -H 2>/dev/null | head -5
run() { name=$1; np=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print(n, "N=%d value %.4g e2e %.4g ms/step %.3f kernel_ms %.4f numa %s h2d %d" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms_per_launch"], d.get("host_numa_node_rank0"), d["e2e"]["h2d_bytes_per_step"]))
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/{n}.err").read()[-2500:])
PY
}
run bench_n8_headline 8 --no-cpu-baseline --steps 5
run bench_n8_headline_e2ehost 8 --no-cpu-baseline --steps 5 --e2e-fanout host
run bench_n4_headline 4 --no-cpu-baseline --steps 5
