mkdir -p gpurun_out
for env in "" "NCCL_MIN_NCHANNELS=32" "NCCL_MIN_NCHANNELS=32 NCCL_P2P_NVL_CHUNKSIZE=4194304" "NCCL_ALGO=Ring NCCL_MIN_NCHANNELS=32"; do
  echo "== env: $env"
  env $env timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/bench_n2.err | tail -1 | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g ms/step %.4f kern %.4f e2e %.4g'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch'],d['e2e']['value']))"
done
