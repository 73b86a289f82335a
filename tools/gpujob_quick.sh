mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tc.json')); print('value',d['value'],'ms',d['ms_per_step'],'kern',d['roofline']['kernel_ms_per_launch'],'frac',d['roofline']['frac'],'chk',d['e2e']['pcm_checksum'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])"
tail -3 gpurun_out/bench_tc.err
