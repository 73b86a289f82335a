mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python profiles/tools/tc_role_stamps.py > gpurun_out/stamps.txt 2>&1; head -48 gpurun_out/stamps.txt
timeout 90 python bench.py --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tc.json')); print('value',d['value'],'ms',d['ms_per_step'],'kern',d['roofline']['kernel_ms_per_launch'],'frac',d['roofline']['frac'],'chk',d['e2e']['pcm_checksum'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])"
tail -3 gpurun_out/bench_tc.err
