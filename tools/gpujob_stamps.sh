timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -x -q 2>&1 | tail -8
python profiles/tools/tc_role_stamps.py > gpurun_out/stamps.txt 2>&1; cat gpurun_out/stamps.txt | head -48
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tc.json')); print('value',d['value'],'ms',d['ms_per_step'],'kern',d['roofline']['kernel_ms_per_launch'],'frac',d['roofline']['frac'],'chk',d['e2e']['pcm_checksum'])"
tail -3 gpurun_out/bench_tc.err
