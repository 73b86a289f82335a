mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -x -q 2>&1 | tail -2
for t in 0 32; do
  GPUCHAN_TC_TUNE=$t timeout 90 python bench.py --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_d.json')); print('tune $t kern %.4f ms chk %d'%(d['roofline']['kernel_ms_per_launch'], d['e2e']['pcm_checksum']))"
  GPUCHAN_TC_TUNE=$t GPUCHAN_DEBUG_STAMPS=1 GPUCHAN_DEBUG_SKIP=3 timeout 90 python bench.py --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_d.json')); print('tune $t skip 3 kern %.4f ms'%(d['roofline']['kernel_ms_per_launch']))"
done
