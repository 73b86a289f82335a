mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_pager.py -x -q 2>&1 | tail -3
for t in ${TUNES:-0 1 2 3 4}; do
  GPUCHAN_TC_TUNE=$t timeout 90 python bench.py --no-cpu-baseline > gpurun_out/bench_tune$t.json 2> gpurun_out/bench_tune$t.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_tune$t.json')); print('tune $t value %.4g ms %.4f kern %.4f frac %.4f chk %d'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac'],d['e2e']['pcm_checksum']))"
done
