mkdir -p gpurun_out
for sl in ${SLEEPS}; do
  GPUCHAN_TC_SLEEP=$sl timeout 90 python bench.py --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_s.json')); print('sleep $sl value %.4g kern %.4f frac %.4f chk %d'%(d['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac'],d['e2e']['pcm_checksum']))"
done
GPUCHAN_TC_TUNE=19 timeout 90 python bench.py --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_s.json')); print('suspended waits: value %.4g kern %.4f frac %.4f chk %d'%(d['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac'],d['e2e']['pcm_checksum']))"
