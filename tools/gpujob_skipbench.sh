mkdir -p gpurun_out
for t in ${TUNES:-8}; do for f in 0 1 2 3; do
  GPUCHAN_TC_TUNE=$t GPUCHAN_DEBUG_STAMPS=1 GPUCHAN_DEBUG_SKIP=$f timeout 90 python bench.py --no-cpu-baseline > gpurun_out/bench_skip$f.json 2> gpurun_out/bench_skip$f.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_skip$f.json')); print('tune $t skip $f kern %.4f ms'%(d['roofline']['kernel_ms_per_launch']))"
done; done
