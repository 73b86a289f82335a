# diagnostics: kernel time with parts of the epilogue compiled out (make -C tsl-sdr_b200/csrc diag)
mkdir -p gpurun_out
for d in 0 1 3 7; do
  lib=""; [ $d != 0 ] && lib=$PWD/tsl-sdr_b200/libtslb200_diag$d.so
  TSLB200_LIB=$lib timeout 90 python bench.py --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_d.json')); print('diag $d kern %.4f ms'%(d['roofline']['kernel_ms_per_launch']))"
done
for f in 1 3; do
  GPUCHAN_DEBUG_STAMPS=1 GPUCHAN_DEBUG_SKIP=$f timeout 90 python bench.py --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_d.json')); print('skip $f kern %.4f ms'%(d['roofline']['kernel_ms_per_launch']))"
done
