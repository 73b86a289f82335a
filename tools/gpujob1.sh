set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
nproc >> gpurun_out/smi.txt; lscpu | grep "Model name" >> gpurun_out/smi.txt
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 120 python bench.py > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
timeout 120 python bench.py --engine 1 --no-cpu-baseline > gpurun_out/bench_imad.json 2> gpurun_out/bench_imad.err
timeout 120 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fir_fm -s 3 -c 2 -o gpurun_out/prof_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_tc.json; cat gpurun_out/bench_ref.json
