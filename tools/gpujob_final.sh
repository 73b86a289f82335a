#!/bin/bash
# round 2 final pass on one B200: tests, smoke, both bench arms, c2, launch list, ncu --set full of the headline kernel, decoder chain, shape survey
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; tail -c 400 gpurun_out/bench_reference_arm.json
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_headline.json 2> gpurun_out/bench_headline.err; tail -c 2600 gpurun_out/bench_headline.json; tail -3 gpurun_out/bench_headline.err
timeout 300 python bench.py --config c2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 900 gpurun_out/bench_c2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_headline.csv python bench.py --steps 2 --warmup 3 --submits 2 --no-cpu-baseline > gpurun_out/launches_headline.log 2>&1; tail -3 gpurun_out/launches_headline.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fir_fm -s 10 -c 2 -o gpurun_out/prof_headline python bench.py --steps 1 --warmup 3 --submits 2 --no-cpu-baseline > gpurun_out/ncu_headline.log 2>&1; tail -2 gpurun_out/ncu_headline.log
timeout 300 python tools/bench_pager.py > gpurun_out/pager_chain.jsonl 2> gpurun_out/pager_chain.err; cat gpurun_out/pager_chain.jsonl
BATCH_LOG2=25 timeout 600 python tools/bench_configs.py > gpurun_out/shapes_survey.jsonl 2> gpurun_out/shapes_survey.err; wc -l gpurun_out/shapes_survey.jsonl
