#!/bin/bash
# round 2: decoder-side rewrite check + measurements: GPU tests, pager chain timing, shape survey, ncu of the headline kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_pager.py > gpurun_out/pager_chain.jsonl 2> gpurun_out/pager_chain.err; cat gpurun_out/pager_chain.jsonl; tail -3 gpurun_out/pager_chain.err
BATCH_LOG2=25 timeout 600 python tools/bench_configs.py > gpurun_out/shapes_survey.jsonl 2> gpurun_out/shapes_survey.err; cat gpurun_out/shapes_survey.jsonl; tail -3 gpurun_out/shapes_survey.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fir_fm -s 8 -c 1 -o gpurun_out/prof_headline python bench.py --steps 1 --warmup 3 --submits 2 --no-cpu-baseline > gpurun_out/ncu_headline.log 2>&1; tail -2 gpurun_out/ncu_headline.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pocsag_kernel -s 1 -c 1 -o gpurun_out/prof_pocsag python tools/bench_pager.py > gpurun_out/ncu_pocsag.log 2>&1; tail -2 gpurun_out/ncu_pocsag.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_headline.csv python bench.py --steps 2 --warmup 3 --submits 2 --no-cpu-baseline > gpurun_out/launches_headline.log 2>&1; tail -1 gpurun_out/launches_headline.log | head -c 300
ls -la gpurun_out | head -40
