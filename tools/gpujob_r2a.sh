#!/bin/bash
# round 2, first full pass: GPU tests (1 GPU), bench at the headline shape and at c2, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_headline.json 2> gpurun_out/bench_headline.err; tail -c 3000 gpurun_out/bench_headline.json; tail -5 gpurun_out/bench_headline.err
timeout 300 python bench.py --config c2 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 1500 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
