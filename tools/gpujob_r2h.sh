#!/bin/bash
# round 2: warp-per-channel FLEX check (all GPU tests), racecheck of both decoder kernels, pager chain timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_pager.py tests/test_gpu_flex.py -x -q -k "pocsag_fsm or dc_blocker or flex_dc or flex_noise" > gpurun_out/sanitizer_pager_racecheck.log 2>&1; echo "pager racecheck rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_pager_racecheck.log | tail -4
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_flex.py -x -q > gpurun_out/sanitizer_flex_memcheck.log 2>&1; echo "flex memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_flex_memcheck.log | tail -3
timeout 300 python tools/bench_pager.py > gpurun_out/pager_chain.jsonl 2> gpurun_out/pager_chain.err; cat gpurun_out/pager_chain.jsonl; tail -3 gpurun_out/pager_chain.err
