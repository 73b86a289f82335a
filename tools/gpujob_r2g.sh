#!/bin/bash
# round 2: compute-sanitizer over the rewritten kernels (pager bank, packed epilogue, relay / multi-device bank) + the new math test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_math.py -x -q 2>&1 | tail -3
export PYTHONFAULTHANDLER=1
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_pager.py -x -q -k "pocsag or resampler_matches or dc_blocker or whole_chain" > gpurun_out/sanitizer_pager_$tool.log 2>&1; echo "pager $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Race|Error" gpurun_out/sanitizer_pager_$tool.log | tail -4
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_multi.py -x -q -k "not full_size and not across_processes" > gpurun_out/sanitizer_chan_memcheck.log 2>&1; echo "chan memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitizer_chan_memcheck.log | tail -4
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_tc.py tests/test_gpu_pager.py -x -q -k "tc or pocsag_fsm" > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitizer_synccheck.log | tail -4
