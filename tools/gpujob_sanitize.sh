#!/bin/bash
# compute-sanitizer on the fused kernel at HEAD (TMA tap image, two channel groups per CTA): memcheck and synccheck
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "tc and (noise_one_shot or chunked or steady_state or tensor_core_engine_shapes)" > gpurun_out/sanitizer_chan_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_chan_memcheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "tc and (noise_one_shot or steady_state)" > gpurun_out/sanitizer_chan_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/sanitizer_chan_synccheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_fullsize_chain.py -x -q > gpurun_out/sanitizer_fullsize_memcheck.log 2>&1; echo "fullsize memcheck rc=$?"; tail -4 gpurun_out/sanitizer_fullsize_memcheck.log
