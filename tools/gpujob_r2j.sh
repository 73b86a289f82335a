#!/bin/bash
# pipe microbenchmark + packed-vs-scalar FP variants of the epilogue
mkdir -p gpurun_out
timeout 120 tools/pipe_microbench > gpurun_out/pipe_microbench.txt 2>&1; cat gpurun_out/pipe_microbench.txt
bash tools/gpujob_var.sh
