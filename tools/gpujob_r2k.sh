#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/pipe_microbench2 > gpurun_out/pipe_microbench2.txt 2>&1; cat gpurun_out/pipe_microbench2.txt
bash tools/gpujob_var.sh
