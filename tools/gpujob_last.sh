#!/bin/bash
mkdir -p gpurun_out
VARNOTEST=1 VARCFGS="headline c2" VARLIBS="tsl-sdr_b200/libtslb200.so tsl-sdr_b200/libtslb200_A.so tsl-sdr_b200/libtslb200.so" bash tools/gpujob_var.sh
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
