timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "8bit" 2>&1 | tail -15
