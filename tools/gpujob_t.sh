timeout 1200 python -m pytest tests/test_gpu_pager.py -x -q -k "mueller" 2>&1 | tail -15
