timeout 1200 python -m pytest tests/test_gpu_flex.py tests/test_gpu_pager.py -x -q 2>&1 | tail -15
