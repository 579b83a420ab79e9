python -m pytest tests/test_gpu_streaming.py tests/test_gpu_modules.py tests/test_gpu_full.py -q -k skim 2>&1 | tail -15
