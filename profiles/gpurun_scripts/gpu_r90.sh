timeout 900 python -m pytest tests/test_gpu_modules.py -q -k "verbose" 2>&1 | tail -30
