timeout 600 python -m pytest tests/test_gpu_modules.py -q -x -k "sdr" 2>&1 | tail -12
