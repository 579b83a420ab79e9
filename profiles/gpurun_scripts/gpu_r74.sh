timeout 900 python -m pytest tests/test_gpu_modules.py -q -x -k "unet" 2>&1 | tail -30
