mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -q -rfE --tb=short -p no:cacheprovider -k "verbose" 2>&1 | tail -15
