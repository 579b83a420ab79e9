mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py -q -rfE --tb=short -p no:cacheprovider -s -k "skim" 2>&1 | tail -25
