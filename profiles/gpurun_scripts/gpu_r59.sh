python -m pytest tests/test_gpu_ops.py -q -k "lstm" 2>&1 | tail -15
