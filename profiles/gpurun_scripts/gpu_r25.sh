mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_ops.py -q -rfE --tb=short -p no:cacheprovider 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_full.py tests/test_gpu_modules.py -q -rfE --tb=short -p no:cacheprovider -s 2>&1 | grep -E "cfg|passed|failed|Error" | tail -12
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r25_bench_cfg3.log 2>&1; tail -1 gpurun_out/r25_bench_cfg3.log | cut -c1-200
