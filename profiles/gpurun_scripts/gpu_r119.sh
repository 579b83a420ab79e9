mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py tests/test_gpu_full.py tests/test_gpu_streaming.py -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r119_bench_cfg4.log 2>&1; tail -1 gpurun_out/r119_bench_cfg4.log | cut -c1-200
timeout 600 python profiles/gpurun_scripts/model_breakdown.py cfg4 > gpurun_out/r119_cfg4_breakdown.txt 2>&1; grep "gemm_simt" gpurun_out/r119_cfg4_breakdown.txt | cut -c1-80,150-215
echo done
