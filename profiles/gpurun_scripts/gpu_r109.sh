mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r109_bench_cfg4.log 2>&1; tail -1 gpurun_out/r109_bench_cfg4.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r109_bench_cfg2.log 2>&1; tail -1 gpurun_out/r109_bench_cfg2.log | cut -c1-200
timeout 600 python profiles/gpurun_scripts/model_breakdown.py cfg4 > gpurun_out/r109_cfg4_breakdown.txt 2>&1; tail -24 gpurun_out/r109_cfg4_breakdown.txt | cut -c1-80,150-215 | head -10
echo done
