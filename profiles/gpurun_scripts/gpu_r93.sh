mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 300 python profiles/gpurun_scripts/enc_bench.py 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_full.py tests/test_gpu_modules.py -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r93_bench_cfg2.log 2>&1; tail -1 gpurun_out/r93_bench_cfg2.log | cut -c1-300
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r93_bench_cfg3.log 2>&1; tail -1 gpurun_out/r93_bench_cfg3.log | cut -c1-200
echo done
