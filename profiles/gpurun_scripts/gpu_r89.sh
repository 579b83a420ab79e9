mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py -q -x -k "dparn or attention" 2>&1 | tail -4
timeout 900 python bench.py --workload ns_dparn_v0 --steps 10 --warmup 3 > gpurun_out/r89_bench_dparn.log 2>&1; tail -1 gpurun_out/r89_bench_dparn.log | cut -c1-200
