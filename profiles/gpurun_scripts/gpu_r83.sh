mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py -q -s -k "dpcrn" 2>&1 | grep -v "^$" | tail -12
timeout 900 python bench.py --workload ns_dpcrn_v0 --steps 10 --warmup 3 > gpurun_out/r83_bench_dpcrn.log 2>&1; tail -1 gpurun_out/r83_bench_dpcrn.log | cut -c1-250
