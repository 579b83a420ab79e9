mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -rfE --tb=short -p no:cacheprovider -k "lstm" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_full.py -q -rfE --tb=short -p no:cacheprovider -s -k "cfg3" 2>&1 | grep -E "cfg|passed|failed" | tail -4
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r24_bench_cfg3.log 2>&1; tail -1 gpurun_out/r24_bench_cfg3.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 70 --csv --log-file gpurun_out/r24_launches_cfg3.csv python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r24_ncu1.log 2>&1
