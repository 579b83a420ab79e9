mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_streaming.py -q -rfE --tb=short -p no:cacheprovider -s 2>&1 | grep -E "err|passed|failed|Error" | tail -8
timeout 600 python bench.py --workload cfg5 --steps 200 --warmup 5 > gpurun_out/r52_bench_cfg5.log 2>&1; tail -1 gpurun_out/r52_bench_cfg5.log | cut -c1-1100
