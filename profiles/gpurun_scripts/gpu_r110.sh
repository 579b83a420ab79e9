mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py -q -x -p no:cacheprovider -k "dprnn or skim or split or cfg3 or veve or full" 2>&1 | tail -3
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r110_bench_cfg3.log 2>&1; tail -1 gpurun_out/r110_bench_cfg3.log | cut -c1-200
timeout 600 python profiles/gpurun_scripts/model_breakdown.py cfg3 > gpurun_out/r110_cfg3_breakdown.txt 2>&1; tail -24 gpurun_out/r110_cfg3_breakdown.txt | cut -c1-80,150-215 | grep -i "segment\|merge"
echo done
