mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py tests/test_gpu_streaming.py -q -x -p no:cacheprovider 2>&1 | tail -5
PS_LN_PAIR=0 timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r96_bench_cfg3_nopair.log 2>&1; tail -1 gpurun_out/r96_bench_cfg3_nopair.log | cut -c1-200
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r96_bench_cfg3.log 2>&1; tail -1 gpurun_out/r96_bench_cfg3.log | cut -c1-200
timeout 600 python profiles/gpurun_scripts/model_breakdown.py cfg3 > gpurun_out/r96_cfg3_breakdown.txt 2>&1; tail -20 gpurun_out/r96_cfg3_breakdown.txt | cut -c1-180
echo done
