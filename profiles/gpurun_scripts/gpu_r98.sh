mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_ops.py tests/test_gpu_graph.py -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 1200 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py tests/test_gpu_streaming.py -q -x -p no:cacheprovider 2>&1 | tail -5
PS_LN_PAIR=0 timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r98_bench_cfg3_nopair.log 2>&1; tail -1 gpurun_out/r98_bench_cfg3_nopair.log | cut -c1-200
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r98_bench_cfg3.log 2>&1; tail -1 gpurun_out/r98_bench_cfg3.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r98_bench_cfg2.log 2>&1; tail -1 gpurun_out/r98_bench_cfg2.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg2', d['ms_per_step'], d['value'], d['e2e'], d['roofline']['avg_launch_ms'])"
PS_DW_LB4=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r98_bench_cfg2_lb4.log 2>&1; tail -1 gpurun_out/r98_bench_cfg2_lb4.log | cut -c1-200
PS_DW_TC=256 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r98_bench_cfg2_tc256.log 2>&1; tail -1 gpurun_out/r98_bench_cfg2_tc256.log | cut -c1-200
PS_DW_LB4=1 PS_DW_TC=256 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r98_bench_cfg2_lb4_tc256.log 2>&1; tail -1 gpurun_out/r98_bench_cfg2_lb4_tc256.log | cut -c1-200
timeout 600 python profiles/gpurun_scripts/model_breakdown.py tse_skim_v0_causal > gpurun_out/r98_skim_breakdown.txt 2>&1; tail -22 gpurun_out/r98_skim_breakdown.txt | cut -c1-200
timeout 600 python bench.py --workload tse_skim_v0_causal --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r98_bench_skim.log 2>&1; tail -1 gpurun_out/r98_bench_skim.log | cut -c1-200
echo done
