mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_ops.py -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py tests/test_gpu_streaming.py tests/test_gpu_graph.py -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r100_bench_cfg2.log 2>&1; tail -1 gpurun_out/r100_bench_cfg2.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg2', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['sequential_api_value'], d['roofline']['avg_launch_ms'])"
timeout 600 python profiles/gpurun_scripts/model_breakdown.py tse_skim_v0_causal > gpurun_out/r100_skim_breakdown.txt 2>&1; tail -22 gpurun_out/r100_skim_breakdown.txt | cut -c1-200
timeout 600 python bench.py --workload tse_skim_v0_causal --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r100_bench_skim.log 2>&1; tail -1 gpurun_out/r100_bench_skim.log | cut -c1-200
PS_CUDA_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_kernel" -s 2 -c 2 -o gpurun_out/r100_prof_lstm python profiles/gpurun_scripts/skim_once.py tse_skim_v0_causal 1 > gpurun_out/r100_ncu_lstm.log 2>&1; tail -2 gpurun_out/r100_ncu_lstm.log
ls -la gpurun_out/r100_prof_lstm.ncu-rep
echo done
