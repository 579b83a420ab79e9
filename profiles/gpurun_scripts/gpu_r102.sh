mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -p no:cacheprovider -k lstm 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py tests/test_gpu_streaming.py -q -x -p no:cacheprovider -k "skim or mel or rnn or verbose" 2>&1 | tail -5
timeout 600 python profiles/gpurun_scripts/model_breakdown.py tse_skim_v0_causal > gpurun_out/r102_skim_breakdown.txt 2>&1; tail -22 gpurun_out/r102_skim_breakdown.txt | cut -c1-200 | head -8
timeout 600 python bench.py --workload tse_skim_v0_causal --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r102_bench_skim.log 2>&1; tail -1 gpurun_out/r102_bench_skim.log | cut -c1-200
PS_CUDA_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_kernel" -s 2 -c 2 -o gpurun_out/r102_prof_lstm python profiles/gpurun_scripts/skim_once.py tse_skim_v0_causal 1 > gpurun_out/r102_ncu_lstm.log 2>&1; tail -2 gpurun_out/r102_ncu_lstm.log
echo done
