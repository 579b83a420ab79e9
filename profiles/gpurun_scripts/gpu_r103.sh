mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -p no:cacheprovider -k lstm 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py tests/test_gpu_streaming.py -q -x -p no:cacheprovider -k "skim or mel or rnn" 2>&1 | tail -3
PS_LSTM_SPT16=1 timeout 600 python bench.py --workload tse_skim_v0_causal --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r103_bench_skim_spt16.log 2>&1; tail -1 gpurun_out/r103_bench_skim_spt16.log | cut -c1-200
timeout 600 python bench.py --workload tse_skim_v0_causal --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r103_bench_skim.log 2>&1; tail -1 gpurun_out/r103_bench_skim.log | cut -c1-200
timeout 600 python profiles/gpurun_scripts/model_breakdown.py tse_skim_v0_causal > gpurun_out/r103_skim_breakdown.txt 2>&1; tail -22 gpurun_out/r103_skim_breakdown.txt | cut -c1-200 | head -6
echo done
