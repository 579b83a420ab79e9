mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -p no:cacheprovider -k lstm 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py tests/test_gpu_streaming.py tests/test_gpu_graph.py -q -x -p no:cacheprovider -k "skim or mel or rnn or stream" 2>&1 | tail -3
timeout 600 python bench.py --workload tse_skim_v0_causal --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r106_bench_skim.log 2>&1; tail -1 gpurun_out/r106_bench_skim.log | cut -c1-200
timeout 600 python bench.py --workload tse_skim_v2_causal --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r106_bench_skim_v2.log 2>&1; tail -1 gpurun_out/r106_bench_skim_v2.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('skim_v2', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['sequential_api_value'])"
timeout 600 python profiles/gpurun_scripts/model_breakdown.py tse_skim_v0_causal > gpurun_out/r106_skim_breakdown.txt 2>&1; tail -22 gpurun_out/r106_skim_breakdown.txt | cut -c1-200 | head -5
echo done
