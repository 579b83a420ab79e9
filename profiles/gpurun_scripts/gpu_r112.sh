mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graph.py -q -x -p no:cacheprovider 2>&1 | tail -3
for W in cfg3 tse_skim_v0_causal cfg4_gated cfg2 cfg1; do
timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r112_bench_$W.log 2>&1; tail -1 gpurun_out/r112_bench_$W.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$W', round(d['ms_per_step'],2), round(d['value']), 'stream', round(d['e2e']['value']), 'seq', round(d['e2e']['sequential_api_value']))"
done
echo done
