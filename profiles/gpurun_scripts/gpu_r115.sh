mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graph.py -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r115_bench_cfg2.log 2>&1; tail -1 gpurun_out/r115_bench_cfg2.log | cut -c1-200
for W in cfg1 cfg3 cfg4 cfg4_gated tse_unet_tcn_v0 ns_dpcrn_v0 ns_dparn_v0 tse_skim_v0_causal tse_skim_v2_causal; do
timeout 600 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/r115_bench_$W.log 2>&1; tail -1 gpurun_out/r115_bench_$W.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$W', round(d['ms_per_step'],2), round(d['value']), 'stream', round(d['e2e']['value']), 'seq', round(d['e2e']['sequential_api_value']))"
done
echo done
