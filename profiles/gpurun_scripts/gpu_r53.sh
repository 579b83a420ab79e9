mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_streaming.py tests/test_gpu_modules.py -q -rfE --tb=short -p no:cacheprovider 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_full.py -q -rfE --tb=short -p no:cacheprovider -s 2>&1 | grep -E "cfg|passed|failed|Error" | tail -10
timeout 600 python bench.py --workload cfg5 --steps 200 --warmup 5 > gpurun_out/r53_bench_cfg5.log 2>&1; tail -1 gpurun_out/r53_bench_cfg5.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("cfg5", round(j["value"]), j["ms_per_step"], j["latency_ms"])'
for W in cfg1 cfg4 cfg2; do
timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r53_bench_$W.log 2>&1; echo "$W: $(tail -1 gpurun_out/r53_bench_$W.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["value"]), round(j["ms_per_step"],3), round(j["e2e"]["value"]), j["clocks"]["sm_mhz"])')"
done
