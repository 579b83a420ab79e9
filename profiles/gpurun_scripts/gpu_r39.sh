mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -rfE --tb=short -p no:cacheprovider 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_full.py tests/test_gpu_modules.py -q -rfE --tb=short -p no:cacheprovider -s 2>&1 | grep -E "cfg|passed|failed|Error" | tail -12
for W in cfg2 cfg3 cfg4; do
timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r39_bench_$W.log 2>&1; echo "$W: $(tail -1 gpurun_out/r39_bench_$W.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["value"], j["ms_per_step"], j["e2e"]["value"], j["clocks"])')"
done
