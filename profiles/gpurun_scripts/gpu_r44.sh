mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -rfE --tb=short -p no:cacheprovider 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_full.py tests/test_gpu_modules.py -q -rfE --tb=short -p no:cacheprovider -s -k "cfg3 or dprnn or DPRNN" 2>&1 | grep -E "cfg|passed|failed|Error" | tail -6
for i in 1 2; do
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r44_bench_cfg3.log 2>&1; echo "cfg3: $(tail -1 gpurun_out/r44_bench_cfg3.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["value"]), round(j["ms_per_step"],3), j["gpu_launches"], j["clocks"]["sm_mhz"])')"
done
PS_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 70 --csv --log-file gpurun_out/r44_launches_cfg3.csv python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r44_ncu1.log 2>&1
