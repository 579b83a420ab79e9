mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -rfE --tb=short -p no:cacheprovider -k "lstm" > gpurun_out/r21_lstm.log 2>&1; echo "lstm exit $?" >> gpurun_out/r21_lstm.log
tail -30 gpurun_out/r21_lstm.log
if grep -q "lstm exit 0" gpurun_out/r21_lstm.log; then
timeout 600 python -m pytest tests/test_gpu_full.py tests/test_gpu_modules.py -q -rfE --tb=short -p no:cacheprovider -s -k "cfg3 or dprnn or DPRNN" 2>&1 | grep -E "cfg|passed|failed" | tail -6
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r21_bench_cfg3.log 2>&1; tail -1 gpurun_out/r21_bench_cfg3.log | cut -c1-300
fi
