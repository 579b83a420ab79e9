mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q -rfE --tb=short -p no:cacheprovider -x > gpurun_out/r8_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/r8_tc.log
tail -25 gpurun_out/r8_tc.log
if grep -q "tc exit 0" gpurun_out/r8_tc.log; then
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py tests/test_gpu_streaming.py tests/test_gpu_full.py -q -rfE --tb=short -p no:cacheprovider -s > gpurun_out/r8_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/r8_full.log
grep -E "cfg|passed|failed|exit" gpurun_out/r8_full.log | tail -14
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench_pair.log 2>&1; echo "bench exit $?" >> gpurun_out/r8_bench_pair.log
tail -2 gpurun_out/r8_bench_pair.log | cut -c1-1600
PS_TC_KERNEL=single timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench_single.log 2>&1
tail -1 gpurun_out/r8_bench_single.log | cut -c1-300
fi
