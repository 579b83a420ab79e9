mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_streaming.py -q -rfE --tb=short -p no:cacheprovider -s > gpurun_out/r3_tc.log 2>&1; TC=$?
echo "tc exit $TC" >> gpurun_out/r3_tc.log
tail -25 gpurun_out/r3_tc.log
timeout 1200 python -m pytest tests/test_gpu_full.py -q -rfE --tb=short -p no:cacheprovider -s -k "auto" > gpurun_out/r3_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3_full.log
grep -E "cfg|passed|failed|exit" gpurun_out/r3_full.log | tail -12
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench_auto.log 2>&1; echo "bench exit $?" >> gpurun_out/r3_bench_auto.log
tail -2 gpurun_out/r3_bench_auto.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 80 -c 3 -o gpurun_out/r3_prof_gemm_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r3_ncu2.log 2>&1
tail -3 gpurun_out/r3_ncu2.log
