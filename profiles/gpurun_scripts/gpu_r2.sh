mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q -rfE --tb=short -p no:cacheprovider > gpurun_out/r2_tc.log 2>&1; TC=$?
echo "tc exit $TC" >> gpurun_out/r2_tc.log
tail -15 gpurun_out/r2_tc.log
if [ $TC -eq 0 ]; then
  timeout 1200 python -m pytest tests -m gpu -q -rfE --tb=short -p no:cacheprovider -s > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu.log
  grep -E "cfg|passed|failed|exit" gpurun_out/r2_pytest_gpu.log | tail -20
  timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_auto.log 2>&1; echo "bench exit $?" >> gpurun_out/r2_bench_auto.log
  tail -2 gpurun_out/r2_bench_auto.log
  timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 513 -c 342 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu1.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 80 -c 2 -o gpurun_out/r2_prof_gemm_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu2.log 2>&1
  tail -3 gpurun_out/r2_ncu2.log
else
  timeout 600 python bench.py --steps 3 --warmup 3 --gemm-backend simt > gpurun_out/r2_bench_simt.log 2>&1
fi
ls -la gpurun_out | tail -12
