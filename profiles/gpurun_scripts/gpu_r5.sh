mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_gemm_tc.py tests/test_gpu_modules.py -q -rfE --tb=short -p no:cacheprovider > gpurun_out/r5_units.log 2>&1; echo "units exit $?" >> gpurun_out/r5_units.log
tail -6 gpurun_out/r5_units.log
timeout 900 python -m pytest tests/test_gpu_full.py -q -rfE --tb=short -p no:cacheprovider -s -k "auto" > gpurun_out/r5_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/r5_full.log
grep -E "cfg|passed|failed|exit" gpurun_out/r5_full.log | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r5_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/r5_bench.log
tail -2 gpurun_out/r5_bench.log
PS_DWCONV_STREAMING=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r5_bench_dwstream.log 2>&1
tail -1 gpurun_out/r5_bench_dwstream.log
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r5_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 513 -c 342 --csv --log-file gpurun_out/r5_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r5_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|dwconv_tile" -s 100 -c 5 -o gpurun_out/r5_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r5_ncu2.log 2>&1
tail -2 gpurun_out/r5_ncu2.log
