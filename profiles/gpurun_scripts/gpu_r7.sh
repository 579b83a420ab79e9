mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_gemm_tc.py tests/test_gpu_modules.py tests/test_gpu_streaming.py -q -rfE --tb=short -p no:cacheprovider -x > gpurun_out/r7_units.log 2>&1; echo "units exit $?" >> gpurun_out/r7_units.log
tail -6 gpurun_out/r7_units.log
timeout 900 python -m pytest tests/test_gpu_full.py -q -rfE --tb=short -p no:cacheprovider -s > gpurun_out/r7_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/r7_full.log
grep -E "cfg|passed|failed|exit" gpurun_out/r7_full.log | tail -12
for BK in 32 64; do
PS_TC_BK=$BK timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench_bk$BK.log 2>&1; echo "bench exit $?" >> gpurun_out/r7_bench_bk$BK.log
tail -2 gpurun_out/r7_bench_bk$BK.log | cut -c1-1500
done
timeout 600 python bench.py --workload cfg5 --steps 200 --warmup 5 > gpurun_out/r7_bench_cfg5.log 2>&1; tail -2 gpurun_out/r7_bench_cfg5.log | cut -c1-1200
timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r7_bench_cfg1.log 2>&1; tail -1 gpurun_out/r7_bench_cfg1.log | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 100 -c 3 -o gpurun_out/r7_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r7_ncu2.log 2>&1
tail -2 gpurun_out/r7_ncu2.log
