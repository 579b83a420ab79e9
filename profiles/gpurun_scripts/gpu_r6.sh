mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_gemm_tc.py tests/test_gpu_modules.py -q -rfE --tb=short -p no:cacheprovider > gpurun_out/r6_units.log 2>&1; echo "units exit $?" >> gpurun_out/r6_units.log
tail -6 gpurun_out/r6_units.log
timeout 900 python -m pytest tests/test_gpu_full.py -q -rfE --tb=short -p no:cacheprovider -s -k "auto" > gpurun_out/r6_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/r6_full.log
grep -E "cfg|passed|failed|exit" gpurun_out/r6_full.log | tail -8
for BK in 32 64; do
PS_TC_BK=$BK timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench_bk$BK.log 2>&1; echo "bench exit $?" >> gpurun_out/r6_bench_bk$BK.log
tail -2 gpurun_out/r6_bench_bk$BK.log
done
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r6_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 513 -c 342 --csv --log-file gpurun_out/r6_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r6_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|dwconv_tile" -s 100 -c 5 -o gpurun_out/r6_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r6_ncu2.log 2>&1
tail -2 gpurun_out/r6_ncu2.log
timeout 600 python bench.py --workload cfg5 --steps 200 --warmup 5 > gpurun_out/r6_bench_cfg5.log 2>&1; tail -2 gpurun_out/r6_bench_cfg5.log
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 > gpurun_out/r6_bench_cfg3.log 2>&1; tail -2 gpurun_out/r6_bench_cfg3.log
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/r6_bench_cfg4.log 2>&1; tail -2 gpurun_out/r6_bench_cfg4.log
timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 5 > gpurun_out/r6_bench_cfg1.log 2>&1; tail -2 gpurun_out/r6_bench_cfg1.log
