mkdir -p gpurun_out
for BK in 32 64; do
  PS_TC_BK=$BK timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q -rfE --tb=short -p no:cacheprovider > gpurun_out/r4_tc_bk$BK.log 2>&1; echo "tc bk$BK exit $?" >> gpurun_out/r4_tc_bk$BK.log
  tail -4 gpurun_out/r4_tc_bk$BK.log
  PS_TC_BK=$BK timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r4_bench_bk$BK.log 2>&1; echo "bench exit $?" >> gpurun_out/r4_bench_bk$BK.log
  tail -2 gpurun_out/r4_bench_bk$BK.log
done
timeout 900 python -m pytest tests/test_gpu_full.py -q -rfE --tb=short -p no:cacheprovider -s -k "auto" > gpurun_out/r4_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/r4_full.log
grep -E "cfg|passed|failed|exit" gpurun_out/r4_full.log | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 80 -c 3 -o gpurun_out/r4_prof_gemm_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r4_ncu2.log 2>&1
tail -2 gpurun_out/r4_ncu2.log
