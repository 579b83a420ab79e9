mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q -rfE --tb=short -p no:cacheprovider -x > gpurun_out/r10_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/r10_tc.log
tail -5 gpurun_out/r10_tc.log
if grep -q "tc exit 0" gpurun_out/r10_tc.log; then
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r10_bench_pair.log 2>&1; echo "bench exit $?" >> gpurun_out/r10_bench_pair.log
tail -2 gpurun_out/r10_bench_pair.log | cut -c1-1600
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair_kernel" -s 100 -c 3 -o gpurun_out/r10_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r10_ncu2.log 2>&1
tail -2 gpurun_out/r10_ncu2.log
fi
