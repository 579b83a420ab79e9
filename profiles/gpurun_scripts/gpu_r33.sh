mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfE --tb=short -p no:cacheprovider -x > gpurun_out/r33_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r33_pytest_gpu.log
tail -5 gpurun_out/r33_pytest_gpu.log
for W in cfg2 cfg1 cfg3 cfg4; do
timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r33_bench_$W.log 2>&1; echo "$W: $(tail -1 gpurun_out/r33_bench_$W.log | cut -c1-330)"
done
