mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rfE --tb=short -p no:cacheprovider > gpurun_out/r70_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r70_pytest_gpu.log
tail -3 gpurun_out/r70_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r70_bench_cfg2.log 2>&1; tail -1 gpurun_out/r70_bench_cfg2.log | cut -c1-300
for W in cfg1 cfg3 cfg4; do
timeout 600 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/r70_bench_$W.log 2>&1; echo "$W: $(tail -1 gpurun_out/r70_bench_$W.log | cut -c1-200)"
done
timeout 600 python bench.py --workload cfg5 --steps 200 --warmup 5 > gpurun_out/r70_bench_cfg5.log 2>&1; tail -1 gpurun_out/r70_bench_cfg5.log | cut -c1-200
echo done
