mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfE --tb=short -p no:cacheprovider -x > gpurun_out/r20_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r20_pytest_gpu.log
tail -4 gpurun_out/r20_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r20_bench.log 2>&1; tail -1 gpurun_out/r20_bench.log | cut -c1-2000
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r20_bench_cfg3.log 2>&1; tail -1 gpurun_out/r20_bench_cfg3.log | cut -c1-400
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r20_bench_cfg4.log 2>&1; tail -1 gpurun_out/r20_bench_cfg4.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 900 --csv --log-file gpurun_out/r20_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r20_ncu1.log 2>&1
