mkdir -p gpurun_out
timeout 600 python bench.py --workload cfg4_gated --steps 10 --warmup 3 > gpurun_out/r73_bench_cfg4_gated.log 2>&1; tail -1 gpurun_out/r73_bench_cfg4_gated.log | cut -c1-250
export PS_CUDA_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r73_launches_cfg4_gated.csv python bench.py --workload cfg4_gated --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r73_ncu.log 2>&1
tail -1 gpurun_out/r73_ncu.log | cut -c1-100
