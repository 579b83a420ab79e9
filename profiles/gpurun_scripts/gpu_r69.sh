mkdir -p gpurun_out
export PS_CUDA_GRAPH=0
timeout 300 python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 70 --csv --log-file gpurun_out/r69_launches_cfg3.csv python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r69_ncu.log 2>&1
tail -1 gpurun_out/r69_ncu.log | cut -c1-200
