mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 200 --csv --log-file gpurun_out/r22_launches_cfg3.csv python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r22_ncu1.log 2>&1
tail -1 gpurun_out/r22_ncu1.log | cut -c1-100
