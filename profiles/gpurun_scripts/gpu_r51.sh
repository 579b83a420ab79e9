mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/r51_launches_cfg5.csv python bench.py --workload cfg5 --steps 50 --warmup 5 > gpurun_out/r51_ncu1.log 2>&1
tail -2 gpurun_out/r51_ncu1.log | cut -c1-300
