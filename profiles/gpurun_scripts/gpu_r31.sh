mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 180 --csv --log-file gpurun_out/r31_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r31_ncu1.log 2>&1
tail -1 gpurun_out/r31_ncu1.log | cut -c1-100
