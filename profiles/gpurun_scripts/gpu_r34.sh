mkdir -p gpurun_out
PS_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/r34_launches_cfg4.csv python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r34_ncu1.log 2>&1
PS_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 180 --csv --log-file gpurun_out/r34_launches_cfg1.csv python bench.py --workload cfg1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r34_ncu2.log 2>&1
