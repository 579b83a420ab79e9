mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -p no:cacheprovider -k "dwconv" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r37_bench.log 2>&1; tail -1 gpurun_out/r37_bench.log | cut -c1-200
PS_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 180 --csv --log-file gpurun_out/r37_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r37_ncu1.log 2>&1
