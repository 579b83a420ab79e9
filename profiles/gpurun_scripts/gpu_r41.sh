mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py -q -rfE --tb=short -p no:cacheprovider 2>&1 | tail -4
PS_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 180 --csv --log-file gpurun_out/r41_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r41_ncu1.log 2>&1
