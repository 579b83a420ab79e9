mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py -q -k "unet" 2>&1 | tail -3
timeout 900 python bench.py --workload tse_unet_tcn_v0 --steps 10 --warmup 3 > gpurun_out/r80_bench_unet.log 2>&1; tail -1 gpurun_out/r80_bench_unet.log | cut -c1-300
