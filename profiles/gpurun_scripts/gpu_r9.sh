mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r9_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair_kernel" -s 100 -c 3 -o gpurun_out/r9_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r9_ncu2.log 2>&1
tail -2 gpurun_out/r9_ncu2.log
