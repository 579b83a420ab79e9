mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r17_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair_kernel" -s 100 -c 2 -o gpurun_out/r17_prof0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r17_ncu0.log 2>&1
PS_PAIR_DBG=5 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair_kernel" -s 100 -c 2 -o gpurun_out/r17_prof5 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r17_ncu5.log 2>&1
PS_PAIR_DBG=7 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair_kernel" -s 100 -c 2 -o gpurun_out/r17_prof7 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r17_ncu7.log 2>&1
tail -1 gpurun_out/r17_ncu7.log
