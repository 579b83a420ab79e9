mkdir -p gpurun_out
PS_CUDA_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dwconv_tile_kernel" -s 30 -c 3 -o gpurun_out/r49_prof_dw python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r49_ncu.log 2>&1
tail -1 gpurun_out/r49_ncu.log | cut -c1-100
