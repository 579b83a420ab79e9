mkdir -p gpurun_out
PS_CUDA_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dwconv_tile_kernel" -s 30 -c 3 -o gpurun_out/r121_prof_dwconv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r121_ncu.log 2>&1; tail -2 gpurun_out/r121_ncu.log
ls -la gpurun_out/r121_prof_dwconv.ncu-rep
echo done
