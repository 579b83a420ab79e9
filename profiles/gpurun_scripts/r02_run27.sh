# round 2, run 27: ncu --set full of the three kernels of a DPRNN pass (cfg3): gx GEMM (gemm_wide), lstm_tc, projection + LN (gemm_rows)
mkdir -p gpurun_out
PS_CUDA_GRAPH=0 timeout 600 python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run27_plain.log 2>&1 &&
PS_CUDA_GRAPH=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"lstm_tc|gemm_rows|gemm_wide" -s 42 -c 6 -o gpurun_out/r02_run27_cfg3 python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run27_ncu.log 2>&1; tail -3 gpurun_out/r02_run27_ncu.log
ls -la gpurun_out/*.ncu-rep
