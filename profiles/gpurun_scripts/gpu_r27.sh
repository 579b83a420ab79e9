mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_tc_kernel" -s 4 -c 1 -o gpurun_out/r27_prof_lstm python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r27_ncu.log 2>&1
tail -1 gpurun_out/r27_ncu.log | cut -c1-100
