mkdir -p gpurun_out
for F in 1 0 1 0; do
for W in cfg2 cfg1 cfg4; do
PS_FUSE_FINALIZE=$F timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r43_bench_${W}_$F.log 2>&1; echo "fuse=$F $W: $(tail -1 gpurun_out/r43_bench_${W}_$F.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["value"]), round(j["ms_per_step"],3), j["gpu_launches"], j["clocks"]["sm_mhz"])')"
done; done
