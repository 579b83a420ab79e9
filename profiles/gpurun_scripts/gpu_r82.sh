mkdir -p gpurun_out
for W in cfg4 cfg4_gated tse_unet_tcn_v0; do
timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r82_bench_$W.log 2>&1; echo "$W: $(tail -1 gpurun_out/r82_bench_$W.log | cut -c1-160)"
done
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
