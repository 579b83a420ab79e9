mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rfE --tb=short -p no:cacheprovider > gpurun_out/r86_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r86_pytest_gpu.log
tail -3 gpurun_out/r86_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r86_smoke.log 2>&1; tail -1 gpurun_out/r86_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r86_bench_cfg2.log 2>&1; tail -1 gpurun_out/r86_bench_cfg2.log | cut -c1-200
for W in cfg1 cfg3 cfg4 cfg4_gated tse_unet_tcn_v0 ns_dpcrn_v0; do
timeout 600 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/r86_bench_$W.log 2>&1; echo "$W: $(tail -1 gpurun_out/r86_bench_$W.log | cut -c1-160)"
done
timeout 600 python bench.py --workload cfg5 --steps 200 --warmup 5 > gpurun_out/r86_bench_cfg5.log 2>&1; tail -1 gpurun_out/r86_bench_cfg5.log | cut -c1-160
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r86_bench_reference.log 2>&1; tail -1 gpurun_out/r86_bench_reference.log | cut -c1-200
PS_CUDA_GRAPH=0 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r86_plain.log 2>&1 && PS_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 180 --csv --log-file gpurun_out/r86_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r86_ncu1.log 2>&1
echo done
