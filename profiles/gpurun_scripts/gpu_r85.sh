mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py -q -s -k "unet or dpcrn" 2>&1 | grep -v "^$" | tail -12
for W in tse_unet_tcn_v0 ns_dpcrn_v0; do
timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r85_bench_$W.log 2>&1; echo "$W: $(tail -1 gpurun_out/r85_bench_$W.log | cut -c1-160)"
done
python profiles/gpurun_scripts/model_breakdown.py tse_unet_tcn_v0 2>&1 | grep -v Warn | cut -c1-72,120-200 | tail -16 | tee gpurun_out/r85_unet_breakdown.txt
