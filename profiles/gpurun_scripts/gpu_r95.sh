mkdir -p gpurun_out
timeout 600 python profiles/gpurun_scripts/model_breakdown.py tse_skim_v0_causal > gpurun_out/r95_skim_breakdown.txt 2>&1; tail -24 gpurun_out/r95_skim_breakdown.txt | cut -c1-200
timeout 600 python bench.py --workload tse_skim_v0_causal --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r95_bench_skim.log 2>&1; tail -1 gpurun_out/r95_bench_skim.log | cut -c1-200
echo done
