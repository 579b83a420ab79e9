# final evidence of the round: tests, smoke, every bench workload, reference arm, launch list, one ncu --set full capture
R=${R:-r99}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfE --tb=short -p no:cacheprovider > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${R}_pytest_gpu.log
tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_cfg2.log 2>&1; tail -1 gpurun_out/${R}_bench_cfg2.log | cut -c1-200
for W in cfg1 cfg3 cfg4 cfg4_gated tse_unet_tcn_v0 ns_dpcrn_v0 ns_dparn_v0 tse_skim_v0_causal tse_skim_v2_causal; do
timeout 600 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/${R}_bench_$W.log 2>&1; echo "$W: $(tail -1 gpurun_out/${R}_bench_$W.log | cut -c1-160)"
done
timeout 600 python bench.py --workload cfg5 --steps 200 --warmup 5 > gpurun_out/${R}_bench_cfg5.log 2>&1; tail -1 gpurun_out/${R}_bench_cfg5.log | cut -c1-160
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.log 2>&1; tail -1 gpurun_out/${R}_bench_reference.log | cut -c1-200
PS_CUDA_GRAPH=0 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_plain.log 2>&1 && PS_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 180 --csv --log-file gpurun_out/${R}_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu1.log 2>&1
PS_CUDA_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair_kernel" -s 100 -c 3 -o gpurun_out/${R}_prof_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu2.log 2>&1
ls -la gpurun_out/${R}_prof_gemm.ncu-rep
echo done
