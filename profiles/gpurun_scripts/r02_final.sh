# round 2 evidence call (one box, one GPU): tests, smoke, every bench workload, reference arm, launch list of a cfg2 step,
# ncu --set full of the dominant kernels.  R = run tag.
R=${R:-r02_final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfE --tb=short -p no:cacheprovider > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${R}_pytest_gpu.log
tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_cfg2.json 2> gpurun_out/${R}_bench_cfg2.err; tail -1 gpurun_out/${R}_bench_cfg2.json | cut -c1-220
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err; tail -1 gpurun_out/${R}_bench_reference.json | cut -c1-220
for W in cfg1 cfg3 cfg4; do
steps=20; [ $W = cfg1 ] && steps=300; [ $W = cfg4 ] && steps=60
timeout 900 python bench.py --workload $W --steps $steps --warmup 5 > gpurun_out/${R}_bench_$W.json 2> gpurun_out/${R}_bench_$W.err; echo "$W: $(tail -1 gpurun_out/${R}_bench_$W.json | cut -c1-200)"
done
for W in cfg1b cfg4_gated tse_unet_tcn_v0 ns_dpcrn_v0 ns_dparn_v0 tse_skim_v0_causal tse_skim_v2_causal; do
timeout 900 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_$W.json 2> gpurun_out/${R}_bench_$W.err; echo "$W: $(tail -1 gpurun_out/${R}_bench_$W.json | cut -c1-200)"
done
timeout 900 python bench.py --workload cfg5 --steps 200 --warmup 5 > gpurun_out/${R}_bench_cfg5.json 2> gpurun_out/${R}_bench_cfg5.err; tail -1 gpurun_out/${R}_bench_cfg5.json | cut -c1-200
PS_CUDA_GRAPH=0 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_plain.log 2>&1 &&
PS_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/${R}_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu1.log 2>&1
PS_CUDA_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_wide|dwconv_tma" -s 120 -c 4 -o gpurun_out/${R}_prof_cfg2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu2.log 2>&1
ls -la gpurun_out/${R}_prof_cfg2.ncu-rep
echo done
