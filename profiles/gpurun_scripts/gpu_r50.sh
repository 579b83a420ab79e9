mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfE --tb=short -p no:cacheprovider > gpurun_out/r50_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r50_pytest_gpu.log
tail -3 gpurun_out/r50_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r50_bench_cfg2.log 2>&1; tail -1 gpurun_out/r50_bench_cfg2.log | cut -c1-300
for W in cfg1 cfg3 cfg4; do
timeout 600 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/r50_bench_$W.log 2>&1; echo "$W: $(tail -1 gpurun_out/r50_bench_$W.log | cut -c1-200)"
done
timeout 600 python bench.py --workload cfg5 --steps 200 --warmup 5 > gpurun_out/r50_bench_cfg5.log 2>&1; tail -1 gpurun_out/r50_bench_cfg5.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r50_bench_reference.log 2>&1; tail -1 gpurun_out/r50_bench_reference.log | cut -c1-200
PS_CUDA_GRAPH=0 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r50_plain.log 2>&1 && PS_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 180 --csv --log-file gpurun_out/r50_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r50_ncu1.log 2>&1
PS_CUDA_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair_kernel" -s 30 -c 3 -o gpurun_out/r50_prof_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r50_ncu2.log 2>&1
PS_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 70 --csv --log-file gpurun_out/r50_launches_cfg3.csv python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r50_ncu3.log 2>&1
PS_CUDA_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_tc_kernel" -s 4 -c 1 -o gpurun_out/r50_prof_lstm python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r50_ncu4.log 2>&1
echo done
