# round 2, run 1: phase-A validation (hygiene fixes, new tests, reference arm, new bench keys)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > gpurun_out/r02_run1_gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r02_run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run1_pytest.log
tail -5 gpurun_out/r02_run1_pytest.log
python __graft_entry__.py --smoke > gpurun_out/r02_run1_smoke.log 2>&1; tail -3 gpurun_out/r02_run1_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_run1_bench_cfg2.json 2> gpurun_out/r02_run1_bench_cfg2.err; tail -c 600 gpurun_out/r02_run1_bench_cfg2.json
python bench.py --steps 10 --warmup 3 --workload cfg3 > gpurun_out/r02_run1_bench_cfg3.json 2> gpurun_out/r02_run1_bench_cfg3.err
python bench.py --steps 10 --warmup 3 --workload cfg1b > gpurun_out/r02_run1_bench_cfg1b.json 2> gpurun_out/r02_run1_bench_cfg1b.err
python bench.py --steps 50 --warmup 3 --workload cfg5 > gpurun_out/r02_run1_bench_cfg5.json 2> gpurun_out/r02_run1_bench_cfg5.err
python bench.py --steps 20 --warmup 3 --workload cfg1 --no-cpu-baseline > gpurun_out/r02_run1_bench_cfg1.json 2> gpurun_out/r02_run1_bench_cfg1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_run1_bench_reference.json 2> gpurun_out/r02_run1_bench_reference.err
for f in cfg3 cfg1b cfg5 cfg1 reference; do echo "== $f"; cut -c1-330 gpurun_out/r02_run1_bench_$f.json; tail -2 gpurun_out/r02_run1_bench_$f.err; done
