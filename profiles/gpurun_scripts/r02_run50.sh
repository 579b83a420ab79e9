# round 2, run 50: last call on the final tree - the GPU suite, smoke, the default bench line and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfE --tb=short -p no:cacheprovider > gpurun_out/r02_run50_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_run50_pytest_gpu.log; tail -3 gpurun_out/r02_run50_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_run50_smoke.log 2>&1; tail -1 gpurun_out/r02_run50_smoke.log
timeout 900 python bench.py > gpurun_out/r02_run50_bench_cfg2.json 2> gpurun_out/r02_run50_bench_cfg2.err; tail -1 gpurun_out/r02_run50_bench_cfg2.json | cut -c1-200
timeout 900 python bench.py --impl reference > gpurun_out/r02_run50_bench_reference.json 2> gpurun_out/r02_run50_bench_reference.err; tail -1 gpurun_out/r02_run50_bench_reference.json | cut -c1-200
