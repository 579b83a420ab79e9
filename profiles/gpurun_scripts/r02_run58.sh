# round 2, run 58: the whole GPU suite and smoke on the final tree (after the U-Net tap-buffer change)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfE --tb=short -p no:cacheprovider > gpurun_out/r02_run58_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_run58_pytest_gpu.log; tail -3 gpurun_out/r02_run58_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_run58_smoke.log 2>&1; tail -1 gpurun_out/r02_run58_smoke.log
