mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r46_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r46_smoke.log; tail -5 gpurun_out/r46_smoke.log
