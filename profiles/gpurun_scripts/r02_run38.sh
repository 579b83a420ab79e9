# round 2, run 38: compute-sanitizer memcheck over the kernels added this round (few-channel GEMM incl. streamed weights, few-tile
# path, TMA depthwise conv, statistics merge) - plain run first
mkdir -p gpurun_out
T="tests/test_gpu_gemm_tc.py -k rows_kernel_or_few_tiles"
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -x -k "rows_kernel or few_tiles" > gpurun_out/r02_run38_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_gemm_tc.py -q -x -k "rows_kernel or few_tiles" > gpurun_out/r02_run38_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r02_run38_memcheck.log
