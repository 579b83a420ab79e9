mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_ops.py tests/test_gpu_gemm_tc.py -q -x -k "not verbose and not wrappers" -p no:cacheprovider > gpurun_out/r91_plain.log 2>&1 && \
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_modules.py tests/test_gpu_ops.py tests/test_gpu_gemm_tc.py -q -x -k "not verbose and not wrappers" -p no:cacheprovider > gpurun_out/r91_memcheck.log 2>&1
echo "exit $?"; tail -3 gpurun_out/r91_plain.log; grep -c "Invalid\|out of bounds" gpurun_out/r91_memcheck.log; tail -12 gpurun_out/r91_memcheck.log | cut -c1-200
