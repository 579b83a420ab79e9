mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_streaming.py -q -x -p no:cacheprovider 2>&1 | tail -3
PS_GEMM_NO_FEW_ROWS=1 timeout 600 python bench.py --workload cfg5 --steps 200 --warmup 5 > gpurun_out/r117_bench_cfg5_before.log 2>&1; tail -1 gpurun_out/r117_bench_cfg5_before.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg5 before', d['ms_per_step'], d['latency_ms'])"
timeout 600 python bench.py --workload cfg5 --steps 200 --warmup 5 > gpurun_out/r117_bench_cfg5.log 2>&1; tail -1 gpurun_out/r117_bench_cfg5.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg5 after', d['ms_per_step'], d['latency_ms'])"
echo done
