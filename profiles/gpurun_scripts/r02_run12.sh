# round 2, run 12: persistent hop kernel (ps_stream_hop) - streaming tests, then cfg5 bench with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_streaming.py -x -q > gpurun_out/r02_run12_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run12_pytest.log; tail -12 gpurun_out/r02_run12_pytest.log
for v in 1 0; do
PS_STREAM_HOP=$v timeout 600 python bench.py --workload cfg5 --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run12_bench_cfg5_hop$v.json 2> gpurun_out/r02_run12_bench_cfg5_hop$v.err; tail -2 gpurun_out/r02_run12_bench_cfg5_hop$v.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_run12_bench_cfg5_hop$v.json"))
    print("PS_STREAM_HOP=$v", round(d["ms_per_step"],3), "ms/hop at 256 streams", d["latency_ms"], "launches", d["gpu_launches"], "frac", d["roofline"]["frac"])
except Exception as e: print("failed", e)
PY
done
