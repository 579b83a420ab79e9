# round 2, run 13: hop kernel after the latency trims - tests, phase timeline, cfg5 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_streaming.py -x -q > gpurun_out/r02_run13_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run13_pytest.log; tail -4 gpurun_out/r02_run13_pytest.log
PS_B200_LIB=$PWD/puresound_b200/libpuresound_b200_exp.so python profiles/hop_probe.py 2>&1 | grep -v "^S=.*timeline -" | tail -18 | tee gpurun_out/r02_run13_hop_timeline.txt
timeout 600 python bench.py --workload cfg5 --steps 100 --warmup 3 > gpurun_out/r02_run13_bench_cfg5.json 2> gpurun_out/r02_run13_bench_cfg5.err; tail -2 gpurun_out/r02_run13_bench_cfg5.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run13_bench_cfg5.json"))
print(round(d["ms_per_step"],3), "ms/hop at 256 streams", d["latency_ms"], "launches", d["gpu_launches"], "frac", d["roofline"]["frac"], d["cpu_baseline"])
PY
