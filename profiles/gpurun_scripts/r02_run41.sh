# round 2, run 41: wide GEMM epilogue with the residual requested one chunk ahead (out_conv form) - tests, isolated probe, cfg2 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_full.py -q -x > gpurun_out/r02_run41_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_run41_pytest.log
PROBE_REPS=30 python profiles/gemm_probe.py 2>&1 | tail -1
PROBE_REPS=30 PROBE_B=8 python profiles/gemm_probe.py 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run41_bench_cfg2.json 2> gpurun_out/r02_run41_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run41_bench_cfg2.json")); r=d["roofline"]
print("cfg2", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["clocks"])
for o in [r]+r["other_kernels"]:
    print("    ", o["kernel"][:80], o["bound"], "frac", round(o["frac"],3), round(o["avg_launch_ms"],4), "ms share", round(o["share_of_step"],3))
PY
