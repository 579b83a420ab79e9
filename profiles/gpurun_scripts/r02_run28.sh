# round 2, run 28: LSTM gx loads requested two chunks ahead across the step boundary - whole GPU suite, cfg3 / veve / dpcrn benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_run28_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run28_pytest.log; tail -6 gpurun_out/r02_run28_pytest.log
for w in cfg3 ns_dpcrn_v0; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run28_bench_${w}.json 2> gpurun_out/r02_run28_bench.err || tail -3 gpurun_out/r02_run28_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_run28_bench_${w}.json")); r=d["roofline"]
    print("$w", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["clocks"]["sm_mhz"])
    for o in ([r]+r["other_kernels"]) if r else []:
        print("    ", o["kernel"][:80], o["bound"], "frac", round(o["frac"],3), round(o["avg_launch_ms"],4), "ms share", round(o["share_of_step"],3))
except Exception as e: print("$w failed", e)
PY
done
