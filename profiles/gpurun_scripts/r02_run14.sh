# round 2, run 14: programmatic dependent launch on the hot-path kernels - full GPU test-suite, then A/B on cfg1 / cfg2 / cfg4
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_run14_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run14_pytest.log; tail -6 gpurun_out/r02_run14_pytest.log
for w in cfg1 cfg4 cfg2; do for v in 0 1; do
steps=200; [ $w = cfg2 ] && steps=10; [ $w = cfg4 ] && steps=40
PS_PDL=$v python bench.py --workload $w --steps $steps --warmup 5 --no-cpu-baseline > gpurun_out/r02_run14_bench_${w}_pdl$v.json 2> gpurun_out/r02_run14_bench.err || tail -3 gpurun_out/r02_run14_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_run14_bench_${w}_pdl$v.json")); r=d["roofline"]
    print("$w PS_PDL=$v", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s; e2e", round(d["e2e"]["value"],1), "; top kernel", round(r["avg_launch_ms"],4), "ms; clocks", d["clocks"])
except Exception as e: print("$w $v failed", e)
PY
done; done
