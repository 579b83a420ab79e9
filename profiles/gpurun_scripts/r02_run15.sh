# round 2, run 15: 256-channel tile split for few-frame GEMMs (batch 1) - tests + cfg1 / cfg4 A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_run15_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run15_pytest.log; tail -4 gpurun_out/r02_run15_pytest.log
for w in cfg1 cfg4; do for v in 0 1; do
steps=300; [ $w = cfg4 ] && steps=60
PS_PAIR_SPLIT=$v python bench.py --workload $w --steps $steps --warmup 5 --no-cpu-baseline > gpurun_out/r02_run15_bench_${w}_split$v.json 2> gpurun_out/r02_run15_bench.err || tail -3 gpurun_out/r02_run15_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_run15_bench_${w}_split$v.json")); r=d["roofline"]
    print("$w PS_PAIR_SPLIT=$v", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s; e2e", round(d["e2e"]["value"],1), "; top kernel", r["kernel"][:40], round(r["avg_launch_ms"],4), "ms; clocks", d["clocks"])
except Exception as e: print("$w $v failed", e)
PY
done; done
