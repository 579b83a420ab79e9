# round 2, run 18: 64-k stages for one-block GEMMs (PS_PAIR_SUB) + cached signature walk - tests, A/B on cfg3 / cfg1b / cfg4, cfg1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_run18_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run18_pytest.log; tail -4 gpurun_out/r02_run18_pytest.log
for w in cfg3 cfg1b cfg4; do for v in 1 2; do
steps=10; [ $w = cfg4 ] && steps=60
PS_PAIR_SUB=$v python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/r02_run18_bench_${w}_sub$v.json 2> gpurun_out/r02_run18_bench.err || tail -3 gpurun_out/r02_run18_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_run18_bench_${w}_sub$v.json")); r=d["roofline"]
    print("$w PS_PAIR_SUB=$v", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["clocks"]["sm_mhz"])
    for o in [r]+r["other_kernels"]:
        print("    ", o["kernel"][:80], o["bound"], "frac", round(o["frac"],3), round(o["avg_launch_ms"],4), "ms share", round(o["share_of_step"],3))
except Exception as e: print("$w $v failed", e)
PY
done; done
python bench.py --workload cfg1 --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run18_bench_cfg1.json 2> gpurun_out/r02_run18_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run18_bench_cfg1.json"))
print("cfg1", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s; e2e", round(d["e2e"]["value"],1), d["clocks"])
PY
