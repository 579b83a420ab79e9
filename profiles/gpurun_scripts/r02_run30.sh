# round 2, run 30: few-tile policy as the default - whole GPU suite, cfg1 / cfg2 / cfg4 / cfg4_gated / cfg3 steps
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_run30_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run30_pytest.log; tail -6 gpurun_out/r02_run30_pytest.log
for w in cfg1 cfg4 cfg4_gated cfg2 tse_skim_v0_causal; do
steps=20; [ $w = cfg1 ] && steps=300; [ $w = cfg4 ] && steps=60
python bench.py --workload $w --steps $steps --warmup 5 --no-cpu-baseline > gpurun_out/r02_run30_bench_${w}.json 2> gpurun_out/r02_run30_bench.err || tail -3 gpurun_out/r02_run30_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run30_bench_${w}.json"))
print("$w", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s e2e", round(d["e2e"]["value"],1), d["clocks"])
PY
done
