# round 2, run 33: statistics merge on raw fp64 moments (no divisions, one barrier) - whole GPU suite, cfg1 / cfg4 / cfg2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_run33_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run33_pytest.log; tail -5 gpurun_out/r02_run33_pytest.log
for w in cfg1 cfg4 cfg2; do
steps=20; [ $w = cfg1 ] && steps=300; [ $w = cfg4 ] && steps=60
python bench.py --workload $w --steps $steps --warmup 5 --no-cpu-baseline > gpurun_out/r02_run33_bench_${w}.json 2> gpurun_out/r02_run33_bench.err || tail -3 gpurun_out/r02_run33_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run33_bench_${w}.json"))
print("$w", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s e2e", round(d["e2e"]["value"],1), d["clocks"])
PY
done
