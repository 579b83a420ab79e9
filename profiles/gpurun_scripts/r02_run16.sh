# round 2, run 16: cfg3 / cfg4 / cfg1b status after the round's kernel changes; launch list of one cfg2 step (ncu)
mkdir -p gpurun_out
for w in cfg3 cfg4 cfg1b; do
python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run16_bench_$w.json 2> gpurun_out/r02_run16_bench.err || tail -3 gpurun_out/r02_run16_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run16_bench_$w.json")); r=d["roofline"]
print("$w", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s; e2e", round(d["e2e"]["value"],1), d["clocks"])
for o in [r]+r["other_kernels"]:
    print("    ", o["kernel"][:90], o["bound"], "frac", round(o["frac"],3), round(o["avg_launch_ms"],4), "ms share", round(o["share_of_step"],3))
PY
done
PS_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/r02_run16_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run16_ncu.log 2>&1; tail -2 gpurun_out/r02_run16_ncu.log | cut -c1-200
