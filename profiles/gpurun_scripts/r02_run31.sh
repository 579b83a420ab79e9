# round 2, run 31: batch 1 (cfg1) - depthwise-conv kernel A/B and the launch list of one eager forward
mkdir -p gpurun_out
for v in 1 0; do
PS_DW_TMA=$v python bench.py --workload cfg1 --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run31_bench_cfg1_dwtma$v.json 2> gpurun_out/r02_run31_bench.err || tail -3 gpurun_out/r02_run31_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run31_bench_cfg1_dwtma$v.json"))
print("cfg1 PS_DW_TMA=$v", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["clocks"])
PY
done
PS_CUDA_GRAPH=0 python bench.py --workload cfg1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run31_plain.log 2>&1 &&
PS_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r02_run31_launches_cfg1.csv python bench.py --workload cfg1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run31_ncu.log 2>&1; tail -2 gpurun_out/r02_run31_ncu.log
python profiles/launch_summary.py gpurun_out/r02_run31_launches_cfg1.csv | head -30
