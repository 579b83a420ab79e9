# round 2, run 48: do the pollers cost power?  cfg2 step with longer mbarrier suspend hints / relaxed-wait sleeps (A/B libraries built
# with -DPS_MBAR_HINT_NS / -DPS_RELAXED_SLEEP_NS; release defaults 2000 / 64 ns), same box, two rounds
mkdir -p gpurun_out
for rep in 1 2; do for v in release a b c; do
if [ $v = release ]; then unset PS_B200_LIB; else export PS_B200_LIB=$PWD/puresound_b200/libps_ab_$v.so; fi
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run48_bench_cfg2_${v}_$rep.json 2> gpurun_out/r02_run48_bench.err || tail -2 gpurun_out/r02_run48_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run48_bench_cfg2_${v}_$rep.json")); r=d["roofline"]
print("cfg2 lib=$v rep $rep", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "gemm", round(r["avg_launch_ms"],4), d["clocks"]["sm_mhz"])
PY
done; done
