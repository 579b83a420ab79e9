# round 2, run 42: same-box A/B of the out_conv epilogue with the residual requested one chunk ahead (PS_WIDE_RESPF)
mkdir -p gpurun_out
for rep in 1 2; do for v in 0 1; do
PS_WIDE_RESPF=$v PROBE_REPS=50 python profiles/gemm_probe.py 2>&1 | tail -1 | sed "s/^/RESPF=$v | /"
PS_WIDE_RESPF=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run42_bench_cfg2_respf${v}_$rep.json 2> gpurun_out/r02_run42_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run42_bench_cfg2_respf${v}_$rep.json")); r=d["roofline"]
print("cfg2 RESPF=$v rep $rep", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "gemm", round(r["avg_launch_ms"],4), d["clocks"]["sm_mhz"])
PY
done; done
