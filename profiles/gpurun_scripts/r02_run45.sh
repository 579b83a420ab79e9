# round 2, run 45: EXPERIMENT - consecutive big launches walk their tensors in alternating directions (PS_ORDER_ALT), same-box A/B
mkdir -p gpurun_out
PS_ORDER_ALT=1 timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_full.py tests/test_gpu_ops.py -q -x 2>&1 | tail -2
for rep in 1 2; do for v in 0 1; do
PS_ORDER_ALT=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run45_bench_cfg2_alt${v}_$rep.json 2> gpurun_out/r02_run45_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run45_bench_cfg2_alt${v}_$rep.json")); r=d["roofline"]
dw=[o for o in r["other_kernels"] if "dwconv" in o["kernel"]][0]
print("cfg2 ORDER_ALT=$v rep $rep", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "gemm", round(r["avg_launch_ms"],4), "dwconv", round(dw["avg_launch_ms"],4), d["clocks"]["sm_mhz"])
PY
done; done
