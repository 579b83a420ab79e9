# round 2, run 24: gemm_rows_kernel with the 4-deep half-block register pipeline and early residual loads
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q > gpurun_out/r02_run24_pytest_gemm.log 2>&1; echo "gemm rc=$?"; tail -5 gpurun_out/r02_run24_pytest_gemm.log
for w in cfg3 cfg2; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run24_bench_${w}.json 2> gpurun_out/r02_run24_bench.err || tail -3 gpurun_out/r02_run24_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_run24_bench_${w}.json")); r=d["roofline"]
    print("$w", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["clocks"]["sm_mhz"])
    for o in ([r]+r["other_kernels"]) if r else []:
        print("    ", o["kernel"][:80], o["bound"], "frac", round(o["frac"],3), round(o["avg_launch_ms"],4), "ms share", round(o["share_of_step"],3))
except Exception as e: print("$w failed", e)
PY
done
