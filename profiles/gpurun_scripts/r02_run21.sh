# round 2, run 21: few-channel tcgen05 GEMM (gemm_rows_kernel) - unit tests, whole suite, A/B on cfg3 / cfg1b / cfg2 / U-Net
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q > gpurun_out/r02_run21_pytest_gemm.log 2>&1; echo "gemm rc=$?"; tail -15 gpurun_out/r02_run21_pytest_gemm.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_run21_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run21_pytest.log; tail -6 gpurun_out/r02_run21_pytest.log
for w in cfg3 cfg1b cfg2 tse_unet_tcn_v0; do for v in 0 1; do
PS_GEMM_ROWS=$v timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run21_bench_${w}_rows$v.json 2> gpurun_out/r02_run21_bench.err || tail -3 gpurun_out/r02_run21_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_run21_bench_${w}_rows$v.json")); r=d["roofline"]
    print("$w PS_GEMM_ROWS=$v", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["clocks"]["sm_mhz"])
    for o in ([r]+r["other_kernels"]) if r else []:
        print("    ", o["kernel"][:80], o["bound"], "frac", round(o["frac"],3), round(o["avg_launch_ms"],4), "ms share", round(o["share_of_step"],3))
except Exception as e: print("$w $v failed", e)
PY
done; done
