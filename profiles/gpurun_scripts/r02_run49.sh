# round 2, run 49: few-channel GEMM on 129 ... 256 channels (PS_GEMM_ROWS_M256 opt-in) - tests, A/B on cfg4 / cfg4_gated / skim / cfg2
mkdir -p gpurun_out
PS_GEMM_ROWS_M256=1 timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_full.py tests/test_gpu_modules.py -q -x > gpurun_out/r02_run49_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_run49_pytest.log
for w in cfg4 cfg4_gated tse_skim_v0_causal ns_dparn_v0; do for v in 0 1; do
steps=10; [ $w = cfg4 ] && steps=60; [ $w = cfg4_gated ] && steps=40
PS_GEMM_ROWS_M256=$v python bench.py --workload $w --steps $steps --warmup 5 --no-cpu-baseline > gpurun_out/r02_run49_bench_${w}_m256_$v.json 2> gpurun_out/r02_run49_bench.err || tail -3 gpurun_out/r02_run49_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_run49_bench_${w}_m256_$v.json")); r=d["roofline"]
    print("$w M256=$v", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["clocks"]["sm_mhz"])
    for o in ([r]+r["other_kernels"])[:4]:
        print("    ", o["kernel"][:70], o["bound"], "frac", round(o["frac"],3), round(o["avg_launch_ms"],4), "ms share", round(o["share_of_step"],3))
except Exception as e: print("$w $v failed", e)
PY
done; done
