# round 2, run 36: streamed-weight mode of gemm_rows_kernel (PS_GEMM_ROWS_WS A/B), clock sampler started before the warm-up
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q > gpurun_out/r02_run36_pytest_gemm.log 2>&1; echo "gemm rc=$?"; tail -5 gpurun_out/r02_run36_pytest_gemm.log
for w in cfg1b tse_unet_tcn_v0 ns_dpcrn_v0 ns_dparn_v0; do for v in 0 1; do
PS_GEMM_ROWS_WS=$v timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run36_bench_${w}_ws$v.json 2> gpurun_out/r02_run36_bench.err || tail -3 gpurun_out/r02_run36_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_run36_bench_${w}_ws$v.json")); r=d["roofline"]
    print("$w PS_GEMM_ROWS_WS=$v", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["clocks"])
    for o in ([r]+r["other_kernels"]) if r else []:
        print("    ", o["kernel"][:80], o["bound"], "frac", round(o["frac"],3), round(o["avg_launch_ms"],4), "ms share", round(o["share_of_step"],3))
except Exception as e: print("$w $v failed", e)
PY
done; done
python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run36_bench_cfg1_short.json 2> gpurun_out/r02_run36_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_run36_bench_cfg1_short.json')); print('cfg1 20 steps', round(d['ms_per_step'],3), d['clocks'])"
