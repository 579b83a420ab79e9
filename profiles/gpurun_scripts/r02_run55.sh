# round 2, run 55: stats_region fast path - tests, U-Net / DPCRN steps
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py tests/test_gpu_full.py -q -x > gpurun_out/r02_run55_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_run55_pytest.log
for w in tse_unet_tcn_v0 ns_dpcrn_v0; do
python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run55_bench_$w.json 2> gpurun_out/r02_run55_bench.err || tail -3 gpurun_out/r02_run55_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run55_bench_$w.json"))
print("$w", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["clocks"]["sm_mhz"])
PY
done
python profiles/gpurun_scripts/model_breakdown.py tse_unet_tcn_v0 2>&1 | grep -E "stats_region|gated_kernel<4|Self CUDA" | cut -c1-72,110-200
