# round 2, run 57: U-Net tap buffers with only their padding zeroed - coverage check with NaN-poisoned buffers, then plain tests and steps
mkdir -p gpurun_out
PS_UNET_POISON=1 timeout 1200 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py -q -x -k "unet or dpcrn or dparn or verbose" > gpurun_out/r02_run57_pytest_poison.log 2>&1; echo "poison rc=$?"; tail -3 gpurun_out/r02_run57_pytest_poison.log
timeout 1200 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py tests/test_gpu_graph.py -q -x > gpurun_out/r02_run57_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_run57_pytest.log
for w in tse_unet_tcn_v0 ns_dpcrn_v0 ns_dparn_v0; do
python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run57_bench_$w.json 2> gpurun_out/r02_run57_bench.err || tail -3 gpurun_out/r02_run57_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run57_bench_$w.json"))
print("$w", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["clocks"]["sm_mhz"])
PY
done
