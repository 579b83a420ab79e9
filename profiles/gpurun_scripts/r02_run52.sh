# round 2, run 52: GatedTCN zeroes only the pad rows of its padded buffers - module / full-size tests, cfg4_gated and tse_unet_tcn_v0 steps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py tests/test_gpu_graph.py -q -x > gpurun_out/r02_run52_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_run52_pytest.log
for w in cfg4_gated tse_unet_tcn_v0; do
python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run52_bench_$w.json 2> gpurun_out/r02_run52_bench.err || tail -3 gpurun_out/r02_run52_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run52_bench_$w.json"))
print("$w", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["clocks"]["sm_mhz"])
PY
done
