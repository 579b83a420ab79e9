# round 2, run 32: fused statistics finalize re-measured at few-frame sizes (the TMA depthwise conv now supports it)
mkdir -p gpurun_out
PS_FUSE_FINALIZE=1 timeout 900 python -m pytest tests/test_gpu_full.py tests/test_gpu_modules.py tests/test_gpu_ops.py -q -x 2>&1 | tail -3
for w in cfg1 cfg4; do for v in 0 1; do
steps=60; [ $w = cfg1 ] && steps=300
PS_FUSE_FINALIZE=$v python bench.py --workload $w --steps $steps --warmup 5 --no-cpu-baseline > gpurun_out/r02_run32_bench_${w}_fuse$v.json 2> gpurun_out/r02_run32_bench.err || tail -3 gpurun_out/r02_run32_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run32_bench_${w}_fuse$v.json"))
print("$w PS_FUSE_FINALIZE=$v", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["gpu_launches"]//d["steps"], "launches/step", d["clocks"]["sm_mhz"])
PY
done; done
