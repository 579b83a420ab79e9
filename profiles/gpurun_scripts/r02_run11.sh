# round 2, run 11: TMA sliding-window depthwise conv - tests, A/B in the cfg2 step, launch list; multi-device capture fix
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sharded.py -x -q > gpurun_out/r02_run11_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run11_pytest.log; tail -6 gpurun_out/r02_run11_pytest.log
for v in 0 1; do
PS_DW_TMA=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run11_bench_cfg2_dwtma$v.json 2> gpurun_out/r02_run11_bench_cfg2.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run11_bench_cfg2_dwtma$v.json")); r=d["roofline"]
dw=[x for x in r["other_kernels"] if "dwconv" in x["kernel"]][0]
print("PS_DW_TMA=$v cfg2", round(d["ms_per_step"],2), "ms/step; gemm", round(r["avg_launch_ms"],4), "; dwconv", round(dw["avg_launch_ms"],4), "ms frac", round(dw["frac"],3), d["clocks"])
PY
done
timeout 900 python -m pytest tests/test_gpu_full.py tests/test_gpu_modules.py -x -q > gpurun_out/r02_run11_pytest_full.log 2>&1; tail -3 gpurun_out/r02_run11_pytest_full.log
