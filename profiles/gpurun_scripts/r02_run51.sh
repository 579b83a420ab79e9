# round 2, run 51: packed fp32 operations in the TMA depthwise conv - tests, cfg2 step, isolated kernel time under ncu; guard-band test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_gemm_tc.py tests/test_gpu_full.py -q -x > gpurun_out/r02_run51_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02_run51_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run51_bench_cfg2.json 2> gpurun_out/r02_run51_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run51_bench_cfg2.json")); r=d["roofline"]
print("cfg2", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["clocks"])
for o in [r]+r["other_kernels"]:
    print("    ", o["kernel"][:80], o["bound"], "frac", round(o["frac"],3), round(o["avg_launch_ms"],4), "ms share", round(o["share_of_step"],3))
PY
PS_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum --clock-control none -k regex:dwconv_tma -s 30 -c 8 --csv --log-file gpurun_out/r02_run51_dwconv_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run51_ncu.log 2>&1
grep -E "gpu__time_duration|issue_active|inst_executed" gpurun_out/r02_run51_dwconv_ncu.csv | awk -F'","' '{print $(NF-2), $NF}' | head -24
