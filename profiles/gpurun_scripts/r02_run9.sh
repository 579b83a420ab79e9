# round 2, run 9: wide GEMM v2 + sequential bulk L2 prefetch of the next tile (operand rows + residual rows)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q -k wide > gpurun_out/r02_run9_pytest_gemm.log 2>&1; tail -2 gpurun_out/r02_run9_pytest_gemm.log
python profiles/gemm_probe.py 2>&1 | tail -1 | tee gpurun_out/r02_run9_probe.txt
PS_TC_WIDE=0 python profiles/gemm_probe.py 2>&1 | tail -1 | tee -a gpurun_out/r02_run9_probe.txt
for i in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run9_bench_cfg2_$i.json 2> gpurun_out/r02_run9_bench_cfg2.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run9_bench_cfg2_$i.json")); r=d["roofline"]
print("cfg2", round(d["ms_per_step"],2), "ms/step; gemm", round(r["avg_launch_ms"],4), "ms frac", round(r["frac"],3), "issued", round(r["issued_frac"],3), d["clocks"])
PY
done
