# round 2, run 7: wide GEMM v2 (TMA raw ring + three rings + residual prefetch): tests, isolated probe, stream-removal, cfg2 bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q > gpurun_out/r02_run7_pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run7_pytest_gemm.log; tail -6 gpurun_out/r02_run7_pytest_gemm.log
python profiles/gemm_probe.py 2>&1 | tail -1 | tee gpurun_out/r02_run7_probe.txt
export PS_B200_LIB=$PWD/puresound_b200/libpuresound_b200_exp.so
for dbg in 0 16 1 2 4 8 20 31; do PS_WIDE_DBG=$dbg timeout 120 python profiles/gemm_probe.py 2>&1 | tail -1; done | tee -a gpurun_out/r02_run7_probe.txt
unset PS_B200_LIB
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run7_bench_cfg2.json 2> gpurun_out/r02_run7_bench_cfg2.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run7_bench_cfg2.json")); r=d["roofline"]
print("cfg2", round(d["ms_per_step"],2), "ms/step; gemm", round(r["avg_launch_ms"],4), "ms frac", round(r["frac"],3), "issued", round(r["issued_frac"],3), d["clocks"])
PY
timeout 600 python -m pytest tests/test_gpu_full.py -x -q -k "benched or full_size" > gpurun_out/r02_run7_pytest_full.log 2>&1; tail -3 gpurun_out/r02_run7_pytest_full.log
