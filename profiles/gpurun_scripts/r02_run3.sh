# round 2, run 3: wide (256-frame) pair GEMM - correctness, then A/B against the 128-frame kernel on cfg2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q > gpurun_out/r02_run3_pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run3_pytest_gemm.log; tail -15 gpurun_out/r02_run3_pytest_gemm.log
timeout 900 python -m pytest tests -m gpu -q -k "benched or speaker_embedding or td_tse or full_size_parity or sharded or graph" > gpurun_out/r02_run3_pytest_sel.log 2>&1; echo "rc=$?" >> gpurun_out/r02_run3_pytest_sel.log; tail -8 gpurun_out/r02_run3_pytest_sel.log
PS_TC_WIDE=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run3_bench_cfg2_narrow.json 2> gpurun_out/r02_run3_bench_cfg2_narrow.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run3_bench_cfg2_wide.json 2> gpurun_out/r02_run3_bench_cfg2_wide.err
PS_TC_WIDE=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run3_bench_cfg2_narrow2.json 2> gpurun_out/r02_run3_bench_cfg2_narrow2.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run3_bench_cfg2_wide2.json 2> gpurun_out/r02_run3_bench_cfg2_wide2.err
for f in narrow wide narrow2 wide2; do python - <<PY
import json
d=json.load(open("gpurun_out/r02_run3_bench_cfg2_$f.json")); r=d["roofline"]
print("$f", round(d["ms_per_step"],2), "ms/step; gemm", round(r["avg_launch_ms"],4), "ms frac", round(r["frac"],3), "issued", round(r["issued_frac"],3), d["clocks"])
PY
done
