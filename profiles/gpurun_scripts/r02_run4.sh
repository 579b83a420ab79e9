# round 2, run 4: stream-removal experiments on the wide GEMM (experiments build; results are garbage, only times matter) + ncu of the release kernel
mkdir -p gpurun_out
export PS_B200_LIB=$PWD/puresound_b200/libpuresound_b200_exp.so
for dbg in 0 1 2 4 8 16 3 7 15 31 23; do
  PS_WIDE_DBG=$dbg python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run4_dbg$dbg.json 2> gpurun_out/r02_run4_dbg$dbg.err || tail -3 gpurun_out/r02_run4_dbg$dbg.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_run4_dbg$dbg.json")); r=d["roofline"]
    g=[r]+r["other_kernels"]
    g=[x for x in g if "512 x 512)" in x["kernel"]][0]
    print("dbg=$dbg step", round(d["ms_per_step"],2), "gemm", round(g["avg_launch_ms"],4), d["clocks"]["sm_mhz"])
except Exception as e: print("dbg=$dbg failed", e)
PY
done
unset PS_B200_LIB
PS_CUDA_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_wide -s 30 -c 3 -o gpurun_out/r02_run4_gemm_wide python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run4_ncu.log 2>&1; tail -3 gpurun_out/r02_run4_ncu.log
ls -la gpurun_out/*.ncu-rep
