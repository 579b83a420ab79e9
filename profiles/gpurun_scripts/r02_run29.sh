# round 2, run 29: one-block tiles with 64-k stages out of the two-block image for few-tile 512-channel GEMMs (PS_PAIR_FROM2): isolated sweep + cfg1 / cfg4 steps
mkdir -p gpurun_out
for shape in "1 3999 512 512" "2 3999 512 512" "4 3999 512 512" "8 3999 512 512" "16 3999 512 512" "32 3999 512 512" "64 497 512 256" "64 747 512 256"; do
set -- $shape
for thr in 0 1000; do
PROBE_REPS=100 PROBE_B=$1 PROBE_T=$2 PROBE_M=$3 PROBE_K=$4 PS_PAIR_FROM2=$thr python profiles/gemm_probe.py 2>&1 | tail -1
done; done | tee gpurun_out/r02_run29_from2_probe.txt
PS_PAIR_FROM2=1000 timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_full.py -q -x 2>&1 | tail -3
for w in cfg1 cfg4; do for thr in 0 1000; do
steps=60; [ $w = cfg1 ] && steps=300
PS_PAIR_FROM2=$thr python bench.py --workload $w --steps $steps --warmup 5 --no-cpu-baseline > gpurun_out/r02_run29_bench_${w}_from2_$thr.json 2> gpurun_out/r02_run29_bench.err || tail -3 gpurun_out/r02_run29_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run29_bench_${w}_from2_$thr.json"))
print("$w PS_PAIR_FROM2=$thr", round(d["ms_per_step"],3), "ms/step", round(d["value"],1), "audio-s/s", d["clocks"]["sm_mhz"])
PY
done; done
