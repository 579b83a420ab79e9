# round 2, run 37: which CTA-pair variant serves the write-bound LSTM input projection (646,400 x 1024 x 128) best
mkdir -p gpurun_out
for shape in "32 20200 1024 128" "64 3999 512 128" "32 20200 256 128"; do
set -- $shape
for env in "PS_TC_WIDE=1 PS_PAIR_FROM2=10" "PS_TC_WIDE=0 PS_PAIR_FROM2=0" "PS_TC_WIDE=1 PS_PAIR_FROM2=1000000"; do
env $env PROBE_REPS=30 PROBE_B=$1 PROBE_T=$2 PROBE_M=$3 PROBE_K=$4 python profiles/gemm_probe.py 2>&1 | tail -1 | sed "s/^/$env | /"
done; done | tee gpurun_out/r02_run37_gx_gemm_probe.txt
