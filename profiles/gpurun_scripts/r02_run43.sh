# round 2, run 43: out_conv form (residual) on the three CTA-pair variants at the cfg2 shape, same box
mkdir -p gpurun_out
for env in "PS_TC_WIDE=1 PS_PAIR_FROM2=10" "PS_TC_WIDE=0 PS_PAIR_FROM2=0" "PS_TC_WIDE=1 PS_PAIR_FROM2=1000000" "PS_TC_WIDE=1 PS_PAIR_FROM2=10"; do
env $env PROBE_REPS=50 python profiles/gemm_probe.py 2>&1 | tail -1 | sed "s/^/$env | /"
done | tee gpurun_out/r02_run43_out_conv_variants.txt
