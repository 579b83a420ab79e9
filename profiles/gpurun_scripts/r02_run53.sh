# round 2, run 53: per-kernel device time of one eager forward of the U-Net / DPCRN recipes (where do memsets and copies stand)
mkdir -p gpurun_out
for w in tse_unet_tcn_v0 ns_dpcrn_v0; do
python profiles/gpurun_scripts/model_breakdown.py $w > gpurun_out/r02_run53_breakdown_$w.txt 2>&1; grep -E "Name|ps::|Memset|Memcpy|at::|Self CUDA" gpurun_out/r02_run53_breakdown_$w.txt | cut -c1-72,110-200 | head -22
done
