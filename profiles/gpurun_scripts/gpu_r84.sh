mkdir -p gpurun_out
python profiles/gpurun_scripts/model_breakdown.py ns_dpcrn_v0 2>&1 | grep -v Warn | cut -c1-72,120-200 | tail -22 | tee gpurun_out/r84_dpcrn_breakdown.txt
