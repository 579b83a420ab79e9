mkdir -p gpurun_out
python profiles/gpurun_scripts/model_breakdown.py ns_dparn_v0 2>&1 | grep -v Warn | cut -c1-72,120-200 | tail -22 | tee gpurun_out/r88_dparn_breakdown.txt
