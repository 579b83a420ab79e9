bash profiles/gpurun_scripts/gpu_r96.sh
bash profiles/gpurun_scripts/gpu_r95.sh
