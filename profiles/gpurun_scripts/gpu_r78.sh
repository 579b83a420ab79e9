mkdir -p gpurun_out
export PS_CUDA_GRAPH=0
python profiles/gpurun_scripts/unet_once.py tse_unet_tcn_v0 64 2 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r78_unet_launches.csv python profiles/gpurun_scripts/unet_once.py tse_unet_tcn_v0 64 2 > gpurun_out/r78_ncu.log 2>&1
tail -1 gpurun_out/r78_ncu.log
