mkdir -p gpurun_out
export PS_CUDA_GRAPH=0
python profiles/gpurun_scripts/veve_once.py veve_dprnn_v0_causal 2 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r60_veve_launches.csv python profiles/gpurun_scripts/veve_once.py veve_dprnn_v0_causal 2 > gpurun_out/r60_ncu.log 2>&1
tail -2 gpurun_out/r60_ncu.log
