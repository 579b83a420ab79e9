mkdir -p gpurun_out
export PS_LSTM_SPQ=8
timeout 120 python profiles/gpurun_scripts/lstm_spq_bench.py 128 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lstm_tc_kernel" -s 5 -c 1 -o gpurun_out/r65_lstm python profiles/gpurun_scripts/lstm_spq_bench.py 128 > gpurun_out/r65_ncu.log 2>&1
tail -2 gpurun_out/r65_ncu.log
