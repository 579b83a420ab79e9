mkdir -p gpurun_out
for H in 128 64; do for s in 2 4 8 12 16; do PS_LSTM_SPQ=$s python profiles/gpurun_scripts/lstm_spq_bench.py $H; done; done 2>&1 | grep "H=" | tee gpurun_out/r62_lstm_spq.txt
