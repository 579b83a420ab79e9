for H in 128 64; do for s in 4 8; do PS_LSTM_DBG=1 PS_LSTM_SPQ=$s timeout 120 python profiles/gpurun_scripts/lstm_spq_bench.py $H; done; done 2>&1 | grep "H="
