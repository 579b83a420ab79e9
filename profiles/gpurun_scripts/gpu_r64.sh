mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "lstm" 2>&1 | tail -5
timeout 300 python -m pytest tests/test_gpu_modules.py -q -k "dprnn or verbose" 2>&1 | tail -3
for H in 128 64; do for s in 2 4 6 8; do PS_LSTM_SPQ=$s timeout 120 python profiles/gpurun_scripts/lstm_spq_bench.py $H; done; done 2>&1 | grep "H=" | tee gpurun_out/r64_lstm_spq.txt
