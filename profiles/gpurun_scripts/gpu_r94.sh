mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py -q -x -p no:cacheprovider -k "mel or rnn_and_mel" -s 2>&1 | tail -25
for W in tse_skim_v2_causal; do
timeout 600 python profiles/gpurun_scripts/veve_once.py $W 3 2>&1 | tail -2
done
echo done
