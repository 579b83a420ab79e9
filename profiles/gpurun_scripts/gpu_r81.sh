export PS_STFT_EXACT=0
timeout 900 python -m pytest tests/test_gpu_full.py tests/test_gpu_modules.py -q -s -k "cfg4 or unet or wrappers or speech_cfg4 or conv_stft" 2>&1 | grep -v "^$" | tail -24
