mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -x -q -k "verbose_probe" > gpurun_out/r02_run19_a.log 2>&1; tail -3 gpurun_out/r02_run19_a.log
grep -n "^puresound_b200\|^tests\|Error" gpurun_out/r02_run19_a.log | head -30
echo "== PS_PDL=1"
PS_PDL=1 timeout 600 python -m pytest tests/test_gpu_modules.py -x -q -k "verbose_probe" 2>&1 | tail -3
echo "== again default"
timeout 600 python -m pytest tests/test_gpu_modules.py -q -k "verbose_probe" 2>&1 | tail -8
