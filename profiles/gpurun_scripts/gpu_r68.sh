mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "lstm" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py tests/test_gpu_streaming.py -q -k "dprnn or verbose or cfg3 or veve or skim" 2>&1 | tail -3
python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r68_cfg3.json 2> gpurun_out/r68_cfg3.err; python -c "
import json;d=json.load(open('gpurun_out/r68_cfg3.json'));print('cfg3',d['value'],d['ms_per_step'],d['e2e']['value'])"
bash -c "$(sed -n '/^python - <<.PY. > gpurun_out\/r58_veve.log/,/^PY$/p' profiles/gpurun_scripts/gpu_r58.sh | sed 's/r58_veve/r68_veve/')"; tail -1 gpurun_out/r68_veve.log
