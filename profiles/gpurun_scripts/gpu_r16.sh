mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q -p no:cacheprovider -x 2>&1 | tail -2
for D in 0 5 69 64 68; do
PS_PAIR_DBG=$D timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r16_dbg$D.log 2>&1
echo "dbg=$D $(tail -1 gpurun_out/r16_dbg$D.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["ms_per_step"], j["roofline"]["avg_launch_ms"], j["clocks"])')"
done
