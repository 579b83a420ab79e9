mkdir -p gpurun_out
for D in 0 16 32 48 21 38 6; do
PS_PAIR_DBG=$D timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r14_dbg$D.log 2>&1
echo "dbg=$D $(tail -1 gpurun_out/r14_dbg$D.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["ms_per_step"], j["roofline"]["avg_launch_ms"], j["clocks"])')"
done
