mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active,temperature.gpu --format=csv -lms 100 > gpurun_out/r19_clocks.csv &
SMI=$!
sleep 1
timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r19_bench.log 2>&1
sleep 1
PS_PAIR_DBG=7 timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r19_bench7.log 2>&1
kill $SMI
tail -1 gpurun_out/r19_bench.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["ms_per_step"], j["roofline"]["avg_launch_ms"], j["clocks"])'
tail -1 gpurun_out/r19_bench7.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["ms_per_step"], j["roofline"]["avg_launch_ms"], j["clocks"])'
