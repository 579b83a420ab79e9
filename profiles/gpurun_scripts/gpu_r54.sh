mkdir -p gpurun_out
cp puresound_b200/libpuresound_b200.so /tmp/lib_cur.so
for V in hint base hint base; do
cp gpurun_ab/lib_$V.so puresound_b200/libpuresound_b200.so
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r54_bench_$V.log 2>&1; echo "$V: $(tail -1 gpurun_out/r54_bench_$V.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["value"]), round(j["ms_per_step"],3), round(j["roofline"]["avg_launch_ms"],4), j["clocks"]["sm_mhz"])')"
done
cp /tmp/lib_cur.so puresound_b200/libpuresound_b200.so
