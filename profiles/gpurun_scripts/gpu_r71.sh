mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r71_bench_4gpu.log 2>&1
grep '^{' gpurun_out/r71_bench_4gpu.log | tail -1 | cut -c1-400
