mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r32_bench_2gpu.log 2>&1; echo "exit $?" >> gpurun_out/r32_bench_2gpu.log
tail -3 gpurun_out/r32_bench_2gpu.log | cut -c1-700
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r32_ref_2gpu.log 2>&1; echo "exit $?" >> gpurun_out/r32_ref_2gpu.log
tail -2 gpurun_out/r32_ref_2gpu.log | cut -c1-400
