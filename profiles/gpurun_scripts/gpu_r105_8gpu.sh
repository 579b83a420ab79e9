mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r105_bench_cfg2_8gpu.log 2>&1; tail -1 gpurun_out/r105_bench_cfg2_8gpu.log | cut -c1-400
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r105_bench_cfg2_1gpu_same_box.log 2>&1; tail -1 gpurun_out/r105_bench_cfg2_1gpu_same_box.log | cut -c1-300
echo done
