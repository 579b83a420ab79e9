# round 2, run 46 (2 GPUs): the bench under torchrun with the driver's launch line and a short timed region (sampler extension path)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_run46_bench_cfg2_2gpu.json 2> gpurun_out/r02_run46_bench_2gpu.err; echo "rc=$?"
tail -1 gpurun_out/r02_run46_bench_cfg2_2gpu.json | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29734 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02_run46_bench_reference_2gpu.json 2> gpurun_out/r02_run46_ref_2gpu.err; echo "ref rc=$?"
tail -1 gpurun_out/r02_run46_bench_reference_2gpu.json | cut -c1-300
