# round 2, run 10 (2 GPUs): ShardedSeparator across two devices + strong-scaling bench at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -q > gpurun_out/r02_run10_pytest_sharded.log 2>&1; tail -3 gpurun_out/r02_run10_pytest_sharded.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_run10_bench_cfg2_2gpu.json 2> gpurun_out/r02_run10_bench_cfg2_2gpu.err
tail -c 1500 gpurun_out/r02_run10_bench_cfg2_2gpu.json; tail -3 gpurun_out/r02_run10_bench_cfg2_2gpu.err
python - <<'PY' > gpurun_out/r02_run10_sharded_probe.txt 2>&1
import time, torch
from puresound_b200 import recipes, testing, ops
from puresound_b200.sharding import ShardedSeparator
ops.require_device()
torch.manual_seed(0)
m = recipes.baseline_config("cfg2").eval(); testing.perturb_(m, seed=1); m = m.to("cuda:0")
x = testing.noisy_speech(64, 64000, seed=1234)[0].pin_memory()
for devs in ([0], [0, 1]):
    sep = ShardedSeparator(m, devs)
    for _ in range(4): y = sep.inference(x, reuse_output=True)
    t0 = time.perf_counter()
    for _ in range(10): y = sep.inference(x, reuse_output=True)
    ms = (time.perf_counter() - t0) * 100
    print(f"ShardedSeparator devices={devs}: {ms:.2f} ms per 64 x 4 s batch host-to-host = {256/(ms/1e3):.0f} audio-s/s")
PY
cat gpurun_out/r02_run10_sharded_probe.txt
