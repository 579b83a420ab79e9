# round 2, run 35 (8 GPUs): strong-scaling bench at N = 8 / 4 / 2 / 1 on one box (one batch of 64 split over the ranks; the weak
# figure beside it), the sharded tests on two GPUs, ShardedSeparator over 1 / 2 / 4 / 8 devices with the deviation from one device
mkdir -p gpurun_out
for n in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_run35_bench_cfg2_${n}gpu.json 2> gpurun_out/r02_run35_bench_${n}gpu.err
python - <<PY
import json
try:
    txt=open("gpurun_out/r02_run35_bench_cfg2_${n}gpu.json").read(); d=json.loads([l for l in txt.splitlines() if l.startswith("{")][0])
    print("N=$n", d["scaling"], round(d["value"],1), "audio-s/s", round(d["ms_per_step"],3), "ms/step; e2e", round(d["e2e"]["value"],1), "; weak", d["weak"], d["clocks"])
except Exception as e: print("N=$n failed", e)
PY
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run35_bench_cfg2_1gpu.json 2> gpurun_out/r02_run35_bench_1gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run35_bench_cfg2_1gpu.json")); print("N=1", round(d["value"],1), "audio-s/s", round(d["ms_per_step"],3), "ms/step; e2e", round(d["e2e"]["value"],1), d["clocks"])
PY
timeout 600 python -m pytest tests/test_gpu_sharded.py -q > gpurun_out/r02_run35_pytest_sharded.log 2>&1; tail -3 gpurun_out/r02_run35_pytest_sharded.log
python - <<'PY' > gpurun_out/r02_run35_sharded_probe.txt 2>&1
import time, torch
from puresound_b200 import recipes, testing, ops
from puresound_b200.sharding import ShardedSeparator
ops.require_device()
torch.manual_seed(0)
m = recipes.baseline_config("cfg2").eval(); testing.perturb_(m, seed=1); m = m.to("cuda:0")
x = testing.noisy_speech(64, 64000, seed=1234)[0].pin_memory()
ref = None
for devs in ([0], [0, 1], [0, 1, 2, 3], list(range(8))):
    sep = ShardedSeparator(m, devs)
    for _ in range(4): y = sep.inference(x, reuse_output=True)
    if ref is None: ref = y.clone()
    dev = float((y - ref).abs().max())
    t0 = time.perf_counter()
    for _ in range(20): y = sep.inference(x, reuse_output=True)
    ms = (time.perf_counter() - t0) * 50
    print(f"ShardedSeparator devices={len(devs)}: {ms:.2f} ms per 64 x 4 s host batch (H2D + forward + D2H) = {256/(ms/1e3):.0f} audio-s/s; max |y - y(1 device)| = {dev:.3e}")
    del sep
PY
cat gpurun_out/r02_run35_sharded_probe.txt | tail -5
