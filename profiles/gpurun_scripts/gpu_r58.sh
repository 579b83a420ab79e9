mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py -q -k "lstm or dprnn or skim" 2>&1 | tail -15
python -m pytest tests/test_gpu_full.py -q 2>&1 | tail -5
python - <<'PY' > gpurun_out/r58_veve.log 2>&1
import torch, time
from puresound_b200 import recipes, testing, ops
ops.require_device()
torch.manual_seed(0)
m = recipes.init_model("veve_dprnn_v0_causal", verbose=False).eval()
testing.perturb_(m, seed=1)
m = m.to("cuda")
mix = testing.noisy_speech(32, 160000, seed=1)[0].cuda()
enr = testing.noisy_speech(32, 96000, seed=2)[0].cuda()
for _ in range(4): y = m.inference(mix, enr)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): y = m.inference(mix, enr)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 5 * 1e3
print(f"veve_dprnn_v0_causal 32 x (10 s mix + 6 s enroll): {ms:.2f} ms/step = {32*10/(ms/1e3):.0f} audio-s/s")
PY
tail -3 gpurun_out/r58_veve.log
python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r58_cfg3.json 2> gpurun_out/r58_cfg3.err; tail -c 600 gpurun_out/r58_cfg3.json
