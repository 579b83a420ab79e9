mkdir -p gpurun_out
python -m pytest tests/test_gpu_full.py -q -s -k "real_speech" 2>&1 | grep -v "^$" | tail -12
python - <<'PY' > gpurun_out/r61_skim.log 2>&1
import torch, time
from puresound_b200 import recipes, testing, ops
ops.require_device()
for name in ("tse_skim_v0_causal", "tse_skim_v0"):
    torch.manual_seed(0)
    m = recipes.init_model(name, verbose=False).eval()
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
    print(f"{name} 32 x (10 s mix + 6 s enroll): {ms:.2f} ms/step = {32*10/(ms/1e3):.0f} audio-s/s")
PY
tail -3 gpurun_out/r61_skim.log
