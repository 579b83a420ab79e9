mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r47_veve.log 2>&1
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
tail -3 gpurun_out/r47_veve.log
PS_CUDA_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/r47_launches_veve.csv python - <<'PY' > gpurun_out/r47_ncu.log 2>&1
import torch
from puresound_b200 import recipes, testing, ops
torch.manual_seed(0)
m = recipes.init_model("veve_dprnn_v0_causal", verbose=False).eval().to("cuda")
mix = testing.noisy_speech(32, 160000, seed=1)[0].cuda()
enr = testing.noisy_speech(32, 96000, seed=2)[0].cuda()
for _ in range(3): y = m.inference(mix, enr)
torch.cuda.synchronize()
PY
