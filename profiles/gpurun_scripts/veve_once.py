"""One eager forward of veve_dprnn_v0_causal at 32 x (10 s + 6 s) (for ncu launch lists; PS_CUDA_GRAPH=0)."""
import sys
import torch
sys.path.insert(0, ".")
from puresound_b200 import recipes, testing, ops
ops.require_device()
name = sys.argv[1] if len(sys.argv) > 1 else "veve_dprnn_v0_causal"
torch.manual_seed(0)
m = recipes.init_model(name, verbose=False).eval()
testing.perturb_(m, seed=1)
m = m.to("cuda")
mix = testing.noisy_speech(32, 160000, seed=1)[0].cuda()
enr = testing.noisy_speech(32, 96000, seed=2)[0].cuda()
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    y = m.inference(mix, enr)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
